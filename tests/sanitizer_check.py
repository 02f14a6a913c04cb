"""Small end-to-end run for `compute-sanitizer` (memcheck / racecheck / synccheck), by hand on the GPU box:

    compute-sanitizer --tool memcheck python tests/sanitizer_check.py

Exercises every kernel once on small shapes: insert (both normalise kernels), gemv and tensor-core scans with
several levels, dense level, warp-pivot / block-pivot / bitonic / radix selects, single-CTA and CTA-pair tensor-core scans, finalize, wide rescoring, exact repair, row filter.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import flat_search as fs  # noqa: E402

pkg = importlib.import_module("autostyle-tts_b200")


def main():
    rng = np.random.default_rng(0)
    ok = True
    only = [int(x) for x in os.environ.get("SAN_CASES", "").split(",") if x.strip()]   # e.g. SAN_CASES=6,7 for racecheck
    for case, (n, d, nq, k, metric, force) in enumerate([(6000, 64, 2, 10, "COSINE", 0), (6000, 64, 40, 10, "COSINE", 0), (5000, 100, 20, 100, "IP", 0),
                                         (3000, 6148, 3, 5, "COSINE", 0), (4000, 64, 5, 10, "COSINE", 2), (40000, 32, 260, 10, "COSINE", 1),
                                         (130000, 32, 1, 100, "COSINE", 0), (9000, 32, 2, 50, "IP", 0), (70000, 64, 100, 10, "COSINE", 0)]):
        if only and case not in only:
            continue
        X = rng.standard_normal((n, d)).astype(np.float32)
        Q = rng.standard_normal((nq, d)).astype(np.float32)
        ids = np.arange(n, dtype=np.int64)
        st = pkg.Store(d, metric, capacity=n // 2)
        st.insert(X, ids)
        st.set_option("force_repair", force)
        got_ids, _ = st.search(Q, k)
        exp_ids, _, _ = fs.search(X, ids, Q, k, metric)
        good = np.array_equal(got_ids, exp_ids)
        mask = rng.random(n) < 0.3
        st.set_filter(mask)
        f_ids, _ = st.search(Q, k)
        rows = np.nonzero(mask)[0]
        ef_ids, _, _ = fs.search(X[rows], rows.astype(np.int64), Q, k, metric)
        good &= np.array_equal(f_ids, ef_ids)
        st.set_filter(None)
        print(n, d, nq, k, metric, force, "OK" if good else "MISMATCH", flush=True)
        ok &= good
        st.close()
    print("ALL OK" if ok else "FAILURES")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
