"""Bring-up / regression script for the tensor-core scan (run by hand under `timeout` on the GPU
box: a wrong barrier protocol hangs instead of failing).  Not collected by pytest.

    timeout 300 python tests/gemm_check.py [cta_group ...]
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import flat_search as fs  # noqa: E402

pkg = importlib.import_module("autostyle-tts_b200")


def run(cg, n, d, nq, k, metric="COSINE", seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    pick = rng.integers(0, n, size=max(1, nq // 4))
    Q[: pick.size] = X[pick] + 0.05 * rng.standard_normal((pick.size, d)).astype(np.float32)
    ids = np.arange(n, dtype=np.int64) * 2 + 1
    st = pkg.Store(d, metric, capacity=n)
    st.insert(X, ids)
    st.set_option("scan_path", 2)
    st.set_option("cta_group", cg)
    t0 = time.time()
    got_ids, got_d = st.search(Q, k)
    dt = time.time() - t0
    exp = fs.search_large(X, ids, Q, k, metric) if n > 20000 else fs.search(X, ids, Q, k, metric)
    ok_ids = np.array_equal(got_ids, exp[0])
    fin = np.isfinite(exp[1])
    ok_d = bool(np.all(np.abs(got_d[fin] - exp[1][fin]) <= 1e-5 * np.maximum(1, np.abs(exp[1][fin]))))
    rep, unc = st.stat("repaired_queries"), st.stat("uncertified_queries")
    print(f"cg={cg} n={n} d={d} nq={nq} k={k} {metric}: ids={'OK' if ok_ids else 'MISMATCH'} scores={'OK' if ok_d else 'BAD'} "
          f"repaired={rep} uncertified={unc} levels={st.stat('last_levels')} path={st.stat('last_scan_path')} t={dt*1e3:.1f}ms", flush=True)
    if not ok_ids:
        bad = np.argwhere(got_ids != exp[0])
        print("   first mismatches:", bad[:5].tolist(), got_ids[bad[0][0]][:5], exp[0][bad[0][0]][:5])
    st.close()
    return ok_ids and ok_d and unc == 0, rep


if __name__ == "__main__":
    groups = [int(a) for a in sys.argv[1:]] or [1, 2]
    all_ok = True
    for cg in groups:
        for cfg in [(1000, 64, 20, 10), (5000, 768, 64, 10), (2560, 128, 300, 5), (70000, 256, 130, 10),
                    (300000, 128, 40, 10), (20000, 1024, 257, 100), (4000, 100, 33, 1, "IP")]:
            ok, rep = run(cg, *cfg)
            all_ok &= ok
    print("ALL OK" if all_ok else "FAILURES")
    sys.exit(0 if all_ok else 1)
