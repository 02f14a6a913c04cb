"""How much would a per-group (256-row) residual bound tighten the certificate's slack?  (task 7.)
The slack is eps = |q_scan| * r + ..., r = max_j |bf16(x^_j) - x^_j| over the STORE (store.cu, gstat[0]).  A per-group r_g
can only help where r_g is well under the store-wide maximum.  CPU-only measurement on the synthetic rows of the bench
(host replay of avs_fill_synthetic) and on the reference's shipped database (tests/golden/f1_*).
    python profiles/r02/residual_spread.py > profiles/r02/residual_spread.json
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def bf16_round(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16          # round to nearest even
    return r.astype(np.uint32).view(np.float32)


def spread(X):
    X = X.astype(np.float64)
    Xn = (X / np.linalg.norm(X, axis=1, keepdims=True)).astype(np.float32)
    res = np.linalg.norm(bf16_round(Xn).astype(np.float64) - Xn.astype(np.float64), axis=1)
    n = res.shape[0] // 256 * 256
    out = {"rows": int(res.shape[0]), "r_max_store": float(res.max()), "r_median_row": float(np.median(res))}
    if n:
        g = res[:n].reshape(-1, 256).max(axis=1)
        out.update({"r_group_max_median": float(np.median(g)), "r_group_max_min": float(g.min()),
                    "slack_ratio_group_median_over_store": float(np.median(g) / res.max()),
                    "slack_ratio_best_group_over_store": float(g.min() / res.max())})
    return out


def main():
    synth = importlib.import_module("autostyle-tts_b200.synth")
    out = {}
    for dim in (768, 1024, 3072):
        out[f"synthetic_d{dim}"] = spread(synth.synth_rows(42, 0, 65536, dim))
    try:
        X = np.load(os.path.join(ROOT, "tests", "golden", "f1_vectors_fp16.npy"))
        out["reference_db_f1_130x6144"] = spread(X)
    except Exception as e:                                   # fixture name differs: report, do not fail
        out["reference_db_f1"] = {"error": str(e)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
