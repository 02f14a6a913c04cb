"""Per-level phase times inside the persistent tensor-core scan kernel (option "trace": globaltimer stamps of CTA 0).
    python profiles/r02/trace_levels.py [--rows N] [--dim D] [--k K] [--batches 1,1024] [--opt key=value ...]
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--metric", default="COSINE")
    ap.add_argument("--batches", default="1,8,128,1024")
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--per-cta", action="store_true", help="add the distribution of per-CTA finish times per level")
    a = ap.parse_args()
    pkg = importlib.import_module("autostyle-tts_b200")
    synth = importlib.import_module("autostyle-tts_b200.synth")
    st = pkg.Store(a.dim, a.metric, capacity=a.rows)
    st.fill_synthetic(42, 0, a.rows)
    for kv in a.opt:
        key, val = kv.split("=")
        st.set_option(key, int(val))
    st.set_option("trace", 1)
    batches = [int(b) for b in a.batches.split(",")]
    Q = torch.from_numpy(synth.planted_queries(43, 42, a.rows, max(batches), a.dim)).cuda()
    out = []
    for b in batches:
        for _ in range(5):
            st.search(Q[:b], a.k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            st.search(Q[:b], a.k)
        e1.record()
        torch.cuda.synchronize()
        L = st.stat("last_levels")
        t = [st.stat(f"trace:{i}") for i in range(4 * L + 2)]
        n_lv = max(i for i in range(L + 1) if t[4 * i] != 0) if any(t) else 0
        rows = []
        for l in range(n_lv):
            s0, s1, s2, s3, s4 = t[4 * l], t[4 * l + 1], t[4 * l + 2], t[4 * l + 3], t[4 * l + 4]
            rows.append({"level": l, "scan_us": (s1 - s0) / 1e3, "wait_grid_us": (s2 - s1) / 1e3, "select_us": (s3 - s2) / 1e3,
                         "wait_thresholds_us": (s4 - s3) / 1e3})
        if a.per_cta:
            for l in range(n_lv):
                fin = np.array([st.stat(f"trace:{64 + l * 256 + c}") for c in range(148)], dtype=np.float64)
                fin = fin[fin > 0]
                if fin.size:
                    rel = (fin - t[4 * l]) / 1e3
                    order = np.argsort(rel)
                    rows[l]["cta_finish_us"] = {"min": float(rel.min()), "p25": float(np.percentile(rel, 25)), "median": float(np.median(rel)),
                                                "p75": float(np.percentile(rel, 75)), "max": float(rel.max()), "ctas": int(fin.size),
                                                "slowest_ctas": [int(x) for x in order[-6:]], "fastest_ctas": [int(x) for x in order[:6]],
                                                "even_median": float(np.median(rel[0::2])), "odd_median": float(np.median(rel[1::2]))}
                # epilogue accounting (cycles, warp 4 and warp 11 of a few CTAs): waiting for the accumulator, tfull -> tempty
                # arrive (the part the MMA can wait for), arrive -> end of tile
                base = 64 + 12 * 256
                acc = []
                for cta in (0, 1, 2, 3, 40, 41, 100, 101):
                    for w in (0, 4):
                        v = [st.stat(f"trace:{base + (l * 256 + cta) * 8 + w + i}") for i in range(4)]
                        if v[3]:
                            acc.append({"cta": cta, "warp": 4 if w == 0 else 11, "tiles": v[3], "wait_cyc_per_tile": v[0] / v[3],
                                        "crit_cyc_per_tile": v[1] / v[3], "tail_cyc_per_tile": v[2] / v[3]})
                rows[l]["epilogue_cycles"] = acc
        SEL = 64 + 12 * 256 * 9                    # AVS_TRACE_SEL: phase stamps of CTA 0's lean level select
        for l in range(n_lv):
            v = [st.stat(f"trace:{SEL + l * 16 + i}") for i in range(6)]
            if v[0] and v[5] and v[5] >= v[0] >= t[4 * l]:
                rows[l]["cta0_fast_select_phases_us"] = [(v[i + 1] - v[i]) / 1e3 if v[i + 1] and v[i] else None for i in range(5)]
        out.append({"batch": b, "ms_per_search": e0.elapsed_time(e1) / 20, "levels_total": L, "levels_in_persistent_kernel": n_lv,
                    "kernel_us": (t[4 * n_lv] - t[0]) / 1e3 if n_lv else None, "per_level": rows})
    print(json.dumps({"rows": a.rows, "dim": a.dim, "k": a.k, "opts": a.opt, "trace": out}))
    st.close()


if __name__ == "__main__":
    main()
