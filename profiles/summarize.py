"""Turns the .ncu-rep captures brought back in gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize.py r01 [subdirectory of gpurun_out/]

Writes, per capture, the metrics the roofline claims rest on (duration, DRAM bytes, tensor-pipe activity,
registers, clocks) as `profiles/<round>/<name>.metrics.txt`, the per-launch list of one default bench.py run as
`launches_default.csv` / `launches_default.summary.txt`, and `traffic.json` (DRAM bytes per launch of the dominant
kernels, read by bench.py for roofline.traffic).
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg", "sm__pipe_tensor_cycles_active_realtime", "dram__cycles_active.avg",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max.per_second",
        "derived__lts__lts2xbar_bytes.sum.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed")


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, v, u in zip(hdr, vals, units)}


def main(rnd, sub=""):
    src, dst = os.path.join(ROOT, "gpurun_out", sub), os.path.join(ROOT, "profiles", rnd)
    os.makedirs(dst, exist_ok=True)
    try:                                   # captures are not repeated every run: keep the entries already committed
        traffic = json.load(open(os.path.join(dst, "traffic.json")))
    except Exception:
        traffic = {}
    for name, key in (("prof_k3_final", "scan_gemm"), ("prof_k2_final", "scan_gemv"), ("prof_k1", "normalize_rows")):
        rep = os.path.join(src, name + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        m = raw_metrics(rep)
        with open(os.path.join(dst, name + ".metrics.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none, one launch of the dominant kernel ({name})\n")
            for h in sorted(m):
                if any(h == w or h.startswith(w) for w in WANT):
                    f.write(f"{h} = {m[h][0]} {m[h][1]}\n")

        def to_bytes(k):
            v, u = m[k]
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        traffic[key] = {"dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
                        "duration_us_under_ncu": float(m["gpu__time_duration.sum"][0]) * (1e3 if m["gpu__time_duration.sum"][1] == "ms" else 1),
                        "capture": f"profiles/{rnd}/{name}.metrics.txt",
                        "workload": "C2 1M x 768, final level" + (", batch 1024" if key == "scan_gemm" else ", batch 1" if key == "scan_gemv" else "")}
    with open(os.path.join(dst, "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    lst = os.path.join(src, "launches_default.csv")
    if os.path.exists(lst):
        lines = [l for l in open(lst) if not l.startswith("==")]
        open(os.path.join(dst, "launches_default.csv"), "w").writelines(lines)
        rows = [(r["Kernel Name"].split("(")[0].replace("void ", ""), float(r["Metric Value"]) / 1e3) for r in csv.DictReader(lines)]
        agg = {}
        for n, v in rows:
            a = agg.setdefault(n, [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(v for _, v in rows)
        with open(os.path.join(dst, "launches_default.summary.txt"), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 1 --no-cpu-baseline\n")
            f.write("# (cold-cache, serialised: compare SHARES, not absolutes)\n")
            for n, (c, v) in sorted(agg.items(), key=lambda t: -t[1][1]):
                f.write(f"{n:50s} launches {c:4d}  total {v:10.1f} us  share {100 * v / tot:5.1f} %\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01", sys.argv[2] if len(sys.argv) > 2 else "")
