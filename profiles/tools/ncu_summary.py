"""Summarise .ncu-rep captures into the small text files kept under profiles/rNN/ (the reports themselves stay in gpurun_out/).
    python profiles/tools/ncu_summary.py OUT_DIR REP [REP ...]     # writes OUT_DIR/<rep>.metrics.txt
"""
import csv
import io
import os
import re
import subprocess
import sys

KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?|"
                  r"gpu__dram_throughput.*|sm__pipe_tensor_cycles_active.*|sm__throughput\.avg\.pct.*|lts__throughput\.avg\.pct.*|"
                  r"derived__lts__lts2xbar_bytes.*|launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit.*|waves_per_multiprocessor)|"
                  r"sm__cycles_elapsed\.(avg|max)\.per_second|sm__warps_active\.avg\.pct.*|smsp__inst_executed\.sum|"
                  r"sm__inst_executed_pipe_tensor.*|l1tex__data_pipe_tc_wavefronts_mem_shared\.sum.*|smsp__cycles_active\.avg|"
                  r"lts__t_sector_hit_rate\.pct|lts__t_bytes\.sum|l1tex__t_bytes\.sum)$")


def main():
    out_dir, reps = sys.argv[1], sys.argv[2:]
    os.makedirs(out_dir, exist_ok=True)
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        name = os.path.splitext(os.path.basename(rep))[0]
        with open(os.path.join(out_dir, name + ".metrics.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none, {len(rows) - 2} launch(es) captured ({name}.ncu-rep, read with ncu -i --page raw --csv)\n")
            for r in rows[2:]:
                d = dict(zip(hdr, r))
                f.write(f"Kernel Name = {d.get('Kernel Name', '')}\n")
                for k, u in zip(hdr, units):
                    if KEEP.match(k):
                        f.write(f"{k} = {d[k]} {u}\n")
                f.write("\n")
        print("wrote", os.path.join(out_dir, name + ".metrics.txt"))


if __name__ == "__main__":
    main()
