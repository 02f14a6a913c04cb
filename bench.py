#!/usr/bin/env python
"""Headline benchmark: queries/sec of exact top-k search on BASELINE.json's configurations on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--rows R] [--dim D] [--k K]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference     # the reference's CPU path (FLAT search restated; oracle/)

One JSON line on stdout (rank 0).
  headline (`value`, `e2e`, `roofline`, `cpu_baseline`): configs[1] = C2, 1 M x 768 cosine top-10, batch 1024 (and batch 1),
      row-sharded over the N ranks (strong scaling: the database is fixed).
  `legs.c5_weak`:   configs[4] per GPU - 12.5 M x 768 rows PER RANK, cosine top-100, batch 1024 and 1 (N = 8 is the
                    100 M-row north-star database itself).
  `legs.c3_strong`: configs[2] - 10 M x 1024, IP top-10, batch 1 and 4096, row-sharded over the N ranks.
`value` = whole-job QPS with queries already resident in HBM; `e2e` = the same through the C-ABI host call (host
buffers, H2D + D2H inside the timed region; N > 1: every rank copies 1/N of the batch and the slices are all-gathered over
NVLink peer memory); `roofline` = the dominant scan kernel timed with CUDA events on its launch stream, plus the
whole-step fraction; `cpu_baseline` = the oracle's CPU path on this box's host cores on a bounded sample.
Parity inside the run: the hits of the TIMED batch itself are compared with the float64 oracle on a 32-query sample.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner under torchrun), so
# the real stdout is kept aside for the result line and fd 1 is pointed at stderr for everything else.
RESULT_OUT = sys.stdout


def _reserve_stdout():
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


METRIC = "queries/sec top-10 exact search"
SEED_DB, SEED_Q = 42, 43
RTOL = 1e-5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--metric", default="COSINE")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scan-path", type=int, default=0, help="0 auto, 1 gemv, 2 gemm")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (repeatable)")
    ap.add_argument("--only-batch", action="store_true", help="skip the extra batch-1 measurement")
    ap.add_argument("--sweep", default="", help="comma-separated batch sizes measured on the same store; adds a `sweep` list to the JSON line")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: ncclAllGather + merge instead of the fused peer-memory kernel")
    ap.add_argument("--legs", default="auto", help="auto (C5 weak, C3 strong and C4 weak legs on the default workload), none, or a comma list of c5_weak,c3_strong,c4_weak")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 3 s back-to-back loop")
    ap.add_argument("--leg-rows-scale", type=float, default=1.0, help="testing: scale the legs' row counts")
    return ap.parse_args()


def workload_tag(rows, dim):
    return {(1_000_000, 768): "C2", (10_000_000, 1024): "C3", (12_500_000, 768): "C5 shard"}.get((rows, dim), "custom")


def workload_name(a):
    return f"{workload_tag(a.rows, a.dim)}: {a.rows} x {a.dim} synthetic unit-norm rows, {a.metric} top-{a.k}, batch {a.batch} (and batch 1)"


def config_of(a):
    """Workload description shared verbatim by both arms (`--impl ours` / `--impl reference`)."""
    dpad = (a.dim + 63) // 64 * 64
    return {"workload": workload_name(a), "rows": a.rows, "dim": a.dim, "k": a.k, "batch": a.batch, "metric_type": a.metric,
            "l2_policy": f"inputs larger than L2: a step streams {a.rows * dpad * 2 / 1e6:.0f} MB of bf16 rows (L2 126 MB)"}


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class Clocks:
    """SM clock / throttle-reason sampler running DURING the timed regions: NVML polled every ~4 ms from a
    thread (the timed regions are tens of milliseconds), `nvidia-smi -lms` as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("hw_power_brake_slowdown", 0x80), ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nvml, self.stop_flag, self.max_mhz = index, [], None, None, False, None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except (ValueError, IndexError):
                pass
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self._physical_index())], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.rows.append((time.time(), mhz, mask))
            except Exception:
                pass
            time.sleep(0.004)

    def _pump(self):
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            try:
                mask = 0
                for (name, bit), v in zip((("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                                           ("sw_power_cap", 0x4)), p[4:8]):
                    if v.lower().startswith("active"):
                        mask |= bit
                self.max_mhz = float(p[2])
                self.rows.append((time.time(), float(p[1]), mask))
            except (ValueError, IndexError):
                continue

    def window(self, lo, hi):
        w = [mhz for t, mhz, m in self.rows if lo <= t <= hi]
        return float(np.median(w)) if w else None

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, windows):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"], "samples": 0}
        time.sleep(0.06)
        sm, mask = [], 0
        for t, mhz, m in self.rows:
            if any(lo <= t <= hi for lo, hi in windows):
                sm.append(mhz)
                mask |= m
        reasons = sorted(name for name, bit in self.REASONS if mask & bit)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_mhz_min": float(min(sm)) if sm else None,
                "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's CPU path (numpy fp32 BLAS FLAT search)
# ------------------------------------------------------------------------------------------------
def host_rows(a, lo, hi):
    synth = importlib.import_module("autostyle-tts_b200.synth")
    out = np.empty((hi - lo, a.dim), dtype=np.float32)
    for s in range(lo, hi, 65536):
        e = min(hi, s + 65536)
        out[s - lo:e - lo] = synth.synth_rows(SEED_DB, s, e - s, a.dim)
    return out


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else (os.cpu_count() or 1)
    except Exception:
        return os.cpu_count() or 1


def cpu_flat_time(Xn, Q, k, reps=1):
    from oracle import flat_search as fs
    fs.cpu_flat_baseline(Xn[:65536], Q[:2], k)        # BLAS warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        fs.cpu_flat_baseline(Xn, Q, k)
    return (time.perf_counter() - t0) / reps


def run_reference(a):
    """The reference's own CPU implementation of the path (Milvus Lite is not installable: the oracle's CPU port, all
    host threads), same `config`, `metric`, `unit`; each step = a bounded sample of the workload (128 of the batch's
    queries against ALL rows); --steps / --warmup are honoured exactly."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    synth = importlib.import_module("autostyle-tts_b200.synth")
    try:                                               # torchrun exports OMP_NUM_THREADS=1: use every host core anyway
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass
    X = host_rows(a, 0, a.rows)                        # unit-norm rows: the cached FLAT/cosine index content
    nq = min(a.batch, 128)                             # bounded sample of the batch per step
    Q = synth.planted_queries(SEED_Q, SEED_DB, a.rows, a.batch, a.dim)[:nq]
    steps, warm = max(1, a.steps), max(0, a.warmup)
    from oracle import flat_search as fs
    for _ in range(warm):
        fs.cpu_flat_baseline(X, Q, a.k)
    t0 = time.perf_counter()
    for _ in range(steps):
        fs.cpu_flat_baseline(X, Q, a.k)
    dt = (time.perf_counter() - t0) / steps
    qps = nq / dt
    sample = f"{nq} of {a.batch} queries x all {a.rows} rows per step, fp32 OpenBLAS sgemm + argpartition"
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": a.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_of(a),
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cpu_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=RESULT_OUT, flush=True)


def ncu_traffic(rows, dim, k, batch, path, world):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of THIS workload on one
    GPU (profiles/rNN/traffic.json).  ncu cannot run under the driver's bench, so the figure is only quoted when the
    capture's workload is exactly the one measured (single GPU, same rows/dim/k/batch/kernel); otherwise null + reason."""
    import glob
    if world != 1:
        return None, "null at N > 1: the committed ncu capture is of the single-GPU shard size"
    if (rows, dim, k) != (1_000_000, 768, 10):
        return None, "no committed ncu capture for this workload"
    key = ("scan_gemm" if (path == 2 and batch == 1024) else "scan_gemm_m128_b1" if (path == 2 and batch == 1)
           else "scan_gemm_m128" if (path == 2 and batch == 64) else "scan_gemv" if (path == 1 and batch == 1) else None)
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*", "traffic.json")))
    if not key or not files:
        return None, "no committed ncu capture for this kernel variant"
    try:
        table = json.load(open(files[-1]))
    except Exception:
        return None, "traffic.json unreadable"
    if key not in table:
        return None, f"no committed ncu capture for this kernel variant in {os.path.relpath(files[-1], ROOT)}"
    return (table[key]["dram_bytes_per_launch"],
            f"ncu --set full capture of this workload with this round's kernels (not measured in this run: ncu cannot run under the "
            f"driver), {os.path.relpath(files[-1], ROOT)} [{key}]")


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Env:
    pass


def sample_slots(nq, pl_slots, n=32):
    """Query slots checked against the oracle: planted ones first (true neighbours exist), then evenly spaced."""
    take = [int(s) for s in np.sort(pl_slots)[: n // 2]]
    for s in np.linspace(0, nq - 1, num=min(nq, 4 * n)).astype(int):
        if len(take) >= min(n, nq):
            break
        if int(s) not in take:
            take.append(int(s))
    return np.asarray(sorted(take), dtype=np.int64)


def gather_rows_to_rank0(env, ss, dim):
    """The whole database on rank 0's host, read back from the device stores (bit-identical to the host replay of the
    generator, tests/test_gpu_parity.py::test_synthetic_fill_is_replayable_and_full_size_config)."""
    torch, dist = env.torch, env.dist
    st = ss.store
    n_loc = len(st)
    if env.world == 1:
        return np.concatenate([st.get_rows(lo, min(131072, n_loc - lo)) for lo in range(0, n_loc, 131072)])
    per = -(-ss.n_total // env.world)
    mine = torch.zeros((per, dim), dtype=torch.float32, device="cuda")
    if n_loc:
        mine[:n_loc] = st.get_rows_device(0, n_loc)
    parts = [torch.empty_like(mine) for _ in range(env.world)] if env.rank == 0 else None
    dist.gather(mine, parts, dst=0)
    if env.rank != 0:
        return None
    X = torch.cat(parts)[: ss.n_total].cpu().numpy()
    del parts
    return X


def recheck_at_scale(env, ss, cfg, Qs, got_ids, got_sc):
    """Parity check for stores too large for a host-side float64 pass.  Every rank scores the sample queries against ALL
    of its rows with a plain fp32 GEMM (torch, chunks of the fp32 master), keeps k+24 candidates per query with their
    rows; rank 0 re-evaluates the union in the oracle's float64 arithmetic, orders it by (score desc, id asc) and
    compares ids and scores with the engine's hits.  The fp32 GEMM only nominates candidates: its window is checked to
    end at least 5e-5 under the k-th exact score (its own error is ~1e-6)."""
    torch, dist = env.torch, env.dist
    from oracle import flat_search as fs
    st = ss.store
    n_loc, k, dim = len(st), cfg["k"], cfg["dim"]
    kc = k + 24
    Qd = torch.from_numpy(np.ascontiguousarray(Qs)).cuda()
    if cfg["metric"] == "COSINE":
        Qd = Qd / Qd.norm(dim=1, keepdim=True)
    best_s = torch.full((Qs.shape[0], 0), 0.0, device="cuda")
    best_id = torch.zeros((Qs.shape[0], 0), dtype=torch.int64, device="cuda")
    best_v = torch.zeros((Qs.shape[0], 0, dim), device="cuda")
    chunk = 1 << 20
    for lo in range(0, n_loc, chunk):
        n = min(chunk, n_loc - lo)
        Xc = st.get_rows_device(lo, n)
        S = Qd @ Xc.T
        if cfg["metric"] == "COSINE":
            S /= Xc.norm(dim=1).clamp_min(1e-30)[None, :]
        kk = min(kc, n)
        s_top, i_top = torch.topk(S, kk, dim=1)
        v_top = Xc[i_top]                                              # [nq, kk, dim]
        best_s = torch.cat([best_s, s_top], dim=1)
        best_id = torch.cat([best_id, i_top + (ss.lo + lo)], dim=1)
        best_v = torch.cat([best_v, v_top], dim=1)
        if best_s.shape[1] > kc:
            s2, j = torch.topk(best_s, kc, dim=1)
            best_s, best_id = s2, torch.gather(best_id, 1, j)
            best_v = torch.gather(best_v, 1, j[:, :, None].expand(-1, -1, dim))
        del Xc, S
    pad = kc - best_s.shape[1]
    if pad > 0:                                                       # tiny shards: pad with never-winning entries
        best_s = torch.cat([best_s, torch.full((Qs.shape[0], pad), -float("inf"), device="cuda")], dim=1)
        best_id = torch.cat([best_id, torch.full((Qs.shape[0], pad), -1, dtype=torch.int64, device="cuda")], dim=1)
        best_v = torch.cat([best_v, torch.zeros((Qs.shape[0], pad, dim), device="cuda")], dim=1)
    if env.world > 1:
        gs = [torch.empty_like(best_s) for _ in range(env.world)] if env.rank == 0 else None
        gi = [torch.empty_like(best_id) for _ in range(env.world)] if env.rank == 0 else None
        gv = [torch.empty_like(best_v) for _ in range(env.world)] if env.rank == 0 else None
        dist.gather(best_s.contiguous(), gs, dst=0)
        dist.gather(best_id.contiguous(), gi, dst=0)
        dist.gather(best_v.contiguous(), gv, dst=0)
        if env.rank != 0:
            return None
        best_s, best_id, best_v = torch.cat(gs, dim=1), torch.cat(gi, dim=1), torch.cat(gv, dim=1)
    cs, cid, cv = best_s.cpu().numpy(), best_id.cpu().numpy(), best_v.cpu().numpy()
    ok, max_rel, min_margin = True, 0.0, float("inf")
    for i in range(Qs.shape[0]):
        valid = cid[i] >= 0
        ids_i, X_i = cid[i][valid], cv[i][valid]
        s = fs.scores64(X_i, Qs[i], cfg["metric"])
        top = fs.order_topk(s, ids_i, k)
        exp_ids, exp_s = ids_i[top], s[top].astype(np.float32)
        kk = exp_ids.shape[0]
        ok &= bool(np.array_equal(got_ids[i][:kk], exp_ids)) and bool(np.all(got_ids[i][kk:] == -1))
        rel = np.abs(got_sc[i][:kk] - exp_s) / np.maximum(1.0, np.abs(exp_s))
        max_rel = max(max_rel, float(rel.max()) if kk else 0.0)
        # every rank's candidate window must end well under the k-th exact score, or a better row could hide behind it
        per_rank_floor = cs[i].reshape(env.world, -1).min(axis=1) if env.world > 1 else cs[i].min(keepdims=True)
        n_cand_min = min(kc, n_loc) if env.world == 1 else kc
        if kk == k and n_cand_min >= kc:
            min_margin = min(min_margin, float(s[top][-1] - per_rank_floor.max()))
    return {"method": "fp32 GEMM over all rows nominates k+24 candidates per rank; oracle float64 re-evaluation and (score desc, id asc) ordering of the union",
            "queries": int(Qs.shape[0]), "ids_match": bool(ok), "max_score_rel_err": max_rel,
            "candidate_window_margin": None if min_margin == float("inf") else min_margin,
            "sound": bool(min_margin > 5e-5) if min_margin != float("inf") else True}


def measure(a, env, cfg):
    """Fill one database, time the batches of `cfg`, check the timed hits against the oracle, free the store."""
    torch, dist = env.torch, env.dist
    synth = importlib.import_module("autostyle-tts_b200.synth")
    sharded = importlib.import_module("autostyle-tts_b200.sharded")
    rank, world, local = env.rank, env.world, env.local
    rows, dim, k, metric = cfg["rows"], cfg["dim"], cfg["k"], cfg["metric"]
    steps, warmup = cfg["steps"], cfg["warmup"]
    dpad = (dim + 63) // 64 * 64

    ss = sharded.ShardedStore(dim, metric, rows, rank, world, device=local, p2p=not a.no_p2p)
    st = ss.store
    ss.fill_synthetic(SEED_DB)
    st.set_option("scan_path", a.scan_path)
    for kv in a.opt:
        key, val = kv.split("=")
        st.set_option(key, int(val))
    torch.cuda.synchronize()
    rows_local = len(st)
    batches = sorted(set(cfg["batches"]))
    Qh, pl_slots, pl_rows = synth.planted_queries(SEED_Q, SEED_DB, rows, max(batches), dim, return_planted=True)
    q_pinned = torch.from_numpy(Qh).pin_memory()
    q_dev = q_pinned.cuda(non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n_steps, n_warm):
        for _ in range(n_warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_lo = time.time()
        e0.record()
        for _ in range(n_steps):
            fn()
        e1.record()
        barrier()
        t_hi = time.time()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / n_steps, (t_lo, t_hi)

    def per_step_latency(fn, n):
        """p10 / median / p90 of single search calls, each bracketed by its own CUDA events."""
        out = []
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            out.append(e0.elapsed_time(e1))
        p10, p50, p90 = (float(x) for x in np.percentile(out, [10, 50, 90]))
        return {"p10": p10, "p50": p50, "p90": p90, "calls": n}

    results, last_hits = {}, {}
    headline_windows = []                       # timed regions of the headline numbers (device-resident + e2e), per batch
    for batch in batches:                       # small batch first: it is not the one that heats the chip
        q = q_dev[:batch]
        qh = q_pinned[:batch].numpy()

        def fn_dev(batch=batch, q=q):
            last_hits[batch] = ss.search(q, k)
        # device-resident throughput: the headline loop carries no event pairs inside a search (an event record between
        # two kernels of a search serialises their programmatic dependent launch)
        l0 = st.stat("kernel_launches")
        ms, win = timed(fn_dev, steps, warmup)
        launches = (st.stat("kernel_launches") - l0) // (steps + warmup) * steps
        # live roofline of the dominant scan kernel: the same K steps again, CUDA event pairs around every scan launch
        # on its launch stream (and around the exchange kernel at N > 1)
        st.scan_timing(1)
        ms_ev, win_ev = timed(fn_dev, steps, 1)
        scan_ms, scan_n = st.scan_timing(-1)
        exch_us = st.stat("exchange_us") if world > 1 else None
        st.scan_timing(0)
        env.windows.append(win_ev)
        lat = per_step_latency(fn_dev, max(5, min(steps, 30)))
        env.windows.append(win)
        headline_windows.append(win)
        path, levels = st.stat("last_scan_path"), st.stat("last_levels")
        rows_final = st.stat("last_final_rows")     # rows the final (dense) level visits
        flops = 2.0 * batch * rows_final * dpad
        passes = (batch + 7) // 8 if path == 1 else 1
        nbytes = float(passes) * rows_final * dpad * 2
        # the tensor-core scan is HBM-bound for small batches: report against whichever roof binds
        tensor_bound = path == 2 and flops / (env.tc_burst * 1e12) > nbytes / (env.hbm_peak * 1e9)
        mhz = env.clocks.window(*win) if rank == 0 else None
        if tensor_bound:
            work, step_work = flops, 2.0 * batch * rows_local * dpad
            # B200_PROFILING.md: burst cuBLAS figure for a kernel timed alone / a short region at full clocks, the
            # sustained one when the timed region ran throttled under the 1 kW cap (decided from the clocks sampled in it)
            capped = bool(mhz and env.clocks.max_mhz and mhz < 0.9 * env.clocks.max_mhz)
            peak = env.tc_sustained if capped else env.tc_burst
            ach = work / (scan_ms * 1e-3) / 1e12 if scan_ms else None
            roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "peak_kind": ("cuBLAS bf16 sustained (region ran power-capped at %.0f MHz)" % mhz) if capped else "cuBLAS bf16 burst",
                    "frac_of_burst_peak": ach / env.tc_burst if ach else None,
                    "frac_of_sustained_peak": ach / env.tc_sustained if ach else None,
                    "step_achieved": step_work / (ms * 1e-3) / 1e12}
        else:
            work, step_work = nbytes, float(passes) * rows_local * dpad * 2
            peak = env.hbm_peak
            roof = {"bound": "hbm", "achieved": work / (scan_ms * 1e-3) / 1e9 if scan_ms else None, "peak": peak,
                    "unit": "GB/s", "step_achieved": step_work / (ms * 1e-3) / 1e9}
        roof["frac"] = roof["achieved"] / roof["peak"] if roof["achieved"] else None
        roof["step_frac"] = roof["step_achieved"] / roof["peak"]      # whole search step (all levels, selects, rescoring[, exchange]) against the same roof
        if roof["bound"] == "hbm" and roof["frac"] and roof["frac"] > 1.0:
            roof["note"] = "peak is the copy (read + write) bandwidth; this kernel only reads and streams faster than a copy does"
        traffic, traffic_src = ncu_traffic(rows, dim, k, batch, path, world)
        roof.update({"traffic": traffic, "traffic_source": traffic_src, "peak_source": env.peak_src,
                     "kernel": "scan_gemm (tcgen05)" if path == 2 else "scan_gemv",
                     "kernel_ms": scan_ms, "timed_launch_groups": scan_n, "kernel_ms_region_ms_per_step": ms_ev,
                     "algorithmic_work_per_launch_group": work,
                     "algorithmic_work_per_step": step_work, "sm_mhz_in_region": mhz})
        # end to end through the C-ABI host call: pinned host queries in, host hits out (N > 1: avs_search_sharded_host)
        fn_host = (lambda qh=qh: ss.search(qh, k))
        ms_e, win_e = timed(fn_host, steps, warmup)
        env.windows.append(win_e)
        headline_windows.append(win_e)
        h2d = int(batch * dim * 4) if world == 1 else int(-(-batch // world) * dim * 4)
        e2e = {"value": batch / (ms_e * 1e-3), "unit": "queries/s", "ms_per_step": ms_e,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(batch * k * (20 if world == 1 else 12)),
               "api": "avs_search_host (C-ABI, host buffers)" if world == 1 else
                      "avs_search_sharded_host (C-ABI, host buffers; per rank: H2D of 1/N of the batch, NVLink peer all-gather, D2H of ids+scores)"}
        results[batch] = {"latency_ms": lat, "qps": batch / (ms * 1e-3), "ms": ms, "launches": int(launches),
                          "launches_per_search": int(launches) // steps, "roofline": roof, "e2e": e2e,
                          "scan_path": path, "levels": levels, "kprime": st.stat("last_kprime"), "exchange_kernel_us": exch_us}

    # ---- parity of the TIMED hits themselves (the last call of the device-timed loop of every batch) ----
    parity = {}
    main_b = cfg["main_batch"]
    ids_t, sc_t = last_hits[main_b]
    got_ids, got_sc = ids_t.cpu().numpy(), sc_t.cpu().numpy()
    in_batch = pl_slots < main_b
    planted_ok = float(np.mean(got_ids[pl_slots[in_batch], 0] == pl_rows[in_batch])) if in_batch.any() else None
    slots = sample_slots(main_b, pl_slots[in_batch])
    X_host = None
    if cfg.get("full_oracle"):
        X_host = gather_rows_to_rank0(env, ss, dim)
        if rank == 0:
            from oracle import flat_search as fs
            exp_ids, exp_d, _ = fs.search_large(X_host, np.arange(rows), Qh[slots], k + 1, metric)
            ok_ids = bool(np.array_equal(got_ids[slots], exp_ids[:, :k]))
            rel = np.abs(got_sc[slots] - exp_d[:, :k]) / np.maximum(1.0, np.abs(exp_d[:, :k]))
            # the envelope inside which an fp32 engine (Milvus Lite's SIMD FLAT) could legitimately order differently:
            # relative gap between the k-th and (k+1)-th exact scores over the sample
            gap = (exp_d[:, k - 1].astype(np.float64) - exp_d[:, k].astype(np.float64)) / np.maximum(1e-30, np.abs(exp_d[:, k - 1]))
            parity = {"parity_ids_match_oracle": ok_ids, "max_score_rel_err": float(rel.max()), "queries_checked": int(slots.size),
                      "what": f"hits of the timed batch-{main_b} call itself vs oracle.search_large (float64) on {int(slots.size)} of its queries",
                      "k_gap_rel_min": float(gap.min()), "k_gap_rel_p01": float(np.percentile(gap, 1)), "k_gap_rel_median": float(np.median(gap))}
            # where that gap exceeds 1e-6 the fp32-accumulate oracle variants must give the same id list as the engine
            f32_ok, n_f32 = True, 0
            for i in np.nonzero(gap > 1e-6)[0][:8]:
                for variant in ("normalize_then_dot", "dot_then_divide"):
                    v_ids, _, _ = fs.search(X_host, np.arange(rows), Qh[slots[i]][None, :], k, metric, accum="f32", variant=variant)
                    f32_ok &= bool(np.array_equal(v_ids[0], got_ids[slots[i]]))
                    n_f32 += 1
            parity["f32_variants_agree_where_gap_gt_1e-6"] = bool(f32_ok)
            parity["f32_variant_checks"] = n_f32
    else:
        rc = recheck_at_scale(env, ss, cfg, Qh[slots], got_ids[slots], got_sc[slots])
        if rank == 0:
            parity = {"parity_ids_match_oracle": bool(rc["ids_match"] and rc["sound"]), **rc,
                      "what": f"hits of the timed batch-{main_b} call itself, {int(slots.size)} of its queries"}
    barrier()       # rank 0 alone ran the oracle above: nobody enters the next collective search until it is back
    # batch-1 hits of the timed call: the single query is slot 0 of the batch
    if 1 in last_hits and main_b != 1:
        b1_ids = last_hits[1][0].cpu().numpy()
        ids_again, _ = ss.search(q_dev[:main_b], k)
        parity["batch1_equals_batch_row0"] = bool(np.array_equal(b1_ids[0], ids_again.cpu().numpy()[0]))
    stats = {"uncertified_queries": st.stat("uncertified_queries"), "repaired_queries": st.stat("repaired_queries"),
             "wide_rescored_queries": st.stat("wide_rescored_queries"), "queries": st.stat("queries"),
             "p2p_timeouts": st.stat("p2p_timeouts") if world > 1 else 0}

    # ---- sustained leg: >= 3 s of back-to-back searches (the serving loop lives under the power cap) ----
    sustained = None
    if cfg.get("sustained") and not a.no_sustained:
        b = main_b
        q = q_dev[:b]
        n_loop = max(50, int(3000.0 / results[b]["ms"]) + 1)
        ms_s, win_s = timed(lambda: ss.search(q, k), n_loop, 3)
        env.windows.append(win_s)
        mhz_s = env.clocks.window(*win_s) if rank == 0 else None
        r = results[b]["roofline"]
        peak_s = env.tc_sustained if r["bound"] == "tensor" else env.hbm_peak
        ach = r["algorithmic_work_per_step"] / (ms_s * 1e-3) / (1e12 if r["bound"] == "tensor" else 1e9)
        sustained = {"seconds": n_loop * ms_s * 1e-3, "steps": n_loop, "ms_per_step": ms_s, "value": b / (ms_s * 1e-3), "unit": "queries/s",
                     "sm_mhz_in_region": mhz_s, "step_achieved": ach, "peak": peak_s,
                     "peak_kind": "cuBLAS bf16 sustained" if r["bound"] == "tensor" else "HBM copy", "step_frac": ach / peak_s}

    # ---- same-box CPU baseline (rank 0, N = 1, main workload only) ----
    cpu = None
    if cfg.get("cpu_baseline") and rank == 0 and world == 1 and not a.no_cpu_baseline:
        X = X_host if X_host is not None else gather_rows_to_rank0(env, ss, dim)
        nq_cpu = min(main_b, 128)
        dt = cpu_flat_time(X, Qh[:nq_cpu], k, reps=8)
        cpu = {"value": nq_cpu / dt, "unit": "queries/s", "cores": cpu_threads(), "kind": "port",
               "sample": f"{nq_cpu} of {main_b} queries x all {rows_local} rows, mean of 8 passes ({dt:.2f} s each), numpy fp32 sgemm + argpartition"}
        dt1 = cpu_flat_time(X, Qh[:1], k, reps=20)
        cpu["batch1_value"] = 1.0 / dt1
        cpu["host_cpu_count"] = os.cpu_count()
        try:                                   # the 1-core figure next to the all-cores one
            from threadpoolctl import threadpool_limits
            with threadpool_limits(limits=1):
                cpu["value_1core"] = 16 / cpu_flat_time(X, Qh[:16], k)
        except Exception:
            cpu["value_1core"] = None
        del X
    del X_host
    exchange = ("fused peer-memory push+merge kernel (NVLink P2P)" if ss.p2p else "ncclAllGather + merge") if world > 1 else None
    ss.close()
    torch.cuda.empty_cache()
    return {"results": results, "parity": parity, "planted_top1_match": planted_ok, "stats": stats, "sustained": sustained,
            "cpu": cpu, "rows_local": rows_local, "exchange": exchange, "windows": headline_windows}


def measure_c1(env):
    """BASELINE.json configs[0] / SURVEY.md section 8(d) C1: the reference's own style database (tests/golden/f1_*: the
    130 x 6144 rows of milvus/milvus_demo.db) through the reference's own call shape - `MilvusClient.search(collection,
    data=[vec.tolist()], limit=5, output_fields=[...])`, one utterance at a time like milvus/search_json.py:382-449 and
    the self-query loop of milvus/RAG.py:567-582.  Too small for a roofline: per-call latency (host wall clock around
    the whole Python call, dict results included), parity against the golden top-5 lists, and the CPU port beside it."""
    pkg = importlib.import_module("autostyle-tts_b200")
    from oracle import flat_search as fs
    G = os.path.join(ROOT, "tests", "golden")
    X = np.load(os.path.join(G, "f1_vectors_fp16.npy")).astype(np.float32)
    rows = json.load(open(os.path.join(G, "f1_rows.json"), encoding="utf-8"))
    kat = np.load(os.path.join(G, "f1_kat.npz"))
    pks, meta = rows["pks"], rows["meta"]
    n, name = X.shape[0], "embeddings_biographies_collection"
    client = pkg.MilvusClient(":memory:", device=env.local)
    client.create_collection(collection_name=name, dimension=X.shape[1])
    client.insert(collection_name=name, data=[{"id": int(pks[i]), "file_id": meta[i]["file_id"], "vector": X[i].tolist(),
                                                "text": meta[i]["text"]} for i in range(n)])
    row_of = {m["file_id"]: j for j, m in enumerate(meta)}
    queries = [X[i].tolist() for i in range(n)]                       # the reference passes Python lists
    for i in range(5):
        client.search(collection_name=name, data=[queries[i]], limit=5, output_fields=["file_id", "text"])
    lat, ok, max_rel = [], True, 0.0
    for i in range(n):
        t0 = time.perf_counter()
        r = client.search(collection_name=name, data=[queries[i]], limit=5, output_fields=["file_id", "text"])
        lat.append((time.perf_counter() - t0) * 1e3)
        got_rows = [row_of[h["entity"]["file_id"]] for h in r[0]]
        ok &= got_rows == kat["self_rows"][i].tolist() and [h["id"] for h in r[0]] == kat["self_pk_ids"][i].tolist()
        max_rel = max(max_rel, float(np.max(np.abs(np.array([h["distance"] for h in r[0]], np.float32) - kat["self_dist"][i]))))
    t0 = time.perf_counter()
    rb = client.search(collection_name=name, data=X, limit=5, output_fields=["file_id", "text"])   # all 130 in one call
    batch_ms = (time.perf_counter() - t0) * 1e3
    ok &= all([row_of[h["entity"]["file_id"]] for h in rb[i]] == kat["self_rows"][i].tolist() for i in range(n))
    rp = client.search(collection_name=name, data=kat["pert_queries"], limit=5, metric_type="COSINE", output_fields=["file_id"])
    ok &= all([row_of[h["entity"]["file_id"]] for h in rp[i]] == kat["pert_rows"][i].tolist() for i in range(len(rp)))
    client.close()
    Xn = X / np.linalg.norm(X.astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    cpu_lat = []
    for i in range(n):                                                # the CPU port, one query per call as well
        q = (X[i] / np.float32(np.linalg.norm(X[i].astype(np.float64))))[None, :]
        t0 = time.perf_counter()
        fs.cpu_flat_baseline(Xn, q, 5)
        cpu_lat.append((time.perf_counter() - t0) * 1e3)
    p10, p50, p90 = (float(x) for x in np.percentile(lat, [10, 50, 90]))
    return {"config": {"workload": "C1: the reference's shipped style database (130 x 6144, tests/golden/f1_*), COSINE top-5, one query per call",
                       "rows": n, "dim": int(X.shape[1]), "k": 5, "batch": 1},
            "api": "MilvusClient.search(collection_name=, data=[list], limit=5, output_fields=[file_id, text]) - Python dicts out",
            "latency_ms": {"p10": p10, "p50": p50, "p90": p90, "calls": n}, "qps": 1e3 / p50,
            "one_call_all_130_queries_ms": batch_ms,
            "cpu_port_latency_ms_p50": float(np.median(cpu_lat)), "cpu_port_qps": 1e3 / float(np.median(cpu_lat)),
            "parity_ids_match_golden": bool(ok), "max_abs_distance_diff_vs_golden": max_rel,
            "what": "top-5 rows, primary keys and distances of all 130 self-queries (KAT-2), the batched call and the 64 perturbed queries (C1) against tests/golden/f1_kat.npz"}


def box_calibration(env):
    """What THIS box's GPU delivers right now on the two library yardsticks the roofs in MEASURED_PEAKS.json were taken with:
    a cuBLAS bf16 GEMM (8192^3, ~60 ms of back-to-back launches: a burst figure) and a device-to-device copy (1 GiB).
    Context for `roofline.frac` only - the roofs stay the pool-wide measured peaks.  It separates "this box is slow" from
    "this build is slow": a 30 % regression of the compute-bound path was first misread as box-to-box spread until the
    same-box yardstick and an A/B against the previous library said otherwise (profiles/r02/c26).  Nothing of the product
    runs here (cuBLAS via torch is the yardstick)."""
    torch = env.torch
    out = {}
    try:
        a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
        b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
        for _ in range(5):
            (a @ b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(80):
            (a @ b)
        e1.record()
        torch.cuda.synchronize()
        out["cublas_bf16_tflops"] = 80 * 2.0 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
        del a, b
        src = torch.empty(1 << 30, device="cuda", dtype=torch.uint8)
        dst = torch.empty_like(src)
        for _ in range(3):
            dst.copy_(src)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            dst.copy_(src)
        e1.record()
        torch.cuda.synchronize()
        out["copy_gbs"] = 20 * 2.0 * (1 << 30) / (e0.elapsed_time(e1) * 1e-3) / 1e9
        del src, dst
        torch.cuda.empty_cache()
        out["what"] = "torch bf16 matmul 8192^3 x 80 (cuBLAS) and 1 GiB device-to-device copy x 20 (read + write bytes), CUDA events, this GPU, before the searches"
    except Exception as e:
        out["error"] = f"{type(e).__name__}: {e}"
    return out


def leg_summary(cfg, m):
    out = {"config": {"rows_total": cfg["rows"], "rows_per_gpu": m["rows_local"], "dim": cfg["dim"], "k": cfg["k"],
                      "metric_type": cfg["metric"], "scaling": cfg["scaling"], "steps": cfg["steps"], "warmup": cfg["warmup"]},
           "batches": {}, "parity": m["parity"], "planted_top1_match": m["planted_top1_match"], **m["stats"]}
    for b, r in m["results"].items():
        out["batches"][str(b)] = {"qps": r["qps"], "ms_per_step": r["ms"], "latency_ms": r["latency_ms"], "e2e": r["e2e"],
                                  "gpu_launches_per_search": r["launches_per_search"], "levels": r["levels"], "kprime": r["kprime"],
                                  "scan_path": {1: "gemv", 2: "gemm"}.get(r["scan_path"]), "roofline": r["roofline"],
                                  "exchange_kernel_us": r["exchange_kernel_us"]}
    return out


def run_ours(a):
    import torch
    import torch.distributed as dist
    env = Env()
    env.torch, env.dist = torch, dist
    env.rank, env.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    env.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    if not os.path.exists(os.path.join(ROOT, "autostyle-tts_b200", "libavs.so")) and env.local == 0:
        importlib.import_module("autostyle-tts_b200.build").build()
    torch.cuda.set_device(env.local)
    if env.world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", env.local))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    env.hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    env.tc_burst = float(peaks.get("bf16_tflops", 1590.0))
    env.tc_sustained = float(peaks.get("bf16_tflops_sustained", 1400.0))
    env.peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    env.clocks = Clocks(env.local)
    env.windows = []
    if env.rank == 0:
        env.clocks.start()
    calib = box_calibration(env) if env.rank == 0 else None

    sweep = sorted({int(b) for b in a.sweep.split(",") if b.strip()})
    main_cfg = {"name": "main", "rows": a.rows, "dim": a.dim, "k": a.k, "metric": a.metric, "scaling": "strong",
                "batches": ({a.batch} if a.only_batch else {a.batch, 1}) | set(sweep), "main_batch": a.batch,
                "steps": a.steps, "warmup": a.warmup, "full_oracle": a.rows * a.dim <= 1_100_000 * 1024,
                "sustained": True, "cpu_baseline": True}
    default_workload = (a.rows, a.dim, a.k, a.metric) == (1_000_000, 768, 10, "COSINE")
    want = [] if a.legs == "none" else (["c5_weak", "c3_strong", "c4_weak"] if a.legs == "auto" else [x for x in a.legs.split(",") if x])
    if a.legs == "auto" and not default_workload:
        want = []
    sc = a.leg_rows_scale
    leg_steps, leg_warm = max(3, min(a.steps, 10)), max(3, min(a.warmup, 3))
    leg_cfgs = {
        "c5_weak": {"name": "c5_weak", "rows": int(12_500_000 * sc) * env.world, "dim": 768, "k": 100, "metric": "COSINE", "scaling": "weak",
                    "batches": {1, 1024}, "main_batch": 1024, "steps": leg_steps, "warmup": leg_warm},
        "c3_strong": {"name": "c3_strong", "rows": int(10_000_000 * sc), "dim": 1024, "k": 10, "metric": "IP", "scaling": "strong",
                      "batches": {1, 4096}, "main_batch": 4096, "steps": leg_steps, "warmup": leg_warm},
        # BASELINE.json configs[3]: 20 M x 3072 cosine top-50 over 8 GPUs = 2.5 M rows per GPU (weak scaling: N = 8 is the config itself)
        "c4_weak": {"name": "c4_weak", "rows": int(2_500_000 * sc) * env.world, "dim": 3072, "k": 50, "metric": "COSINE", "scaling": "weak",
                    "batches": {1, 1024}, "main_batch": 1024, "steps": leg_steps, "warmup": leg_warm},
    }

    m = measure(a, env, main_cfg)
    legs = {}
    if a.legs == "auto" and default_workload:
        if env.world == 1:
            try:
                legs["c1_reference_db"] = measure_c1(env)
            except Exception as e:
                legs["c1_reference_db"] = {"error": f"{type(e).__name__}: {e}"}
        else:
            legs["c1_reference_db"] = None       # 130 rows do not shard: measured at N = 1 only
    for name in want:
        try:
            legs[name] = leg_summary(leg_cfgs[name], measure(a, env, leg_cfgs[name]))
        except Exception as e:                       # a leg must never take the headline down with it
            legs[name] = {"error": f"{type(e).__name__}: {e}"}
            try:
                torch.cuda.empty_cache()
            except Exception:
                pass

    if env.rank == 0:
        clk = env.clocks.summary(m["windows"])          # clocks during the headline's timed regions (C2, device-resident and e2e)
        clk["all_timed_regions"] = env.clocks.summary(env.windows)   # legs and the >= 3 s sustained loop included
        env.clocks.stop()
        results = m["results"]
        main = results[a.batch]
        cpu = m["cpu"]
        if cpu is not None:
            cpu["parity_ids_match_oracle"] = m["parity"].get("parity_ids_match_oracle")
        line = {"metric": METRIC, "value": main["qps"], "unit": "queries/s", "n_gpus": env.world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": main["ms"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config_of(a),
                "engine": {"sharding": f"rows/{env.world}" if env.world > 1 else "none", "exchange": m["exchange"],
                           "arith": "bf16 operands, fp32 accumulate scan; float64 rescoring of the candidates",
                           "scan_path": {1: "gemv", 2: "gemm"}.get(main["scan_path"]), "levels": main["levels"],
                           "oversample_kprime": main["kprime"], "gpu_launches_per_search": main["launches_per_search"],
                           "exchange_kernel_us": main["exchange_kernel_us"]},
                "latency_ms": main["latency_ms"], "clocks": clk, "box_calibration": calib, "e2e": main["e2e"], "gpu_launches": main["launches"], "roofline": main["roofline"],
                "cpu_baseline": cpu, "parity": m["parity"], "parity_ids_match_oracle": m["parity"].get("parity_ids_match_oracle"),
                "planted_top1_match": m["planted_top1_match"], **m["stats"], "sustained": m["sustained"], "legs": legs}
        if 1 in results and a.batch != 1:
            b1 = results[1]
            line["batch1"] = {"value": b1["qps"], "unit": "queries/s", "ms_per_step": b1["ms"], "latency_ms": b1["latency_ms"], "e2e": b1["e2e"],
                              "gpu_launches": b1["launches"], "gpu_launches_per_search": b1["launches_per_search"], "roofline": b1["roofline"],
                              "exchange_kernel_us": b1["exchange_kernel_us"]}
        if sweep:
            line["sweep"] = [{"batch": b, "value": results[b]["qps"], "ms_per_step": results[b]["ms"],
                              "e2e": results[b]["e2e"]["value"] if results[b]["e2e"] else None,
                              "scan_path": {1: "gemv", 2: "gemm"}.get(results[b]["scan_path"]), "levels": results[b]["levels"],
                              "bound": results[b]["roofline"]["bound"], "achieved": results[b]["roofline"]["achieved"],
                              "peak": results[b]["roofline"]["peak"], "frac": results[b]["roofline"]["frac"], "step_frac": results[b]["roofline"]["step_frac"],
                              "kernel_ms": results[b]["roofline"]["kernel_ms"]} for b in sweep]
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    if env.world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    _reserve_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
