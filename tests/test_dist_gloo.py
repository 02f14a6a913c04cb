"""CPU suite, world_size 2 over gloo: the host-side plumbing of the sharded search — id exchange
from rank 0, contiguous row sharding, and that (local top-k -> all-gather -> merge) equals the
global answer.  The arithmetic here is the oracle's; the GPU path does the same exchange with one
ncclAllGather (csrc/comm.cu) and is checked in test_gpu_parity.py."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    import torch
    import torch.distributed as dist
    from oracle import flat_search as fs
    sh = importlib.import_module("autostyle-tts_b200.sharded")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        uid = sh.exchange_unique_id(lambda: bytes(range(128)), rank)
        assert uid == bytes(range(128))
        rng = np.random.default_rng(123)                     # same data on every rank
        n, d, k = 1001, 32, 7
        X = rng.standard_normal((n, d)).astype(np.float32)
        X[900] = X[17]                                       # a cross-shard exact tie
        ids = rng.permutation(n).astype(np.int64)
        Q = rng.standard_normal((4, d)).astype(np.float32)
        Q[0] = X[17]
        lo, hi = sh.shard_range(n, rank, world)
        p_ids, _, p_rows = fs.search(X[lo:hi], ids[lo:hi], Q, k, "COSINE")
        s64 = np.stack([fs.scores64(X[lo:hi], Q[i], "COSINE")[p_rows[i]] for i in range(4)])
        mine = torch.from_numpy(np.concatenate([s64, p_ids.astype(np.float64)], axis=1))
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        parts = [(g[:, k:].numpy().astype(np.int64), g[:, :k].numpy()) for g in gathered]
        m_ids, m_s = fs.merge_shards(parts, k)
        g_ids, g_d, _ = fs.search(X, ids, Q, k, "COSINE")
        ok = np.array_equal(m_ids, g_ids) and np.array_equal(m_s.astype(np.float32), g_d)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_merge_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True


def test_reference_arm_under_torchrun_prints_one_line():
    """The driver launches the reference arm like ours (torchrun, N ranks): rank 0 alone works and prints the
    ONE JSON line, the other ranks exit 0 without output."""
    import json
    import subprocess
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--rows", "3000",
           "--dim", "64", "--batch", "16", "--steps", "1", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-800:]
    lines = out.stdout.strip().splitlines()
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
