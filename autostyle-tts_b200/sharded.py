"""Row-sharded multi-GPU search: one process per GPU, rank r holds rows
[r*ceil(N/W), (r+1)*ceil(N/W)) of the database, every rank sees the full query batch, local exact
top-k, one ncclAllGather of [nq, k] (score f64, id i64) over NVLink, identical merge on every rank
(kernels in csrc/comm.cu).  torch.distributed is plumbing only: it carries the
128-byte ncclUniqueId from rank 0 to the others.
"""
from __future__ import annotations

from typing import Optional, Tuple

from .engine import Store


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range of `rank`."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    per = -(-n // world)
    return min(rank * per, n), min((rank + 1) * per, n)


def exchange_unique_id(make_id, rank: int, group=None) -> bytes:
    """Rank 0 creates the id with `make_id()`; everyone returns the same 128 bytes.  Works on any
    torch.distributed backend (gloo in the CPU tests, nccl on the GPU box)."""
    import torch.distributed as dist
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("ncclUniqueId exchange failed")
    return bytes(uid)


class ShardedStore:
    """A Store holding this rank's shard plus the NCCL communicator for the merge."""

    def __init__(self, dim: int, metric: str, n_total: int, rank: int, world: int, device: Optional[int] = None,
                 group=None, p2p: bool = True):
        self.rank, self.world, self.n_total = rank, world, n_total
        self.lo, self.hi = shard_range(n_total, rank, world)
        self.store = Store(dim, metric, capacity=max(self.hi - self.lo, 1), device=rank if device is None else device)
        self.p2p = False
        if world > 1:
            uid = exchange_unique_id(Store.nccl_unique_id, rank, group)
            self.store.comm_init(uid, rank, world)
            if p2p and world <= 8:
                self.p2p = self._connect_peers(group)

    def _connect_peers(self, group) -> bool:
        """Exchange the CUDA IPC handles of the peer exchange regions; all ranks agree on the outcome
        (NCCL all-gather stays the exchange when any pair of GPUs has no peer access)."""
        import torch.distributed as dist
        from .engine import AvsError
        try:
            mine = self.store.p2p_init(self.rank, self.world)
        except AvsError:
            mine = None
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        ok = all(h is not None for h in handles)
        if ok:
            try:
                self.store.p2p_connect(b"".join(handles), self.world)
            except AvsError:
                ok = False
        votes = [None] * self.world
        dist.all_gather_object(votes, ok, group=group)
        ok = all(votes)
        self.store.set_option("p2p_merge", 1 if ok else 0)
        return ok

    def fill_synthetic(self, seed: int):
        """Rank-local slice of the global synthetic stream; ids are global row numbers."""
        self.store.fill_synthetic(seed, self.lo, self.hi - self.lo, id_base=0)

    def insert_local(self, rows, ids):
        self.store.insert(rows, ids)

    def search(self, queries, k: int):
        """Device tensor in -> device tensors out (asynchronous); host array in -> numpy out, end to end
        (`avs_search_sharded_host`: each rank copies 1/world of the batch H2D, NVLink all-gather, search, merge, D2H)."""
        if self.world == 1:
            return self.store.search(queries, k)
        return self.store.search(queries, k, sharded=True)

    def close(self):
        self.store.close()
