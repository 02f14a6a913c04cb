"""TEST INFRASTRUCTURE ONLY — CPU oracle for the FLAT top-k search path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; the product path
(`autostyle-tts_b200/`) must never route through it.

PARITY PINNING — "parity unpinned" against the engine: the reference holds NO asserting test for this boundary
 and its engine (pymilvus -> milvus-lite, unpinned:
`/root/reference/milvus/RAG.py:2`) is an un-vendored third-party dependency
that cannot be installed here.  The oracle is therefore pinned against
(1) the reference's shipped database `milvus/milvus_demo.db` decoded by
`oracle/milvus_db.py` with the reference's own inline self-query check
(`/root/reference/milvus/RAG.py:567-582`: query every stored vector, top-1 must
be itself with distance ~1.0) as the known-answer test, (2) the distance
convention and row schema of the reference's real output
`/root/reference/output_emb/search_results.json` (cosine SIMILARITY, larger is
better, 0.81-0.95), and (3) agreement of three independent arithmetic variants
(f64, fp32 normalise-then-dot, fp32 dot-then-divide) on the top-5 lists.  With
no runnable Milvus Lite the parity claim is "parity unpinned with respect to the
engine; pinned to the reference's shipped artefacts and inline check" — see DESIGN.md.

What is restated (reference call sites):
  * `client.search(collection_name, data=[vec], limit=k, output_fields=[...])`
    `/root/reference/milvus/search_embeddings.py:15-22`,
    `/root/reference/milvus/RAG.py:383-390`,
    `/root/reference/milvus/search_json.py:247-254`,
    `/root/reference/src/search_milvus.py:139-146` (explicit metric_type="COSINE").
  * quick-setup collections default to COSINE
    (`/root/reference/milvus/RAG.py:54-57`, stored index meta in the .db).
  * Milvus Lite serves FLAT (exact) regardless of the requested index
    (`/root/reference/milvus/insert_embeddings.py:66-79` asks for IVF_FLAT).

Published algorithm of the absent engine (knowhere FLAT brute force): for
COSINE the query is L2-normalised, every stored row contributes
ip(q, x)/||x||, a size-k heap keeps the best, results are returned best-first
and the reduce step orders by (distance desc, pk asc).

Score definition used here (and by the CUDA path's rescoring kernel):
  COSINE: s = <x,q> / (||x|| * ||q||)      IP: s = <x,q>
evaluated in float64 from the fp32 inputs, ordered by (s desc, id asc) in
float64, returned as float32.
"""
from __future__ import annotations

import numpy as np

METRICS = ("COSINE", "IP")
_CHUNK = 1 << 15


def _rowdot64(X: np.ndarray, v64: np.ndarray) -> np.ndarray:
    """Deterministic per-row float64 dot: depends only on the row's own content
    (numpy pairwise summation along the contiguous axis), so identical rows get
    bit-identical scores and ties fall through to the id comparison."""
    out = np.empty(X.shape[0], dtype=np.float64)
    for lo in range(0, X.shape[0], _CHUNK):
        blk = X[lo:lo + _CHUNK].astype(np.float64)
        blk *= v64
        out[lo:lo + _CHUNK] = blk.sum(axis=1)
    return out


def row_norms64(X: np.ndarray) -> np.ndarray:
    out = np.empty(X.shape[0], dtype=np.float64)
    for lo in range(0, X.shape[0], _CHUNK):
        blk = X[lo:lo + _CHUNK].astype(np.float64)
        blk *= blk
        out[lo:lo + _CHUNK] = np.sqrt(blk.sum(axis=1))
    return out


def scores64(X: np.ndarray, q: np.ndarray, metric: str, xnorm: np.ndarray | None = None) -> np.ndarray:
    """float64 scores of one query against all rows."""
    q64 = np.asarray(q, dtype=np.float32).astype(np.float64)
    s = _rowdot64(X, q64)
    if metric == "COSINE":
        if xnorm is None:
            xnorm = row_norms64(X)
        qn = np.sqrt((q64 * q64).sum())
        den = xnorm * qn
        with np.errstate(divide="ignore", invalid="ignore"):
            s = np.where(den > 0, s / den, 0.0)
    elif metric != "IP":
        raise ValueError(f"metric must be one of {METRICS}")
    return s


def scores32(X: np.ndarray, q: np.ndarray, metric: str, variant: str = "normalize_then_dot") -> np.ndarray:
    """fp32-accumulate variants: bounds what an fp32 SIMD engine could return."""
    X = np.asarray(X, dtype=np.float32)
    q = np.asarray(q, dtype=np.float32)
    if metric == "IP":
        return X @ q
    xn = np.sqrt((X * X).sum(axis=1, dtype=np.float32))
    qn = np.sqrt((q * q).sum(dtype=np.float32))
    if variant == "normalize_then_dot":
        return ((X / xn[:, None]) @ (q / qn)).astype(np.float32)
    if variant == "dot_then_divide":
        return ((X @ q) / (xn * qn)).astype(np.float32)
    raise ValueError(variant)


def order_topk(s: np.ndarray, ids: np.ndarray, k: int, dedup_pk: bool = False):
    """Indices of the best k rows by (score desc, id asc[, row asc])."""
    n = s.shape[0]
    k_eff = min(k, n)
    if k_eff == 0:
        return np.zeros(0, dtype=np.int64)
    if dedup_pk:
        order = np.lexsort((np.arange(n), ids, -s))
        seen, keep = set(), []
        for i in order:
            pk = int(ids[i])
            if pk in seen:
                continue
            seen.add(pk)
            keep.append(i)
            if len(keep) == k_eff:
                break
        return np.asarray(keep, dtype=np.int64)
    if k_eff < n:
        kth = np.partition(s, n - k_eff)[n - k_eff]
        cand = np.nonzero(s >= kth)[0]
    else:
        cand = np.arange(n)
    order = np.lexsort((cand, ids[cand], -s[cand]))
    return cand[order[:k_eff]]


def search(X, ids, Q, k, metric="COSINE", accum="f64", dedup_pk=False, variant="normalize_then_dot"):
    """Exact FLAT search.

    X [N,D] fp32 raw rows, ids [N] int64 primary keys, Q [nq,D] fp32.
    Returns (ids int64[nq,k], dist fp32[nq,k], rows int64[nq,k]); slots past
    min(k,N) hold id -1, dist -inf, row -1 (the tensor-API padding; the dict
    API simply returns fewer hits).
    """
    X = np.ascontiguousarray(X, dtype=np.float32)
    Q = np.atleast_2d(np.asarray(Q, dtype=np.float32))
    ids = np.asarray(ids, dtype=np.int64)
    if metric not in METRICS:
        raise ValueError(f"metric must be one of {METRICS}")
    if X.shape[0] and Q.shape[1] != X.shape[1]:
        raise ValueError(f"query dim {Q.shape[1]} != collection dim {X.shape[1]}")
    nq = Q.shape[0]
    out_ids = np.full((nq, k), -1, dtype=np.int64)
    out_rows = np.full((nq, k), -1, dtype=np.int64)
    out_d = np.full((nq, k), -np.inf, dtype=np.float32)
    if X.shape[0] == 0:
        return out_ids, out_d, out_rows
    xnorm = row_norms64(X) if (metric == "COSINE" and accum == "f64") else None
    for i in range(nq):
        if accum == "f64":
            s = scores64(X, Q[i], metric, xnorm)
        else:
            s = scores32(X, Q[i], metric, variant).astype(np.float64)
        top = order_topk(s, ids, k, dedup_pk)
        out_ids[i, :top.size] = ids[top]
        out_rows[i, :top.size] = top
        out_d[i, :top.size] = s[top].astype(np.float32)
    return out_ids, out_d, out_rows


def search_large(X, ids, Q, k, metric="COSINE", slack=64, xnorm=None):
    """Same result as `search(accum="f64")` for big N: a BLAS float64 pass picks
    k+slack candidates per query, which are then re-evaluated with the
    deterministic per-row arithmetic and ordered exactly.  The BLAS scores differ
    from the deterministic ones by ~1e-15, far below the slack window.  NOT valid when
    more than `slack` rows tie exactly with the k-th score (use `search` then)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    Q = np.atleast_2d(np.asarray(Q, dtype=np.float32))
    ids = np.asarray(ids, dtype=np.int64)
    n, nq = X.shape[0], Q.shape[0]
    kk = min(n, k + slack)
    if xnorm is None and metric == "COSINE":
        xnorm = row_norms64(X)
    best_s = np.full((nq, 0), 0.0)
    best_r = np.zeros((nq, 0), dtype=np.int64)
    Q64 = Q.astype(np.float64)
    for lo in range(0, n, 1 << 17):
        blk = X[lo:lo + (1 << 17)].astype(np.float64)
        S = Q64 @ blk.T
        if metric == "COSINE":
            S /= np.maximum(xnorm[lo:lo + blk.shape[0]], 1e-300)[None, :]
        rows = np.broadcast_to(np.arange(lo, lo + blk.shape[0])[None, :], S.shape)
        best_s = np.concatenate([best_s, S], axis=1)
        best_r = np.concatenate([best_r, rows], axis=1)
        if best_s.shape[1] > kk:
            part = np.argpartition(-best_s, kk - 1, axis=1)[:, :kk]
            best_s = np.take_along_axis(best_s, part, axis=1)
            best_r = np.take_along_axis(best_r, part, axis=1)
    out_ids = np.full((nq, k), -1, dtype=np.int64)
    out_rows = np.full((nq, k), -1, dtype=np.int64)
    out_d = np.full((nq, k), -np.inf, dtype=np.float32)
    for i in range(nq):
        rows = np.sort(best_r[i])
        s = scores64(X[rows], Q[i], metric, None if xnorm is None else xnorm[rows])
        top = order_topk(s, ids[rows], k)
        out_ids[i, :top.size] = ids[rows[top]]
        out_rows[i, :top.size] = rows[top]
        out_d[i, :top.size] = s[top].astype(np.float32)
    return out_ids, out_d, out_rows


def merge_shards(parts, k):
    """K6 restated: parts = list of (ids[nq,k], score64[nq,k]) per shard (padded
    with id -1 / -inf); returns the global (ids, score64) by (score desc, id asc)."""
    ids = np.concatenate([p[0] for p in parts], axis=1)
    sc = np.concatenate([np.asarray(p[1], dtype=np.float64) for p in parts], axis=1)
    nq = ids.shape[0]
    out_i = np.full((nq, k), -1, dtype=np.int64)
    out_s = np.full((nq, k), -np.inf, dtype=np.float64)
    for i in range(nq):
        valid = np.nonzero(~((ids[i] == -1) & np.isneginf(sc[i])))[0]
        order = np.lexsort((ids[i][valid], -sc[i][valid]))[:k]
        sel = valid[order]
        out_i[i, :sel.size] = ids[i][sel]
        out_s[i, :sel.size] = sc[i][sel]
    return out_i, out_s


def shard_bounds(n: int, world: int):
    """Contiguous row sharding: rank r holds [r*ceil(n/W), (r+1)*ceil(n/W))."""
    per = -(-n // world) if world else n
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]


def cpu_flat_baseline(Xn: np.ndarray, Q: np.ndarray, k: int, chunk: int = 1 << 15):
    """The same-box CPU baseline: fp32 BLAS `Q @ Xn.T` on pre-normalised rows,
    chunked over N so a score tile stays cache-resident, with a running k-th-best threshold per query
    (what a FLAT engine's per-query heap amounts to): only scores above the threshold are gathered and
    merged.  Returns (rows[nq,k], scores fp32[nq,k]) ordered by (score desc, row asc)."""
    nq, n = Q.shape[0], Xn.shape[0]
    kk = min(k, n)
    best_s = np.full((nq, kk), -np.inf, dtype=np.float32)
    best_r = np.full((nq, kk), -1, dtype=np.int64)
    thr = np.full(nq, -np.inf, dtype=np.float32)
    for lo in range(0, n, chunk):
        S = Q @ Xn[lo:lo + chunk].T
        if lo == 0 and S.shape[1] >= kk:
            part = np.argpartition(S, S.shape[1] - kk, axis=1)[:, S.shape[1] - kk:]
            best_s = np.take_along_axis(S, part, axis=1)
            best_r = part.astype(np.int64)
            thr = best_s.min(axis=1)
            continue
        qi, ci = np.nonzero(S > thr[:, None])
        if qi.size == 0:
            continue
        vals = S[qi, ci]
        bounds = np.searchsorted(qi, np.arange(nq + 1))
        for q in np.unique(qi):
            a, b = bounds[q], bounds[q + 1]
            s_all = np.concatenate([best_s[q], vals[a:b]])
            r_all = np.concatenate([best_r[q], ci[a:b] + lo])
            if s_all.size > kk:
                sel = np.argpartition(s_all, s_all.size - kk)[s_all.size - kk:]
                s_all, r_all = s_all[sel], r_all[sel]
            best_s[q, :s_all.size], best_r[q, :r_all.size] = s_all, r_all
            thr[q] = s_all.min() if s_all.size == kk else -np.inf
    order = np.lexsort((best_r, -best_s), axis=1)
    return np.take_along_axis(best_r, order, axis=1), np.take_along_axis(best_s, order, axis=1)
