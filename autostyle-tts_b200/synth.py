"""Host replay of the device-side synthetic row generator (csrc/store.cu,
`avs_fill_synthetic`): integer hashing + correctly rounded sqrt/divide only, so
numpy reproduces every row bit for bit.  Benchmark / test tooling — the search
path never calls it.

  z = mix(mix(seed + row*G1) ^ (col+1)*G2)          splitmix64 finaliser
  v = (sum of the four 16-bit fields of z) - 131070  Irwin-Hall(4), symmetric integer
  x = float32(float64(v) / sqrt(float64(sum_c v^2)))  unit-norm row
"""
from __future__ import annotations

import numpy as np

_G1 = np.uint64(0x9E3779B97F4A7C15)
_G2 = np.uint64(0xD1B54A32D192ED03)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _mix(z: np.ndarray) -> np.ndarray:
    z = z ^ (z >> np.uint64(30))
    z = z * _M1
    z = z ^ (z >> np.uint64(27))
    z = z * _M2
    return z ^ (z >> np.uint64(31))


def synth_ints(seed: int, first_row: int, n: int, dim: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        rows = np.arange(first_row, first_row + n, dtype=np.uint64)
        cols = np.arange(1, dim + 1, dtype=np.uint64)
        zr = _mix(np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + rows * _G1)
        z = _mix(zr[:, None] ^ (cols[None, :] * _G2))
    m = np.uint64(0xFFFF)
    v = ((z & m) + ((z >> np.uint64(16)) & m) + ((z >> np.uint64(32)) & m) + (z >> np.uint64(48))).astype(np.int64)
    return v - 131070


def synth_rows(seed: int, first_row: int, n: int, dim: int) -> np.ndarray:
    """Rows [first_row, first_row+n) of stream `seed` as float32 [n, dim], unit norm."""
    v = synth_ints(seed, first_row, n, dim)
    ss = (v * v).sum(axis=1)
    nrm = np.sqrt(ss.astype(np.float64))
    with np.errstate(divide="ignore", invalid="ignore"):
        x = np.where(nrm[:, None] > 0, v.astype(np.float64) / nrm[:, None], 0.0)
    return x.astype(np.float32)


def planted_queries(seed_q: int, seed_db: int, n_db: int, nq: int, dim: int, planted_frac: float = 0.1,
                    noise: float = 0.1, seed_pick: int = 44, return_planted: bool = False):
    """Query set of SURVEY.md section 8d: stream `seed_q` rows, with a `planted_frac` share replaced by
    normalize(x_j + noise * g) for random database rows j (so true neighbours exist)."""
    q = synth_rows(seed_q, 0, nq, dim)
    n_pl = int(round(nq * planted_frac))
    slots = rows = np.zeros(0, dtype=np.int64)
    if n_pl and n_db:
        rng = np.random.default_rng(seed_pick)
        slots = rng.choice(nq, size=n_pl, replace=False)
        rows = rng.integers(0, n_db, size=n_pl)
        g = synth_rows(seed_q ^ 0x5DEECE66D, 1 << 40, n_pl, dim)
        for s, r, gi in zip(slots, rows, g):
            x = synth_rows(seed_db, int(r), 1, dim)[0].astype(np.float64) + noise * gi.astype(np.float64)
            q[s] = (x / np.linalg.norm(x)).astype(np.float32)
    if return_planted:
        return q, np.asarray(slots, dtype=np.int64), np.asarray(rows, dtype=np.int64)
    return q
