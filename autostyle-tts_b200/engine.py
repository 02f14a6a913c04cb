"""ctypes binding of libavs.so — the only road from Python to the CUDA kernels.

There is deliberately no CPU fallback: if the library is missing or no B200 is
visible, construction raises.  PyTorch is used for device buffers and stream
hand-off only (`tensor.data_ptr()`); nothing here computes a score on the host.

C-ABI: /root/repo/include/avs.h (each entry point cites the reference call it
replaces).
"""
from __future__ import annotations

import ctypes
import os
import warnings
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libavs.so")

METRIC_CODES = {"COSINE": 0, "IP": 1}
METRIC_NAMES = {0: "COSINE", 1: "IP"}

# every symbol include/avs.h declares; tests check the built library exports all of them
ABI_SYMBOLS = (
    "avs_create", "avs_destroy", "avs_reserve", "avs_insert", "avs_fill_synthetic", "avs_count", "avs_dim",
    "avs_metric", "avs_get_rows", "avs_get_ids", "avs_save", "avs_load", "avs_set_filter", "avs_search", "avs_search_host", "avs_nccl_unique_id",
    "avs_comm_init", "avs_search_sharded", "avs_search_sharded_host", "avs_p2p_init", "avs_p2p_connect", "avs_set_option", "avs_get_stat", "avs_scan_timing",
    "avs_last_error", "avs_version",
)


class AvsError(RuntimeError):
    """Raised for every non-zero status of the C-ABI (message from avs_last_error)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[avs {code}] {message}")
        self.code = code
        self.message = message


class ExactnessWarning(UserWarning):
    """A search returned hits whose exactness certificate could not be established (see avs_get_stat)."""


_lib = None


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """dlopen libavs.so and declare the prototypes.  Needs no GPU (symbols only)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("AVS_LIB") or LIB_PATH      # AVS_LIB: A/B runs of an older build of the library (profiles/)
    if not os.path.exists(p):
        raise ImportError(
            f"{p} is missing: build it with `python autostyle-tts_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
    vp, i64, i32, u64, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint64, ctypes.c_double
    protos = {
        "avs_create": (i32, [i32, i32, i32, i64, ctypes.POINTER(vp)]),
        "avs_destroy": (i32, [vp]),
        "avs_reserve": (i32, [vp, i64]),
        "avs_insert": (i32, [vp, vp, vp, i64, vp]),
        "avs_fill_synthetic": (i32, [vp, u64, i64, i64, i64, vp]),
        "avs_count": (i64, [vp]),
        "avs_dim": (i32, [vp]),
        "avs_metric": (i32, [vp]),
        "avs_get_rows": (i32, [vp, i64, i64, vp, vp]),
        "avs_get_ids": (i32, [vp, i64, i64, vp, vp]),
        "avs_save": (i32, [vp, ctypes.c_char_p]),
        "avs_load": (i32, [ctypes.c_char_p, i32, ctypes.POINTER(vp)]),
        "avs_set_filter": (i32, [vp, vp, i64]),
        "avs_search": (i32, [vp, vp, i32, i32, vp, vp, vp, vp]),
        "avs_search_host": (i32, [vp, vp, i32, i32, vp, vp, vp]),
        "avs_nccl_unique_id": (i32, [vp]),
        "avs_comm_init": (i32, [vp, vp, i32, i32]),
        "avs_search_sharded": (i32, [vp, vp, i32, i32, vp, vp, vp]),
        "avs_search_sharded_host": (i32, [vp, vp, i32, i32, vp, vp]),
        "avs_p2p_init": (i32, [vp, i32, i32, vp]),
        "avs_p2p_connect": (i32, [vp, vp, i32]),
        "avs_set_option": (i32, [vp, ctypes.c_char_p, i64]),
        "avs_get_stat": (i32, [vp, ctypes.c_char_p, ctypes.POINTER(i64)]),
        "avs_scan_timing": (i32, [vp, i32, ctypes.POINTER(dbl), ctypes.POINTER(i64)]),
        "avs_last_error": (ctypes.c_char_p, []),
        "avs_version": (ctypes.c_char_p, []),
    }
    for name, (res, args) in protos.items():
        if os.environ.get("AVS_LIB") and not hasattr(lib, name):
            continue                                       # an older build does not export the newer entry points
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _check(lib, rc: int):
    if rc != 0:
        raise AvsError(rc, (lib.avs_last_error() or b"").decode("utf-8", "replace"))


def _as_f32_matrix(a, dim: int, what: str) -> np.ndarray:
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if arr.ndim == 1:
        arr = arr.reshape(1, -1)
    if arr.ndim != 2 or arr.shape[1] != dim:
        raise AvsError(-1, f"{what}: expected vectors of dimension {dim}, got array of shape {tuple(arr.shape)}")
    return arr


class Store:
    """One device-resident FLAT collection on one GPU.

    search() accepts torch CUDA tensors (zero-copy fast path, returns device
    tensors) or host data (list / ndarray / CPU tensor -> avs_search_host, which
    stages H2D, runs the pipeline and copies the hits back).
    """

    def __init__(self, dim: int, metric: str = "COSINE", capacity: int = 0, device: int = 0):
        metric = str(metric).upper()
        if metric not in METRIC_CODES:
            raise AvsError(-1, f"metric_type must be COSINE or IP, got {metric!r}")
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        self.dim, self.metric, self.device = int(dim), metric, int(device)
        _check(self._lib, self._lib.avs_create(self.device, self.dim, METRIC_CODES[metric], int(capacity),
                                               ctypes.byref(self._h)))
        # AVS_OPTS="key=value,...": engine options applied to every store (A/B runs of the whole test suite under a
        # non-default schedule, profiles/); unset in normal use
        for kv in filter(None, os.environ.get("AVS_OPTS", "").split(",")):
            key, _, val = kv.partition("=")
            self.set_option(key.strip(), int(val))

    # -- bulk snapshot ------------------------------------------------------------------------
    def save(self, path: str):
        """Raw snapshot (ids + fp32 master rows) for 10^7-10^8-row stores; see avs_save in include/avs.h."""
        _check(self._lib, self._lib.avs_save(self._h, os.fsencode(path)))

    @classmethod
    def load(cls, path: str, device: int = 0) -> "Store":
        lib = load_library()
        h = ctypes.c_void_p()
        _check(lib, lib.avs_load(os.fsencode(path), int(device), ctypes.byref(h)))
        self = cls.__new__(cls)
        self._lib, self._h, self.device = lib, h, int(device)
        self.dim, self.metric = int(lib.avs_dim(h)), METRIC_NAMES[int(lib.avs_metric(h))]
        return self

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.avs_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self._lib.avs_count(self._h))

    count = property(__len__)

    # -- build ---------------------------------------------------------------------------------
    def reserve(self, capacity: int):
        _check(self._lib, self._lib.avs_reserve(self._h, int(capacity)))

    def insert(self, rows, ids=None, stream: int = 0):
        """rows: [n, dim] fp32 (ndarray / list / torch tensor on CPU or on this device)."""
        rp, n, keep = self._rows_ptr(rows, "insert")
        ip, keep2 = 0, None
        if ids is not None:
            ip, keep2 = self._ids_ptr(ids, n)
        _check(self._lib, self._lib.avs_insert(self._h, rp, ip or None, n, stream or None))
        if keep is not None or keep2 is not None:
            self.synchronize()  # host staging arrays must outlive the async copy
        return n

    def fill_synthetic(self, seed: int, first_row: int, n: int, id_base: int = 0, stream: int = 0):
        _check(self._lib, self._lib.avs_fill_synthetic(self._h, int(seed), int(first_row), int(n), int(id_base),
                                                       stream or None))

    def get_rows(self, first: int, n: int) -> np.ndarray:
        out = np.empty((int(n), self.dim), dtype=np.float32)
        _check(self._lib, self._lib.avs_get_rows(self._h, int(first), int(n), out.ctypes.data, None))
        return out

    def get_rows_device(self, first: int, n: int):
        """Master rows [first, first+n) as a torch CUDA tensor on the store's device (device-to-device copy)."""
        torch = _torch()
        out = torch.empty((int(n), self.dim), dtype=torch.float32, device=torch.device("cuda", self.device))
        stream = torch.cuda.current_stream(out.device).cuda_stream
        _check(self._lib, self._lib.avs_get_rows(self._h, int(first), int(n), out.data_ptr(), stream or None))
        return out

    def get_ids(self, first: int, n: int) -> np.ndarray:
        out = np.empty(int(n), dtype=np.int64)
        _check(self._lib, self._lib.avs_get_ids(self._h, int(first), int(n), out.ctypes.data, None))
        return out

    # -- search --------------------------------------------------------------------------------
    def search(self, queries, k: int, return_rows: bool = False, sharded: bool = False):
        """-> (ids int64[nq,k], scores fp32[nq,k][, rows int64[nq,k]]).

        Device tensor in -> device tensors out (asynchronous on torch's current stream);
        host data in -> numpy arrays out (synchronous, copies included)."""
        torch = _torch()
        if torch is not None and isinstance(queries, torch.Tensor) and queries.is_cuda:
            q = queries
            if q.dim() == 1:
                q = q.unsqueeze(0)
            if q.dim() != 2 or q.shape[1] != self.dim:
                raise AvsError(-1, f"search: expected query vectors of dimension {self.dim}, got shape {tuple(q.shape)}")
            if q.device.index != self.device:
                raise AvsError(-1, f"search: queries live on cuda:{q.device.index}, the store on cuda:{self.device}")
            q = q.to(torch.float32).contiguous()
            nq = q.shape[0]
            ids = torch.empty((nq, k), dtype=torch.int64, device=q.device)
            sc = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            rows = torch.empty((nq, k), dtype=torch.int64, device=q.device) if return_rows else None
            stream = torch.cuda.current_stream(q.device).cuda_stream
            if sharded:
                _check(self._lib, self._lib.avs_search_sharded(self._h, q.data_ptr(), nq, int(k), ids.data_ptr(),
                                                               sc.data_ptr(), stream or None))
            else:
                _check(self._lib, self._lib.avs_search(self._h, q.data_ptr(), nq, int(k), ids.data_ptr(), sc.data_ptr(),
                                                       rows.data_ptr() if rows is not None else None, stream or None))
            return (ids, sc, rows) if return_rows else (ids, sc)
        if torch is not None and isinstance(queries, torch.Tensor):
            queries = queries.detach().cpu().numpy()
        q = _as_f32_matrix(queries, self.dim, "search")
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.int64)
        sc = np.empty((nq, k), dtype=np.float32)
        if sharded:
            _check(self._lib, self._lib.avs_search_sharded_host(self._h, q.ctypes.data, nq, int(k), ids.ctypes.data, sc.ctypes.data))
            return ids, sc
        rows = np.empty((nq, k), dtype=np.int64)
        _check(self._lib, self._lib.avs_search_host(self._h, q.ctypes.data, nq, int(k), ids.ctypes.data, sc.ctypes.data,
                                                    rows.ctypes.data))
        unproven = self.stat("last_uncertified")
        if unproven:
            warnings.warn(f"{unproven} of {nq} queries could not be proven exact: more rows than the exact-repair slice holds "
                          "(up to 4096) score within 2^-15 of their k-th best (masses of duplicate rows). The hits returned for "
                          "them are the best found, but may differ from the exact top-k in the tied tail", ExactnessWarning)
        return (ids, sc, rows) if return_rows else (ids, sc)

    def set_filter(self, mask):
        """mask: bool array with one entry per stored row (True = may be returned), or None to clear."""
        if mask is None:
            _check(self._lib, self._lib.avs_set_filter(self._h, None, 0))
            return
        m = np.ascontiguousarray(np.asarray(mask, dtype=bool)).reshape(-1)
        bits = np.packbits(m, bitorder="little")
        words = np.zeros((m.shape[0] + 31) // 32 * 4, dtype=np.uint8)
        words[: bits.shape[0]] = bits
        _check(self._lib, self._lib.avs_set_filter(self._h, words.ctypes.data, int(m.shape[0])))

    # -- multi-GPU -----------------------------------------------------------------------------
    @staticmethod
    def nccl_unique_id() -> bytes:
        lib = load_library()
        buf = ctypes.create_string_buffer(128)
        _check(lib, lib.avs_nccl_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        if len(unique_id) != 128:
            raise AvsError(-1, "ncclUniqueId must be 128 bytes")
        buf = ctypes.create_string_buffer(unique_id, 128)
        _check(self._lib, self._lib.avs_comm_init(self._h, buf, int(rank), int(world)))

    def p2p_init(self, rank: int, world: int) -> bytes:
        """Allocates this rank's peer exchange region; returns its 64-byte CUDA IPC handle."""
        buf = ctypes.create_string_buffer(64)
        _check(self._lib, self._lib.avs_p2p_init(self._h, int(rank), int(world), buf))
        return buf.raw

    def p2p_connect(self, handles: bytes, world: int):
        if len(handles) != 64 * world:
            raise AvsError(-1, "p2p_connect needs world x 64 bytes of IPC handles in rank order")
        buf = ctypes.create_string_buffer(handles, len(handles))
        _check(self._lib, self._lib.avs_p2p_connect(self._h, buf, int(world)))

    # -- knobs ---------------------------------------------------------------------------------
    def set_option(self, key: str, value: int):
        _check(self._lib, self._lib.avs_set_option(self._h, key.encode(), int(value)))

    def stat(self, key: str) -> int:
        out = ctypes.c_int64()
        _check(self._lib, self._lib.avs_get_stat(self._h, key.encode(), ctypes.byref(out)))
        return int(out.value)

    def scan_timing(self, enable_reset: int = -1):
        """enable_reset: 1 start/reset, 0 stop, -1 just read.  -> (mean_ms, launches)"""
        ms, n = ctypes.c_double(), ctypes.c_int64()
        _check(self._lib, self._lib.avs_scan_timing(self._h, int(enable_reset), ctypes.byref(ms), ctypes.byref(n)))
        return float(ms.value), int(n.value)

    def synchronize(self):
        torch = _torch()
        if torch is not None and torch.cuda.is_available():
            torch.cuda.synchronize(self.device)

    # -- helpers -------------------------------------------------------------------------------
    def _rows_ptr(self, rows, what):
        torch = _torch()
        if torch is not None and isinstance(rows, torch.Tensor):
            t = rows
            if t.dim() == 1:
                t = t.unsqueeze(0)
            if t.dim() != 2 or t.shape[1] != self.dim:
                raise AvsError(-1, f"{what}: expected vectors of dimension {self.dim}, got shape {tuple(t.shape)}")
            t = t.to(torch.float32).contiguous()
            if t.is_cuda:
                self._keep = t
                return t.data_ptr(), t.shape[0], None
            arr = t.numpy()
            return arr.ctypes.data, arr.shape[0], arr
        arr = _as_f32_matrix(rows, self.dim, what)
        return arr.ctypes.data, arr.shape[0], arr

    def _ids_ptr(self, ids, n):
        torch = _torch()
        if torch is not None and isinstance(ids, torch.Tensor) and ids.is_cuda:
            t = ids.to(torch.int64).contiguous()
            if t.numel() != n:
                raise AvsError(-1, f"insert: {t.numel()} ids for {n} rows")
            self._keep_ids = t
            return t.data_ptr(), None
        if torch is not None and isinstance(ids, torch.Tensor):
            ids = ids.numpy()
        arr = np.ascontiguousarray(np.asarray(ids, dtype=np.int64)).reshape(-1)
        if arr.shape[0] != n:
            raise AvsError(-1, f"insert: {arr.shape[0]} ids for {n} rows")
        return arr.ctypes.data, arr


def _torch():
    try:
        import torch
        return torch
    except Exception:  # pragma: no cover
        return None
