"""`MilvusClient` drop-in for the one path AutoStyle-TTS uses pymilvus for.

Mirrors the call surface the reference exercises (same names, argument meaning
and error behaviour — every failure is an `Exception` subclass, which is what
the reference's `try/except Exception: print; return []` wrappers expect):

  MilvusClient(uri)                              /root/reference/milvus/search_embeddings.py:31
  has_collection / drop_collection               /root/reference/milvus/RAG.py:49-50
  create_collection(name, dimension=)            /root/reference/milvus/RAG.py:54-57   (quick setup: COSINE)
  create_collection(name, schema=)               /root/reference/milvus/insert_embeddings.py:63
  create_index(name, field_name=, index_params=) /root/reference/milvus/insert_embeddings.py:75-79 (accepted, FLAT is served)
  insert(name, data=[{...}])                     /root/reference/milvus/RAG.py:541-544
  search(name, data=[vec], limit=, output_fields=, filter=None, anns_field=, param=, metric_type=, search_params=)
        /root/reference/milvus/search_embeddings.py:15-22, /root/reference/milvus/RAG.py:383-390,
        /root/reference/src/search_milvus.py:139-146
  -> list[list[{"id", "distance", "entity": {...}}]], best first, distance = similarity.

Vectors live on the GPU (autostyle-tts_b200/engine.py -> libavs.so); scalar and
dynamic fields stay on the host keyed by row index and never cross the C-ABI.
When `uri` names a file, collections persist in the Milvus Lite SQLite format
(autostyle-tts_b200/milvus_lite_db.py), so the reference's shipped
`milvus/milvus_demo.db` opens unchanged.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np

from . import milvus_lite_db as mldb
from .engine import AvsError, Store
from .filter_expr import compile_filter
from .schema import CollectionSchema, DataType, FieldSchema, IndexParams, MilvusException

MAX_LIMIT = 16384  # MilvusClient's own ceiling; limits above 256 take the exact fp32 master scan (include/avs.h)


class _Collection:
    def __init__(self, name: str, dim: int, metric: str, pk_name: str, vec_name: str, auto_id: bool,
                 fields: List[Dict[str, Any]], enable_dynamic: bool, index_params: Optional[Dict[str, Any]] = None):
        self.name, self.dim, self.metric = name, int(dim), metric
        self.pk_name, self.vec_name, self.auto_id = pk_name, vec_name, auto_id
        self.fields, self.enable_dynamic = fields, enable_dynamic
        self.index_params = dict(index_params or {})
        self.store: Optional[Store] = None
        self.pks: List[Any] = []
        self.meta: List[Dict[str, Any]] = []   # scalar + dynamic fields per row (no vector)
        self.next_auto = 1
        self.pending: List[np.ndarray] = []    # vectors loaded from disk, not yet on the device
        self.filter_cache: Dict[str, np.ndarray] = {}   # expression -> row mask, valid until the next insert

    @property
    def scalar_field_names(self) -> List[str]:
        return [f["name"] for f in self.fields if f["dtype"] != mldb.DT_FLOAT_VECTOR]


class MilvusClient:
    def __init__(self, uri: str = "./milvus_demo.db", user: str = "", password: str = "", db_name: str = "",
                 token: str = "", timeout: Optional[float] = None, **kwargs):
        self.uri = uri
        dev = kwargs.pop("device", None)
        if dev is None:
            dev = os.environ.get("AVS_DEVICE", os.environ.get("LOCAL_RANK", "0"))
        self.device = int(dev)
        self.dedup_pk = bool(kwargs.pop("dedup_pk", False))
        self._colls: Dict[str, _Collection] = {}
        self._file: Optional[mldb.MilvusLiteFile] = None
        if uri and not str(uri).startswith(("http://", "https://", "tcp://", "unix:")) and uri != ":memory:":
            if not str(uri).endswith(".db"):
                raise MilvusException(f"uri: {uri} is illegal, needs start with [unix, http, https, tcp] or a local file endswith [.db]")
            self._file = mldb.MilvusLiteFile(str(uri))
            for name in self._file.list_collections():
                self._colls[name] = self._load_collection(name)
        elif uri != ":memory:":
            raise MilvusException("this client serves local collections only (a *.db path or ':memory:'), not a remote Milvus server")

    # ------------------------------------------------------------------ collection management
    def has_collection(self, collection_name: str, timeout: Optional[float] = None, **kwargs) -> bool:
        return collection_name in self._colls

    def list_collections(self, **kwargs) -> List[str]:
        return list(self._colls)

    def drop_collection(self, collection_name: str, timeout: Optional[float] = None, **kwargs):
        c = self._colls.pop(collection_name, None)
        if c is not None and c.store is not None:
            c.store.close()
        if self._file is not None:
            self._file.drop_collection(collection_name)

    def create_schema(self, **kwargs) -> CollectionSchema:
        return CollectionSchema([], **kwargs)

    def prepare_index_params(self, field_name: str = "", **kwargs) -> IndexParams:
        p = IndexParams()
        if field_name:
            p.add_index(field_name, **kwargs)
        return p

    def create_collection(self, collection_name: str, dimension: Optional[int] = None, primary_field_name: str = "id",
                          id_type: str = "int", vector_field_name: str = "vector", metric_type: str = "COSINE",
                          auto_id: bool = False, timeout: Optional[float] = None,
                          schema: Optional[CollectionSchema] = None, index_params: Optional[Any] = None, **kwargs):
        if collection_name in self._colls:
            raise MilvusException(f"collection {collection_name} already exists")
        if schema is None:
            if dimension is None:
                raise MilvusException("create_collection needs either `dimension` (quick setup) or `schema`")
            if id_type not in ("int", "string", "str", DataType.INT64, DataType.VARCHAR):
                raise MilvusException(f"unsupported id_type {id_type!r}")
            pk_dt = mldb.DT_INT64 if id_type in ("int", DataType.INT64) else mldb.DT_VARCHAR
            fields = [
                {"name": primary_field_name, "dtype": pk_dt, "is_primary": True, "auto_id": bool(auto_id),
                 "max_length": kwargs.get("max_length") if pk_dt == mldb.DT_VARCHAR else None},
                {"name": vector_field_name, "dtype": mldb.DT_FLOAT_VECTOR, "dim": int(dimension)},
            ]
            metric = str(metric_type or "COSINE").upper()
            coll = _Collection(collection_name, int(dimension), metric, primary_field_name, vector_field_name, bool(auto_id),
                               fields, enable_dynamic=True,
                               index_params={"index_type": "AUTOINDEX", "metric_type": metric})
        else:
            schema.verify()
            pk, vec = schema.primary_field, schema.vector_field
            fields = []
            for f in schema.fields:
                if int(f.dtype) not in mldb.SUPPORTED_TYPES:
                    raise MilvusException(f"field {f.name}: data type {f.dtype.name} is not supported by this store "
                                          "(INT64, VARCHAR, JSON, FLOAT_VECTOR)")
                fields.append({"name": f.name, "dtype": int(f.dtype), "is_primary": f.is_primary,
                               "auto_id": f.auto_id or (f.is_primary and schema.auto_id), "dim": f.dim,
                               "max_length": f.max_length, "description": f.description})
            metric = self._metric_from(index_params) or (str(schema.metric_type).upper() if schema.metric_type else "COSINE")
            coll = _Collection(collection_name, int(vec.dim), metric, pk.name, vec.name,
                               bool(pk.auto_id or schema.auto_id), fields, schema.enable_dynamic_field,
                               index_params={"index_type": "AUTOINDEX", "metric_type": metric})
        if coll.metric not in ("COSINE", "IP"):
            raise MilvusException(f"metric type {coll.metric} is not supported: this store serves COSINE and IP")
        if coll.dim <= 0 or coll.dim > 32768:
            raise MilvusException(f"invalid dimension: {coll.dim}. should be in range 1 ~ 32768")
        self._colls[collection_name] = coll
        if self._file is not None:
            self._file.create_collection(collection_name, coll.fields, coll.enable_dynamic)
            self._write_index_meta(coll)

    def create_index(self, collection_name: str, index_params: Any = None, timeout: Optional[float] = None, **kwargs):
        """Accepted and recorded; the engine always serves exact FLAT search, as Milvus Lite does for
        the reference whatever index it asks for (IVF_FLAT nlist=128 at insert_embeddings.py:66-79)."""
        coll = self._coll(collection_name)
        metric = self._metric_from(index_params)
        if isinstance(index_params, dict):
            coll.index_params.update({k: v for k, v in index_params.items() if k != "params"})
            coll.index_params.update(index_params.get("params") or {})
        if metric and metric != coll.metric:
            if metric not in ("COSINE", "IP"):
                raise MilvusException(f"metric type {metric} is not supported: this store serves COSINE and IP")
            self._rebuild_with_metric(coll, metric)
        coll.index_params["metric_type"] = coll.metric
        if self._file is not None:
            self._write_index_meta(coll)

    def describe_collection(self, collection_name: str, timeout: Optional[float] = None, **kwargs) -> Dict[str, Any]:
        c = self._coll(collection_name)
        fields = []
        for i, f in enumerate(c.fields):
            d = {"field_id": 100 + i, "name": f["name"], "description": f.get("description", ""), "type": DataType(f["dtype"]),
                 "params": {}}
            if f.get("dim") is not None:
                d["params"]["dim"] = f["dim"]
            if f.get("max_length") is not None:
                d["params"]["max_length"] = f["max_length"]
            if f.get("is_primary"):
                d["is_primary"] = True
                d["auto_id"] = bool(f.get("auto_id"))
            fields.append(d)
        return {"collection_name": c.name, "auto_id": c.auto_id, "num_shards": 1, "description": "", "fields": fields,
                "enable_dynamic_field": c.enable_dynamic, "metric_type": c.metric, "index": dict(c.index_params)}

    # src/search_milvus.py:183 calls this (non-upstream) name; keep it working
    get_collection_info = describe_collection

    def get_collection_stats(self, collection_name: str, timeout: Optional[float] = None, **kwargs) -> Dict[str, int]:
        return {"row_count": len(self._coll(collection_name).pks)}

    def load_collection(self, collection_name: str, **kwargs):
        self._ensure_store(self._coll(collection_name))

    def release_collection(self, collection_name: str, **kwargs):
        self._coll(collection_name)

    def flush(self, collection_name: str, **kwargs):
        self._coll(collection_name)

    def close(self):
        for c in self._colls.values():
            if c.store is not None:
                c.store.close()
                c.store = None
        if self._file is not None:
            self._file.close()
            self._file = None

    # ------------------------------------------------------------------ insert
    def insert(self, collection_name: str, data: Union[Dict[str, Any], List[Dict[str, Any]]],
               timeout: Optional[float] = None, partition_name: str = "", **kwargs) -> Dict[str, Any]:
        coll = self._coll(collection_name)
        if isinstance(data, dict):
            data = [data]
        if not isinstance(data, (list, tuple)):
            raise MilvusException("wrong type of argument 'data', expected 'Dict' or 'List[Dict]'")
        if len(data) == 0:
            return {"insert_count": 0, "ids": []}
        declared = {f["name"] for f in coll.fields}
        # vectors: one bulk conversion when every row carries a well-formed vector (the common case; a 10^6-row insert
        # must not convert row by row), the per-row path only to produce the reference's error messages
        vecs = None
        try:
            bulk = np.asarray([row[coll.vec_name] for row in data], dtype=np.float32)
            if bulk.ndim == 2 and bulk.shape == (len(data), coll.dim):
                vecs = np.ascontiguousarray(bulk)
        except (KeyError, TypeError, ValueError):
            vecs = None
        bulk_ok = vecs is not None
        if not bulk_ok:
            vecs = np.empty((len(data), coll.dim), dtype=np.float32)
        pks, metas, rows_for_disk, dyn_for_disk = [], [], [], []
        for i, row in enumerate(data):
            if not isinstance(row, dict):
                raise MilvusException(f"wrong type of argument 'data[{i}]', expected 'Dict', got '{type(row).__name__}'")
            if coll.vec_name not in row:
                raise MilvusException(f"Insert missed an field `{coll.vec_name}` to collection without set nullable==true or set default_value")
            if bulk_ok:
                v = vecs[i]
            else:
                v = np.asarray(row[coll.vec_name], dtype=np.float32).reshape(-1)
                if v.shape[0] != coll.dim:
                    raise MilvusException(f"the length({v.shape[0]}) of float data should divide the dim({coll.dim})")
                vecs[i] = v
            if coll.auto_id:
                if coll.pk_name in row:
                    raise MilvusException(f"Attempt to insert an unexpected field `{coll.pk_name}` to collection without enabling dynamic field"
                                          if not coll.enable_dynamic else f"auto_id is enabled: do not pass `{coll.pk_name}`")
                pk = coll.next_auto + i
            else:
                if coll.pk_name not in row:
                    raise MilvusException(f"Insert missed an field `{coll.pk_name}` to collection without set nullable==true or set default_value")
                pk = row[coll.pk_name]
            pks.append(pk)
            meta: Dict[str, Any] = {}
            dyn: Dict[str, Any] = {}
            for k, val in row.items():
                if k in (coll.vec_name, coll.pk_name):
                    continue
                if k in declared:
                    meta[k] = val
                elif coll.enable_dynamic:
                    meta[k] = val
                    dyn[k] = val
                else:
                    raise MilvusException(f"Attempt to insert an unexpected field `{k}` to collection without enabling dynamic field")
            for f in coll.fields:
                if f["name"] not in (coll.vec_name, coll.pk_name) and f["name"] not in meta:
                    raise MilvusException(f"Insert missed an field `{f['name']}` to collection without set nullable==true or set default_value")
            metas.append(meta)
            if self._file is not None:
                disk = dict(meta)
                disk[coll.pk_name] = pk
                disk[coll.vec_name] = v
                rows_for_disk.append(disk)
                dyn_for_disk.append(dyn if coll.enable_dynamic else None)
        if not np.all(np.isfinite(vecs)):
            raise MilvusException("float vector contains NaN or Inf")
        id_arr = self._ids_for_device(coll, pks)
        store = self._ensure_store(coll)
        try:
            store.insert(vecs, id_arr)
        except AvsError as e:
            raise MilvusException(e.message, e.code) from e
        coll.pks.extend(pks)
        coll.meta.extend(metas)
        coll.filter_cache.clear()
        if coll.auto_id:
            coll.next_auto += len(data)
        if self._file is not None:
            self._file.append(coll.name, coll.fields, coll.pk_name, rows_for_disk, dyn_for_disk)
        return {"insert_count": len(data), "ids": list(pks), "cost": 0}

    # ------------------------------------------------------------------ bulk snapshot
    def save_snapshot(self, collection_name: str, path: str):
        """Raw snapshot of a collection for 10^7-10^8-row stores: `path` = ids + fp32 vectors straight from the device
        (avs_save), `path + ".meta.json"` = schema, primary keys and scalar / dynamic fields.  The per-row Milvus Lite
        file (`MilvusClient("x.db")`, what the reference re-opens: /root/reference/milvus/search.py:197-210) stays the
        persistence of small collections."""
        import json
        coll = self._coll(collection_name)
        try:
            self._ensure_store(coll).save(path)
        except AvsError as e:
            raise MilvusException(e.message, e.code) from e
        with open(path + ".meta.json", "w", encoding="utf-8") as f:
            json.dump({"name": coll.name, "dim": coll.dim, "metric": coll.metric, "pk_name": coll.pk_name, "vec_name": coll.vec_name,
                       "auto_id": coll.auto_id, "fields": coll.fields, "enable_dynamic": coll.enable_dynamic,
                       "index_params": coll.index_params, "next_auto": coll.next_auto, "pks": coll.pks, "meta": coll.meta}, f)

    def load_snapshot(self, path: str, collection_name: Optional[str] = None) -> str:
        """Restores a collection saved by save_snapshot into this client (device store rebuilt by avs_load: the bf16 scan
        copy and norms come from the normalise-on-insert kernel, not from disk).  Returns the collection name."""
        import json
        with open(path + ".meta.json", encoding="utf-8") as f:
            m = json.load(f)
        name = collection_name or m["name"]
        if name in self._colls:
            raise MilvusException(f"collection {name} already exists")
        coll = _Collection(name, m["dim"], m["metric"], m["pk_name"], m["vec_name"], m["auto_id"], m["fields"], m["enable_dynamic"],
                           index_params=m.get("index_params"))
        coll.pks, coll.meta, coll.next_auto = m["pks"], m["meta"], m.get("next_auto", 1)
        try:
            coll.store = Store.load(path, device=self.device)
        except AvsError as e:
            raise MilvusException(e.message, e.code) from e
        if len(coll.store) != len(coll.pks):
            coll.store.close()
            raise MilvusException(f"snapshot {path}: {len(coll.store)} vectors but {len(coll.pks)} metadata rows")
        self._colls[name] = coll
        return name

    # ------------------------------------------------------------------ search
    def search(self, collection_name: str, data: Any = None, filter: Optional[str] = "", limit: int = 10,
               output_fields: Optional[Sequence[str]] = None, search_params: Optional[Dict[str, Any]] = None,
               timeout: Optional[float] = None, partition_names: Optional[List[str]] = None,
               anns_field: Optional[str] = None, **kwargs) -> List[List[Dict[str, Any]]]:
        coll = self._coll(collection_name)
        if data is None:
            raise MilvusException("search needs `data` (one query vector or a list of them)")
        if anns_field not in (None, "", coll.vec_name):
            raise MilvusException(f"failed to get field schema by name: fieldName({anns_field}) not found")
        want = kwargs.pop("metric_type", None) or (search_params or {}).get("metric_type") or (kwargs.get("param") or {}).get("metric_type")
        if want and str(want).upper() != coll.metric:
            raise MilvusException(f"metric type not match: invalid parameter[expected={coll.metric}][actual={str(want).upper()}]")
        limit = int(limit)
        if limit < 1 or limit > MAX_LIMIT:
            raise MilvusException(f"limit {limit} is out of range [1, {MAX_LIMIT}]")
        q = self._as_queries(data, coll.dim)
        nq = q.shape[0]
        n_rows = len(coll.pks)
        if n_rows == 0:
            return [[] for _ in range(nq)]
        store = self._ensure_store(coll)
        mask = self._filter_mask(coll, filter)      # scalar expression -> row bitmap applied inside the scan
        if mask is not None and not mask.any():
            return [[] for _ in range(nq)]
        k = limit
        try:
            if mask is not None:
                store.set_filter(mask)
            while True:
                _, dist, rows = store.search(q, min(k, MAX_LIMIT), return_rows=True)
                if not self.dedup_pk or k >= min(n_rows, MAX_LIMIT):
                    break
                if all(len({coll.pks[r] for r in rows[i] if r >= 0}) >= min(limit, n_rows) for i in range(nq)):
                    break
                k = min(k * 2, MAX_LIMIT)
        except AvsError as e:
            raise MilvusException(e.message, e.code) from e
        finally:
            if mask is not None:
                store.set_filter(None)
        names = self._resolve_output_fields(coll, output_fields)
        fetch_vec = coll.vec_name in names
        out: List[List[Dict[str, Any]]] = []
        for i in range(nq):
            hits: List[Dict[str, Any]] = []
            seen = set()
            for j in range(rows.shape[1]):
                r = int(rows[i, j])
                if r < 0:
                    break
                pk = coll.pks[r]
                if self.dedup_pk:
                    if pk in seen:
                        continue
                    seen.add(pk)
                ent: Dict[str, Any] = {}
                meta = coll.meta[r]
                for nme in names:
                    if nme == coll.vec_name:
                        continue
                    if nme == coll.pk_name:
                        ent[nme] = pk
                    elif nme in meta:
                        ent[nme] = meta[nme]
                if fetch_vec:
                    ent[coll.vec_name] = store.get_rows(r, 1)[0].tolist()
                hits.append({"id": pk, "distance": float(dist[i, j]), "entity": ent})
                if len(hits) == limit:
                    break
            out.append(hits)
        return out

    def search_tensors(self, collection_name: str, queries, limit: int = 10, metric_type: Optional[str] = None,
                       return_rows: bool = False):
        """Benchmark-grade surface (BASELINE.json north_star): queries [nq, D] as a torch CUDA tensor
        (returns device tensors, asynchronous) or host array (returns numpy arrays); no Python dicts.
        -> ids int64 [nq, limit], distances fp32 [nq, limit] (+ row indices for metadata lookup)."""
        coll = self._coll(collection_name)
        if metric_type and str(metric_type).upper() != coll.metric:
            raise MilvusException(f"metric type not match: invalid parameter[expected={coll.metric}][actual={str(metric_type).upper()}]")
        try:
            return self._ensure_store(coll).search(queries, int(limit), return_rows=return_rows)
        except AvsError as e:
            raise MilvusException(e.message, e.code) from e

    # ------------------------------------------------------------------ get / query by primary key
    def get(self, collection_name: str, ids: Union[list, str, int], output_fields: Optional[Sequence[str]] = None,
            **kwargs) -> List[Dict[str, Any]]:
        coll = self._coll(collection_name)
        if not isinstance(ids, (list, tuple)):
            ids = [ids]
        wanted = set(ids)
        names = self._resolve_output_fields(coll, output_fields if output_fields is not None else ["*"])
        out = []
        for r, pk in enumerate(coll.pks):
            if pk in wanted:
                ent = {coll.pk_name: pk}
                ent.update({k: v for k, v in coll.meta[r].items() if k in names})
                if coll.vec_name in names:
                    ent[coll.vec_name] = self._ensure_store(coll).get_rows(r, 1)[0].tolist()
                out.append(ent)
        return out

    def query(self, collection_name: str, filter: str = "", output_fields: Optional[Sequence[str]] = None,
              ids: Optional[Union[list, str, int]] = None, limit: Optional[int] = None, **kwargs) -> List[Dict[str, Any]]:
        if ids is not None:
            res = self.get(collection_name, ids, output_fields)
        else:
            coll = self._coll(collection_name)
            names = self._resolve_output_fields(coll, output_fields if output_fields is not None else ["*"])
            mask = self._filter_mask(coll, filter)
            res = []
            for r, pk in enumerate(coll.pks):
                if mask is not None and not mask[r]:
                    continue
                ent = {coll.pk_name: pk}
                ent.update({k: v for k, v in coll.meta[r].items() if k in names})
                res.append(ent)
                if limit is not None and len(res) >= limit:
                    break
        return res[:limit] if limit is not None else res

    # ------------------------------------------------------------------ internals
    @staticmethod
    def _filter_mask(coll: _Collection, expr: Optional[str]) -> Optional[np.ndarray]:
        """Evaluates a scalar filter expression over the host-side fields -> bool mask per row (None = no filter)."""
        if expr in (None, ""):
            return None
        if not isinstance(expr, str):
            raise MilvusException(f"wrong type of argument 'filter', expected 'str', got '{type(expr).__name__}'")
        cached = coll.filter_cache.get(expr)
        if cached is not None and cached.shape[0] == len(coll.pks):
            return cached
        pred = compile_filter(expr)
        mask = np.asarray(pred.rows(coll.meta, coll.pk_name, coll.pks), dtype=bool)
        if len(coll.filter_cache) >= 16:
            coll.filter_cache.pop(next(iter(coll.filter_cache)))
        coll.filter_cache[expr] = mask
        return mask

    def _coll(self, name: str) -> _Collection:
        c = self._colls.get(name)
        if c is None:
            raise MilvusException(f"collection not found[collection={name}]", code=100)
        return c

    @staticmethod
    def _metric_from(index_params: Any) -> Optional[str]:
        if index_params is None:
            return None
        entries = index_params if isinstance(index_params, (list, tuple)) else [index_params]
        for e in entries:
            if isinstance(e, dict):
                m = e.get("metric_type") or (e.get("params") or {}).get("metric_type")
                if m:
                    return str(m).upper()
        return None

    def _write_index_meta(self, coll: _Collection):
        vec_idx = next(i for i, f in enumerate(coll.fields) if f["name"] == coll.vec_name)
        params = dict(coll.index_params)
        params.setdefault("index_type", "AUTOINDEX")
        params["metric_type"] = coll.metric
        params["dim"] = coll.dim
        self._file.write_index(coll.name, 100 + vec_idx, coll.vec_name, params)

    def _load_collection(self, name: str) -> _Collection:
        schema, index = self._file.read_meta(name)
        fields = [f for f in schema["fields"] if not f["is_dynamic"]]
        pk = next((f for f in fields if f["is_primary"]), None)
        vec = next((f for f in fields if f["dtype"] == mldb.DT_FLOAT_VECTOR), None)
        if pk is None or vec is None or not vec.get("dim"):
            raise MilvusException(f"collection {name} in {self.uri}: no primary key / float vector field")
        metric = str(index.get("metric_type", "COSINE")).upper()
        coll = _Collection(name, vec["dim"], metric, pk["name"], vec["name"], bool(pk["auto_id"]), fields,
                           bool(schema["enable_dynamic_field"]), index_params=index)
        for ent in self._file.load_rows(name):
            coll.pks.append(ent.get(pk["name"]))
            meta = {k: v for k, v in ent.items() if k not in (pk["name"], vec["name"], mldb.META_FIELD)}
            meta.update(ent.get(mldb.META_FIELD) or {})
            coll.meta.append(meta)
            coll.pending.append(np.asarray(ent[vec["name"]], dtype=np.float32))
        if coll.auto_id and coll.pks:
            coll.next_auto = max(int(p) for p in coll.pks) + 1
        return coll

    def _ids_for_device(self, coll: _Collection, pks: List[Any]) -> np.ndarray:
        """The device needs an int64 tie-break key per row: the primary key itself when it is an
        integer, otherwise the insertion order (string keys compare on the host only)."""
        if all(isinstance(p, (int, np.integer)) and not isinstance(p, bool) for p in pks):
            return np.asarray(pks, dtype=np.int64)
        base = len(coll.pks)
        return np.arange(base, base + len(pks), dtype=np.int64)

    def _ensure_store(self, coll: _Collection) -> Store:
        if coll.store is None:
            try:
                coll.store = Store(coll.dim, coll.metric, capacity=max(1024, len(coll.pending)), device=self.device)
                if coll.pending:
                    vecs = np.stack(coll.pending).astype(np.float32)
                    saved, coll.pks = coll.pks, []
                    ids = self._ids_for_device(coll, saved)
                    coll.pks = saved
                    coll.store.insert(vecs, ids)
                    coll.pending = []
            except AvsError as e:
                coll.store = None
                raise MilvusException(e.message, e.code) from e
        return coll.store

    def _rebuild_with_metric(self, coll: _Collection, metric: str):
        old = coll.store
        coll.metric = metric
        if old is None:
            return
        n = len(old)
        vecs = old.get_rows(0, n) if n else np.zeros((0, coll.dim), np.float32)
        ids = old.get_ids(0, n) if n else np.zeros(0, np.int64)
        old.close()
        coll.store = Store(coll.dim, metric, capacity=max(1024, n), device=self.device)
        if n:
            coll.store.insert(vecs, ids)

    @staticmethod
    def _as_queries(data: Any, dim: int) -> np.ndarray:
        try:
            import torch
            if isinstance(data, torch.Tensor):
                data = data.detach().to("cpu", torch.float32).numpy()
        except ImportError:  # pragma: no cover
            pass
        try:
            q = np.asarray(data, dtype=np.float32)
        except (ValueError, TypeError) as e:
            raise MilvusException(f"`data` must be a vector or a list of vectors of equal length: {e}") from e
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.ndim != 2 or q.shape[1] != dim:
            raise MilvusException(f"vector dimension mismatch, expected vector size(byte) {dim * 4}, actual {q.shape[-1] * 4 if q.ndim else 0}.")
        if not np.all(np.isfinite(q)):
            raise MilvusException("query vector contains NaN or Inf")
        return np.ascontiguousarray(q)

    @staticmethod
    def _resolve_output_fields(coll: _Collection, output_fields: Optional[Sequence[str]]) -> List[str]:
        if not output_fields:
            return []
        names: List[str] = []
        for f in output_fields:
            if f == "*":
                names.extend(n for n in coll.scalar_field_names if n not in names)
                if coll.enable_dynamic:
                    for m in coll.meta[:4096]:
                        names.extend(k for k in m if k not in names)
            elif f not in names:
                names.append(f)
        declared = {f["name"] for f in coll.fields}
        if not coll.enable_dynamic:
            for n in names:
                if n not in declared:
                    raise MilvusException(f"field {n} not exist")
        return names
