"""Reader/writer for the Milvus Lite SQLite file the reference's scripts share
between processes: `MilvusClient("milvus_demo.db")` builds it in
/root/reference/milvus/RAG.py:46-57,541-544 and every search script re-opens it
(/root/reference/milvus/search_embeddings.py:31, /root/reference/milvus/search.py:197-210,
/root/reference/milvus/search_json.py:337-350).  Keeping that file format is what
lets the drop-in client open the reference's shipped `milvus/milvus_demo.db`
unchanged and persist new collections the same way.

Wire format (protobuf, hand-rolled — no protobuf dependency):
  collection_meta(id, collection_name, meta_type in {schema,index}, blob_field, string_field)
    schema blob: name=1, fields=4{fieldID=1,name=2,is_primary=3,description=4,data_type=5,
                 type_params=6{key=1,value=2},autoID=8,is_dynamic=12}, enable_dynamic_field=5
    index blob:  fieldID=1, indexID=2, index_name=3, params=5{key=1,value=2}
  "<collection>"(id INTEGER PK, milvus_id VARCHAR, data BLOB)
    data: fields=1{type=1,field_name=2,scalars=3,vectors=4,field_id=5}, num_rows=2
      Int64 -> scalars.long_data(3).data(1) packed varint; VarChar -> scalars.string_data(6).data(1)
      JSON  -> scalars.json_data(9).data(1);  FloatVector -> vectors{dim=1, float_vector(2).data(1)}
"""
from __future__ import annotations

import json
import os
import sqlite3
import time
from typing import Any, Dict, Iterable, List, Optional, Tuple

import numpy as np

DT_INT64, DT_VARCHAR, DT_JSON, DT_FLOAT_VECTOR = 5, 21, 23, 101
SUPPORTED_TYPES = (DT_INT64, DT_VARCHAR, DT_JSON, DT_FLOAT_VECTOR)
META_FIELD = "$meta"


# ---- protobuf primitives ---------------------------------------------------------------------
def _enc_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _tag(fno: int, wt: int) -> bytes:
    return _enc_varint((fno << 3) | wt)


def _f_varint(fno: int, v: int) -> bytes:
    return _tag(fno, 0) + _enc_varint(v)


def _f_bytes(fno: int, b: bytes) -> bytes:
    return _tag(fno, 2) + _enc_varint(len(b)) + b


def _dec_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes) -> Iterable[Tuple[int, int, Any]]:
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _dec_varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _dec_varint(buf, pos)
        elif wt == 2:
            ln, pos = _dec_varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 1:
            val, pos = buf[pos:pos + 8], pos + 8
        elif wt == 5:
            val, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"corrupt Milvus Lite blob: wire type {wt}")
        yield fno, wt, val


def _get(buf: bytes, fno: int, default=None):
    for f, _, v in _fields(buf):
        if f == fno:
            return v
    return default


def _s64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


# ---- schema / index meta ---------------------------------------------------------------------
def encode_schema(name: str, fields: List[Dict[str, Any]], enable_dynamic: bool) -> bytes:
    out = _f_bytes(1, name.encode())
    fid = 100
    for f in fields:
        msg = _f_varint(1, fid) + _f_bytes(2, f["name"].encode())
        if f.get("is_primary"):
            msg += _f_varint(3, 1)
        if f.get("description"):
            msg += _f_bytes(4, f["description"].encode())
        msg += _f_varint(5, int(f["dtype"]))
        if f.get("dim") is not None:
            msg += _f_bytes(6, _f_bytes(1, b"dim") + _f_bytes(2, str(int(f["dim"])).encode()))
        if f.get("max_length") is not None:
            msg += _f_bytes(6, _f_bytes(1, b"max_length") + _f_bytes(2, str(int(f["max_length"])).encode()))
        if f.get("auto_id"):
            msg += _f_varint(8, 1)
        out += _f_bytes(4, msg)
        fid += 1
    if enable_dynamic:
        out += _f_bytes(4, _f_varint(1, fid) + _f_bytes(2, META_FIELD.encode()) + _f_bytes(4, b"dynamic schema") +
                        _f_varint(5, DT_JSON) + _f_varint(12, 1))
    out += _f_bytes(4, _f_bytes(2, b"RowID") + _f_bytes(4, b"row id") + _f_varint(5, DT_INT64))
    out += _f_bytes(4, _f_varint(1, 1) + _f_bytes(2, b"Timestamp") + _f_bytes(4, b"time stamp") + _f_varint(5, DT_INT64))
    if enable_dynamic:
        out += _f_varint(5, 1)
    return out


def decode_schema(blob: bytes) -> Dict[str, Any]:
    info: Dict[str, Any] = {"name": "", "fields": [], "enable_dynamic_field": False}
    for fno, _, v in _fields(blob):
        if fno == 1:
            info["name"] = v.decode()
        elif fno == 5:
            info["enable_dynamic_field"] = bool(v)
        elif fno == 4:
            f: Dict[str, Any] = {"field_id": 0, "name": "", "is_primary": False, "description": "", "dtype": 0,
                                 "auto_id": False, "is_dynamic": False, "dim": None, "max_length": None}
            for g, _, w in _fields(v):
                if g == 1:
                    f["field_id"] = w
                elif g == 2:
                    f["name"] = w.decode()
                elif g == 3:
                    f["is_primary"] = bool(w)
                elif g == 4:
                    f["description"] = w.decode()
                elif g == 5:
                    f["dtype"] = w
                elif g == 6:
                    k, val = _get(w, 1, b"").decode(), _get(w, 2, b"").decode()
                    if k in ("dim", "max_length"):
                        f[k] = int(val)
                elif g == 8:
                    f["auto_id"] = bool(w)
                elif g == 12:
                    f["is_dynamic"] = bool(w)
            if f["name"] in ("RowID", "Timestamp"):
                continue
            info["fields"].append(f)
    return info


def encode_index(field_id: int, field_name: str, params: Dict[str, Any]) -> bytes:
    out = _f_varint(1, field_id) + _f_varint(2, int(time.time_ns()) & ((1 << 62) - 1)) + _f_bytes(3, field_name.encode())
    for k, v in params.items():
        out += _f_bytes(5, _f_bytes(1, str(k).encode()) + _f_bytes(2, str(v).encode()))
    return out + _f_varint(6, 1)


def decode_index(blob: bytes) -> Dict[str, str]:
    params: Dict[str, str] = {}
    for fno, _, v in _fields(blob):
        if fno == 5:
            params[_get(v, 1, b"").decode()] = _get(v, 2, b"").decode()
    return params


# ---- entities --------------------------------------------------------------------------------
def encode_entity(schema_fields: List[Dict[str, Any]], row: Dict[str, Any], dynamic: Optional[Dict[str, Any]],
                  tso: int) -> bytes:
    out = b""
    fid = 100
    for f in schema_fields:
        name, dt = f["name"], int(f["dtype"])
        head = _f_varint(1, dt) + _f_bytes(2, name.encode())
        val = row[name]
        if dt == DT_INT64:
            body = _f_bytes(3, _f_bytes(3, _f_bytes(1, _enc_varint(int(val)))))
        elif dt == DT_VARCHAR:
            body = _f_bytes(3, _f_bytes(6, _f_bytes(1, str(val).encode("utf-8"))))
        elif dt == DT_JSON:
            body = _f_bytes(3, _f_bytes(9, _f_bytes(1, json.dumps(val, ensure_ascii=False).encode("utf-8"))))
        elif dt == DT_FLOAT_VECTOR:
            vec = np.ascontiguousarray(val, dtype="<f4")
            body = _f_bytes(4, _f_varint(1, vec.shape[0]) + _f_bytes(2, _f_bytes(1, vec.tobytes())))
        else:
            raise ValueError(f"field {name}: data type {dt} is not supported by this store")
        out += _f_bytes(1, head + body + _f_varint(5, fid))
        fid += 1
    if dynamic is not None:
        raw = json.dumps(dynamic, ensure_ascii=False, separators=(",", ":")).encode("utf-8")
        out += _f_bytes(1, _f_varint(1, DT_JSON) + _f_bytes(2, META_FIELD.encode()) +
                        _f_bytes(3, _f_bytes(9, _f_bytes(1, raw))) + _f_varint(5, fid))
    sysval = _f_bytes(3, _f_bytes(3, _f_bytes(1, _enc_varint(tso))))
    out += _f_bytes(1, _f_varint(1, DT_INT64) + _f_bytes(2, b"RowID") + sysval)
    out += _f_bytes(1, _f_varint(1, DT_INT64) + _f_bytes(2, b"Timestamp") + sysval + _f_varint(5, 1))
    return out + _f_varint(2, 1)


def decode_entity(blob: bytes) -> Dict[str, Any]:
    row: Dict[str, Any] = {}
    for fno, _, fd in _fields(blob):
        if fno != 1:
            continue
        dt, name, scalars, vectors = 0, "", None, None
        for g, _, w in _fields(fd):
            if g == 1:
                dt = w
            elif g == 2:
                name = w.decode()
            elif g == 3:
                scalars = w
            elif g == 4:
                vectors = w
        if name in ("RowID", "Timestamp"):
            continue
        if dt == DT_INT64:
            packed = _get(_get(scalars or b"", 3, b""), 1, b"")
            row[name] = _s64(_dec_varint(packed, 0)[0]) if packed else None
        elif dt == DT_VARCHAR:
            raw = _get(_get(scalars or b"", 6, b""), 1, b"")
            row[name] = raw.decode("utf-8")
        elif dt == DT_JSON:
            raw = _get(_get(scalars or b"", 9, b""), 1, b"")
            row[name] = json.loads(raw.decode("utf-8")) if raw else {}
        elif dt == DT_FLOAT_VECTOR:
            raw = _get(_get(vectors or b"", 2, b""), 1, b"")
            row[name] = np.frombuffer(raw, dtype="<f4")
    return row


# ---- file level ------------------------------------------------------------------------------
class MilvusLiteFile:
    def __init__(self, path: str):
        self.path = path
        new = not os.path.exists(path)
        d = os.path.dirname(os.path.abspath(path))
        os.makedirs(d, exist_ok=True)
        self.con = sqlite3.connect(path)
        self.con.text_factory = bytes  # `data` may carry TEXT affinity in files written by Milvus Lite
        if new or not self._has_table("collection_meta"):
            self.con.execute("CREATE TABLE IF NOT EXISTS collection_meta (id INTEGER PRIMARY KEY, "
                             "collection_name VARCHAR(1024), meta_type VARCHAR(1024), blob_field BLOB, "
                             "string_field VARCHAR(1024))")
            self.con.commit()

    def _has_table(self, name: str) -> bool:
        r = self.con.execute("select count(*) from sqlite_master where type='table' and name=?", (name,)).fetchone()
        return bool(r and r[0])

    @staticmethod
    def _txt(v) -> str:
        return v.decode("utf-8") if isinstance(v, (bytes, bytearray)) else str(v)

    def close(self):
        try:
            self.con.close()
        except Exception:
            pass

    def list_collections(self) -> List[str]:
        rows = self.con.execute("select distinct collection_name from collection_meta where meta_type='schema'").fetchall()
        return [self._txt(r[0]) for r in rows]

    def read_meta(self, name: str) -> Tuple[Dict[str, Any], Dict[str, str]]:
        schema, index = None, {}
        for mt, blob in self.con.execute("select meta_type, blob_field from collection_meta where collection_name=?", (name,)):
            if self._txt(mt) == "schema":
                schema = decode_schema(bytes(blob))
            elif self._txt(mt) == "index":
                index = decode_index(bytes(blob))
        if schema is None:
            raise KeyError(name)
        return schema, index

    def create_collection(self, name: str, fields: List[Dict[str, Any]], enable_dynamic: bool):
        pk = next(f["name"] for f in fields if f.get("is_primary"))
        self.con.execute("insert into collection_meta (collection_name, meta_type, blob_field, string_field) values (?,?,?,?)",
                         (name, "schema", sqlite3.Binary(encode_schema(name, fields, enable_dynamic)), pk))
        self.con.execute(f'CREATE TABLE IF NOT EXISTS "{name}" (id INTEGER PRIMARY KEY, milvus_id VARCHAR(1024), data BLOB)')
        self.con.commit()

    def write_index(self, name: str, field_id: int, field_name: str, params: Dict[str, Any]):
        self.con.execute("delete from collection_meta where collection_name=? and meta_type='index'", (name,))
        self.con.execute("insert into collection_meta (collection_name, meta_type, blob_field, string_field) values (?,?,?,?)",
                         (name, "index", sqlite3.Binary(encode_index(field_id, field_name, params)), field_name))
        self.con.commit()

    def drop_collection(self, name: str):
        self.con.execute("delete from collection_meta where collection_name=?", (name,))
        self.con.execute(f'DROP TABLE IF EXISTS "{name}"')
        self.con.commit()

    def append(self, name: str, fields: List[Dict[str, Any]], pk_name: str, rows: List[Dict[str, Any]],
               dynamics: List[Optional[Dict[str, Any]]]):
        tso = (int(time.time() * 1000) << 18)
        payload = [(str(r[pk_name]), sqlite3.Binary(encode_entity(fields, r, d, tso))) for r, d in zip(rows, dynamics)]
        self.con.executemany(f'insert into "{name}" (milvus_id, data) values (?,?)', payload)
        self.con.commit()

    def load_rows(self, name: str) -> List[Dict[str, Any]]:
        if not self._has_table(name):
            return []
        out = []
        for (blob,) in self.con.execute(f'select data from "{name}" order by id'):
            out.append(decode_entity(bytes(blob)))
        return out
