"""TEST INFRASTRUCTURE ONLY — CPU oracle, never imported by the product path.

Decoder for the Milvus Lite SQLite file the reference ships and re-opens
(`/root/reference/milvus/milvus_demo.db`, opened by `MilvusClient(db_path)` at
`/root/reference/milvus/search_embeddings.py:31`, written by
`/root/reference/milvus/RAG.py:538-548`).

Milvus Lite itself is a third-party dependency that is absent from
`/root/reference` (no requirements file; only the hint
`# pip install -U pymilvus ...` at `/root/reference/milvus/RAG.py:2`), so the
wire format below was recovered from the shipped file:

  collection_meta(id, collection_name, meta_type in {schema,index}, blob_field, string_field)
  "<collection>"(id INTEGER PK, milvus_id VARCHAR, data BLOB)
  data = protobuf { repeated FieldData fields = 1; uint32 num_rows = 2 }
  FieldData { type=1; field_name=2; scalars=3; vectors=4; field_id=5 }
    Int64     -> scalars.long_data(3).data(1)   packed varints
    FloatVec  -> vectors.dim(1), vectors.float_vector(2).data(1) packed LE fp32
    JSON      -> scalars.json_data(9).data(1)   UTF-8 bytes
    VarChar   -> scalars.string_data(6).data(1) UTF-8 bytes

This is an independent restatement kept separate from the product-side reader
(`autostyle-tts_b200/milvus_lite_db.py`) so that the two can be tested against
each other.
"""
from __future__ import annotations

import json
import sqlite3

import numpy as np

DT_INT64, DT_VARCHAR, DT_JSON, DT_FLOAT_VECTOR = 5, 21, 23, 101


def _varint(buf: bytes, pos: int):
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def walk(buf: bytes):
    """Yield (field_number, wire_type, value) for one protobuf message level."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            val, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, val


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _first(buf: bytes, fno: int):
    for f, _, v in walk(buf):
        if f == fno:
            return v
    return None


def decode_entity(blob: bytes) -> dict:
    """One row blob -> {field_name: python value}; `$meta` JSON is merged in flat."""
    row: dict = {}
    for fno, _, fd in walk(blob):
        if fno != 1:
            continue
        ftype, fname, scalars, vectors = 0, "", None, None
        for f, _, v in walk(fd):
            if f == 1:
                ftype = v
            elif f == 2:
                fname = v.decode("utf-8")
            elif f == 3:
                scalars = v
            elif f == 4:
                vectors = v
        if ftype == DT_INT64 and scalars is not None:
            packed = _first(_first(scalars, 3) or b"", 1) or b""
            pos, vals = 0, []
            while pos < len(packed):
                v, pos = _varint(packed, pos)
                vals.append(_signed64(v))
            row[fname] = vals[0] if vals else None
        elif ftype == DT_FLOAT_VECTOR and vectors is not None:
            dim = _first(vectors, 1)
            raw = _first(_first(vectors, 2), 1)
            vec = np.frombuffer(raw, dtype="<f4")
            if vec.shape[0] != dim:
                raise ValueError(f"vector length {vec.shape[0]} != dim {dim}")
            row[fname] = vec
        elif ftype == DT_JSON and scalars is not None:
            raw = _first(_first(scalars, 9), 1)
            row[fname] = json.loads(raw.decode("utf-8")) if raw else {}
        elif ftype == DT_VARCHAR and scalars is not None:
            raw = _first(_first(scalars, 6), 1)
            row[fname] = raw.decode("utf-8") if raw is not None else ""
    return row


def list_collections(path: str) -> list:
    con = sqlite3.connect(f"file:{path}?mode=ro", uri=True)
    try:
        rows = con.execute("select distinct collection_name from collection_meta").fetchall()
    finally:
        con.close()
    return [r[0] for r in rows]


def load_collection(path: str, name: str):
    """Returns (pks int64[N], vectors fp32[N,D], meta list[dict]) in SQLite rowid order."""
    con = sqlite3.connect(f"file:{path}?mode=ro", uri=True)
    con.text_factory = bytes  # `data` has TEXT affinity: length() stops at NUL otherwise
    try:
        rows = con.execute(f'select id, milvus_id, data from "{name}" order by id').fetchall()
    finally:
        con.close()
    pks, vecs, meta = [], [], []
    for _, milvus_id, blob in rows:
        ent = decode_entity(bytes(blob))
        pk = ent.get("id")
        if pk is None:
            pk = int(milvus_id)
        pks.append(pk)
        vecs.append(ent["vector"])
        m = dict(ent.get("$meta") or {})
        for k, v in ent.items():
            if k not in ("id", "vector", "$meta", "RowID", "Timestamp"):
                m[k] = v
        meta.append(m)
    d = vecs[0].shape[0] if vecs else 0
    return (np.asarray(pks, dtype=np.int64),
            np.stack(vecs).astype(np.float32) if vecs else np.zeros((0, d), np.float32),
            meta)
