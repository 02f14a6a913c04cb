"""Schema helper types the reference imports from pymilvus
(`from pymilvus import FieldSchema, CollectionSchema, DataType`,
/root/reference/milvus/insert_embeddings.py:13, /root/reference/milvus/RAG.py:30)
and builds at /root/reference/milvus/insert_embeddings.py:52-60.  Same names,
argument meaning and enum numbers as pymilvus; unknown keyword arguments are
tolerated because the reference passes some that upstream does not define
(`CollectionSchema(..., metric_type="COSINE")`, insert_embeddings.py:60).
"""
from __future__ import annotations

import enum
from typing import Any, Dict, List, Optional


class MilvusException(Exception):
    """Same role as pymilvus.MilvusException: every reference call site catches
    `Exception`, prints and returns [] (/root/reference/milvus/search_embeddings.py:24-27)."""

    def __init__(self, message: str = "", code: int = 1):
        super().__init__(message)
        self.code = code
        self.message = message

    def __str__(self):
        return f"<MilvusException: (code={self.code}, message={self.message})>"


class DataType(enum.IntEnum):
    NONE = 0
    BOOL = 1
    INT8 = 2
    INT16 = 3
    INT32 = 4
    INT64 = 5
    FLOAT = 10
    DOUBLE = 11
    STRING = 20
    VARCHAR = 21
    ARRAY = 22
    JSON = 23
    BINARY_VECTOR = 100
    FLOAT_VECTOR = 101
    FLOAT16_VECTOR = 102
    BFLOAT16_VECTOR = 103
    SPARSE_FLOAT_VECTOR = 104
    UNKNOWN = 999


class FieldSchema:
    def __init__(self, name: str, dtype: DataType, description: str = "", **kwargs):
        self.name = name
        self.dtype = DataType(dtype)
        self.description = description
        self.is_primary = bool(kwargs.pop("is_primary", False))
        self.auto_id = bool(kwargs.pop("auto_id", False))
        self.max_length = kwargs.pop("max_length", None)
        self.dim = kwargs.pop("dim", None)
        self.is_dynamic = bool(kwargs.pop("is_dynamic", False))
        self.params = dict(kwargs)
        if self.dtype == DataType.FLOAT_VECTOR and self.dim is not None:
            self.dim = int(self.dim)
            if self.dim <= 0:
                raise MilvusException(f"invalid dimension {self.dim} for field {name}")

    def to_dict(self) -> Dict[str, Any]:
        d: Dict[str, Any] = {"name": self.name, "type": self.dtype, "description": self.description}
        params = {}
        if self.dim is not None:
            params["dim"] = self.dim
        if self.max_length is not None:
            params["max_length"] = self.max_length
        if params:
            d["params"] = params
        if self.is_primary:
            d["is_primary"] = True
            d["auto_id"] = self.auto_id
        return d

    def __repr__(self):
        return f"FieldSchema({self.to_dict()})"


class CollectionSchema:
    def __init__(self, fields: Optional[List[FieldSchema]] = None, description: str = "", **kwargs):
        self.fields: List[FieldSchema] = list(fields or [])
        self.description = description
        self.enable_dynamic_field = bool(kwargs.pop("enable_dynamic_field", False))
        self.auto_id = bool(kwargs.pop("auto_id", False))
        # not an upstream argument, but the reference passes it (insert_embeddings.py:60)
        self.metric_type = kwargs.pop("metric_type", None)
        self.extra = dict(kwargs)

    def add_field(self, field_name: str, datatype: DataType, **kwargs) -> "CollectionSchema":
        self.fields.append(FieldSchema(field_name, datatype, **kwargs))
        return self

    @property
    def primary_field(self) -> Optional[FieldSchema]:
        for f in self.fields:
            if f.is_primary:
                return f
        return None

    @property
    def vector_field(self) -> Optional[FieldSchema]:
        for f in self.fields:
            if f.dtype == DataType.FLOAT_VECTOR:
                return f
        return None

    def verify(self):
        pk, vec = self.primary_field, self.vector_field
        if pk is None:
            raise MilvusException("Schema must have a primary key field.")
        if pk.dtype not in (DataType.INT64, DataType.VARCHAR):
            raise MilvusException("Primary key type must be DataType.INT64 or DataType.VARCHAR.")
        if vec is None or not vec.dim:
            raise MilvusException("Schema must have a FLOAT_VECTOR field with a dim.")
        if sum(1 for f in self.fields if f.dtype == DataType.FLOAT_VECTOR) != 1:
            raise MilvusException("exactly one FLOAT_VECTOR field is supported")

    def to_dict(self) -> Dict[str, Any]:
        return {"auto_id": self.auto_id or bool(self.primary_field and self.primary_field.auto_id),
                "description": self.description, "fields": [f.to_dict() for f in self.fields],
                "enable_dynamic_field": self.enable_dynamic_field}

    def __repr__(self):
        return f"CollectionSchema({self.to_dict()})"


class IndexParams(list):
    """`MilvusClient.prepare_index_params()` result: a list of index descriptions."""

    def add_index(self, field_name: str, index_type: str = "", index_name: str = "", **kwargs):
        entry = {"field_name": field_name, "index_type": index_type or "AUTOINDEX", "index_name": index_name}
        entry.update(kwargs)
        self.append(entry)
        return self
