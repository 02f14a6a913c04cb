"""B200-native exact (FLAT) vector search behind the `pymilvus.MilvusClient` surface that
AutoStyle-TTS uses (`from pymilvus import MilvusClient, FieldSchema, CollectionSchema, DataType`).

The directory name carries a hyphen, so import it through the `autostyle_tts_b200` alias module at
the repository root (or `importlib.import_module("autostyle-tts_b200")`).
"""
from .client import MAX_LIMIT, MilvusClient
from .engine import ABI_SYMBOLS, LIB_PATH, AvsError, ExactnessWarning, Store, load_library
from .schema import CollectionSchema, DataType, FieldSchema, IndexParams, MilvusException

__all__ = ["MilvusClient", "FieldSchema", "CollectionSchema", "DataType", "IndexParams", "MilvusException",
           "Store", "AvsError", "ExactnessWarning", "load_library", "ABI_SYMBOLS", "LIB_PATH", "MAX_LIMIT"]
