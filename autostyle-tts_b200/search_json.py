"""Batch front end shaped like the reference's production query script
(`/root/reference/milvus/search_json.py:313-465`): read dialogue lines, search the style database,
write the JSONL that `tts_with_rag.py` consumes (`/root/reference/tts_with_rag.py:77-96`):

    {"zh_text", "speaker", "retrieved_file_id", "retrieved_text", "distance", "whisper"}

The reference computes every line's 6144-d query embedding with its fine-tuned LLM inside the loop
(`search_json.py:382-411`) and issues one batch-1 search per line.  The LLM is out of scope here
(SURVEY.md section 2, row 5), so this front end takes the embeddings **precomputed** in the input JSONL
(field `embedding` / `vector`) and batches all lines into ONE search call.

    python -m autostyle_tts_b200.search_json --input_json lines.jsonl --output_json search_results.json \
        --db_path milvus_demo.db --prefix /data/style_db/

Error behaviour mirrors the reference (`search_json.py:431-449`): a line without a usable embedding or
without hits is written with retrieved_file_id "N/A" and distance null instead of aborting the run.
"""
from __future__ import annotations

import argparse
import json
from typing import Any, Dict, Iterable, List, Optional

import numpy as np

from .client import MilvusClient

DEFAULT_COLLECTION = "embeddings_biographies_collection"   # search_json.py:340, RAG.py:457


def read_input_json(path: str) -> List[Dict[str, Any]]:
    """JSON array or JSON-lines, like the reference's reader (`search_json.py:264-310`)."""
    with open(path, "r", encoding="utf-8") as f:
        text = f.read().strip()
    if not text:
        return []
    if text[0] == "[":
        return list(json.loads(text))
    return [json.loads(line) for line in text.splitlines() if line.strip()]


def retrieve_styles(client: MilvusClient, rows: Iterable[Dict[str, Any]], collection_name: str = DEFAULT_COLLECTION,
                    prefix: str = "", top_k: int = 1, embedding_field: Optional[str] = None) -> List[Dict[str, Any]]:
    rows = list(rows)
    dim = client.describe_collection(collection_name)["fields"]
    dim = next(f["params"]["dim"] for f in dim if "dim" in f.get("params", {}))
    vecs, where = [], []
    for i, r in enumerate(rows):
        v = r.get(embedding_field) if embedding_field else (r.get("embedding", r.get("vector")))
        try:
            arr = np.asarray(v, dtype=np.float32).reshape(-1)
        except (TypeError, ValueError):
            continue
        if arr.shape[0] == dim and np.all(np.isfinite(arr)):
            vecs.append(arr)
            where.append(i)
    hits_per_row: Dict[int, List[Dict[str, Any]]] = {}
    if vecs:
        res = client.search(collection_name=collection_name, data=np.stack(vecs), limit=top_k, filter=None,
                            output_fields=["file_id", "text"])
        hits_per_row = {i: h for i, h in zip(where, res)}
    out = []
    for i, r in enumerate(rows):
        rec = {"zh_text": r.get("zh_text", r.get("text", "")), "speaker": r.get("speaker", ""),
               "whisper": r.get("whisper", "")}
        hits = hits_per_row.get(i) or []
        if hits:
            top = hits[0]
            rec["retrieved_file_id"] = f"{prefix}{top['entity'].get('file_id', '')}"
            rec["retrieved_text"] = top["entity"].get("text", "")
            rec["distance"] = top["distance"]
            if top_k > 1:
                rec["alternatives"] = [{"retrieved_file_id": f"{prefix}{h['entity'].get('file_id', '')}",
                                        "distance": h["distance"]} for h in hits[1:]]
        else:
            rec["retrieved_file_id"] = "N/A"
            rec["retrieved_text"] = "N/A"
            rec["distance"] = None
        out.append(rec)
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description="Retrieve style samples for dialogue lines (precomputed embeddings)")
    ap.add_argument("--input_json", required=True)
    ap.add_argument("--output_json", required=True)
    ap.add_argument("--db_path", default="milvus_demo.db")
    ap.add_argument("--collection", default=DEFAULT_COLLECTION)
    ap.add_argument("--prefix", default="")
    ap.add_argument("--top_k", type=int, default=1)
    ap.add_argument("--embedding_field", default=None)
    a = ap.parse_args(argv)
    client = MilvusClient(a.db_path)
    if not client.has_collection(collection_name=a.collection):
        raise SystemExit(f"Collection '{a.collection}' does not exist in '{a.db_path}'.")
    results = retrieve_styles(client, read_input_json(a.input_json), a.collection, a.prefix, a.top_k, a.embedding_field)
    with open(a.output_json, "w", encoding="utf-8") as f:
        for rec in results:
            f.write(json.dumps(rec, ensure_ascii=False) + "\n")
    print(f"Saved {len(results)} search results to '{a.output_json}'.")
    client.close()


if __name__ == "__main__":
    main()
