"""Regenerates the golden fixtures in this directory from the reference's shipped
artefacts.  Runs only where `/root/reference` exists (the authoring container);
the fixtures it writes are committed so the GPU box never needs the reference.

  python tests/golden/make_golden.py

Inputs  (reference, read-only):
  /root/reference/milvus/milvus_demo.db          the 130 x 6144 style database (F1)
  /root/reference/output_emb/search_results.json 64 rows of real search output (F2)
Outputs (committed):
  f1_vectors_fp16.npy     130 x 6144 vectors; every value is fp16-exact (asserted), so
                          fp16 storage is lossless and halves the fixture
  f1_rows.json            pk + {file_id, text} per row, collection + index meta
  f1_kat.npz              KAT-1/KAT-2: top-5 of the 130 self-queries (rows, ids, fp32 dist),
                          C1 perturbed queries (64 x 6144) and their top-5
  f2_search_results_schema.json   key set + distance range + retrieved basenames of F2
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import flat_search, milvus_db  # noqa: E402

REF = "/root/reference"
COLL = "embeddings_biographies_collection"


def main():
    db = os.path.join(REF, "milvus", "milvus_demo.db")
    pks, X, meta = milvus_db.load_collection(db, COLL)
    assert X.shape == (130, 6144), X.shape
    X16 = X.astype(np.float16)
    assert np.array_equal(X16.astype(np.float32), X), "vectors are not fp16-exact"
    np.save(os.path.join(HERE, "f1_vectors_fp16.npy"), X16)
    with open(os.path.join(HERE, "f1_rows.json"), "w", encoding="utf-8") as f:
        json.dump({"collection": COLL, "dim": 6144, "metric_type": "COSINE",
                   "collections_in_db": sorted(milvus_db.list_collections(db)),
                   "pks": pks.tolist(), "meta": meta}, f, ensure_ascii=False, indent=0)

    # KAT-1 / KAT-2: self-queries, three arithmetic variants must agree on the top-5 rows.
    ids_row = np.arange(130, dtype=np.int64)        # unique ids = row index (pk has duplicates)
    i64, d64, r64 = flat_search.search(X, ids_row, X, 5, "COSINE")
    _, _, r32a = flat_search.search(X, ids_row, X, 5, "COSINE", accum="f32", variant="normalize_then_dot")
    _, _, r32b = flat_search.search(X, ids_row, X, 5, "COSINE", accum="f32", variant="dot_then_divide")
    assert np.array_equal(r64, r32a) and np.array_equal(r64, r32b), "KAT-2 unstable"
    assert np.array_equal(r64[:, 0], np.arange(130)), "KAT-1: self not top-1"
    assert np.all(np.abs(d64[:, 0] - 1.0) <= 1e-6)
    # with the real (duplicated) primary keys as tie-break key and as returned id
    ipk, dpk, rpk = flat_search.search(X, pks, X, 5, "COSINE")
    assert np.array_equal(rpk, r64)                 # no exact score ties -> same rows

    # C1 perturbed queries: x_i + 0.05*||x_i||*g/sqrt(D), seed 1234
    rng = np.random.default_rng(1234)
    pick = rng.integers(0, 130, size=64)
    g = rng.standard_normal((64, 6144)).astype(np.float32)
    nrm = np.linalg.norm(X[pick].astype(np.float64), axis=1).astype(np.float32)
    Qp = (X[pick] + 0.05 * nrm[:, None] * g / np.sqrt(np.float32(6144))).astype(np.float32)
    ip, dp, rp = flat_search.search(X, ids_row, Qp, 5, "COSINE")
    _, _, rp32 = flat_search.search(X, ids_row, Qp, 5, "COSINE", accum="f32")
    assert np.array_equal(rp, rp32)
    assert np.array_equal(rp[:, 0], pick)
    # IP metric on the raw (un-normalised) vectors
    _, dip, rip = flat_search.search(X, ids_row, Qp, 5, "IP")
    np.savez_compressed(os.path.join(HERE, "f1_kat.npz"),
                        self_rows=r64, self_dist=d64, self_pk_ids=ipk,
                        pert_pick=pick, pert_queries=Qp, pert_rows=rp, pert_dist=dp,
                        pert_ip_rows=rip, pert_ip_dist=dip)

    rows = [json.loads(l) for l in open(os.path.join(REF, "output_emb", "search_results.json"),
                                        encoding="utf-8") if l.strip()]
    dist = [r["distance"] for r in rows if isinstance(r.get("distance"), (int, float))]
    base = sorted({os.path.basename(str(r["retrieved_file_id"])) for r in rows})
    db_base = {os.path.basename(m["file_id"]) for m in meta}
    with open(os.path.join(HERE, "f2_search_results_schema.json"), "w", encoding="utf-8") as f:
        json.dump({"n_rows": len(rows), "keys": sorted(rows[0].keys()),
                   "distance_min": min(dist), "distance_max": max(dist),
                   "retrieved_basenames": base,
                   "all_retrieved_in_f1": all(b in db_base for b in base)}, f, ensure_ascii=False, indent=1)
    off = flat_search.scores64(X, X[0], "COSINE")
    print("F1", X.shape, "norms", float(np.linalg.norm(X, axis=1).min()), float(np.linalg.norm(X, axis=1).max()),
          "distinct pks", len(set(pks.tolist())), "F2 rows", len(rows), "dist", min(dist), max(dist),
          "all in F1:", all(b in db_base for b in base), "row0 2nd best", float(np.sort(off)[-2]))


if __name__ == "__main__":
    main()
