"""Sharded raw Truth-Vault format.

The reference's vault is ONE pickle holding an (N,512) ndarray plus Python lists
(train_clip_detective.py:515-573 writes it, misinfo_forensics.py:214-246 reads it); at 10 M
rows that is a 20 GB `pickle.load` on every rank.  This module writes the same information
as a directory each rank can open lazily:

    manifest.json            {"format": "mmf-vault-1", "n_rows", "dim", "dtype", "shards": [{"file","row0","rows"}],
                              "metadata_file", "source"}
    shard-00000.npy ...      raw row blocks (np.load(mmap_mode="r") -> no copy until the rows are uploaded)
    metadata.jsonl           one {"title","url","date"} object per row (host side only; strings never go to the GPU)

`open_vault_dir(dir, rank, world)` maps just the rows of that rank's ShardPlan slice, so a
row-sharded deployment never materialises the whole vault on any host.  The legacy pickle
layouts stay readable through vault.read_vault_dict / `convert_pickle`.
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .vault import ShardPlan, read_vault_dict

FORMAT = "mmf-vault-1"


def save_vault_dir(path: str, embeddings: np.ndarray, metadata: Optional[Sequence[dict]] = None,
                   rows_per_shard: int = 1 << 20, dtype=None, source: str = "") -> dict:
    """Write `embeddings` ((N,dim), any float dtype) + metadata as a sharded directory."""
    emb = np.asarray(embeddings)
    if emb.ndim != 2:
        raise ValueError("embeddings must be 2-D")
    if metadata is not None and len(metadata) != emb.shape[0]:
        raise ValueError(f"metadata has {len(metadata)} records for {emb.shape[0]} rows")
    dtype = np.dtype(dtype or emb.dtype)
    os.makedirs(path, exist_ok=True)
    shards = []
    for i, row0 in enumerate(range(0, emb.shape[0], rows_per_shard)):
        block = np.ascontiguousarray(emb[row0:row0 + rows_per_shard], dtype=dtype)
        name = f"shard-{i:05d}.npy"
        np.save(os.path.join(path, name), block)
        shards.append({"file": name, "row0": row0, "rows": int(block.shape[0])})
    manifest = {"format": FORMAT, "n_rows": int(emb.shape[0]), "dim": int(emb.shape[1]), "dtype": dtype.name,
                "shards": shards, "metadata_file": None, "source": source}
    if metadata is not None:
        manifest["metadata_file"] = "metadata.jsonl"
        with open(os.path.join(path, "metadata.jsonl"), "w", encoding="utf-8") as fh:
            for m in metadata:
                fh.write(json.dumps({"title": m["title"], "url": m.get("url", "N/A"), "date": m.get("date", "N/A")},
                                    ensure_ascii=False) + "\n")
    with open(os.path.join(path, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    return manifest


def convert_pickle(pickle_path: str, out_dir: str, rows_per_shard: int = 1 << 20) -> dict:
    """Legacy guardian_embeddings.pkl (either layout) -> sharded directory."""
    with open(pickle_path, "rb") as fh:
        data = pickle.load(fh)
    emb, meta = read_vault_dict(data)
    if emb is None:
        raise ValueError(f"{pickle_path}: unknown vault pickle layout")
    return save_vault_dir(out_dir, emb, meta, rows_per_shard, source=os.path.basename(pickle_path))


def read_manifest(path: str) -> dict:
    with open(os.path.join(path, "manifest.json")) as fh:
        m = json.load(fh)
    if m.get("format") != FORMAT:
        raise ValueError(f"{path}: not a {FORMAT} directory")
    return m


def open_vault_dir(path: str, rank: int = 0, world: int = 1) -> Tuple[np.ndarray, int, int]:
    """Rows of this rank's ShardPlan slice as one array (memory-mapped blocks, concatenated only if the
    slice spans several files).  Returns (rows, row_offset, n_total)."""
    m = read_manifest(path)
    lo, hi = ShardPlan(m["n_rows"], world).bounds(rank)
    parts: List[np.ndarray] = []
    for sh in m["shards"]:
        a, b = max(lo, sh["row0"]), min(hi, sh["row0"] + sh["rows"])
        if a < b:
            block = np.load(os.path.join(path, sh["file"]), mmap_mode="r")
            parts.append(block[a - sh["row0"]:b - sh["row0"]])
    if not parts:
        rows = np.empty((0, m["dim"]), dtype=np.dtype(m["dtype"]))
    else:
        rows = parts[0] if len(parts) == 1 else np.concatenate(parts)
    return rows, lo, m["n_rows"]


def read_metadata(path: str, rows: Optional[Iterable[int]] = None) -> List[dict]:
    """All metadata records, or only those of the given global row ids (one pass over the file)."""
    m = read_manifest(path)
    if not m["metadata_file"]:
        return []
    want = None if rows is None else set(int(r) for r in rows)
    out = {} if want is not None else []
    with open(os.path.join(path, m["metadata_file"]), encoding="utf-8") as fh:
        for i, line in enumerate(fh):
            if want is None:
                out.append(json.loads(line))
            elif i in want:
                out[i] = json.loads(line)
    return out if want is None else [out[int(r)] for r in rows]


def generate_embeddings_database(model_path: str = "clip_detective_best.pth", json_file: str = "vector_db_seed.json",
                                 output_file: str = "guardian_embeddings.pkl", *, clip_model=None, processor=None,
                                 articles: Optional[Sequence[dict]] = None, batch_size: int = 64, device=None,
                                 val_accuracy=None, vault_dir: Optional[str] = None, rows_per_shard: int = 1 << 20) -> dict:
    """Batched form of the reference's vault writer (train_clip_detective.py:457-607).

    The reference encodes ONE article per CLIP forward; this encodes `batch_size` at a time (the encoders stay PyTorch
    producers) and writes the same database: keys article_ids / text_contents / image_paths / image_embeddings /
    text_embeddings / metadata{model_path,total_articles,embedding_dim,val_accuracy}, rows L2-normalised (:556-557),
    the pickle at `output_file` plus `<output>_summary.json`; with `vault_dir` also the sharded raw directory of this
    module.  Articles whose image cannot be opened are reported and skipped, as in the reference (:589-591).
    clip_model: anything callable as model(input_ids=..., pixel_values=..., ...) -> .image_embeds / .text_embeds
    (a CLIPModel, or the reference's CLIPDetective.clip); by default loaded from the reference's paths."""
    import torch
    from PIL import Image
    print("\n" + "=" * 60 + "\nGenerating Embeddings Database\n" + "=" * 60)
    if clip_model is None or processor is None:
        from transformers import CLIPModel, CLIPProcessor
        clip_dir = r"C:\Users\Lenovo\OneDrive\Desktop\hack\models\clip-vit-b32"
        processor = processor or CLIPProcessor.from_pretrained(clip_dir)
        if clip_model is None:
            clip_model = CLIPModel.from_pretrained(clip_dir)
            ckpt = torch.load(model_path, map_location="cpu", weights_only=False)
            clip_model.load_state_dict({k[len("clip."):]: v for k, v in ckpt["model_state_dict"].items() if k.startswith("clip.")},
                                       strict=False)
            val_accuracy = ckpt.get("val_accuracy", val_accuracy)
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    if hasattr(clip_model, "to"):
        clip_model = clip_model.to(device)
    if hasattr(clip_model, "eval"):
        clip_model.eval()
    if articles is None:
        with open(json_file, "r", encoding="utf-8") as fh:
            articles = json.load(fh)
    print(f"Found {len(articles)} articles")
    db = {"article_ids": [], "text_contents": [], "image_paths": [], "image_embeddings": [], "text_embeddings": [],
          "metadata": {"model_path": model_path, "total_articles": len(articles), "embedding_dim": None, "val_accuracy": val_accuracy}}
    with torch.no_grad():
        for b0 in range(0, len(articles), batch_size):
            batch, images = [], []
            for article in articles[b0:b0 + batch_size]:
                try:
                    images.append(Image.open(article["image_local_path"]).convert("RGB"))
                    batch.append(article)
                except Exception as e:                       # the reference skips the article and goes on
                    print(f"\nError processing {article['article_id']}: {e}")
            if not batch:
                continue
            inputs = processor(text=[a["text_content"] for a in batch], images=images, return_tensors="pt", padding=True,
                               truncation=True, max_length=77)
            inputs = {k: v.to(device) for k, v in inputs.items()}
            out = clip_model(**inputs, return_dict=True)
            img = out.image_embeds.float().cpu().numpy()
            txt = out.text_embeds.float().cpu().numpy()
            img = img / np.linalg.norm(img, axis=1, keepdims=True)       # :556-557, row-wise
            txt = txt / np.linalg.norm(txt, axis=1, keepdims=True)
            db["article_ids"] += [a["article_id"] for a in batch]
            db["text_contents"] += [a["text_content"] for a in batch]
            db["image_paths"] += [a["image_local_path"] for a in batch]
            db["image_embeddings"].append(img)
            db["text_embeddings"].append(txt)
    dim = 512
    db["image_embeddings"] = np.concatenate(db["image_embeddings"]) if db["image_embeddings"] else np.zeros((0, dim), np.float32)
    db["text_embeddings"] = np.concatenate(db["text_embeddings"]) if db["text_embeddings"] else np.zeros((0, dim), np.float32)
    db["metadata"]["embedding_dim"] = int(db["image_embeddings"].shape[1])
    print(f"\n✓ Generated embeddings for {len(db['article_ids'])} articles")
    with open(output_file, "wb") as fh:
        pickle.dump(db, fh)
    size_mb = os.path.getsize(output_file) / 1e6
    summary = {"total_articles": len(db["article_ids"]), "embedding_dimension": db["metadata"]["embedding_dim"],
               "model_val_accuracy": db["metadata"]["val_accuracy"], "database_size_mb": size_mb,
               "sample_articles": db["article_ids"][:5]}
    with open(output_file.replace(".pkl", "_summary.json"), "w", encoding="utf-8") as fh:
        json.dump(summary, fh, indent=2)
    if vault_dir is not None:
        _, meta = read_vault_dict(db)
        save_vault_dir(vault_dir, db["image_embeddings"], meta, rows_per_shard=rows_per_shard, source=output_file)
    print(f"✓ Saved embeddings database ({size_mb:.2f} MB)")
    return db
