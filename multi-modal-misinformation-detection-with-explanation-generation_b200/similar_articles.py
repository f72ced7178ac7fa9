"""Drop-in for the reference's search_similar_articles (train_clip_detective.py:610-688):
text- or image-query search of the Guardian embedding database, the similarity + top-k on the B200 through the
same Truth-Vault kernels (mmf_vault_load / mmf_vault_search) instead of np.dot + a full argsort.

Difference to the reference, documented: the database rows are re-normalised in fp32 at upload (the vault kernels
search unit rows).  The writer already normalised them (:556-557), so for fp32 databases this changes nothing beyond
rounding (~1e-7); an fp16 database (built under CUDA autocast) stores rows whose norm is 1 +- 5e-4, and the
reference's similarities carry that factor -- ours do not (tested at the 1e-2 band like the fp16 vault)."""
from __future__ import annotations

import pickle
from typing import List, Optional

import numpy as np
import torch
from PIL import Image

from .engine import Engine

# The article database lives in its OWN library handle, one per (database object, embedding kind, device): loading it
# into a caller's Engine would silently replace the Truth Vault resident there (MisinfoForensics / TruthVault share that
# handle) and re-upload + re-normalise the database on every query.  Small LRU: a handle owns HBM.
_INDEX_CACHE: "dict[tuple, tuple]" = {}
_INDEX_CACHE_MAX = 4


def _article_index(embeddings_db: dict, key: str, device) -> Engine:
    dev = torch.device(device)
    ck = (id(embeddings_db), key, str(dev))
    hit = _INDEX_CACHE.get(ck)
    if hit is not None and hit[0] is embeddings_db:
        _INDEX_CACHE[ck] = _INDEX_CACHE.pop(ck)                    # most recently used last
        return hit[1]
    eng = Engine(dev)
    eng.vault_load(np.asarray(embeddings_db[key]), mode="fp32")
    _INDEX_CACHE[ck] = (embeddings_db, eng)                        # keeps the dict alive, so id() stays unique
    while len(_INDEX_CACHE) > _INDEX_CACHE_MAX:
        _INDEX_CACHE.pop(next(iter(_INDEX_CACHE)))[1].close()
    return eng


def search_similar_articles(query_text: Optional[str] = None, query_image_path: Optional[str] = None,
                            embeddings_db_path: str = "guardian_embeddings.pkl", top_k: int = 5, search_mode: str = "text",
                            *, clip_model=None, processor=None, engine: Optional[Engine] = None, embeddings_db: Optional[dict] = None,
                            clip_model_dir: str = r"C:\Users\Lenovo\OneDrive\Desktop\hack\models\clip-vit-b32",
                            clip_weights: str = "clip_detective_best.pth", device: str = "cuda") -> List[dict]:
    """Positional / keyword arguments up to search_mode are the reference's.  The keyword-only ones let a caller inject an
    already-loaded CLIP model (anything with get_text_features / get_image_features), its processor, an Engine (used for
    its DEVICE only: the database is kept in a private, cached handle, never in the caller's, whose resident Truth Vault
    stays untouched) and an in-memory database dict (offline use, tests, repeated queries: the same dict object is
    uploaded once); by default everything is loaded like the reference does.  Returns the reference's records: rank, article_id, similarity, text (first 100 characters + '...'), image_path."""
    print(f"\nSearching for similar articles (mode: {search_mode}, top_k: {top_k})...")
    dev = engine.device if engine is not None else device
    if embeddings_db is None:
        with open(embeddings_db_path, "rb") as fh:
            embeddings_db = pickle.load(fh)
    if clip_model is None or processor is None:
        from transformers import CLIPModel, CLIPProcessor
        processor = processor or CLIPProcessor.from_pretrained(clip_model_dir)
        if clip_model is None:
            clip_model = CLIPModel.from_pretrained(clip_model_dir)
            ckpt = torch.load(clip_weights, map_location="cpu", weights_only=False)
            # the checkpoint is a CLIPDetective state dict: its CLIP weights live under the 'clip.' prefix
            sd = {k[len("clip."):]: v for k, v in ckpt["model_state_dict"].items() if k.startswith("clip.")}
            clip_model.load_state_dict(sd, strict=False)
    if hasattr(clip_model, "to"):
        clip_model = clip_model.to(dev)
    if hasattr(clip_model, "eval"):
        clip_model.eval()

    def feats(x):
        return x if torch.is_tensor(x) else x.pooler_output       # transformers >= 5

    with torch.no_grad():
        if search_mode == "text" and query_text:
            inputs = processor(text=[query_text], return_tensors="pt", padding=True, truncation=True)
            inputs = {k: v.to(dev) for k, v in inputs.items()}
            query_embed = feats(clip_model.get_text_features(**inputs))
            db_key = "text_embeddings"
        elif search_mode == "image" and query_image_path:
            image = Image.open(query_image_path).convert("RGB")
            inputs = processor(images=[image], return_tensors="pt")
            inputs = {k: v.to(dev) for k, v in inputs.items()}
            query_embed = feats(clip_model.get_image_features(**inputs))
            db_key = "image_embeddings"
        else:
            raise ValueError("Invalid search mode or missing query")
    # query normalisation, similarities and top-k (train_clip_detective.py:657-664) on the device
    index = engine.article_index(embeddings_db, db_key) if hasattr(engine, "article_index") else \
        _article_index(embeddings_db, db_key, dev)
    scores, rows, _ = index.vault_search(query_embed.reshape(1, -1).float(), int(top_k))
    scores, rows = scores[0].tolist(), rows[0].tolist()
    print(f"\nTop {top_k} similar articles:")
    print("-" * 60)
    results = []
    for i, (sim, idx) in enumerate(zip(scores, rows)):
        if idx < 0:
            break                                                  # top_k > database size: the reference returns what exists
        result = {"rank": i + 1, "article_id": embeddings_db["article_ids"][idx], "similarity": float(sim),
                  "text": embeddings_db["text_contents"][idx][:100] + "...", "image_path": embeddings_db["image_paths"][idx]}
        results.append(result)
        print(f"{i + 1}. [{result['similarity']:.4f}] {result['article_id']}")
        print(f"   {result['text']}")
        print()
    return results
