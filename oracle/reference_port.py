"""CPU restatement of the reference's scoring hot path.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it follows (paths relative to
/root/reference).  The arithmetic libraries are the reference's own: torch CPU
ops for the cosine / fusion maths and NumPy for the Truth-Vault search
(misinfo_forensics.py:443-450 is NumPy, not torch).  De-facto version pin (the
reference's requirements.txt is unpinned): torch 2.11.0, numpy 2.3.5.

Nothing in the product package imports this module (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = [
    "VAULT_THRESHOLD", "MATCH_THRESHOLD", "FAKE_THRESHOLD", "FUSION_KEYS", "FUSION_ORDER",
    "normalise_rows", "cosine_pairs", "clip_match_label", "clip_explanation",
    "vault_search_as_shipped", "vault_normalise", "vault_search_batched", "order_key64",
    "discrepancy_rule", "merge_topk", "fusion_weights_from_checkpoint", "fusion_forward",
    "fusion_verdict", "fallback_verdict", "assemble_verdict", "video_aggregate",
    "read_vault_dict", "similar_articles_as_shipped",
]

VAULT_THRESHOLD = 0.85   # misinfo_forensics.py:464,468
MATCH_THRESHOLD = 0.25   # clip_similarity_engine.py:18
FAKE_THRESHOLD = 0.5     # misinfo_forensics.py:605,892
# nn.Sequential indices of the three Linear layers, misinfo_forensics.py:83-90
FUSION_KEYS = ("0.weight", "0.bias", "3.weight", "3.bias", "5.weight", "5.bias")
# misinfo_forensics.py:587-593
FUSION_ORDER = ("ai_score", "misinfo_score", "deepfake_score", "clip_similarity", "vault_discrepancy")


# --------------------------------------------------------------------------- cosine
def normalise_rows(x: torch.Tensor) -> torch.Tensor:
    """x / x.norm(dim=-1, keepdim=True), no eps -- misinfo_forensics.py:400-401, :439, :481."""
    return x / x.norm(dim=-1, keepdim=True)


def cosine_pairs(a, b, scalar_loop: bool = False) -> np.ndarray:
    """Row-wise normalise-then-dot, the batched form of misinfo_forensics.py:399-404
    (and clip_similarity_engine.py:103-108, same scalar with the operands swapped).

    scalar_loop=True replays the reference literally, one (1,512)@(512,1) matmul and
    one .item() per pair; the vectorised form differs only in summation order (~1e-7).
    """
    a = torch.as_tensor(np.asarray(a), dtype=torch.float32)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float32)
    if scalar_loop:
        out = np.empty(a.shape[0], dtype=np.float32)
        for i in range(a.shape[0]):
            t = normalise_rows(a[i:i + 1])
            m = normalise_rows(b[i:i + 1])
            out[i] = (t @ m.T).item()
        return out
    return (normalise_rows(a) * normalise_rows(b)).sum(-1).numpy()


def clip_match_label(similarity: float, threshold: float = MATCH_THRESHOLD) -> str:
    """clip_similarity_engine.py:111."""
    return "Match" if similarity >= threshold else "Mismatch"


def clip_explanation(similarity: float, label: str) -> str:
    """Tier selection of clip_similarity_engine.py:163-174 (returns the tier name only;
    the sentences themselves are presentation)."""
    if label == "Match":
        return "strong" if similarity >= 0.7 else ("moderate" if similarity >= 0.5 else "weak")
    return "strong_mismatch" if similarity < 0.1 else "mismatch"


# --------------------------------------------------------------------------- vault
def vault_search_as_shipped(vault: np.ndarray, image_embed: np.ndarray, top_k: int = 5):
    """The numeric core of MisinfoForensics.search_vault exactly as shipped,
    misinfo_forensics.py:438-464: torch-normalise the query, renormalise the WHOLE vault
    in the vault's own dtype, NumPy matvec, full argsort, last k reversed, > 0.85 rule.

    Returns (indices int64 (k,), similarities (k,), vault_discrepancy python float).
    """
    e = torch.as_tensor(np.asarray(image_embed), dtype=torch.float32).reshape(1, -1)
    q = normalise_rows(e).cpu().numpy()[0]                                  # :439-440
    vn = vault / np.linalg.norm(vault, axis=1, keepdims=True)               # :443-445
    sims = vn @ q                                                           # :446
    idx = np.argsort(sims)[-top_k:][::-1]                                   # :449
    top = sims[idx]                                                         # :450
    max_similarity = float(top[0])                                          # :463
    disc = max_similarity if max_similarity > VAULT_THRESHOLD else 0.0      # :464
    return idx.astype(np.int64), top, disc


def vault_normalise(vault: np.ndarray) -> np.ndarray:
    """misinfo_forensics.py:443-445 hoisted out of the per-query loop (V is immutable
    after load, so the result is identical for every query)."""
    return vault / np.linalg.norm(vault, axis=1, keepdims=True)


def order_key64(scores: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """Total order used to make top-k deterministic: score descending, NaN ranked above
    everything (np.argsort puts NaN last, so [-k:][::-1] returns it FIRST --
    misinfo_forensics.py:449), equal scores -> higher index first (what a stable
    ascending argsort followed by [::-1] gives).  Larger key == better."""
    s = np.asarray(scores, dtype=np.float32) + np.float32(0.0)       # -0.0 -> +0.0
    u = s.view(np.uint32).astype(np.uint64)
    neg = (u >> np.uint64(31)) != 0
    key = np.where(neg, u ^ np.uint64(0xFFFFFFFF), u ^ np.uint64(0x80000000))
    key = np.where(np.isnan(s), np.uint64(0xFFFFFFFF), key)
    return (key << np.uint64(32)) | np.asarray(idx, dtype=np.uint64)


def vault_search_batched(vault: np.ndarray, queries: np.ndarray, top_k: int,
                         vault_is_normalised: bool = False, chunk: int = 64,
                         row_offset: int = 0, queries_normalised: bool = False):
    """Batched restatement of misinfo_forensics.py:438-450: per query identical to
    vault_search_as_shipped except that ties / NaN follow order_key64 (the shipped
    np.argsort is unstable, so tie order there is implementation-defined).

    Returns (idx int64 (Q,k'), scores float32 (Q,k'), discrepancy float32 (Q,)),
    k' = min(k, N) as in the reference (argsort[-k:] of a shorter array).
    """
    vn = vault if vault_is_normalised else vault_normalise(vault)
    vn32 = np.ascontiguousarray(vn, dtype=np.float32)   # numpy promotes fp16 @ fp32 -> fp32
    q = torch.as_tensor(np.asarray(queries), dtype=torch.float32).reshape(-1, vault.shape[1])
    qn = (q if queries_normalised else normalise_rows(q)).numpy()
    n = vn32.shape[0]
    k = min(int(top_k), n)
    nq = qn.shape[0]
    out_i = np.empty((nq, k), dtype=np.int64)
    out_s = np.empty((nq, k), dtype=np.float32)
    ar = np.arange(n, dtype=np.uint64)
    for c0 in range(0, nq, chunk):
        s = qn[c0:c0 + chunk] @ vn32.T                                      # :446 batched
        for r in range(s.shape[0]):
            keys = order_key64(s[r], ar)
            if k < n:
                part = np.argpartition(keys, n - k)[n - k:]
            else:
                part = np.arange(n)
            sel = part[np.argsort(keys[part])[::-1]]
            out_i[c0 + r] = sel
            out_s[c0 + r] = s[r, sel]
    disc = discrepancy_rule(out_s[:, 0]) if k > 0 else np.zeros(nq, np.float32)
    return out_i + int(row_offset), out_s, disc


def discrepancy_rule(max_similarity: np.ndarray, threshold: float = VAULT_THRESHOLD) -> np.ndarray:
    """misinfo_forensics.py:463-464 -- the compare is Python double vs an fp32 value:
    float(s) > 0.85  (so the fp32 value 0.85f = 0x3F59999A, which is > 0.85, passes;
    NaN fails)."""
    s = np.asarray(max_similarity, dtype=np.float32)
    return np.where(s.astype(np.float64) > float(threshold), s, np.float32(0.0)).astype(np.float32)


def merge_topk(idx_parts, score_parts, top_k: int):
    """Global top-k of the union of per-shard top-k lists.  idx are global
    row ids.  Pure selection under order_key64, so the result is independent of how the
    vault was sharded."""
    idx = np.concatenate(idx_parts, axis=1)
    sc = np.concatenate(score_parts, axis=1)
    nq = idx.shape[0]
    k = min(top_k, idx.shape[1])
    out_i = np.empty((nq, k), np.int64)
    out_s = np.empty((nq, k), np.float32)
    for r in range(nq):
        order = np.argsort(order_key64(sc[r], idx[r]))[::-1][:k]
        out_i[r] = idx[r, order]
        out_s[r] = sc[r, order]
    return out_i, out_s


def read_vault_dict(vault_data: dict):
    """Reader semantics of misinfo_forensics.py:222-238: returns (embeddings, metadata)
    or (None, None) for an unknown format."""
    if "embeddings" in vault_data:
        return vault_data["embeddings"], vault_data["metadata"]
    if "image_embeddings" in vault_data:
        texts = vault_data.get("text_contents", [])
        paths = vault_data["image_paths"] if texts else []
        meta = [{"title": texts[i] if i < len(texts) else "Unknown",
                 "url": paths[i] if i < len(paths) else "N/A",
                 "date": "N/A"} for i in range(len(texts))]
        return vault_data["image_embeddings"], meta
    return None, None


def similar_articles_as_shipped(db_embeddings: np.ndarray, query_embed: np.ndarray, top_k: int = 5):
    """Numeric core of search_similar_articles, train_clip_detective.py:657-664: the query is normalised with NumPy,
    the database rows are used AS STORED (the writer normalised them, :556-557 -- no renormalisation here, unlike
    search_vault), full descending argsort, first k.  Returns (indices int64 (k,), similarities (k,))."""
    q = np.asarray(query_embed)
    q = q / np.linalg.norm(q)                                               # :657
    sims = np.dot(db_embeddings, q)                                         # :660
    idx = np.argsort(sims)[::-1][:top_k]                                    # :663
    return idx.astype(np.int64), sims[idx]


# --------------------------------------------------------------------------- fusion judge
def fusion_weights_from_checkpoint(ckpt: dict) -> dict:
    """The two .pth layouts of train_fusion_judge.py:259-267: 'fusion_layer_state_dict'
    (keys 0.weight ... 5.bias) or 'full_model_state_dict' with the 'fusion_layer.' prefix
    (what misinfo_forensics.py:182 loads)."""
    if "fusion_layer_state_dict" in ckpt:
        sd = ckpt["fusion_layer_state_dict"]
        return {k: torch.as_tensor(sd[k]).float() for k in FUSION_KEYS}
    sd = ckpt["full_model_state_dict"]
    return {k: torch.as_tensor(sd["fusion_layer." + k]).float() for k in FUSION_KEYS}


def fusion_forward(weights: dict, x) -> np.ndarray:
    """fusion_layer in eval mode (Dropout = identity), misinfo_forensics.py:83-90,106-108,
    then softmax(dim=1), :597-598.  x (B,5) fp32 -> probs (B,2) fp32 [real, fake]."""
    x = torch.as_tensor(np.asarray(x), dtype=torch.float32).reshape(-1, 5)
    with torch.no_grad():
        h = torch.relu(torch.nn.functional.linear(x, weights["0.weight"], weights["0.bias"]))
        h = torch.relu(torch.nn.functional.linear(h, weights["3.weight"], weights["3.bias"]))
        logits = torch.nn.functional.linear(h, weights["5.weight"], weights["5.bias"])
        return torch.softmax(logits, dim=1).numpy()


def fusion_verdict(weights: dict, scores: dict) -> dict:
    """misinfo_forensics.py:575-615 for one sample."""
    x = [[float(scores.get(k, 0.0)) for k in FUSION_ORDER]]
    p = fusion_forward(weights, np.asarray(x, dtype=np.float32))
    real_prob, fake_prob = float(p[0, 0]), float(p[0, 1])
    verdict = 1 if fake_prob > FAKE_THRESHOLD else 0
    return {"verdict": verdict, "confidence": fake_prob if verdict == 1 else real_prob,
            "fake_probability": fake_prob, "real_probability": real_prob}


def fallback_verdict(scores: dict, has_text: bool, has_visual: bool) -> dict:
    """misinfo_forensics.py:884-899 (missing-modality rule, Python double arithmetic)."""
    if has_text and not has_visual:
        fake_prob = float(scores.get("misinfo_score", 0.0))
    elif has_visual and not has_text:
        fake_prob = float(max(scores.get("deepfake_score", 0.0), scores.get("vault_discrepancy", 0.0)))
    else:
        fake_prob = 0.5
    fake_prob = max(0.0, min(1.0, fake_prob))
    real_prob = 1.0 - fake_prob
    verdict = 1 if fake_prob > FAKE_THRESHOLD else 0
    return {"verdict": verdict, "confidence": fake_prob if verdict == 1 else real_prob,
            "fake_probability": fake_prob, "real_probability": real_prob}


def assemble_verdict(weights: dict, scores: dict, has_text: bool, has_visual: bool) -> dict:
    """misinfo_forensics.py:866-900: fusion only when text AND (image or video)."""
    if has_text and has_visual:
        return fusion_verdict(weights, scores)
    return fallback_verdict(scores, has_text, has_visual)


def video_aggregate(deepfake_scores, clip_sims, vault_results):
    """misinfo_forensics.py:546-572: mean deepfake, mean clip (0.0 if none), the FIRST frame
    with the strictly largest vault_discrepancy (all-zero keeps the initial empty result)."""
    best = {"vault_discrepancy": 0.0, "matches": [], "text_similarity": 0.0}
    best_frame = -1
    for i, v in enumerate(vault_results):
        if float(v.get("vault_discrepancy", 0.0)) > float(best.get("vault_discrepancy", 0.0)):
            best, best_frame = v, i
    return {"deepfake_score": float(np.mean(deepfake_scores)),
            "clip_similarity": float(np.mean(clip_sims)) if len(clip_sims) else 0.0,
            "vault_discrepancy": float(best.get("vault_discrepancy", 0.0)),
            "text_similarity": float(best.get("text_similarity", 0.0)),
            "vault_matches": best.get("matches", []),
            "best_frame_index": best_frame}
