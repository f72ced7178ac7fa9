"""Seeded synthetic inputs for the scoring hot path: no datasets or
weights exist offline, so benchmarks and tests run on 512-d CLIP-ViT-B/32-shaped
embeddings and default-init fusion weights.  NumPy only; nothing here is timed."""
from __future__ import annotations

import numpy as np

D = 512
VAULT_SEED, QUERY_SEED, FUSION_SEED, SCORE_SEED = 1234, 5678, 0, 42
PLANT_COSINES = (0.80, 0.849, 0.851, 0.90, 0.99)


def vault_rows(n_rows: int, row_offset: int = 0, dim: int = D, seed: int = VAULT_SEED,
               normalised: bool = True, block: int = 65536) -> np.ndarray:
    """Rows [row_offset, row_offset+n_rows) of the synthetic vault.  Generated per 64Ki-row
    block from (seed, block index), so a shard can be produced without the rest of the
    vault ever existing on the host.  Rows are L2-normalised like the reference's vault
    writer does (train_clip_detective.py:556)."""
    out = np.empty((n_rows, dim), np.float32)
    r = row_offset
    end = row_offset + n_rows
    while r < end:
        b = r // block
        lo, hi = b * block, (b + 1) * block
        g = np.random.default_rng([seed, b])
        blk = g.standard_normal((block, dim), dtype=np.float32)
        s, e = r - lo, min(end, hi) - lo
        out[r - row_offset:r - row_offset + (e - s)] = blk[s:e]
        r = lo + e
    if normalised:
        out /= np.linalg.norm(out, axis=1, keepdims=True)
    return out


def planted_query(row: np.ndarray, cos: float, g: np.random.Generator) -> np.ndarray:
    """A vector with cosine `cos` to `row`: cos*r + sin*u, u a unit vector orthogonal to r."""
    r = row.astype(np.float64) / np.linalg.norm(row.astype(np.float64))
    u = g.standard_normal(row.shape[0])
    u -= (u @ r) * r
    u /= np.linalg.norm(u)
    return (cos * r + np.sqrt(max(0.0, 1 - cos * cos)) * u).astype(np.float32)


def queries(n_q: int, n_vault: int, dim: int = D, seed: int = QUERY_SEED, plant_frac: float = 0.1,
            vault_seed: int = VAULT_SEED, scale: bool = True):
    """90% N(0,1) queries (no vault match -> discrepancy 0) and 10% planted near-duplicates
    of uniformly drawn vault rows at the cosines in PLANT_COSINES (both sides of the 0.85
    rule).  Returns (Q (n_q,dim) fp32 un-normalised, planted_row (n_q,) int64, -1 if none,
    planted_cos (n_q,) fp32)."""
    g = np.random.default_rng(seed)
    q = g.standard_normal((n_q, dim), dtype=np.float32)
    rows = np.full(n_q, -1, np.int64)
    cosv = np.zeros(n_q, np.float32)
    n_plant = int(round(n_q * plant_frac)) if n_vault > 0 else 0
    if n_plant:
        which = g.choice(n_q, n_plant, replace=False)
        blocks = {}                         # 64Ki-row blocks of the seeded vault already generated in this call: a planted
                                            # row costs one block (33M normals), not one block per planted query

        def vault_row(row: int, block: int = 65536) -> np.ndarray:
            b = row // block
            if b not in blocks:
                if len(blocks) >= 6:
                    blocks.pop(next(iter(blocks)))
                blocks[b] = np.random.default_rng([vault_seed, b]).standard_normal((block, dim), dtype=np.float32)
            out = blocks[b][row - b * block:row - b * block + 1].copy()
            out /= np.linalg.norm(out, axis=1, keepdims=True)       # the very operations of vault_rows(1, row, dim, vault_seed)
            return out[0]
        for j, qi in enumerate(which):
            row = int(g.integers(0, n_vault))
            c = PLANT_COSINES[j % len(PLANT_COSINES)]
            q[qi] = planted_query(vault_row(row), c, g)
            rows[qi], cosv[qi] = row, c
    if scale:   # embeddings reach the hot path un-normalised; exercise the normalise step
        q *= g.uniform(0.5, 20.0, size=(n_q, 1)).astype(np.float32)
    return q, rows, cosv


def caption_image_pairs(n: int, dim: int = D, seed: int = QUERY_SEED + 1, corr_frac: float = 0.1):
    """Two independent N(0,1) (n,dim) sets plus a correlated subset with cosine ~0.25 +- 0.01
    to straddle the Match threshold (clip_similarity_engine.py:18)."""
    g = np.random.default_rng(seed)
    a = g.standard_normal((n, dim), dtype=np.float32) * 3.0
    b = g.standard_normal((n, dim), dtype=np.float32) * 0.7
    for i in g.choice(n, int(round(n * corr_frac)), replace=False):
        c = 0.25 + float(g.uniform(-0.01, 0.01))
        b[i] = planted_query(a[i], c, g) * 2.5
    return a, b


def fusion_state_dict(seed: int = FUSION_SEED) -> dict:
    """nn.Linear default init of the fusion judge (misinfo_forensics.py:83-90) under
    torch.manual_seed(seed), as a state dict with the reference's key names."""
    import torch
    import torch.nn as nn
    torch.manual_seed(seed)
    layer = nn.Sequential(nn.Linear(5, 64), nn.ReLU(), nn.Dropout(0.2),
                          nn.Linear(64, 32), nn.ReLU(), nn.Linear(32, 2))
    return {k: v.detach().clone() for k, v in layer.state_dict().items()}


def head_scores(n: int, seed: int = SCORE_SEED) -> np.ndarray:
    """(n,3) U(0,1): ai_score, misinfo_score, deepfake_score columns of the fusion input."""
    return np.random.default_rng(seed).uniform(0, 1, size=(n, 3)).astype(np.float32)
