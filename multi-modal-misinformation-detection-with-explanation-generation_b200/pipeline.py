"""Batched scoring pipeline: everything downstream of the encoders for a batch of samples,
device-resident from the embeddings to the verdict (the batched form of
misinfo_forensics.py:396-404, :438-464, :587-608, :866-900)."""
from __future__ import annotations

import torch

from .engine import Engine, VAULT_THRESHOLD


def score_batch(engine: Engine, vault, text_embeds, image_embeds, head_scores, modality=None, top_k: int = 5,
                algo: str = "auto"):
    """text_embeds/image_embeds (B,512), head_scores (B,3) = [ai, misinfo, deepfake] (host or
    device; host tensors are copied in, pinned memory makes that asynchronous), modality (B,)
    uint8 (bit0 text, bit1 visual; default both), vault: a TruthVault or None.
    Per row the results equal the scalar path (MisinfoForensics.analyze) on the same producer
    outputs.  Returns a dict of device tensors.

    No torch arithmetic: a vault resident in `engine` (or none) is ONE library call (mmf_score_batch); a row-sharded
    vault is three (cosine, sharded search with its all-gather, score assembly + verdict)."""
    sharded = vault is not None and getattr(vault, "world", 1) > 1
    one_call = hasattr(engine, "score_batch") and not sharded and (
        (vault is not None and getattr(vault, "engine", None) is engine) or
        (vault is None and not getattr(engine, "vault_rows", 0)))        # (an engine may hold a vault the caller did not pass)
    if one_call:
        return engine.score_batch(text_embeds, image_embeds, head_scores, modality, top_k, VAULT_THRESHOLD, algo)
    dev = engine.device
    t = torch.as_tensor(text_embeds).to(dev, torch.float32, non_blocking=True)
    im = torch.as_tensor(image_embeds).to(dev, torch.float32, non_blocking=True)
    b = im.shape[0]
    sim = engine.cosine_pairs(t, im)
    if vault is not None:
        vs, vr, disc = vault.search(im, top_k, VAULT_THRESHOLD, algo)
    else:
        vs = torch.full((b, top_k), float("nan"), device=dev)
        vr = torch.full((b, top_k), -1, dtype=torch.int64, device=dev)
        disc = torch.zeros(b, device=dev)
    if hasattr(engine, "verdict_assemble"):
        x, probs, verdict, conf = engine.verdict_assemble(head_scores, modality, sim, disc)
        return {"clip_similarity": sim, "vault_discrepancy": disc, "vault_scores": vs, "vault_rows": vr,
                "scores": x, "probs": probs, "verdict": verdict, "confidence": conf}
    # engines without the fused tail (the CPU test double of tests/cpu_engine.py): same rules, tensor by tensor
    hs = torch.as_tensor(head_scores).to(dev, torch.float32, non_blocking=True)
    if modality is None:
        mod = torch.full((b,), 3, dtype=torch.uint8, device=dev)
        x = torch.cat([hs, sim[:, None], disc[:, None]], dim=1)
    else:
        mod = torch.as_tensor(modality).to(dev, torch.uint8, non_blocking=True)
        has_text, has_vis = (mod & 1).bool(), (mod & 2).bool()
        zero = torch.zeros_like(sim)
        sim = torch.where(has_text & has_vis, sim, zero)          # analyze() skips the steps whose
        disc = torch.where(has_vis, disc, zero)                   # modality is missing -> 0.0
        x = torch.stack([torch.where(has_text, hs[:, 0], zero), torch.where(has_text, hs[:, 1], zero),
                         torch.where(has_vis, hs[:, 2], zero), sim, disc], dim=1)
    probs, verdict, conf = engine.verdict_batch(x, mod)
    return {"clip_similarity": sim, "vault_discrepancy": disc, "vault_scores": vs, "vault_rows": vr,
            "scores": x, "probs": probs, "verdict": verdict, "confidence": conf}
