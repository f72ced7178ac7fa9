#!/usr/bin/env python
"""C5 (BASELINE.json configs[4]): full batched analyze on one GPU / N data-parallel replicas --
random-init PyTorch producers (RoBERTa-base text heads, EfficientNet-B0 deepfake head, CLIP ViT-B/32 towers,
SURVEY.md 8d: weights and datasets are not available offline) feeding the B200 scoring kernels
(caption/image cosine, Truth-Vault top-k + discrepancy, fusion judge) through mmf_b200.score_batch.

    python tools/bench_c5.py [--batch 256] [--rows 1000000] [--steps 10] [--warmup 3] [--dry-run]
    torchrun --nproc-per-node 8 tools/bench_c5.py          # 8 replicas, vault replicated, no collective

Prints one JSON line: samples/s (whole job), and how the step time splits between the encoders (out of scope
of this repo: stock PyTorch) and the scoring hot path (this repo's kernels).  --dry-run builds the producers
and one tiny CPU forward only (no GPU needed: checks shapes and the offline random-init recipe).

Measured on one B200 (round 2, bf16 autocast encoders): 5 476 samples/s; the scoring hot path is 0.30 ms of a 46.7 ms
step (0.65 %): this workload is encoder-bound, the kernels of this repo are off its critical path.  bench.py carries
the same measurement as its "c5" record at every N."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_producers(seed: int = 0):
    """The reference's producer stack (misinfo_forensics.py:43-108, :210-211) with random-init weights."""
    from torchvision import models
    from transformers import CLIPConfig, CLIPModel, RobertaConfig, RobertaModel
    torch.manual_seed(seed)
    clip = CLIPModel(CLIPConfig())                     # default config == ViT-B/32, projection 512
    roberta = RobertaModel(RobertaConfig(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, pad_token_id=1))
    hidden = roberta.config.hidden_size

    def head():
        return torch.nn.Sequential(torch.nn.Linear(hidden, 256), torch.nn.ReLU(), torch.nn.Dropout(0.3), torch.nn.Linear(256, 2))
    ai_head, misinfo_head = head(), head()
    eff = models.efficientnet_b0(weights=None)
    eff.classifier = torch.nn.Sequential(torch.nn.Dropout(0.2), torch.nn.Linear(1280, 2))
    return clip.eval(), roberta.eval(), ai_head.eval(), misinfo_head.eval(), eff.eval()


@torch.no_grad()
def encode(prod, pixel_values, clip_ids, roberta_ids):
    """-> text_embeds (B,512), image_embeds (B,512), head scores (B,3) = [ai, misinfo, deepfake] fake-probabilities
    (softmax[:, 1], misinfo_forensics.py:290-296, :332-336, :361-364)."""
    clip, roberta, ai_head, misinfo_head, eff = prod
    out = clip(input_ids=clip_ids, pixel_values=pixel_values)
    cls = roberta(input_ids=roberta_ids).last_hidden_state[:, 0, :]
    ai = torch.softmax(ai_head(cls), dim=1)[:, 1]
    mis = torch.softmax(misinfo_head(cls), dim=1)[:, 1]
    deep = torch.softmax(eff(pixel_values), dim=1)[:, 1]
    return out.text_embeds.float(), out.image_embeds.float(), torch.stack([ai, mis, deep], dim=1).float()


def synthetic_inputs(batch, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    px = torch.randn(batch, 3, 224, 224, device=device, generator=g)
    clip_ids = torch.randint(0, 49406, (batch, 77), device=device, generator=g)
    clip_ids[:, -1] = 49407                            # EOS last: CLIP pools at the EOS position
    rob_ids = torch.randint(3, 50265, (batch, 128), device=device, generator=g)
    return px, clip_ids, rob_ids


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--top-k", type=int, default=5)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--autocast", action="store_true", help="run the encoders under bf16 autocast (the hot path stays fp32)")
    ap.add_argument("--dry-run", action="store_true")
    args = ap.parse_args()

    prod = build_producers()
    if args.dry_run:
        px, cid, rid = synthetic_inputs(2, torch.device("cpu"), 1)
        t, i, h = encode(prod, px, cid, rid)
        assert t.shape == (2, 512) and i.shape == (2, 512) and h.shape == (2, 3) and bool(((h >= 0) & (h <= 1)).all())
        print(json.dumps({"dry_run": "ok", "text_embeds": list(t.shape), "image_embeds": list(i.shape), "head_scores": list(h.shape)}))
        return

    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import mmf_b200
    from mmf_b200 import synth
    eng = mmf_b200.Engine(dev)
    prod = tuple(m.to(dev) for m in prod)
    g = torch.Generator(device=dev).manual_seed(synth.VAULT_SEED)
    vault = mmf_b200.TruthVault(eng, torch.randn(args.rows, 512, device=dev, generator=g), None, mode="fp32")   # replicated per rank
    eng.fusion_load(synth.fusion_state_dict())
    px, cid, rid = synthetic_inputs(args.batch, dev, 100 + rank)

    def step(events=None):
        if events:
            events[0].record()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.autocast):
            t, i, h = encode(prod, px, cid, rid)
        if events:
            events[1].record()
        out = eng.score_batch(t, i, h, None, args.top_k)          # ONE library call: cosine + vault top-k + verdict
        if events:
            events[2].record()
        return out

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    for e in evs:
        out = step(e)
    verdict = out["verdict"].cpu()                     # the result a caller reads
    torch.cuda.synchronize()
    enc_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    hot_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    step_ms = evs[0][0].elapsed_time(evs[-1][2]) / args.steps
    if world > 1:
        t = torch.tensor([step_ms, enc_ms, hot_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, enc_ms, hot_ms = t.tolist()
    if rank == 0:
        print(json.dumps({
            "metric": "analyze samples/s", "value": args.batch * world / (step_ms * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "data": "synthetic", "dtype": "bf16 autocast encoders + f32 hot path" if args.autocast else "f32",
            "config": {"workload": "C5: batched analyze, random-init RoBERTa-base + EfficientNet-B0 + CLIP ViT-B/32 producers -> "
                                   "cosine + vault top-%d + fusion judge" % args.top_k,
                       "batch_per_gpu": args.batch, "vault_rows_per_gpu": args.rows, "parallelism": "replica x%d, vault replicated, no collective" % world},
            "split": {"encoders_ms": enc_ms, "scoring_hot_path_ms": hot_ms, "hot_path_fraction": hot_ms / max(step_ms, 1e-9)},
            "fake_verdicts": int(verdict.sum())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
