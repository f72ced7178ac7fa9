"""Truth Vault: on-disk formats of the reference (misinfo_forensics.py:214-246 reader,
train_clip_detective.py:515-573 writer), row-sharding across ranks, and the search that
returns the reference's `matches` records.  Strings (titles/urls) stay on the host,
indexed by global row id; only (row id, score) comes back from the device."""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .engine import Engine, VAULT_THRESHOLD


def read_vault_dict(vault_data: dict):
    """(embeddings, metadata) from either pickle layout, (None, None) if unknown:
    {'embeddings', 'metadata': [ {title,url,date} ]}  or the vault-builder layout
    {'image_embeddings', 'text_contents', 'image_paths', ...} whose metadata is synthesised
    (title = text, url = image path, date 'N/A')."""
    if "embeddings" in vault_data:
        return vault_data["embeddings"], vault_data["metadata"]
    if "image_embeddings" in vault_data:
        texts = vault_data.get("text_contents", [])
        meta = []
        for i, title in enumerate(texts):
            paths = vault_data["image_paths"]
            meta.append({"title": title, "url": paths[i] if i < len(paths) else "N/A", "date": "N/A"})
        return vault_data["image_embeddings"], meta
    return None, None


def load_vault_file(path: str):
    with open(path, "rb") as fh:
        data = pickle.load(fh)
    emb, meta = read_vault_dict(data)
    return data, emb, meta


@dataclass(frozen=True)
class ShardPlan:
    """Contiguous row-sharding: rank r owns rows [r*ceil(N/R), min(N,(r+1)*ceil(N/R)))."""
    n_rows: int
    world: int

    @property
    def rows_per_rank(self) -> int:
        return -(-self.n_rows // self.world) if self.world > 0 else 0

    def bounds(self, rank: int) -> Tuple[int, int]:
        per = self.rows_per_rank
        lo = min(self.n_rows, rank * per)
        return lo, min(self.n_rows, lo + per)

    def owner(self, row: int) -> int:
        return int(row) // max(1, self.rows_per_rank)


def exchange_candidates(local_packed: torch.Tensor, group=None) -> torch.Tensor:
    """The path's ONE collective: all-gather of the per-shard packed top-k candidates
    ((Q,k) int64 each) -> (world, Q, k).  NCCL over NVLink on the GPU box; the same call
    runs over gloo in the CPU tests of the host logic."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    nq, k = local_packed.shape
    out = torch.empty((world * nq, k), dtype=local_packed.dtype, device=local_packed.device)
    dist.all_gather_into_tensor(out, local_packed.contiguous(), group=group)   # rank-major concat
    return out.view(world, nq, k)


class TruthVault:
    """A (possibly row-sharded) vault resident in HBM plus its host-side metadata."""

    def __init__(self, engine: Engine, embeddings, metadata: Optional[Sequence[dict]] = None, mode: str = "fp32",
                 rank: int = 0, world: int = 1, group=None, n_total: Optional[int] = None, row_offset: Optional[int] = None,
                 exchange: Optional[str] = None):
        """exchange: how the shards' candidates meet when world > 1 -- "nccl" (one all-gather, the default) or
        "p2p" (stores into the peers' symmetric memory + flag wait fused into the merge kernel, csrc/exchange.cu;
        not yet validated on a multi-GPU box); None reads MMF_EXCHANGE."""
        self.engine = engine
        self.exchange = (exchange or os.environ.get("MMF_EXCHANGE", "nccl")).lower()
        if self.exchange not in ("nccl", "p2p"):
            raise ValueError(f"exchange must be 'nccl' or 'p2p', got {self.exchange!r}")
        self._symm = None          # (tensor, handle, bytes) of the symmetric exchange buffer
        self.metadata = metadata
        self.mode = mode
        self.rank, self.world, self.group = rank, world, group
        n_local = int(embeddings.shape[0])
        if row_offset is None:
            # `embeddings` is the whole vault: keep only this rank's slice
            self.plan = ShardPlan(n_local, world)
            lo, hi = self.plan.bounds(rank)
            shard = embeddings[lo:hi]
            self.n_total = n_local
        else:
            # `embeddings` already is this rank's shard of an n_total-row vault
            self.plan = ShardPlan(int(n_total), world)
            lo, shard = int(row_offset), embeddings
            self.n_total = int(n_total)
        self.row_offset = lo
        engine.vault_load(shard, mode=mode, row_offset=lo)

    def search(self, queries, top_k: int = 5, threshold: float = VAULT_THRESHOLD, algo: str = "auto"):
        """(scores (Q,k), rows (Q,k), discrepancy (Q,)) device tensors; global over all shards."""
        if self.world == 1:
            return self.engine.vault_search(queries, top_k, threshold, algo)
        k_local = min(top_k, max(1, self.plan.rows_per_rank))
        if self.exchange == "p2p":
            nq = int(queries.shape[0]) if hasattr(queries, "shape") and len(queries.shape) == 2 else 1
            self._ensure_peer_buffers(nq, k_local)
            return self.engine.vault_search_exchange(queries, top_k, k_local, threshold, algo)
        packed = self.engine.vault_search_candidates(queries, k_local, algo)
        gathered = exchange_candidates(packed, self.group)
        return self.engine.topk_merge(gathered, top_k, threshold)

    def _ensure_peer_buffers(self, n_queries: int, k_local: int) -> None:
        """(Re)allocate the symmetric exchange buffer when this search needs more than is attached.  Collective:
        every rank runs the same searches in the same order, so all of them get here together."""
        need = self.engine.exchange_layout(self.world, n_queries, k_local)
        if self._symm is not None and self._symm[2] >= need:
            return
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        if self._symm is not None:
            torch.cuda.synchronize(self.engine.device)
            dist.barrier(self.group)                  # nobody still pushes into the buffers being replaced
            self.engine.exchange_detach()
        nbytes = max(need + need // 2, 1 << 20)
        buf = symm_mem.empty(nbytes // 8, dtype=torch.int64, device=self.engine.device)
        group = self.group if self.group is not None else dist.group.WORLD
        hdl = symm_mem.rendezvous(buf, group)
        self.engine.exchange_attach(self.rank, self.world, list(hdl.buffer_ptrs), (nbytes // 8) * 8)
        torch.cuda.synchronize(self.engine.device)
        dist.barrier(self.group)                      # every rank has cleared its flags before anyone pushes
        self._symm = (buf, hdl, (nbytes // 8) * 8)

    def matches(self, scores_row, rows_row) -> List[dict]:
        """The reference's match records (misinfo_forensics.py:452-460) for one query."""
        out = []
        for s, r in zip(scores_row, rows_row):
            r = int(r)
            if r < 0:
                break                       # top_k > n_rows: the reference returns n_rows items
            m = self.metadata[r]
            out.append({"similarity": float(s), "title": m["title"], "url": m.get("url", "N/A"),
                        "date": m.get("date", "N/A")})
        return out
