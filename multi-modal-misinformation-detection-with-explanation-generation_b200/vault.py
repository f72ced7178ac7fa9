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
    """All-gather of the per-shard packed top-k candidates ((Q,k) int64 each) -> (world, Q, k) through
    torch.distributed (exchange="torch"): the path the CPU tests run over gloo.  On the GPU box the default is the
    same collective owned by the library (exchange="nccl": csrc/shard.cu, no torch in the data path)."""
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
        """exchange: how the shards' candidates meet when world > 1 --
        "nccl"  (default) the library's own communicator: local search + ncclAllGather + merge inside
                mmf_vault_search_sharded (csrc/shard.cu); torch.distributed only carries the 128-byte unique id once;
        "p2p"   stores into the peers' symmetric memory + flag wait fused into the merge kernel (csrc/exchange.cu),
                no library collective at all;
        "torch" torch.distributed.all_gather_into_tensor between two library calls (works over gloo: the CPU tests).
        None reads MMF_EXCHANGE, else "nccl" ("torch" for engines without the sharded entry points)."""
        self.engine = engine
        default = "nccl" if hasattr(engine, "vault_search_sharded") else "torch"
        self.exchange = (exchange or os.environ.get("MMF_EXCHANGE", default)).lower()
        if self.exchange not in ("nccl", "p2p", "torch"):
            raise ValueError(f"exchange must be 'nccl', 'p2p' or 'torch', got {self.exchange!r}")
        self._symm = None          # (tensor, handle, bytes) of the symmetric exchange buffer
        self._shard_ready = False
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

    def _ensure_shard_group(self) -> None:
        """Collective, once: rank 0's ncclUniqueId travels through torch.distributed, then every rank joins the
        library's communicator (mmf_shard_init)."""
        if self._shard_ready:
            return
        import torch.distributed as dist
        box = [self.engine.shard_unique_id() if self.rank == 0 else None]
        group = self.group if self.group is not None else dist.group.WORLD
        src = dist.get_global_rank(group, 0) if hasattr(dist, "get_global_rank") else 0
        dist.broadcast_object_list(box, src=src, group=self.group)
        self.engine.shard_init(self.rank, self.world, box[0])
        self._shard_ready = True

    def search(self, queries, top_k: int = 5, threshold: float = VAULT_THRESHOLD, algo: str = "auto"):
        """(scores (Q,k), rows (Q,k), discrepancy (Q,)) device tensors; global over all shards."""
        if self.world == 1:
            return self.engine.vault_search(queries, top_k, threshold, algo)
        if self.exchange == "nccl":
            self._ensure_shard_group()
            return self.engine.vault_search_sharded(queries, top_k, threshold, algo)
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
