"""Device-side operators of the scoring hot path: thin torch-tensor wrappers over the C ABI.
torch supplies device memory and streams only; all arithmetic runs in libmmf_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import MMFError

VAULT_THRESHOLD = 0.85   # misinfo_forensics.py:464
MATCH_THRESHOLD = 0.25   # clip_similarity_engine.py:18
_TORCH_DTYPE = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16, torch.float64: _lib.F64}
_NP_DTYPE = {np.dtype(np.float32): _lib.F32, np.dtype(np.float16): _lib.F16, np.dtype(np.float64): _lib.F64}
_ALGO = {"auto": _lib.ALGO_AUTO, "stream": _lib.ALGO_STREAM, "mma": _lib.ALGO_MMA}
_MODE = {"fp32": _lib.VAULT_FP32, "bf16": _lib.VAULT_BF16}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    """One libmmf_b200 handle on one CUDA device.  Not thread-safe (like the reference
    object it serves); create one per (process, device)."""

    def __init__(self, device="cuda"):
        self._h = C.c_void_p()
        self.lib = _lib.load()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise MMFError(_lib.ERR_NO_DEVICE, f"mmf_b200 runs on sm_100 CUDA devices only, got device {dev}; no CPU fallback")
        if not torch.cuda.is_available():
            raise MMFError(_lib.ERR_NO_DEVICE, "no CUDA device visible; no CPU fallback")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        rc = self.lib.mmf_create(self.device.index, C.byref(self._h))
        if rc != _lib.OK:
            self._h = C.c_void_p()
            raise MMFError(rc, "mmf_create: " + self.lib.mmf_status_string(rc).decode())
        self.vault_rows = 0
        self.vault_row_offset = 0
        self.vault_mode = None
        self._pending = {}          # slot -> host arrays of a submitted batch (kept alive until collected)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.mmf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _check(self, rc: int):
        if rc != _lib.OK:
            raise MMFError(rc, self.lib.mmf_last_error(self._h).decode() or self.lib.mmf_status_string(rc).decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev_f32(self, x, cols: int) -> torch.Tensor:
        t = torch.as_tensor(x)
        t = t.to(device=self.device, dtype=torch.float32).reshape(-1, cols).contiguous()
        return t

    @property
    def launch_count(self) -> int:
        return int(self.lib.mmf_launch_count(self._h))

    # ------------------------------------------------------------------ cosine
    def cosine_pairs(self, a, b, match_threshold: Optional[float] = None):
        """Row-wise normalise-then-dot of two (n,dim) embedding sets.  Returns sim (n,) fp32
        on the device, or (sim, match uint8) when match_threshold is given."""
        a_t = torch.as_tensor(a)
        dim = a_t.shape[-1]
        a_t, b_t = self._dev_f32(a_t, dim), self._dev_f32(b, dim)
        if a_t.shape != b_t.shape:
            raise ValueError(f"cosine_pairs: shapes differ {tuple(a_t.shape)} vs {tuple(b_t.shape)}")
        n = a_t.shape[0]
        sim = torch.empty(n, dtype=torch.float32, device=self.device)
        match = torch.empty(n, dtype=torch.uint8, device=self.device) if match_threshold is not None else None
        self._check(self.lib.mmf_cosine_pairs(self._h, _ptr(a_t), _ptr(b_t), n, dim,
                                              float(match_threshold if match_threshold is not None else 0.0),
                                              _ptr(sim), _ptr(match), self._stream()))
        return sim if match is None else (sim, match)

    # ------------------------------------------------------------------ vault
    def vault_load(self, rows, mode: str = "fp32", row_offset: int = 0):
        """Normalise + upload this rank's vault rows ((n,512) numpy array or torch tensor,
        fp16/bf16/fp32/fp64, host or device)."""
        if isinstance(rows, torch.Tensor):
            t = rows.contiguous()
            if t.dtype not in _TORCH_DTYPE:
                t = t.float()
            if t.dim() != 2:
                raise ValueError("vault rows must be 2-D")
            on_dev = t.is_cuda
            if on_dev and t.device != self.device:
                t = t.to(self.device)
            dt, ptr, n, dim, keep = _TORCH_DTYPE[t.dtype], t.data_ptr(), t.shape[0], t.shape[1], t
            if on_dev:
                torch.cuda.current_stream(self.device).synchronize()
        else:
            arr = np.ascontiguousarray(rows)
            if arr.dtype not in _NP_DTYPE:
                arr = arr.astype(np.float32)
            if arr.ndim != 2:
                raise ValueError("vault rows must be 2-D")
            dt, ptr, n, dim, keep, on_dev = _NP_DTYPE[arr.dtype], arr.ctypes.data, arr.shape[0], arr.shape[1], arr, False
        self._check(self.lib.mmf_vault_load(self._h, C.c_void_p(ptr), int(on_dev), n, dim, dt, _MODE[mode], int(row_offset)))
        del keep
        self.vault_rows, self.vault_row_offset, self.vault_mode = n, int(row_offset), mode

    def vault_unload(self):
        self._check(self.lib.mmf_vault_unload(self._h))
        self.vault_rows, self.vault_mode = 0, None

    def vault_search(self, queries, top_k: int = 5, threshold: float = VAULT_THRESHOLD, algo: str = "auto"
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """queries (Q,512) un-normalised embeddings on the device -> (scores (Q,k) fp32,
        rows (Q,k) int64 global ids, discrepancy (Q,) fp32), all on the device, asynchronous
        on the current stream.  Slots beyond the vault size hold NaN / -1."""
        q = self._dev_f32(queries, 512)
        nq = q.shape[0]
        scores = torch.empty((nq, top_k), dtype=torch.float32, device=self.device)
        rows = torch.empty((nq, top_k), dtype=torch.int64, device=self.device)
        disc = torch.empty(nq, dtype=torch.float32, device=self.device)
        self._check(self.lib.mmf_vault_search(self._h, _ptr(q), nq, int(top_k), float(threshold), _ALGO[algo],
                                              _ptr(scores), _ptr(rows), _ptr(disc), self._stream()))
        return scores, rows, disc

    def vault_search_host(self, queries: np.ndarray, top_k: int = 5, threshold: float = VAULT_THRESHOLD,
                          algo: str = "auto"):
        """Host-buffer entry: numpy in, numpy out, copies and sync included."""
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, 512)
        nq = q.shape[0]
        scores = np.empty((nq, top_k), np.float32)
        rows = np.empty((nq, top_k), np.int64)
        disc = np.empty(nq, np.float32)
        self._check(self.lib.mmf_vault_search_host(self._h, q.ctypes.data, nq, int(top_k), float(threshold), _ALGO[algo],
                                                   scores.ctypes.data, rows.ctypes.data, disc.ctypes.data))
        return scores, rows, disc

    def vault_search_candidates(self, queries, top_k: int, algo: str = "auto") -> torch.Tensor:
        """Local top-k of this rank's shard as packed uint64 candidates (Q,k) (int64 storage)."""
        q = self._dev_f32(queries, 512)
        nq = q.shape[0]
        packed = torch.empty((nq, top_k), dtype=torch.int64, device=self.device)
        self._check(self.lib.mmf_vault_search_candidates(self._h, _ptr(q), nq, int(top_k), _ALGO[algo], _ptr(packed),
                                                         self._stream()))
        return packed

    def topk_merge(self, packed: torch.Tensor, top_k: int, threshold: float = VAULT_THRESHOLD):
        """packed (n_lists, Q, k_in) candidates -> global (scores, rows, discrepancy)."""
        p = packed.to(self.device).contiguous()
        n_lists, nq, k_in = p.shape
        scores = torch.empty((nq, top_k), dtype=torch.float32, device=self.device)
        rows = torch.empty((nq, top_k), dtype=torch.int64, device=self.device)
        disc = torch.empty(nq, dtype=torch.float32, device=self.device)
        self._check(self.lib.mmf_topk_merge(self._h, _ptr(p), n_lists, nq, k_in, int(top_k), float(threshold),
                                            _ptr(scores), _ptr(rows), _ptr(disc), self._stream()))
        return scores, rows, disc

    def verdict_assemble(self, head_scores, modality, clip_similarity: torch.Tensor, vault_discrepancy: torch.Tensor):
        """Score assembly (skipped modalities zeroed) + fusion judge / fallback verdict in one launch.  clip_similarity
        and vault_discrepancy are device fp32 tensors, masked IN PLACE.  Returns (scores5, probs, verdict, confidence)."""
        hs = self._dev_f32(head_scores, 3)
        n = hs.shape[0]
        mod = None if modality is None else torch.as_tensor(modality).to(device=self.device, dtype=torch.uint8).contiguous()
        x = torch.empty((n, 5), dtype=torch.float32, device=self.device)
        probs = torch.empty((n, 2), dtype=torch.float32, device=self.device)
        verdict = torch.empty(n, dtype=torch.int32, device=self.device)
        conf = torch.empty(n, dtype=torch.float32, device=self.device)
        self._check(self.lib.mmf_verdict_assemble(self._h, _ptr(hs), _ptr(mod), n, _ptr(clip_similarity), _ptr(vault_discrepancy),
                                                  _ptr(x), _ptr(probs), _ptr(verdict), _ptr(conf), self._stream()))
        return x, probs, verdict, conf

    def score_batch(self, text_embeds, image_embeds, head_scores, modality=None, top_k: int = 5,
                    threshold: float = VAULT_THRESHOLD, algo: str = "auto") -> dict:
        """The whole hot path for a batch in ONE asynchronous library call on device tensors (host tensors are copied
        in first): dict of device tensors, keys as score_batch_host.  Uses the vault resident in this engine (none:
        zero discrepancy, no matches)."""
        t, im, hs = self._dev_f32(text_embeds, 512), self._dev_f32(image_embeds, 512), self._dev_f32(head_scores, 3)
        n = im.shape[0]
        if t.shape[0] != n or hs.shape[0] != n:
            raise ValueError("score_batch: text / image / head batch sizes differ")
        mod = None if modality is None else torch.as_tensor(modality).to(device=self.device, dtype=torch.uint8).contiguous()
        dev = self.device
        out = {"clip_similarity": torch.empty(n, dtype=torch.float32, device=dev), "vault_discrepancy": torch.empty(n, dtype=torch.float32, device=dev),
               "vault_scores": torch.empty((n, top_k), dtype=torch.float32, device=dev), "vault_rows": torch.empty((n, top_k), dtype=torch.int64, device=dev),
               "scores": torch.empty((n, 5), dtype=torch.float32, device=dev), "probs": torch.empty((n, 2), dtype=torch.float32, device=dev),
               "verdict": torch.empty(n, dtype=torch.int32, device=dev), "confidence": torch.empty(n, dtype=torch.float32, device=dev)}
        self._check(self.lib.mmf_score_batch(self._h, _ptr(t), _ptr(im), _ptr(hs), _ptr(mod), n, int(top_k), float(threshold), _ALGO[algo],
                                             *[_ptr(out[k]) for k in self._BATCH_KEYS], self._stream()))
        return out

    _BATCH_KEYS = ("clip_similarity", "vault_discrepancy", "vault_scores", "vault_rows", "scores", "probs", "verdict", "confidence")

    @staticmethod
    def _host(x, cols, dt):
        a = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)   # a CPU tensor is not copied
        return np.ascontiguousarray(a, dtype=dt).reshape(-1, cols) if cols else np.ascontiguousarray(a, dtype=dt).reshape(-1)

    def score_batch_submit(self, slot: int, text_embeds, image_embeds, head_scores, modality=None, top_k: int = 5,
                           threshold: float = VAULT_THRESHOLD, algo: str = "auto") -> None:
        """Enqueue the whole hot path for one batch with host buffers (mmf_score_batch_submit): H2D, kernels and D2H run
        on the library's own streams, the call returns at once.  Two slots (0, 1): submit batch i+1 before collecting
        batch i and the copies of one overlap the kernels of the other.  The input arrays are kept alive until the
        slot is collected."""
        t, im, hs = self._host(text_embeds, 512, np.float32), self._host(image_embeds, 512, np.float32), self._host(head_scores, 3, np.float32)
        n = im.shape[0]
        if t.shape[0] != n or hs.shape[0] != n:
            raise ValueError("score_batch: text / image / head batch sizes differ")
        mod = None if modality is None else self._host(modality, 0, np.uint8)
        self._check(self.lib.mmf_score_batch_submit(self._h, int(slot), t.ctypes.data, im.ctypes.data, hs.ctypes.data,
                                                    None if mod is None else mod.ctypes.data, n, int(top_k), float(threshold), _ALGO[algo]))
        self._pending[int(slot)] = (t, im, hs, mod, n, int(top_k))

    def score_batch_collect(self, slot: int) -> dict:
        """Wait for the batch submitted in `slot` and return its results as numpy arrays (keys as score_batch_host)."""
        if int(slot) not in self._pending:
            raise MMFError(_lib.ERR_BAD_ARG, f"score_batch_collect: nothing submitted in slot {slot}")
        *_, n, top_k = self._pending.pop(int(slot))
        out = {"clip_similarity": np.empty(n, np.float32), "vault_discrepancy": np.empty(n, np.float32),
               "vault_scores": np.empty((n, top_k), np.float32), "vault_rows": np.empty((n, top_k), np.int64),
               "scores": np.empty((n, 5), np.float32), "probs": np.empty((n, 2), np.float32),
               "verdict": np.empty(n, np.int32), "confidence": np.empty(n, np.float32)}
        self._check(self.lib.mmf_score_batch_collect(self._h, int(slot), *[out[k].ctypes.data for k in self._BATCH_KEYS]))
        return out

    def score_batch_host(self, text_embeds, image_embeds, head_scores, modality=None, top_k: int = 5,
                         threshold: float = VAULT_THRESHOLD, algo: str = "auto"):
        """The whole hot path for a batch in ONE library call with host buffers: (B,512) text / image embeddings
        and (B,3) head scores as numpy arrays or CPU tensors (pinned memory makes the copies asynchronous DMA),
        results as numpy arrays in host memory -- one D2H and one synchronisation instead of one per tensor.
        Same values as mmf_b200.score_batch.  Needs the fusion weights; a vault is optional."""
        self.score_batch_submit(0, text_embeds, image_embeds, head_scores, modality, top_k, threshold, algo)
        return self.score_batch_collect(0)

    def score_stream(self, batches, top_k: int = 5, threshold: float = VAULT_THRESHOLD, algo: str = "auto"):
        """Generator over an iterable of (text, image, head[, modality]) host batches: keeps two batches in flight
        (the copies of one overlap the kernels of the other) and yields one result dict per batch, in order."""
        it, slot, pending = iter(batches), 0, []
        for b in it:
            self.score_batch_submit(slot, *b, top_k=top_k, threshold=threshold, algo=algo)
            pending.append(slot)
            slot ^= 1
            if len(pending) == 2:
                yield self.score_batch_collect(pending.pop(0))
        while pending:
            yield self.score_batch_collect(pending.pop(0))

    # ------------------------------------------------------------------ switches (A/B, triage)
    def set_option(self, name: str, value: int) -> None:
        self._check(self.lib.mmf_set_option(self._h, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        v = C.c_int()
        self._check(self.lib.mmf_get_option(self._h, name.encode(), C.byref(v)))
        return int(v.value)

    # ------------------------------------------------------------------ row-sharded search, library-owned NCCL (csrc/shard.cu)
    def shard_unique_id(self) -> bytes:
        """128 bytes that rank 0 hands to its peers (any out-of-band channel) before shard_init."""
        buf = C.create_string_buffer(_lib.SHARD_ID_BYTES)
        rc = self.lib.mmf_shard_unique_id(buf)
        if rc != _lib.OK:
            raise MMFError(rc, "mmf_shard_unique_id: NCCL library not found (set MMF_NCCL_LIB)")
        return buf.raw

    def shard_init(self, rank: int, world: int, unique_id: Optional[bytes] = None) -> None:
        """Collective: every rank of the shard group calls this with rank 0's unique id (world == 1 needs none)."""
        if world > 1 and (unique_id is None or len(unique_id) != _lib.SHARD_ID_BYTES):
            raise ValueError("shard_init: unique_id must be the %d bytes of rank 0's shard_unique_id()" % _lib.SHARD_ID_BYTES)
        self._check(self.lib.mmf_shard_init(self._h, int(rank), int(world), unique_id))

    def shard_finalize(self) -> None:
        self._check(self.lib.mmf_shard_finalize(self._h))

    def shard_info(self) -> Tuple[int, int, int]:
        r, w, v = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.mmf_shard_info(self._h, C.byref(r), C.byref(w), C.byref(v)))
        return int(r.value), int(w.value), int(v.value)

    def vault_search_sharded(self, queries, top_k: int = 5, threshold: float = VAULT_THRESHOLD, algo: str = "auto"
                             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """vault_search over ALL shards of the group: local search + ONE ncclAllGather of the packed candidates +
        merge, all inside the library on the current stream.  Same outputs on every rank."""
        q = self._dev_f32(queries, 512)
        nq = q.shape[0]
        scores = torch.empty((nq, top_k), dtype=torch.float32, device=self.device)
        rows = torch.empty((nq, top_k), dtype=torch.int64, device=self.device)
        disc = torch.empty(nq, dtype=torch.float32, device=self.device)
        self._check(self.lib.mmf_vault_search_sharded(self._h, _ptr(q), nq, int(top_k), float(threshold), _ALGO[algo],
                                                      _ptr(scores), _ptr(rows), _ptr(disc), self._stream()))
        return scores, rows, disc

    def shard_all_gather(self, packed: torch.Tensor, world: int) -> torch.Tensor:
        """The collective alone (for callers that time the phases): (Q,k) packed candidates -> (world, Q, k)."""
        p = packed.contiguous()
        out = torch.empty((world,) + tuple(p.shape), dtype=p.dtype, device=self.device)
        self._check(self.lib.mmf_shard_all_gather(self._h, _ptr(p), p.numel(), _ptr(out), self._stream()))
        return out

    @property
    def collective_count(self) -> int:
        return int(self.lib.mmf_collective_count(self._h))

    # ------------------------------------------------------------------ peer-memory candidate exchange
    def exchange_layout(self, world: int, n_queries: int, k_in: int) -> int:
        """Bytes of symmetric memory per rank that an exchange of (n_queries, k_in) candidates needs."""
        need = C.c_int64()
        self._check(self.lib.mmf_exchange_layout(int(world), int(n_queries), int(k_in), None, C.byref(need)))
        return int(need.value)

    def exchange_attach(self, rank: int, world: int, peer_ptrs, bytes_per_rank: int) -> None:
        """peer_ptrs[r]: device address of rank r's symmetric buffer as mapped in this process."""
        arr = (C.c_uint64 * int(world))(*[int(x) for x in peer_ptrs])
        self._check(self.lib.mmf_exchange_attach(self._h, int(rank), int(world), arr, int(bytes_per_rank)))

    def exchange_detach(self) -> None:
        self._check(self.lib.mmf_exchange_detach(self._h))

    def vault_search_exchange(self, queries, top_k: int, k_local: int, threshold: float = VAULT_THRESHOLD,
                              algo: str = "auto") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Row-sharded search with the candidate exchange over NVLink peer memory (csrc/exchange.cu): local
        search + push into the peers' buffers + wait + merge, no NCCL call.  Outputs as vault_search."""
        q = self._dev_f32(queries, 512)
        nq = q.shape[0]
        scores = torch.empty((nq, top_k), dtype=torch.float32, device=self.device)
        rows = torch.empty((nq, top_k), dtype=torch.int64, device=self.device)
        disc = torch.empty(nq, dtype=torch.float32, device=self.device)
        self._check(self.lib.mmf_vault_search_exchange(self._h, _ptr(q), nq, int(top_k), int(k_local), float(threshold),
                                                       _ALGO[algo], _ptr(scores), _ptr(rows), _ptr(disc), self._stream()))
        return scores, rows, disc

    # ------------------------------------------------------------------ fusion judge
    def fusion_load(self, state_dict) -> None:
        """state_dict with keys 0.weight,0.bias,3.weight,3.bias,5.weight,5.bias (optionally under
        the 'fusion_layer.' prefix) -- the .pth layouts of train_fusion_judge.py:259-267."""
        keys = ("0.weight", "0.bias", "3.weight", "3.bias", "5.weight", "5.bias")
        shapes = ((64, 5), (64,), (32, 64), (32,), (2, 32), (2,))
        pre = "" if "0.weight" in state_dict else "fusion_layer."
        parts = []
        for k, shp in zip(keys, shapes):
            t = torch.as_tensor(state_dict[pre + k]).detach().to("cpu", torch.float32)
            if tuple(t.shape) != shp:
                raise ValueError(f"fusion weight {k}: shape {tuple(t.shape)} != {shp}")
            parts.append(t.reshape(-1))
        blob = torch.cat(parts).contiguous().numpy()
        assert blob.size == _lib.FUSION_PARAMS
        self._check(self.lib.mmf_fusion_load(self._h, blob.ctypes.data))

    def fusion_forward(self, x):
        """x (n,5) -> (probs (n,2) [real,fake], verdict (n,) int32, confidence (n,))."""
        x_t = self._dev_f32(x, 5)
        n = x_t.shape[0]
        probs = torch.empty((n, 2), dtype=torch.float32, device=self.device)
        verdict = torch.empty(n, dtype=torch.int32, device=self.device)
        conf = torch.empty(n, dtype=torch.float32, device=self.device)
        self._check(self.lib.mmf_fusion_forward(self._h, _ptr(x_t), n, _ptr(probs), _ptr(verdict), _ptr(conf), self._stream()))
        return probs, verdict, conf

    def verdict_batch(self, scores, modality):
        """scores (n,5), modality (n,) uint8 (bit0 text, bit1 visual): fusion judge where both
        modalities are present, the reference's fallback rule elsewhere."""
        x_t = self._dev_f32(scores, 5)
        m_t = torch.as_tensor(modality).to(device=self.device, dtype=torch.uint8).contiguous()
        n = x_t.shape[0]
        probs = torch.empty((n, 2), dtype=torch.float32, device=self.device)
        verdict = torch.empty(n, dtype=torch.int32, device=self.device)
        conf = torch.empty(n, dtype=torch.float32, device=self.device)
        self._check(self.lib.mmf_verdict_batch(self._h, _ptr(x_t), _ptr(m_t), n, _ptr(probs), _ptr(verdict), _ptr(conf), self._stream()))
        return probs, verdict, conf
