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

    def score_batch_host(self, text_embeds, image_embeds, head_scores, modality=None, top_k: int = 5,
                         threshold: float = VAULT_THRESHOLD, algo: str = "auto"):
        """The whole hot path for a batch in ONE library call with host buffers: (B,512) text / image embeddings
        and (B,3) head scores as numpy arrays or CPU tensors (pinned memory makes the copies asynchronous DMA),
        results as numpy arrays in host memory -- one D2H and one synchronisation instead of one per tensor.
        Same values as mmf_b200.score_batch.  Needs the fusion weights; a vault is optional."""
        def host(x, cols, dt):
            a = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)   # a CPU tensor is not copied
            a = np.ascontiguousarray(a, dtype=dt).reshape(-1, cols) if cols else np.ascontiguousarray(a, dtype=dt).reshape(-1)
            return a
        t, im, hs = host(text_embeds, 512, np.float32), host(image_embeds, 512, np.float32), host(head_scores, 3, np.float32)
        n = im.shape[0]
        if t.shape[0] != n or hs.shape[0] != n:
            raise ValueError("score_batch_host: text / image / head batch sizes differ")
        mod = None if modality is None else host(modality, 0, np.uint8)
        out = {"clip_similarity": np.empty(n, np.float32), "vault_discrepancy": np.empty(n, np.float32),
               "vault_scores": np.empty((n, top_k), np.float32), "vault_rows": np.empty((n, top_k), np.int64),
               "scores": np.empty((n, 5), np.float32), "probs": np.empty((n, 2), np.float32),
               "verdict": np.empty(n, np.int32), "confidence": np.empty(n, np.float32)}
        self._check(self.lib.mmf_score_batch_host(
            self._h, t.ctypes.data, im.ctypes.data, hs.ctypes.data, None if mod is None else mod.ctypes.data, n, int(top_k),
            float(threshold), _ALGO[algo], *[out[k].ctypes.data for k in ("clip_similarity", "vault_discrepancy", "vault_scores",
                                                                           "vault_rows", "scores", "probs", "verdict", "confidence")]))
        return out

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
