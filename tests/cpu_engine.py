"""TEST DOUBLE -- an Engine look-alike whose arithmetic is the CPU oracle.

It exists so that the HOST logic of the drop-in layer (mmf_b200.forensics / clip_similarity_engine / pipeline /
vault: argument handling, modality rules, match records, video aggregation, return schemas, error behaviour) can
be exercised against the reference fixtures in the CPU test-suite.  It lives in tests/ and is never imported by the
package; the product Engine has no CPU path (tests/test_host_cpu.py::test_no_cpu_fallback)."""
import numpy as np
import torch

import oracle


class OracleEngine:
    def __init__(self):
        self.device = torch.device("cpu")
        self.launch_count = 0
        self._vault = None
        self._row_offset = 0
        self._weights = None
        self.vault_rows = 0
        self.vault_mode = None

    @staticmethod
    def _np(x, cols):
        t = torch.as_tensor(x).detach().to("cpu", torch.float32)
        return np.ascontiguousarray(t.reshape(-1, cols).numpy())

    def cosine_pairs(self, a, b, match_threshold=None):
        dim = torch.as_tensor(a).shape[-1]
        sim = torch.from_numpy(np.asarray(oracle.cosine_pairs(self._np(a, dim), self._np(b, dim)), np.float32))
        if match_threshold is None:
            return sim
        return sim, torch.from_numpy((sim.numpy().astype(np.float64) >= float(match_threshold)).astype(np.uint8))

    def vault_load(self, rows, mode="fp32", row_offset=0):
        arr = rows.detach().cpu().numpy() if isinstance(rows, torch.Tensor) else np.asarray(rows)
        if arr.ndim != 2 or arr.shape[1] != 512:
            raise ValueError("vault rows must be (n, 512)")
        self._vault = oracle.vault_normalise(arr.astype(np.float32)).astype(np.float32)   # the library normalises in fp32
        self._row_offset, self.vault_rows, self.vault_mode = int(row_offset), arr.shape[0], mode

    def article_index(self, embeddings_db, key):
        """search_similar_articles keeps the article database in a PRIVATE handle (similar_articles._article_index);
        the test double's equivalent: a second OracleEngine, so that this one's resident vault stays untouched."""
        idx = OracleEngine()
        idx.vault_load(np.asarray(embeddings_db[key]), mode="fp32")
        return idx

    def vault_unload(self):
        self._vault, self.vault_rows, self.vault_mode = None, 0, None

    def vault_search(self, queries, top_k=5, threshold=oracle.VAULT_THRESHOLD, algo="auto"):
        q = self._np(queries, 512)
        idx, sc, _ = oracle.vault_search_batched(self._vault, q, top_k, vault_is_normalised=True, row_offset=self._row_offset)
        nq, kk = idx.shape
        rows = np.full((nq, top_k), -1, np.int64)
        scores = np.full((nq, top_k), np.nan, np.float32)
        rows[:, :kk], scores[:, :kk] = idx, sc
        disc = oracle.discrepancy_rule(sc[:, 0], threshold) if kk else np.zeros(nq, np.float32)
        return torch.from_numpy(scores), torch.from_numpy(rows), torch.from_numpy(disc)

    def fusion_load(self, state_dict):
        pre = "" if "0.weight" in state_dict else "fusion_layer."
        self._weights = {k: torch.as_tensor(state_dict[pre + k]).detach().float().cpu() for k in oracle.FUSION_KEYS}

    def fusion_forward(self, x):
        p = np.asarray(oracle.fusion_forward(self._weights, self._np(x, 5)), np.float32)
        verdict = (p[:, 1] > 0.5).astype(np.int32)
        conf = np.where(verdict == 1, p[:, 1], p[:, 0]).astype(np.float32)
        return torch.from_numpy(p), torch.from_numpy(verdict), torch.from_numpy(conf)

    def verdict_batch(self, scores, modality):
        x = self._np(scores, 5)
        mod = torch.as_tensor(modality).to(torch.uint8).numpy()
        p, verdict, conf = (t.numpy().copy() for t in self.fusion_forward(x))
        for i in np.nonzero(mod != 3)[0]:
            d = oracle.fallback_verdict(dict(zip(oracle.FUSION_ORDER, map(float, x[i]))), bool(mod[i] & 1), bool(mod[i] & 2))
            p[i] = (d["real_probability"], d["fake_probability"])
            verdict[i], conf[i] = d["verdict"], d["confidence"]
        return torch.from_numpy(p), torch.from_numpy(verdict), torch.from_numpy(conf)

    # ---- row-sharded search: packed candidates + merge (csrc/topk.cuh: key = order-preserving score << 32 | row)
    def vault_search_candidates(self, queries, top_k, algo="auto"):
        q = self._np(queries, 512)
        idx, sc, _ = oracle.vault_search_batched(self._vault, q, top_k, vault_is_normalised=True, row_offset=self._row_offset)
        packed = np.zeros((q.shape[0], top_k), np.uint64)                 # 0 = empty slot
        packed[:, :idx.shape[1]] = oracle.order_key64(sc, idx)
        return torch.from_numpy(packed.view(np.int64))

    def topk_merge(self, packed, top_k, threshold=oracle.VAULT_THRESHOLD):
        keys = np.ascontiguousarray(packed.cpu().numpy()).view(np.uint64)              # (n_lists, Q, k_in)
        n_lists, nq, k_in = keys.shape
        flat = np.sort(keys.transpose(1, 0, 2).reshape(nq, n_lists * k_in), axis=1)[:, ::-1][:, :top_k]
        if flat.shape[1] < top_k:
            flat = np.concatenate([flat, np.zeros((nq, top_k - flat.shape[1]), np.uint64)], axis=1)
        u = (flat >> np.uint64(32)).astype(np.uint32)
        bits = np.where(u & np.uint32(0x80000000), u & np.uint32(0x7FFFFFFF), ~u).astype(np.uint32)
        scores = np.where(flat != 0, bits.view(np.float32), np.float32(np.nan)).astype(np.float32)
        rows = np.where(flat != 0, (flat & np.uint64(0xFFFFFFFF)).astype(np.int64), -1)
        disc = np.where(flat[:, 0] != 0, oracle.discrepancy_rule(np.nan_to_num(scores[:, 0], nan=0.0), threshold), 0).astype(np.float32)
        return torch.from_numpy(scores), torch.from_numpy(rows), torch.from_numpy(disc)

    # ---- the fused tail and the one-call entries (Engine.verdict_assemble / score_batch / score_batch_host / submit+collect)
    def verdict_assemble(self, head_scores, modality, clip_similarity, vault_discrepancy):
        hs = torch.as_tensor(head_scores).detach().to("cpu", torch.float32).reshape(-1, 3)
        n = hs.shape[0]
        mod = torch.full((n,), 3, dtype=torch.uint8) if modality is None else torch.as_tensor(modality).to(torch.uint8)
        has_text, has_vis = (mod & 1).bool(), (mod & 2).bool()
        zero = torch.zeros(n)
        clip_similarity.copy_(torch.where(has_text & has_vis, clip_similarity, zero))      # masked in place, like the kernel
        vault_discrepancy.copy_(torch.where(has_vis, vault_discrepancy, zero))
        x = torch.stack([torch.where(has_text, hs[:, 0], zero), torch.where(has_text, hs[:, 1], zero),
                         torch.where(has_vis, hs[:, 2], zero), clip_similarity, vault_discrepancy], dim=1)
        probs, verdict, conf = self.verdict_batch(x, mod)
        return x, probs, verdict, conf

    def score_batch(self, text_embeds, image_embeds, head_scores, modality=None, top_k=5, threshold=oracle.VAULT_THRESHOLD, algo="auto"):
        sim = self.cosine_pairs(text_embeds, image_embeds)
        n = sim.shape[0]
        if self._vault is not None:
            vs, vr, disc = self.vault_search(image_embeds, top_k, threshold, algo)
        else:
            vs, vr, disc = torch.full((n, top_k), float("nan")), torch.full((n, top_k), -1, dtype=torch.int64), torch.zeros(n)
        x, probs, verdict, conf = self.verdict_assemble(head_scores, modality, sim, disc)
        return {"clip_similarity": sim, "vault_discrepancy": disc, "vault_scores": vs, "vault_rows": vr, "scores": x,
                "probs": probs, "verdict": verdict, "confidence": conf}

    def score_batch_host(self, text_embeds, image_embeds, head_scores, modality=None, top_k=5,
                         threshold=oracle.VAULT_THRESHOLD, algo="auto"):
        out = self.score_batch(torch.as_tensor(text_embeds), torch.as_tensor(image_embeds), torch.as_tensor(head_scores),
                               modality, top_k, threshold, algo)
        return {k: v.numpy() for k, v in out.items()}

    def score_batch_submit(self, slot, *a, **k):
        if not hasattr(self, "_slots"):
            self._slots = {}
        if slot in self._slots:
            raise RuntimeError(f"slot {slot} has not been collected")
        self._slots[slot] = self.score_batch_host(*a, **k)

    def score_batch_collect(self, slot):
        return self._slots.pop(slot)

    def vault_search_host(self, queries, top_k=5, threshold=oracle.VAULT_THRESHOLD, algo="auto"):
        return tuple(t.numpy() for t in self.vault_search(torch.as_tensor(queries), top_k, threshold, algo))

    def get_option(self, name):
        return {"screen": 1}.get(name, 0)

    def set_option(self, name, value):
        pass

    def close(self):
        pass

    collective_count = 0
