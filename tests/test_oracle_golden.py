"""Pins the oracle (oracle/reference_port.py) to the fixtures produced by the reference's
own methods (tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN

FP32_TOL = 1e-5      # north_star: 1e-5 relative, absolute floor 1e-5 (cosines are O(1)-scaled)


def close(a, b, tol=FP32_TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b)))


def test_cosine_matches_reference():
    g = np.load(os.path.join(GOLDEN, "cosine.npz"))
    for loop in (True, False):
        s = oracle.cosine_pairs(g["text"], g["image"], scalar_loop=loop)
        assert close(s, g["clip_similarity"], 2e-6)
        assert close(s, g["engine_similarity"], 2e-6)
    s = oracle.cosine_pairs(g["text"], g["image"], scalar_loop=True)
    # the literal replay reproduces the fixture to the last ulp or two (same ops; the BLAS
    # kernel picked for the 512-long dot may differ between hosts)
    assert np.max(np.abs(s.astype(np.float64) - g["clip_similarity"])) <= 2.5e-7
    lab = np.array([oracle.clip_match_label(float(x)) == "Match" for x in g["engine_similarity"]])
    assert np.array_equal(lab, g["engine_match"])


@pytest.mark.parametrize("k", [5, 10])
def test_vault_matches_reference(k):
    g = np.load(os.path.join(GOLDEN, "vault.npz"))
    idx, sc, disc = oracle.vault_search_batched(g["vault"], g["queries"], k)
    gi, gs = g[f"idx_k{k}"], g[f"sim_k{k}"]
    assert close(sc, gs, 2e-6)
    assert np.array_equal(idx, gi)          # no near-ties in the fixture
    assert close(disc, g[f"disc_k{k}"], 2e-6)
    assert np.array_equal(disc > 0, g[f"disc_k{k}"] > 0)
    # as-shipped scalar replay
    for r in range(0, len(gi), 7):
        i1, s1, d1 = oracle.vault_search_as_shipped(g["vault"], g["queries"][r], k)
        assert np.array_equal(i1, gi[r]) and close(s1, gs[r], 1e-6) and close(d1, g[f"disc_k{k}"][r], 1e-6)


def test_vault_text_similarity():
    g = np.load(os.path.join(GOLDEN, "vault.npz"))
    disc, tsim = g["disc_k5"], g["tsim_k5"]
    want = np.where(disc > 0, oracle.cosine_pairs(g["text_table"], g["title_table"]), 0.0)
    assert close(want, tsim, 2e-6)
    assert np.array_equal(tsim != 0, disc > 0)


def test_vault_fp16_small_nan_notloaded():
    g = np.load(os.path.join(GOLDEN, "vault.npz"))
    idx, sc, disc = oracle.vault_search_batched(g["vault"].astype(np.float16), g["queries"], 5)
    assert close(sc, g["sim_f16_k5"], 2e-6) and close(disc, g["disc_f16_k5"], 2e-6)
    # fp16-vault reference vs fp32 maths: the documented <=1e-2 band
    i32, s32, _ = oracle.vault_search_batched(g["vault"], g["queries"], 5)
    assert close(s32, g["sim_f16_k5"], 1e-2)
    idx, sc, disc = oracle.vault_search_batched(g["vault"][:3], g["queries"][:8], 5)
    assert idx.shape == (8, 3) and np.all(g["small_nmatch"] == 3)
    assert np.array_equal(idx, g["small_idx"][:, :3]) and close(sc, g["small_sim"][:, :3], 2e-6)
    zv = g["vault"][:50].copy()
    zv[7] = 0
    with np.errstate(all="ignore"):
        idx, sc, disc = oracle.vault_search_batched(zv, g["queries"][:4], 5)
    assert np.all(idx[:, 0] == 7) and np.all(np.isnan(sc[:, 0])) and np.all(disc == 0)
    assert np.array_equal(idx, g["nan_idx"]) and np.all(g["nan_disc"] == 0)
    nl = json.loads(str(g["not_loaded"]))
    assert nl == {"vault_discrepancy": 0.0, "matches": [], "vault_available": False, "text_similarity": 0.0}


def test_order_and_merge_are_shard_invariant():
    g = np.load(os.path.join(GOLDEN, "vault.npz"))
    v, q = g["vault"], g["queries"]
    v[100] = v[50]
    v[200] = v[50]           # exact ties -> higher index first
    q[5] = v[50] * 2
    full_i, full_s, _ = oracle.vault_search_batched(v, q, 10)
    assert list(full_i[5, :3]) == [200, 100, 50]
    for shards in (2, 3, 8):
        bounds = np.linspace(0, len(v), shards + 1).astype(int)
        parts = [oracle.vault_search_batched(v[a:b], q, 10, row_offset=a) for a, b in zip(bounds[:-1], bounds[1:])]
        mi, ms = oracle.merge_topk([p[0] for p in parts], [p[1] for p in parts], 10)
        assert np.array_equal(mi, full_i) and np.array_equal(ms, full_s)
    # stable-argsort restatement of :449 agrees with the key order
    vn = oracle.vault_normalise(v)
    qn = oracle.normalise_rows(torch.from_numpy(q)).numpy()
    for r in (0, 5, 9):
        s = vn @ qn[r]
        assert np.array_equal(np.argsort(s, kind="stable")[-10:][::-1], full_i[r])


def test_discrepancy_threshold_edge():
    f85 = np.float32(0.85)                       # 0x3F59999A > 0.85 in double
    below = np.nextafter(f85, np.float32(0))
    d = oracle.discrepancy_rule(np.array([f85, below, np.nan, 1.0, 0.0], np.float32))
    assert d[0] == f85 and d[1] == 0 and d[2] == 0 and d[3] == 1.0 and d[4] == 0


def test_fusion_matches_reference():
    g = np.load(os.path.join(GOLDEN, "fusion.npz"))
    for tag, pre in (("", "w_"), ("2", "w2_")):
        w = {k: torch.from_numpy(g[pre + k]) for k in oracle.FUSION_KEYS}
        p = oracle.fusion_forward(w, g["x"])
        assert close(p[:, 0], g["real" + tag], 1e-6) and close(p[:, 1], g["fake" + tag], 1e-6)
        for r in range(0, len(p), 37):
            v = oracle.fusion_verdict(w, dict(zip(oracle.FUSION_ORDER, map(float, g["x"][r]))))
            assert v["verdict"] == g["verdict" + tag][r]
            assert close(v["confidence"], g["confidence" + tag][r], 1e-6)
    ck = {"fusion_layer_state_dict": {k: torch.from_numpy(g["w_" + k]) for k in oracle.FUSION_KEYS}}
    ck2 = {"full_model_state_dict": {"fusion_layer." + k: v for k, v in ck["fusion_layer_state_dict"].items()}}
    a, b = oracle.fusion_weights_from_checkpoint(ck), oracle.fusion_weights_from_checkpoint(ck2)
    assert all(torch.equal(a[k], b[k]) for k in oracle.FUSION_KEYS)


def test_analyze_assembly_matches_reference():
    with open(os.path.join(GOLDEN, "analyze_cases.json")) as fh:
        cases = json.load(fh)
    import fakes
    w = {k: v.detach() for k, v in fakes.FakeDetector([.5], [.5], [.5], fusion_seed=5).fusion_layer.state_dict().items()}
    for c in cases["cases"]:
        sc = c["result"]["scores"]
        v = oracle.assemble_verdict(w, sc, c["mode"] in ("both", "text"), c["mode"] in ("both", "image"))
        assert v["verdict"] == c["result"]["verdict"]
        assert close(v["confidence"], c["result"]["confidence"], 1e-6)
        assert close(v["fake_probability"], sc["fake_probability"], 1e-6)
    for v in cases["videos"]:
        assert v["result"]["scores"]["vault_discrepancy"] == v["video"]["vault_discrepancy"]


def test_similar_articles_matches_reference():
    """search_similar_articles (train_clip_detective.py:610-688): fixtures produced by the
    reference's own function (tests/golden/make_golden.py: gen_similar)"""
    import json
    g = np.load(os.path.join(GOLDEN, "similar.npz"))
    with open(os.path.join(GOLDEN, "similar_cases.json")) as fh:
        c = json.load(fh)
    k = c["top_k"]
    for mode, db, queries in (("text", g["text_embeddings"], g["text_queries"]), ("image", g["image_embeddings"], g["image_queries"]),
                              ("text_f16", g["text_embeddings"].astype(np.float16), g["text_queries"])):
        for i, want in enumerate(c["results"][mode]):
            idx, sims = oracle.similar_articles_as_shipped(db, queries[i], k)
            assert [c["article_ids"][j] for j in idx] == [r["article_id"] for r in want], (mode, i)
            assert np.allclose(sims, [r["similarity"] for r in want], atol=5e-7 if mode != "text_f16" else 2e-3), (mode, i)
            assert [r["rank"] for r in want] == list(range(1, k + 1))
            assert want[0]["text"] == c["text_contents"][idx[0]][:100] + "..." and want[0]["image_path"] == c["image_paths"][idx[0]]
