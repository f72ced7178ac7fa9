"""Host logic of the drop-in layer on the CPU: mmf_b200.MisinfoForensics / CLIPSimilarityEngine driven through the
OracleEngine test double (tests/cpu_engine.py) must reproduce the fixtures that the REFERENCE'S OWN methods produced
(tests/golden, SURVEY.md 8b: same signatures, return schemas, error behaviour).  The arithmetic here is the oracle's;
the CUDA arithmetic is checked by the `-m gpu` suite with the same fixtures and the same comparison helpers."""
import json
import os
import sys

import numpy as np
import pytest
import torch

import fakes
from conftest import GOLDEN
from cpu_engine import OracleEngine
from test_gpu_parity import _cmp_result, _load_analyze_fixture
from util import FP32_TOL

import mmf_b200


def _forensics(g, cases, **kw):
    det = fakes.FakeDetector(g["ai"], g["misinfo"], g["deepfake"], fusion_seed=5)
    return mmf_b200.MisinfoForensics(
        detector=det, roberta_tokenizer=fakes.FakeTokenizer(), clip_model=fakes.FakeClipModel(g["image_table"], g["text_table"]),
        clip_processor=fakes.FakeClipProcessor(), vault={"embeddings": g["vault"], "metadata": cases["metadata"]},
        engine=OracleEngine(), **kw)


def test_analyze_host_logic_matches_reference(capsys):
    g, cases = _load_analyze_fixture()
    f = _forensics(g, cases)
    assert f.vault_loaded and len(f.vault_metadata) == len(cases["metadata"])
    for c in cases["cases"]:
        i = c["id"]
        text = fakes.text_for_id(i) if c["mode"] in ("both", "text") else None
        image = fakes.image_for_id(i) if c["mode"] in ("both", "image") else None
        _cmp_result(f.analyze(text=text, image_path=image, verbose=(i % 7 == 0)), c["result"], f"case {i} ({c['mode']})")
    assert capsys.readouterr().out                                        # verbose=True prints the per-step scores
    with pytest.raises(ValueError, match="Provide at least one of"):
        f.analyze(verbose=False)
    sv = f.search_vault(fakes.image_for_id(0), user_caption=fakes.text_for_id(0), top_k=3)
    assert set(sv) == {"vault_discrepancy", "matches", "vault_available", "text_similarity"} and len(sv["matches"]) == 3
    assert set(sv["matches"][0]) == {"similarity", "title", "url", "date"}
    assert set(f.fusion_verdict({"ai_score": 0.3})) == {"verdict", "confidence", "fake_probability", "real_probability"}
    assert set(f.analyze_consistency(fakes.text_for_id(1), fakes.image_for_id(1))) == {"clip_similarity"}


def test_analyze_batch_score_matrix_and_dedup_host_logic():
    g, cases = _load_analyze_fixture()
    f = _forensics(g, cases)
    texts = [fakes.text_for_id(c["id"]) if c["mode"] in ("both", "text") else None for c in cases["cases"]]
    images = [fakes.image_for_id(c["id"]) if c["mode"] in ("both", "image") else None for c in cases["cases"]]
    for got, c in zip(f.analyze_batch(texts, images), cases["cases"]):
        _cmp_result(got, c["result"], f"batch case {c['id']} ({c['mode']})")
    assert f.score_matrix(texts, images).shape == (len(texts), 5)
    f2 = _forensics(g, cases, dedup_clip_encode=False)
    for c in cases["cases"][:6]:
        assert f.analyze(text=texts[c["id"]], image_path=images[c["id"]], verbose=False) == \
               f2.analyze(text=texts[c["id"]], image_path=images[c["id"]], verbose=False)
    with pytest.raises(ValueError):
        f.analyze_batch([None], [None])


def test_video_aggregation_host_logic(monkeypatch):
    g, cases = _load_analyze_fixture()
    f = _forensics(g, cases)
    monkeypatch.setitem(sys.modules, "cv2", fakes.FakeCv2)
    for v in cases["videos"]:
        got = f.analyze_video(v["path"], text=v["text"], max_frames=12, stride_seconds=1.0)
        got.pop("best_frame")
        for key in ("deepfake_score", "clip_similarity", "vault_discrepancy", "text_similarity"):
            assert abs(got[key] - v["video"][key]) <= FP32_TOL, (v["path"], key)
        assert [m["title"] for m in got["vault_matches"]] == [m["title"] for m in v["video"]["vault_matches"]]
        _cmp_result(f.analyze(text=v["text"], video_path=v["path"], verbose=False), v["result"], v["path"])
    with pytest.raises(RuntimeError, match="Could not open video"):
        f.analyze_video("nope.mp4")


def test_vault_missing_and_builder_pickle_host_logic(tmp_path):
    import pickle
    g, cases = _load_analyze_fixture()
    det = fakes.FakeDetector(g["ai"], g["misinfo"], g["deepfake"], fusion_seed=5)
    common = dict(detector=det, roberta_tokenizer=fakes.FakeTokenizer(), clip_processor=fakes.FakeClipProcessor(),
                  clip_model=fakes.FakeClipModel(g["image_table"], g["text_table"]))
    f = mmf_b200.MisinfoForensics(faiss_index_path=str(tmp_path / "missing.pkl"), engine=OracleEngine(), **common)
    assert f.vault_loaded is False
    assert f.search_vault(fakes.image_for_id(0)) == {"vault_discrepancy": 0.0, "matches": [], "vault_available": False,
                                                    "text_similarity": 0.0}
    r = f.analyze(text=fakes.text_for_id(1), image_path=fakes.image_for_id(1), verbose=False)
    assert r["scores"]["vault_discrepancy"] == 0.0 and r["vault_matches"] == []
    n = 300
    p = tmp_path / "guardian_embeddings.pkl"
    with open(p, "wb") as fh:
        pickle.dump({"article_ids": [str(i) for i in range(n)], "text_contents": [fakes.text_for_id(i) for i in range(n)],
                     "image_paths": [f"img/{i}.jpg" for i in range(n)], "image_embeddings": g["vault"][:n].astype(np.float16),
                     "text_embeddings": g["vault"][:n].astype(np.float16), "metadata": {"total_articles": n}}, fh)
    f2 = mmf_b200.MisinfoForensics(faiss_index_path=str(p), engine=OracleEngine(), **common)
    assert f2.vault_loaded and len(f2.vault_metadata) == n and f2.vault_embeddings.dtype == np.float16
    sv = f2.search_vault(fakes.image_for_id(0))
    assert sv["vault_available"] and sv["matches"][0]["date"] == "N/A" and sv["matches"][0]["url"].startswith("img/")


def test_clip_similarity_engine_host_logic(tmp_path):
    g = np.load(os.path.join(GOLDEN, "cosine.npz"))
    e = mmf_b200.CLIPSimilarityEngine(model=fakes.FakeClipModel(g["image"], g["text"]), processor=fakes.FakeClipProcessor(),
                                      threshold=0.25, engine=OracleEngine())
    paths = []
    for i in range(12):
        p = tmp_path / f"{i}.png"
        fakes.image_for_id(i).save(p)
        paths.append(str(p))
    for i, p in enumerate(paths):
        sim, label = e.calculate_similarity(p, fakes.text_for_id(i))
        assert abs(sim - g["engine_similarity"][i]) <= FP32_TOL
        if abs(g["engine_similarity"][i] - 0.25) > FP32_TOL:
            assert (label == "Match") == bool(g["engine_match"][i])
        out = e.analyze_with_explanation(p, fakes.text_for_id(i))
        assert out["label"] == label and out["similarity_score"] == round(sim, 4)
        assert out["explanation"].split("(")[0] == str(g["engine_explanation"][i]).split("(")[0]
    with pytest.raises(FileNotFoundError):
        e.calculate_similarity(str(tmp_path / "missing.png"), "caption #1")
    with pytest.raises(ValueError):
        e.calculate_similarity(paths[0], "")
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not an image")
    with pytest.raises(ValueError):
        e.calculate_similarity(str(bad), "caption #1")
    assert "error" in e.analyze_with_explanation(str(bad), "caption #1")
    json.dumps(e.analyze_with_explanation(paths[0], fakes.text_for_id(0)))


def test_score_batch_host_logic_equals_scalar_fixtures():
    """pipeline.score_batch (modality masks, score assembly, verdict rule) per row == the reference's scalar analyze"""
    g, cases = _load_analyze_fixture()
    f = _forensics(g, cases)
    ns = len(g["ai"])
    head = np.zeros((ns, 3), np.float32)
    for c in cases["cases"]:
        sc = c["result"]["scores"]
        head[c["id"]] = [sc["ai_score"], sc["misinfo_score"], sc["deepfake_score"]]
    mod = np.array([3 if c["mode"] == "both" else 1 if c["mode"] == "text" else 2 for c in cases["cases"]], np.uint8)
    out = f.score_batch(g["text_table"][:ns], g["image_table"][:ns], head, mod, top_k=5)
    assert set(out) == {"clip_similarity", "vault_discrepancy", "vault_scores", "vault_rows", "scores", "probs", "verdict", "confidence"}
    for c in cases["cases"]:
        i, want = c["id"], c["result"]["scores"]
        assert abs(out["clip_similarity"][i].item() - want["clip_similarity"]) <= FP32_TOL
        assert abs(out["vault_discrepancy"][i].item() - want["vault_discrepancy"]) <= FP32_TOL
        assert abs(out["probs"][i, 1].item() - want["fake_probability"]) <= FP32_TOL
        if abs(want["fake_probability"] - 0.5) > FP32_TOL:
            assert out["verdict"][i].item() == want["verdict"]
    # no modality mask = both present; no vault = zero discrepancy, no rows
    out2 = mmf_b200.score_batch(f.engine, None, g["text_table"][:4], g["image_table"][:4], head[:4], None, 5)
    assert torch.all(out2["vault_rows"] == -1) and torch.all(out2["vault_discrepancy"] == 0)


def test_search_similar_articles_host_logic(tmp_path, capsys):
    """drop-in for train_clip_detective.search_similar_articles: records (rank, article_id, similarity, text, image_path)
    equal the reference's; fp16 databases inside the documented 1e-2 band (we renormalise the rows in fp32)"""
    g = np.load(os.path.join(GOLDEN, "similar.npz"))
    with open(os.path.join(GOLDEN, "similar_cases.json")) as fh:
        c = json.load(fh)
    k = c["top_k"]
    db = {"article_ids": c["article_ids"], "text_contents": c["text_contents"], "image_paths": c["image_paths"],
          "image_embeddings": g["image_embeddings"], "text_embeddings": g["text_embeddings"]}
    clip = fakes.FakeClipModel(g["image_queries"], g["text_queries"])
    shared = OracleEngine()
    shared.vault_load(np.eye(4, 512, dtype=np.float32), mode="fp32")          # a resident Truth Vault that must survive
    common = dict(clip_model=clip, processor=fakes.FakeClipProcessor(), engine=shared)
    for i, want in enumerate(c["results"]["text"]):
        got = mmf_b200.search_similar_articles(query_text=fakes.text_for_id(i), top_k=k, search_mode="text", embeddings_db=db, **common)
        assert [(r["rank"], r["article_id"], r["text"], r["image_path"]) for r in got] == \
               [(r["rank"], r["article_id"], r["text"], r["image_path"]) for r in want], i
        assert np.allclose([r["similarity"] for r in got], [r["similarity"] for r in want], atol=FP32_TOL)
    for i, want in enumerate(c["results"]["image"]):
        p = tmp_path / f"q{i}.png"
        fakes.image_for_id(i).save(p)
        got = mmf_b200.search_similar_articles(query_image_path=str(p), top_k=k, search_mode="image", embeddings_db=db, **common)
        assert [r["article_id"] for r in got] == [r["article_id"] for r in want], i
        assert np.allclose([r["similarity"] for r in got], [r["similarity"] for r in want], atol=FP32_TOL)
    db16 = dict(db, text_embeddings=g["text_embeddings"].astype(np.float16))
    for i, want in enumerate(c["results"]["text_f16"][:4]):
        got = mmf_b200.search_similar_articles(query_text=fakes.text_for_id(i), top_k=k, embeddings_db=db16, **common)
        assert got[0]["article_id"] == want[0]["article_id"]
        assert np.allclose([r["similarity"] for r in got], [r["similarity"] for r in want], atol=1e-2)
    # the pickle path, top_k larger than the database, and the reference's error
    import pickle
    small = {key: (v[:3] if not isinstance(v, dict) else v) for key, v in db.items()}
    with open(tmp_path / "db.pkl", "wb") as fh:
        pickle.dump(small, fh)
    got = mmf_b200.search_similar_articles(query_text=fakes.text_for_id(1), embeddings_db_path=str(tmp_path / "db.pkl"), top_k=5, **common)
    assert len(got) == 3 and [r["rank"] for r in got] == [1, 2, 3]
    with pytest.raises(ValueError, match="Invalid search mode or missing query"):
        mmf_b200.search_similar_articles(search_mode="text", embeddings_db=db, **common)
    assert "Top 5 similar articles" in capsys.readouterr().out
    json.dumps(got)
    assert shared.vault_rows == 4                                             # the caller's vault is not replaced


def test_fusion_training_dataset_host_logic(tmp_path, capsys):
    """8f rank 1: the cached (M,5) score matrix == what the reference's FusionTrainingDataset.__getitem__ computes per
    sample (tests/golden/fusion_dataset.npz: output of train_fusion_judge.FusionTrainingDataset itself); missing image
    -> zeros, max_samples, item schema, and a DataLoader batch like the reference trainer's"""
    g, cases = _load_analyze_fixture()
    want = np.load(os.path.join(GOLDEN, "fusion_dataset.npz"))
    f = _forensics(g, cases)
    n = len(g["ai"])
    rows = []
    for i in range(n):
        p = tmp_path / f"s{i}.png"
        if i != int(want["missing"][0]):
            fakes.image_for_id(i).save(p)
        rows.append((fakes.text_for_id(i), str(p), i % 2))
    csv = tmp_path / "Final_Fusion_Train.csv"
    csv.write_text("text,image_path,label\n" + "".join(f"{t},{p},{lab}\n" for t, p, lab in rows))
    for batch_size in (256, 5):
        ds = mmf_b200.FusionTrainingDataset(str(csv), f, batch_size=batch_size)
        assert len(ds) == n
        items = [ds[i] for i in range(n)]
        scores = torch.stack([it["scores"] for it in items]).numpy()
        assert scores.dtype == np.float32 and items[0]["label"].dtype == torch.long and items[0]["scores"].shape == (5,)
        assert np.allclose(scores, want["scores"], atol=FP32_TOL) and not scores[int(want["missing"][0])].any()
        assert np.array_equal(torch.stack([it["label"] for it in items]).numpy(), want["labels"])
    assert "Image not found" in capsys.readouterr().out
    assert len(mmf_b200.FusionTrainingDataset(str(csv), f, max_samples=7)) == 7
    batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=16, shuffle=False)))
    assert batch["scores"].shape == (16, 5) and batch["label"].shape == (16,)
    logits = f.detector.forward_fusion(batch["scores"])            # what the reference trainer does with it (:221)
    assert logits.shape == (16, 2)


def test_individual_weight_files_load_in_the_reference_layout(tmp_path, capsys):
    """ai_head_best.pth / roberta_detective_best.pth are {'model_state_dict': {'ai_head.0.weight', ...},
    'epoch': ...}; the reference filters by branch name and strips the prefix (misinfo_forensics.py:270-283), and accepts
    EfficientNet either as {'model_state_dict': {'efficientnet.…'}} or as a raw state_dict (:286-304)."""
    from transformers import RobertaConfig, RobertaModel
    from mmf_b200.forensics import MultiModalMisinfoDetector, load_individual_weights
    cfg = RobertaConfig(vocab_size=64, hidden_size=32, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64,
                        max_position_embeddings=40, type_vocab_size=1, pad_token_id=1)
    torch.manual_seed(11)
    src = MultiModalMisinfoDetector(roberta=RobertaModel(cfg))          # "trained" weights
    torch.manual_seed(12)
    dst = MultiModalMisinfoDetector(roberta=RobertaModel(cfg))          # freshly initialised
    full = src.state_dict()
    assert not torch.equal(src.ai_head[0].weight, dst.ai_head[0].weight)
    ai, mis, eff = (str(tmp_path / n) for n in ("ai_head_best.pth", "roberta_detective_best.pth", "efficientnet_cifake_best.pth"))
    torch.save({"model_state_dict": {k: v for k, v in full.items() if k.startswith(("roberta.", "ai_head."))}, "epoch": 3}, ai)
    torch.save({"model_state_dict": {k: v for k, v in full.items() if k.startswith(("roberta.", "misinfo_head."))}, "epoch": 4}, mis)
    torch.save({"model_state_dict": {k: v for k, v in full.items() if k.startswith("efficientnet.")}, "epoch": 5}, eff)
    load_individual_weights(dst, ai, mis, eff)
    for a, b in ((src.ai_head, dst.ai_head), (src.misinfo_head, dst.misinfo_head), (src.efficientnet, dst.efficientnet)):
        for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
            assert ka == kb and torch.equal(va, vb), ka
    out = capsys.readouterr().out
    assert "epoch 3" in out and "epoch 4" in out and "epoch 5" in out and "no parameter" not in out
    # raw EfficientNet state_dict form
    torch.manual_seed(13)
    dst2 = MultiModalMisinfoDetector(roberta=RobertaModel(cfg))
    torch.save(src.efficientnet.state_dict(), eff)
    load_individual_weights(dst2, str(tmp_path / "absent.pth"), str(tmp_path / "absent2.pth"), eff)
    assert torch.equal(src.efficientnet.classifier[1].weight, dst2.efficientnet.classifier[1].weight)
    # a checkpoint that matches nothing is reported, not silently ignored
    torch.save({"model_state_dict": {"something.else": torch.zeros(1)}, "epoch": 0}, ai)
    load_individual_weights(dst2, ai, str(tmp_path / "absent2.pth"), str(tmp_path / "absent3.pth"))
    assert "no parameter of the checkpoint matched" in capsys.readouterr().out
