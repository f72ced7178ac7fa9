"""A/B parity of the kernel variants selected by environment switches read at search time: the screened
fp32-exact search (default; MMF_MMA_SCREEN=0 = the 3-pass kernel) and the histogram bound (default;
MMF_MMA_BOUND=pool = the bucket maxima of the first revisions).  The same comparisons were run on a B200 through
the torch-free tools/cabi_selftest (profiles/r01_cabi_selftest.log: 48 bit-identical, 17 within tolerance); this
Python form of them was written when the round's GPU time was spent and has not run on a GPU yet, so it is
skipped unless MMF_EXPERIMENTAL=1:

    MMF_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -x -q"""
import contextlib
import os

import numpy as np
import pytest
import torch

import oracle
from util import BF16_TOL, FP32_TOL, assert_close, assert_topk

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MMF_EXPERIMENTAL") != "1", reason="experimental variants: set MMF_EXPERIMENTAL=1")]

import mmf_b200  # noqa: E402
from mmf_b200 import synth  # noqa: E402

DOUBLE = os.environ.get("MMF_TEST_DOUBLE") == "1"     # CPU check of this file's own Python (see the eng fixture)
DEV = "cpu" if DOUBLE else "cuda"


@pytest.fixture(scope="module")
def eng():
    if os.environ.get("MMF_TEST_DOUBLE") == "1":          # CPU check of THIS FILE's own Python only (tests/cpu_engine.py)
        from cpu_engine import OracleEngine
        yield OracleEngine()
        return
    e = mmf_b200.Engine("cuda:0")
    yield e
    e.close()


@contextlib.contextmanager
def env(**kw):
    """the library reads its switches with getenv() at every search call"""
    old = {k: os.environ.get(k) for k in kw}
    os.environ.update({k: str(v) for k, v in kw.items()})
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def npy(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("n_rows,nq,k", [(33333, 300, 100), (200000, 128, 32), (2000, 40, 256), (150, 3, 200),
                                         (300000, 513, 100), (70001, 129, 17)])
def test_histogram_bound_equals_bucket_pool(eng, mode, n_rows, nq, k):
    """MMF_MMA_BOUND=hist only changes which elements the epilogue filter lets through; the selected top-k
    (exact selection over the candidate lists) must be bit-identical to the default bucket-pool variant"""
    vault = synth.vault_rows(n_rows, seed=n_rows + 3) * np.random.default_rng(3).uniform(0.1, 5, (n_rows, 1)).astype(np.float32)
    q, _, _ = synth.queries(nq, n_rows, seed=nq + 13, plant_frac=0.4, vault_seed=n_rows + 3)
    eng.vault_load(vault, mode=mode)
    with env(MMF_MMA_BOUND="pool", MMF_MMA_SCREEN="0"):
        base = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
    with env(MMF_MMA_BOUND="hist", MMF_MMA_SCREEN="0"):
        got = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
    for a, b, what in zip(got, base, ("scores", "rows", "discrepancy")):
        assert np.array_equal(a, b, equal_nan=True), f"{what} differ between the histogram and the bucket-pool bound"
    ri, rs, rd = oracle.vault_search_batched(vault, q, k)
    kk = ri.shape[1]
    if mode == "fp32":
        assert_topk(got[1][:, :kk], got[0][:, :kk], ri, rs, FP32_TOL, f"hist N={n_rows} Q={nq} k={k}")
    else:
        assert_close(got[0][:, :kk], rs, BF16_TOL, "hist bf16 scores")


def test_histogram_bound_adversarial_orders(eng):
    """ascending scores (every row beats the threshold), all-equal rows, all-negative scores, tiny scores"""
    n_rows, k = 40000, 100
    base = synth.vault_rows(1, seed=1)[0]
    r = np.random.default_rng(5)
    noise = r.standard_normal((n_rows, 512)).astype(np.float32)
    w = np.linspace(-1.0, 3.0, n_rows, dtype=np.float32)[:, None]          # cosine to `base` rises with the row id
    vault = noise + w * base[None, :] * np.sqrt(512)
    q = np.stack([base, -base, base + 0.5 * noise[0], noise[1] * 1e-3] + [noise[i] for i in range(2, 140)])
    for mode, tol in (("fp32", FP32_TOL), ("bf16", BF16_TOL)):
        eng.vault_load(vault, mode=mode)
        with env(MMF_MMA_BOUND="pool"):
            ref = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
        with env(MMF_MMA_BOUND="hist"):
            got = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
        assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, ref)), mode
        ri, rs, _ = oracle.vault_search_batched(vault, q, k)
        assert_close(got[0], rs, tol, f"adversarial {mode}")
    dup = np.repeat(vault[:7], 3000, axis=0)                                  # 3000 copies of each row: ties everywhere
    eng.vault_load(dup, mode="fp32")
    with env(MMF_MMA_BOUND="pool"):
        ref = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
    with env(MMF_MMA_BOUND="hist"):
        got = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, ref))


# ------------------------------------------------------------------------------ screened fp32-exact search
@pytest.mark.parametrize("n_rows,nq,k", [(128, 1, 1), (129, 130, 5), (5000, 16, 10), (40000, 257, 10), (200000, 128, 16),
                                         (150, 3, 12), (1_000_000, 256, 10)])
def test_screened_search_is_exact(eng, n_rows, nq, k):
    """MMF_MMA_SCREEN=1: one f16 pass over the hi planes + exact fp32 re-scoring of everything within the proven
    error band.  The result must be the exact top-k: bit-identical to the streaming kernel (same re-scoring
    arithmetic), and within the fp32 tolerance of the oracle."""
    if n_rows >= 1_000_000:
        g = torch.Generator(device=DEV).manual_seed(3)
        vault = torch.randn(n_rows, 512, device=DEV, generator=g)
        q = torch.randn(nq, 512, device=DEV, generator=g)
        q[:32] = vault[torch.arange(32, device=DEV) * 31_001] + 0.3 * q[:32]
    else:
        vault = synth.vault_rows(n_rows, seed=n_rows + 1) * np.random.default_rng(2).uniform(0.1, 5, (n_rows, 1)).astype(np.float32)
        q, _, _ = synth.queries(nq, n_rows, seed=nq + 11, plant_frac=0.4, vault_seed=n_rows + 1)
    eng.vault_load(vault, mode="fp32")
    exact = [npy(t) for t in eng.vault_search(q, k, algo="stream")]
    with env(MMF_MMA_SCREEN="1"):
        got = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
    for a, b, what in zip(got, exact, ("scores", "rows", "discrepancy")):
        assert np.array_equal(a, b, equal_nan=True), f"screened search: {what} differ from the streaming kernel"
    if n_rows < 1_000_000:
        ri, rs, rd = oracle.vault_search_batched(vault, q, k)
        kk = ri.shape[1]
        assert_topk(got[1][:, :kk], got[0][:, :kk], ri, rs, FP32_TOL, f"screen N={n_rows} Q={nq} k={k}")
        assert_close(got[2], rd, FP32_TOL, "disc")


def test_screened_search_band_overflow_falls_back(eng):
    """thousands of identical rows: the candidate band cannot fit a list, the search flags the overflow and the
    guarded 3-pass kernel redoes the batch -> bit-identical to the default tcgen05 result; ties: higher row id first"""
    base = synth.vault_rows(20000, seed=31)
    vault = np.concatenate([base[:5000], np.repeat(base[7:8], 4000, axis=0), base[5000:]])
    q = np.stack([base[7] * 2.0, base[9], base[11] + 0.1 * base[12]] + [base[100 + i] + base[300 + i] for i in range(140)])
    eng.vault_load(vault, mode="fp32")
    with env(MMF_MMA_SCREEN="0"):
        ref = [npy(t) for t in eng.vault_search(q, 10, algo="mma")]
    with env(MMF_MMA_SCREEN="1"):
        got = [npy(t) for t in eng.vault_search(q, 10, algo="mma")]
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, ref))
    assert list(got[1][0]) == list(range(8999, 8989, -1))
    # moderately clustered vault (bands of a few hundred rows: no overflow expected, exact either way)
    r = np.random.default_rng(4)
    centres = synth.vault_rows(50, seed=32)
    vault = centres[r.integers(0, 50, 60000)] + 0.02 * r.standard_normal((60000, 512)).astype(np.float32)
    q = centres[:40] + 0.02 * r.standard_normal((40, 512)).astype(np.float32)
    eng.vault_load(vault, mode="fp32")
    exact = [npy(t) for t in eng.vault_search(q, 10, algo="stream")]
    with env(MMF_MMA_SCREEN="1"):
        got = [npy(t) for t in eng.vault_search(q, 10, algo="mma")]
    ri, rs, _ = oracle.vault_search_batched(vault, q, 10)
    assert_topk(got[1], got[0], ri, rs, FP32_TOL, "clustered vault")
    assert_close(got[0], exact[0], 5e-6, "clustered vault vs streaming kernel")


# ------------------------------------------------------------------------------ one-call host entry, peer exchange (world 1)
def test_score_batch_host_equals_score_batch(eng):
    """Engine.score_batch_host (mmf_score_batch_host: one library call, host buffers) == mmf_b200.score_batch"""
    n_rows, b, k = 50000, 200, 5
    vault_rows = synth.vault_rows(n_rows, seed=41)
    vault = mmf_b200.TruthVault(eng, vault_rows, None, mode="fp32")
    eng.fusion_load(synth.fusion_state_dict(1))
    q, _, _ = synth.queries(b, n_rows, seed=42, plant_frac=0.3, vault_seed=41)
    text, _ = synth.caption_image_pairs(b, seed=43)
    head = synth.head_scores(b, seed=44)
    for modality in (None, (np.arange(b) % 4).astype(np.uint8)):
        want = mmf_b200.score_batch(eng, vault, text, q, head, modality, k)
        got = eng.score_batch_host(text, q, head, modality, k)
        for key in ("clip_similarity", "vault_discrepancy", "vault_scores", "vault_rows", "scores", "probs", "verdict", "confidence"):
            assert np.array_equal(got[key], npy(want[key]), equal_nan=True), key
    # pinned torch tensors in, and no vault loaded -> zero discrepancy / no rows
    eng.vault_unload()
    pin = (lambda t: t) if DOUBLE else (lambda t: t.pin_memory())
    got = eng.score_batch_host(pin(torch.from_numpy(text)), pin(torch.from_numpy(q)), pin(torch.from_numpy(head)), None, k)
    assert np.all(got["vault_rows"] == -1) and np.all(got["vault_discrepancy"] == 0) and np.all(np.isnan(got["vault_scores"]))


def test_peer_exchange_single_rank_loopback(eng):
    """csrc/exchange.cu with world = 1 (the only rank pushes into its own buffer): push, flag, wait-merge must give
    the plain search result; three calls in a row cover both buffer parities"""
    n_rows, nq = 30000, 40
    vault = synth.vault_rows(n_rows, seed=51)
    q, _, _ = synth.queries(nq, n_rows, seed=52, plant_frac=0.5, vault_seed=51)
    eng.vault_load(vault, mode="fp32")
    need = eng.exchange_layout(1, nq, 100)
    buf = torch.zeros(need // 8 + 16, dtype=torch.int64, device=DEV)
    eng.exchange_attach(0, 1, [buf.data_ptr()], buf.numel() * 8)
    try:
        for k in (10, 100, 5):
            want = [npy(t) for t in eng.vault_search(q, k)]
            for _ in range(3):
                got = [npy(t) for t in eng.vault_search_exchange(q, k, k)]
                assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, want)), k
        with pytest.raises(mmf_b200.MMFError):
            eng.vault_search_exchange(np.zeros((100000, 512), np.float32), 100, 100)     # does not fit the attached buffer
    finally:
        eng.exchange_detach()


def test_c4_shard_size_properties(eng):
    """BASELINE config C4 at the size one rank sees on 8 GPUs (4096 queries x 1.25 M bf16 rows, top-100), through
    size-independent properties: planted rows first at their cosine, scores sorted, a query scaled is the same
    query, a 2-way row split + candidate merge equals the unsplit search bit for bit, both bound variants agree."""
    if DOUBLE:
        pytest.skip("4096 queries x 1.25 M rows is not a CPU-sized problem")
    n_rows, nq, k = 1_250_000, 4096, 100
    g = torch.Generator(device=DEV).manual_seed(11)
    vault = torch.randn(n_rows, 512, device=DEV, generator=g)
    q = torch.randn(nq, 512, device=DEV, generator=g)
    pick = torch.randint(0, n_rows, (400,), device=DEV, generator=g)
    vn = torch.nn.functional.normalize(vault[pick], dim=1)
    noise = torch.nn.functional.normalize(q[:400] - (q[:400] * vn).sum(1, keepdim=True) * vn, dim=1)
    cosv = torch.tensor([0.8, 0.849, 0.851, 0.9, 0.99], device=DEV).repeat(80)
    q[:400] = (cosv[:, None] * vn + torch.sqrt(1 - cosv ** 2)[:, None] * noise) * 3.0
    eng.vault_load(vault, mode="bf16")
    scores, rows, disc = eng.vault_search(q, k)
    assert torch.equal(rows[:400, 0], pick)
    assert torch.allclose(scores[:400, 0], cosv, atol=BF16_TOL)
    safe = (cosv - 0.85).abs() > BF16_TOL
    assert torch.equal((disc[:400] > 0)[safe], (cosv > 0.85)[safe])
    assert torch.all(scores[:, :-1] >= scores[:, 1:]) and torch.all(rows >= 0)
    s2, r2, _ = eng.vault_search(q * 0.01, k)
    # the bf16 operand of q/|q| does not depend on |q| up to an fp32 rounding that can flip a bf16 rounding
    assert torch.equal(r2[:400, 0], rows[:400, 0]) and torch.allclose(s2, scores, atol=BF16_TOL)
    with env(MMF_MMA_BOUND="pool"):
        s3, r3, d3 = eng.vault_search(q, k)
    assert torch.equal(r3, rows) and torch.equal(s3, scores) and torch.equal(d3, disc)
    half = n_rows // 2
    packed = []
    for lo, hi in ((0, half), (half, n_rows)):
        eng.vault_load(vault[lo:hi], mode="bf16", row_offset=lo)
        packed.append(eng.vault_search_candidates(q, k).clone())
    s4, r4, d4 = eng.topk_merge(torch.stack(packed), k)
    assert torch.equal(r4, rows) and torch.equal(s4, scores) and torch.equal(d4, disc)
    eng.vault_unload()


# ------------------------------------------------------------------------------ variants written without a GPU at hand
@pytest.mark.parametrize("n_rows,nq,k", [(1, 1, 1), (31, 3, 5), (1000, 1, 10), (4099, 2, 7), (20000, 5, 10), (70001, 15, 5),
                                         (300000, 8, 16), (1_000_000, 1, 10)])
def test_screened_streaming_kernel_is_exact(eng, n_rows, nq, k):
    """MMF_STREAM_SCREEN=1 (batch-1 path reads only the hi planes, re-scores the band exactly) == exact streaming kernel"""
    g = torch.Generator(device=DEV).manual_seed(n_rows + nq)
    vault = torch.randn(n_rows, 512, device=DEV, generator=g) * (0.1 + torch.rand(n_rows, 1, device=DEV, generator=g) * 4)
    q = torch.randn(nq, 512, device=DEV, generator=g)
    q[0] = vault[n_rows // 2] * 2 + 0.3 * q[0]
    eng.vault_load(vault, mode="fp32")
    want = [npy(t) for t in eng.vault_search(q, k, algo="stream")]
    with env(MMF_STREAM_SCREEN="1"):
        got = [npy(t) for t in eng.vault_search(q, k, algo="stream")]
        again = [npy(t) for t in eng.vault_search(q, k, algo="stream")]
    for a, b, c in zip(got, want, again):
        assert np.array_equal(a, b, equal_nan=True) and np.array_equal(c, b, equal_nan=True)


def test_screened_streaming_kernel_overflow_and_ties(eng):
    base = synth.vault_rows(20000, seed=61)
    vault = np.concatenate([base[:5000], np.repeat(base[7:8], 4000, axis=0), base[5000:]])
    q = np.stack([base[7] * 2.0, base[9], base[11] + 0.1 * base[12]])
    eng.vault_load(vault, mode="fp32")
    want = [npy(t) for t in eng.vault_search(q, 10, algo="stream")]
    with env(MMF_STREAM_SCREEN="1"):
        got = [npy(t) for t in eng.vault_search(q, 10, algo="stream")]
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, want))
    assert list(got[1][0]) == list(range(8999, 8989, -1))


@pytest.mark.parametrize("switches", [{"MMF_MMA_STAGES": "12"}, {"MMF_MMA_PREFETCH": "1"}, {"MMF_MMA_LEAN": "1"},
                                      {"MMF_MMA_STAGES": "12", "MMF_MMA_PREFETCH": "1", "MMF_MMA_LEAN": "1"}])
def test_screened_search_experiments_change_nothing(eng, switches):
    """deeper ring / L2 prefetch / lean launch sequence only change HOW the screening pass runs"""
    for n_rows, nq, k in ((40000, 257, 10), (5000, 16, 10), (500000, 256, 5)):
        g = torch.Generator(device=DEV).manual_seed(n_rows)
        vault = torch.randn(n_rows, 512, device=DEV, generator=g)
        q = torch.randn(nq, 512, device=DEV, generator=g)
        q[:8] = vault[:8] + 0.2 * q[:8]
        eng.vault_load(vault, mode="fp32")
        want = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
        with env(**switches):
            got = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
            again = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
        for a, b, c in zip(got, want, again):
            assert np.array_equal(a, b, equal_nan=True) and np.array_equal(c, b, equal_nan=True), (switches, n_rows)


def test_search_similar_articles_dropin_on_gpu(tmp_path):
    """8f rank 4: search_similar_articles through the real kernels vs the fixtures of the reference's own function"""
    import json
    import fakes
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "similar.npz"))
    with open(os.path.join(GOLDEN, "similar_cases.json")) as fh:
        c = json.load(fh)
    k = c["top_k"]
    db = {"article_ids": c["article_ids"], "text_contents": c["text_contents"], "image_paths": c["image_paths"],
          "image_embeddings": g["image_embeddings"], "text_embeddings": g["text_embeddings"]}
    if DOUBLE:
        from cpu_engine import OracleEngine
        engine = OracleEngine()
        engine.close = lambda: None
    else:
        engine = mmf_b200.Engine("cuda:0")
    common = dict(clip_model=fakes.FakeClipModel(g["image_queries"], g["text_queries"]), processor=fakes.FakeClipProcessor(), engine=engine)
    for i, want in enumerate(c["results"]["text"]):
        got = mmf_b200.search_similar_articles(query_text=fakes.text_for_id(i), top_k=k, embeddings_db=db, **common)
        assert [r["article_id"] for r in got] == [r["article_id"] for r in want], i
        assert np.allclose([r["similarity"] for r in got], [r["similarity"] for r in want], atol=FP32_TOL)
    for i, want in enumerate(c["results"]["image"]):
        p = tmp_path / f"q{i}.png"
        fakes.image_for_id(i).save(p)
        got = mmf_b200.search_similar_articles(query_image_path=str(p), top_k=k, search_mode="image", embeddings_db=db, **common)
        assert [r["article_id"] for r in got] == [r["article_id"] for r in want], i
        assert np.allclose([r["similarity"] for r in got], [r["similarity"] for r in want], atol=FP32_TOL)
    engine.close()
