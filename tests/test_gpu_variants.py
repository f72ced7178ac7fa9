"""GPU parity of the search variants and of the entry points added in round 2, against the CPU oracle
(oracle/reference_port.py) and against the HBM-streaming kernel (an independent CUDA-core implementation):
the screened fp32-exact search and its guarded fallback, the histogram bound (top_k > 16), BASELINE.json's
configurations at their full sizes (C1 exactly, C2 exactly, C4's shard shape), the host-buffer entry points
(one call, submit / collect), the peer-memory exchange on one rank, the sharded raw vault directory, and
search_similar_articles.  Every test runs in the default `-m gpu` suite; variants are selected with
Engine.set_option (mmf_set_option), never with environment switches.

MMF_TEST_DOUBLE=1 runs this file's own Python against the oracle-backed test double on the CPU (tests/cpu_engine.py)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from util import BF16_TOL, FP32_TOL, assert_close, assert_topk

pytestmark = [pytest.mark.gpu]

import mmf_b200  # noqa: E402
from mmf_b200 import synth  # noqa: E402

DOUBLE = os.environ.get("MMF_TEST_DOUBLE") == "1"     # CPU check of this file's own Python (see the eng fixture)
DEV = "cpu" if DOUBLE else "cuda"


def _engine():
    if DOUBLE:
        from cpu_engine import OracleEngine
        return OracleEngine()
    return mmf_b200.Engine("cuda:0")


@pytest.fixture(scope="module")
def eng():
    e = _engine()
    yield e
    e.close()


def npy(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def oracle_topk_torch(vn: torch.Tensor, qn: torch.Tensor, k: int, chunk: int = 512):
    """exact top-k of qn @ vn.T on the CPU for sizes where the per-row argpartition of oracle.vault_search_batched is
    too slow (the same arithmetic: fp32 GEMM; ties by torch.topk's order, so compare with assert_topk)"""
    idx, sc = [], []
    for c0 in range(0, qn.shape[0], chunk):
        s = qn[c0:c0 + chunk] @ vn.T
        v, i = torch.topk(s, k, dim=1)
        idx.append(i)
        sc.append(v)
    return torch.cat(idx).numpy(), torch.cat(sc).numpy()


def oracle_search(vault: np.ndarray, q: np.ndarray, k: int):
    """(rows, scores, discrepancy) of the oracle.  Small problems: oracle.vault_search_batched (per-row argpartition under
    the kernels' total order).  Large ones: the same normalisations and the same fp32 GEMM, top-k by torch.topk -- its tie
    order is unspecified, which assert_topk tolerates (rows must agree only outside near-ties)."""
    if vault.shape[0] * q.shape[0] <= 4_000_000 or k > vault.shape[0]:
        return oracle.vault_search_batched(vault, q, k)
    vn = torch.from_numpy(np.ascontiguousarray(oracle.vault_normalise(vault), dtype=np.float32))
    qn = oracle.normalise_rows(torch.from_numpy(np.asarray(q, np.float32)))
    ri, rs = oracle_topk_torch(vn, qn, k)
    return ri, rs, oracle.discrepancy_rule(rs[:, 0])


# ------------------------------------------------------------------------------ histogram bound (top_k > 16)
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("n_rows,nq,k", [(33333, 300, 100), (200000, 128, 32), (2000, 40, 256), (150, 3, 200),
                                         (300000, 513, 100), (70001, 129, 17)])
def test_histogram_bound_vs_oracle(eng, mode, n_rows, nq, k):
    """top_k > 16: the grid-wide bound comes from the per-query score histogram; the selected top-k must be the
    oracle's (fp32 mode: 1e-5, rows exact outside near-ties; bf16 mode: 1e-2) and the streaming kernel's"""
    vault = synth.vault_rows(n_rows, seed=n_rows + 3) * np.random.default_rng(3).uniform(0.1, 5, (n_rows, 1)).astype(np.float32)
    q, _, _ = synth.queries(nq, n_rows, seed=nq + 13, plant_frac=0.4, vault_seed=n_rows + 3)
    eng.vault_load(vault, mode=mode)
    got = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
    again = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
    for a, b, what in zip(got, again, ("scores", "rows", "discrepancy")):
        assert np.array_equal(a, b, equal_nan=True), f"{what} differ between two runs"
    ri, rs, rd = oracle_search(vault, q, k)
    kk = ri.shape[1]
    if mode == "fp32":
        assert_topk(got[1][:, :kk], got[0][:, :kk], ri, rs, FP32_TOL, f"hist N={n_rows} Q={nq} k={k}")
    else:
        assert_close(got[0][:, :kk], rs, BF16_TOL, "hist bf16 scores")
    assert np.all(got[1][:, kk:] == -1) and np.all(np.isnan(got[0][:, kk:]))       # top_k > N: padding


def test_histogram_bound_adversarial_orders(eng):
    """ascending scores (every row beats the threshold), all-negative scores, tiny scores, thousands of ties"""
    n_rows, k = 40000, 100
    base = synth.vault_rows(1, seed=1)[0]
    r = np.random.default_rng(5)
    noise = r.standard_normal((n_rows, 512)).astype(np.float32)
    w = np.linspace(-1.0, 3.0, n_rows, dtype=np.float32)[:, None]          # cosine to `base` rises with the row id
    vault = noise + w * base[None, :] * np.sqrt(512)
    q = np.stack([base, -base, base + 0.5 * noise[0], noise[1] * 1e-3] + [noise[i] for i in range(2, 140)])
    for mode, tol in (("fp32", FP32_TOL), ("bf16", BF16_TOL)):
        eng.vault_load(vault, mode=mode)
        got = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
        ri, rs, _ = oracle.vault_search_batched(vault, q, k)
        assert_close(got[0], rs, tol, f"adversarial {mode}")
        if mode == "fp32":
            assert_topk(got[1], got[0], ri, rs, tol, "adversarial fp32")
    dup = np.repeat(vault[:7], 3000, axis=0)                                  # 3000 copies of each row: ties everywhere
    eng.vault_load(dup, mode="fp32")
    got = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
    ri, rs, _ = oracle.vault_search_batched(dup, q, k)
    assert_close(got[0], rs, FP32_TOL, "ties")
    assert np.all(got[1] >= 0) and all(len(set(row)) == k for row in got[1])          # k distinct rows per query


# ------------------------------------------------------------------------------ screened fp32-exact search
@pytest.mark.parametrize("n_rows,nq,k", [(128, 1, 1), (129, 130, 5), (5000, 16, 10), (40000, 257, 10), (200000, 128, 16),
                                         (150, 3, 12), (262144, 1024, 10)])
def test_screened_search_is_exact(eng, n_rows, nq, k):
    """fp32-exact vaults, top_k <= 16: one f16 pass over the hi planes + exact fp32 re-scoring of everything within
    the proven error band.  The result must be the exact top-k: bit-identical to the streaming kernel (same re-scoring
    arithmetic) for both epilogue forms, and within the fp32 tolerance of the oracle."""
    vault = synth.vault_rows(n_rows, seed=n_rows + 1) * np.random.default_rng(2).uniform(0.1, 5, (n_rows, 1)).astype(np.float32)
    q, _, _ = synth.queries(nq, n_rows, seed=nq + 11, plant_frac=0.4, vault_seed=n_rows + 1)
    eng.vault_load(vault, mode="fp32")
    exact = [npy(t) for t in eng.vault_search(q, k, algo="stream")]
    for parity in (-1, 0, 1):
        eng.set_option("epi_parity", parity)
        got = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
        for a, b, what in zip(got, exact, ("scores", "rows", "discrepancy")):
            assert np.array_equal(a, b, equal_nan=True), f"screened search (epi_parity={parity}): {what} differ from the streaming kernel"
    eng.set_option("epi_parity", -1)
    ri, rs, rd = oracle_search(vault, q, k)
    kk = ri.shape[1]
    assert_topk(got[1][:, :kk], got[0][:, :kk], ri, rs, FP32_TOL, f"screen N={n_rows} Q={nq} k={k}")
    assert_close(got[2], rd, FP32_TOL, "disc")
    eng.set_option("screen", 0)                       # the 3-pass kernel (also the guarded fallback): within 1e-5
    three = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
    eng.set_option("screen", 1)
    assert_topk(three[1][:, :kk], three[0][:, :kk], ri, rs, FP32_TOL, "3-pass")


def test_c2_full_size_vs_oracle(eng):
    """BASELINE.json configs[1] EXACTLY: 256 queries vs 1 M fp32-exact rows, top-10 -- every (query, rank) against the
    CPU oracle (a 262 GFLOP fp32 GEMM: seconds), not only the planted rows; and bit-identical to the streaming kernel"""
    if DOUBLE:
        pytest.skip("1 M rows x 256 queries is not a CPU-sized problem for the test double")
    n_rows, nq, k = 1_000_000, 256, 10
    g = torch.Generator(device=DEV).manual_seed(3)
    vault = torch.randn(n_rows, 512, device=DEV, generator=g)
    q = torch.randn(nq, 512, device=DEV, generator=g) * 2.5
    q[:32] = vault[torch.arange(32, device=DEV) * 31_001] + 0.3 * q[:32]
    eng.vault_load(vault, mode="fp32")
    got = [npy(t) for t in eng.vault_search(q, k)]
    stream = [npy(t) for t in eng.vault_search(q, k, algo="stream")]
    for a, b, what in zip(got, stream, ("scores", "rows", "discrepancy")):
        assert np.array_equal(a, b, equal_nan=True), what
    vh = vault.cpu()
    vn = vh / vh.norm(dim=1, keepdim=True)            # misinfo_forensics.py:443-445 (fp32)
    qh = q.cpu()
    qn = qh / qh.norm(dim=-1, keepdim=True)           # :439
    ri, rs = oracle_topk_torch(vn, qn, k)
    assert_topk(got[1], got[0], ri, rs, FP32_TOL, "C2 full size")
    assert_close(got[2], oracle.discrepancy_rule(rs[:, 0]), FP32_TOL, "C2 discrepancy")
    assert np.array_equal(got[1][:32, 0], np.arange(32) * 31_001)


def test_c1_exact_vs_oracle(eng):
    """BASELINE.json configs[0] EXACTLY: 1000 queries vs a 100 k-row fp32 vault, top-10, + caption/image cosine + fusion
    judge on the (1000,5) score vectors -- the whole batched path against the oracle, row by row"""
    n_rows, nq, k = 100_000, 1000, 10
    vault = synth.vault_rows(n_rows, seed=71)
    q, prow, _ = synth.queries(nq, n_rows, seed=72, plant_frac=0.1, vault_seed=71)
    text, _ = synth.caption_image_pairs(nq, seed=73)
    head = synth.head_scores(nq, seed=74)
    w = synth.fusion_state_dict(0)
    eng.vault_load(vault, mode="fp32")
    eng.fusion_load(w)
    out = {key: npy(v) for key, v in eng.score_batch(text, q, head, None, k).items()}
    ri, rs, rd = oracle_search(vault, q, k)
    assert_topk(out["vault_rows"], out["vault_scores"], ri, rs, FP32_TOL, "C1 vault")
    assert_close(out["vault_discrepancy"], rd, FP32_TOL, "C1 discrepancy")
    assert np.array_equal(out["vault_rows"][prow >= 0, 0], prow[prow >= 0])
    sim = oracle.cosine_pairs(text, q)
    assert_close(out["clip_similarity"], sim, FP32_TOL, "C1 cosine")
    x = np.concatenate([head, sim[:, None], rd[:, None]], axis=1).astype(np.float32)
    assert_close(out["scores"], x, FP32_TOL, "C1 fusion inputs")
    p = oracle.fusion_forward(w, out["scores"])
    assert_close(out["probs"], p, FP32_TOL, "C1 fusion probabilities")
    sure = np.abs(p[:, 1] - 0.5) > FP32_TOL
    assert np.array_equal(out["verdict"][sure], (p[:, 1] > 0.5).astype(np.int32)[sure])


def test_c4_shape_vs_oracle(eng):
    """The kernel instantiation of BASELINE.json configs[3] (bf16 vault, 4096 queries, top-100, histogram bound) against
    the oracle run on the kernel's own operands (bf16-rounded normalised rows and queries, fp32 accumulate), on a shard
    large enough for the long-strip schedule (4096 x 300 k: a 1.26 TFLOP CPU GEMM)"""
    if DOUBLE:
        pytest.skip("not a CPU-sized problem for the test double")
    n_rows, nq, k = 300_000, 4096, 100
    g = torch.Generator(device=DEV).manual_seed(17)
    vault = torch.randn(n_rows, 512, device=DEV, generator=g)
    q = torch.randn(nq, 512, device=DEV, generator=g) * 1.7
    q[:100] = vault[torch.arange(100, device=DEV) * 2_999] * 2 + 0.5 * q[:100]
    eng.vault_load(vault, mode="bf16")
    scores, rows, disc = [npy(t) for t in eng.vault_search(q, k)]
    vh, qh = vault.cpu(), q.cpu()
    vn = (vh / vh.norm(dim=1, keepdim=True)).bfloat16().float()
    qn = (qh / qh.norm(dim=-1, keepdim=True)).bfloat16().float()
    ri, rs = oracle_topk_torch(vn, qn, k)
    # same operands, fp32 accumulation on both sides -- but the tensor core TRUNCATES when it accumulates:
    # up to 512 * 2^-24 * |score| below the CPU's round-to-nearest sum, 8.4e-5 measured on the planted rows (score ~ 0.97)
    assert_topk(rows, scores, ri, rs, 2e-4, "C4 shape vs the oracle on the bf16 operands")
    assert np.array_equal(rows[:100, 0], np.arange(100) * 2_999)
    safe = np.abs(rs[:, 0] - 0.85) > BF16_TOL
    assert_close(disc[safe], oracle.discrepancy_rule(rs[:, 0])[safe], BF16_TOL, "C4 discrepancy")


def test_c4_shard_size_properties(eng):
    """BASELINE config C4 at the size one rank sees on 8 GPUs (4096 queries x 1.25 M bf16 rows, top-100), through
    size-independent properties: planted rows first at their cosine, scores sorted, a query scaled is the same
    query, a 2-way row split + candidate merge equals the unsplit search bit for bit, two runs agree."""
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


def test_screened_search_band_overflow_falls_back(eng):
    """thousands of identical rows: the candidate band cannot fit a list, the search flags the overflow and the
    guarded 3-pass kernel redoes the batch -> bit-identical to the 3-pass result; ties: higher row id first"""
    base = synth.vault_rows(20000, seed=31)
    vault = np.concatenate([base[:5000], np.repeat(base[7:8], 4000, axis=0), base[5000:]])
    q = np.stack([base[7] * 2.0, base[9], base[11] + 0.1 * base[12]] + [base[100 + i] + base[300 + i] for i in range(140)])
    eng.vault_load(vault, mode="fp32")
    eng.set_option("screen", 0)
    ref = [npy(t) for t in eng.vault_search(q, 10, algo="mma")]
    eng.set_option("screen", 1)
    for _ in range(2):                                 # twice: the flag and both counter sets must reset
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
    got = [npy(t) for t in eng.vault_search(q, 10, algo="mma")]
    ri, rs, _ = oracle.vault_search_batched(vault, q, 10)
    assert_topk(got[1], got[0], ri, rs, FP32_TOL, "clustered vault")
    assert_close(got[0], exact[0], 5e-6, "clustered vault vs streaming kernel")


def test_zero_query_is_the_same_on_both_kernels(eng):
    """an all-zero embedding (q / |q| = NaN; the reference would return arbitrary rows with NaN similarity): both search
    kernels return the same thing for it, and the other queries of the batch are unaffected"""
    n_rows, nq, k = 30000, 40, 10
    vault = synth.vault_rows(n_rows, seed=91)
    q, _, _ = synth.queries(nq, n_rows, seed=92, plant_frac=0.5, vault_seed=91)
    q[3] = 0.0
    eng.vault_load(vault, mode="fp32")
    a = [npy(t) for t in eng.vault_search(q, k, algo="stream")]
    b = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
    keep = np.arange(nq) != 3
    for x, y in zip(a, b):
        assert np.array_equal(x[keep], y[keep], equal_nan=True)
    assert a[2][3] == 0.0 and b[2][3] == 0.0                        # discrepancy: NaN fails `> 0.85`
    assert np.all(np.isnan(a[0][3])) and np.all(np.isnan(b[0][3]))
    want = np.arange(n_rows - 1, n_rows - 1 - k, -1)                # np.argsort of all-NaN keeps the row order: [-k:][::-1]
    assert np.array_equal(a[1][3], want) and np.array_equal(b[1][3], want)
    eng.vault_load(vault, mode="bf16")                              # and the 1-plane bf16 kernels, top_k > 16 included
    for kk in (10, 40):
        c = [npy(t) for t in eng.vault_search(q, kk, algo="mma")]
        assert np.all(np.isnan(c[0][3])) and np.array_equal(c[1][3], np.arange(n_rows - 1, n_rows - 1 - kk, -1)) and c[2][3] == 0.0


# ------------------------------------------------------------------------------ options, host entries, peer exchange (world 1)
def test_options(eng):
    if DOUBLE:
        pytest.skip("library switches")
    assert eng.get_option("screen") == 1 and eng.get_option("epi_parity") == -1
    eng.set_option("screen", 0)
    assert eng.get_option("screen") == 0
    eng.set_option("screen", 1)
    with pytest.raises(mmf_b200.MMFError):
        eng.set_option("no-such-switch", 1)


def test_lockstep_producers_do_not_change_results(eng):
    """option `lockstep` only paces the TMA producers of the tcgen05 search (several query-tile groups sweeping the same vault
    tiles stay within an L2's worth of each other): same rows, same scores, on or off"""
    n_rows, n_q = 300000, 1300                                       # 6 groups of 256 queries: a segment grid + leftover pairs
    vault = synth.vault_rows(n_rows, seed=61)
    q, _, _ = synth.queries(n_q, n_rows, seed=62, vault_seed=61)
    eng.vault_load(vault, mode="bf16")
    assert eng.get_option("lockstep") == 1 or DOUBLE
    for k in (10, 100):
        eng.set_option("lockstep", 1)
        on = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
        eng.set_option("lockstep", 0)
        off = [npy(t) for t in eng.vault_search(q, k, algo="mma")]
        eng.set_option("lockstep", 1)
        for a, b, what in zip(on, off, ("scores", "rows", "discrepancy")):
            assert np.array_equal(a, b, equal_nan=True), f"top-{k}: {what} differ with lock-step producers on / off"


def test_topk_merge_sorted_lists_unsorted_lists_and_duplicates(eng):
    """mmf_topk_merge: the fast path for lists sorted descending (what the search entries write: binary-search ranks), the
    general path for anything else, and duplicate keys across lists -- all against a plain sort of the packed keys"""
    if DOUBLE:
        pytest.skip("kernel paths of the library")
    n_rows, nq, k, world = 40000, 37, 100, 8
    vault = synth.vault_rows(n_rows, seed=71)
    q, _, _ = synth.queries(nq, n_rows, seed=72, vault_seed=71)
    packed = []
    for r in range(world):
        lo, hi = r * n_rows // world, (r + 1) * n_rows // world
        eng.vault_load(vault[lo:hi], mode="bf16", row_offset=lo)
        packed.append(eng.vault_search_candidates(q, k).clone())
    packed = torch.stack(packed)                                     # (world, nq, k) sorted lists
    keys = packed.cpu().numpy().view(np.uint64)

    def reference(kk, top):
        flat = np.sort(kk.transpose(1, 0, 2).reshape(nq, -1), axis=1)[:, ::-1][:, :top]
        return (flat & np.uint64(0xFFFFFFFF)).astype(np.int64)

    for top in (100, 10):
        rows_sorted = npy(eng.topk_merge(packed, top)[1])
        assert np.array_equal(rows_sorted, reference(keys, top)), f"sorted lists, top-{top}"
    g = torch.Generator().manual_seed(5)
    perm = torch.stack([torch.stack([torch.randperm(k, generator=g) for _ in range(nq)]) for _ in range(world)]).to(packed.device)
    shuffled = torch.gather(packed, 2, perm)                         # same keys, lists no longer sorted: general path
    s_a, r_a, d_a = [npy(t) for t in eng.topk_merge(packed, 100)]
    s_b, r_b, d_b = [npy(t) for t in eng.topk_merge(shuffled, 100)]
    assert np.array_equal(r_a, r_b) and np.array_equal(s_a, s_b) and np.array_equal(d_a, d_b)
    twice = torch.cat([packed[:2], packed[:2]])                      # every key twice, in different lists
    r_dup = npy(eng.topk_merge(twice, 100)[1])
    assert np.array_equal(r_dup, reference(twice.cpu().numpy().view(np.uint64), 100)), "duplicate keys across lists"


def test_score_batch_entries_agree(eng):
    """Engine.score_batch (mmf_score_batch: one asynchronous call, device tensors) == Engine.score_batch_host (host buffers)
    == submit / collect with two batches in flight == cosine + search + verdict_assemble, with and without a modality mask"""
    n_rows, b, k = 50000, 200, 5
    vault_rows = synth.vault_rows(n_rows, seed=41)
    vault = mmf_b200.TruthVault(eng, vault_rows, None, mode="fp32")
    eng.fusion_load(synth.fusion_state_dict(1))
    q, _, _ = synth.queries(b, n_rows, seed=42, plant_frac=0.3, vault_seed=41)
    text, _ = synth.caption_image_pairs(b, seed=43)
    head = synth.head_scores(b, seed=44)
    keys = ("clip_similarity", "vault_discrepancy", "vault_scores", "vault_rows", "scores", "probs", "verdict", "confidence")
    for modality in (None, (np.arange(b) % 4).astype(np.uint8)):
        want = mmf_b200.score_batch(eng, vault, text, q, head, modality, k)
        sim = eng.cosine_pairs(text, q)
        vs, vr, disc = eng.vault_search(q, k)
        x, probs, verdict, conf = eng.verdict_assemble(head, modality, sim, disc)
        three = dict(zip(keys, (sim, disc, vs, vr, x, probs, verdict, conf)))
        got = eng.score_batch_host(text, q, head, modality, k)
        eng.score_batch_submit(0, text, q, head, modality, k)
        eng.score_batch_submit(1, text[::-1].copy(), q[::-1].copy(), head[::-1].copy(), None if modality is None else modality[::-1].copy(), k)
        first, second = eng.score_batch_collect(0), eng.score_batch_collect(1)
        for key in keys:
            assert np.array_equal(got[key], npy(want[key]), equal_nan=True), key
            assert np.array_equal(npy(three[key]), npy(want[key]), equal_nan=True), key
            assert np.array_equal(first[key], got[key], equal_nan=True), key
            assert np.array_equal(second[key], got[key][::-1], equal_nan=True), key
        streamed = list(eng.score_stream([(text, q, head, modality)] * 3, top_k=k)) if hasattr(eng, "score_stream") else []
        assert all(np.array_equal(r["probs"], got["probs"]) for r in streamed)
    # the scalar rule of analyze() per row (misinfo_forensics.py:884-899) on the masked batch
    m = (np.arange(b) % 4).astype(np.uint8)
    out = eng.score_batch_host(text, q, head, m, k)
    w = synth.fusion_state_dict(1)
    for i in range(0, b, 7):
        s = dict(zip(oracle.FUSION_ORDER, map(float, out["scores"][i])))
        ref = oracle.assemble_verdict(w, s, bool(m[i] & 1), bool(m[i] & 2))
        assert abs(out["probs"][i, 1] - ref["fake_probability"]) <= FP32_TOL and out["verdict"][i] == ref["verdict"]
    # pinned torch tensors in, and no vault loaded -> zero discrepancy / no rows
    eng.vault_unload()
    pin = (lambda t: t) if DOUBLE else (lambda t: t.pin_memory())
    got = eng.score_batch_host(pin(torch.from_numpy(text)), pin(torch.from_numpy(q)), pin(torch.from_numpy(head)), None, k)
    assert np.all(got["vault_rows"] == -1) and np.all(got["vault_discrepancy"] == 0) and np.all(np.isnan(got["vault_scores"]))
    if not DOUBLE:
        with pytest.raises(mmf_b200.MMFError):
            eng.score_batch_collect(0)                               # nothing pending


def test_peer_exchange_single_rank_loopback(eng):
    """csrc/exchange.cu with world = 1 (the only rank pushes into its own buffer): push, flag, wait-merge must give
    the plain search result; three calls in a row cover both buffer parities, and the batch size changes between
    exchanges (the parities live at fixed offsets: ADVICE r1)"""
    if DOUBLE:
        pytest.skip("peer memory")
    n_rows, nq = 30000, 40
    vault = synth.vault_rows(n_rows, seed=51)
    q, _, _ = synth.queries(nq, n_rows, seed=52, plant_frac=0.5, vault_seed=51)
    eng.vault_load(vault, mode="fp32")
    need = eng.exchange_layout(1, nq, 100)
    buf = torch.zeros(need // 8 + 16, dtype=torch.int64, device=DEV)
    eng.exchange_attach(0, 1, [buf.data_ptr()], buf.numel() * 8)
    try:
        for fused in (1, 0):
            eng.set_option("fused_push", fused)
            for k in (10, 100, 5):
                for n in (nq, 17, nq):
                    want = [npy(t) for t in eng.vault_search(q[:n], k)]
                    got = [npy(t) for t in eng.vault_search_exchange(q[:n], k, k)]
                    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, want)), (k, n, fused)
        with pytest.raises(mmf_b200.MMFError):
            eng.vault_search_exchange(np.zeros((100000, 512), np.float32), 100, 100)     # does not fit the attached buffer
    finally:
        eng.set_option("fused_push", 1)
        eng.exchange_detach()


def test_sharded_world1_entry_is_the_plain_search(eng):
    """mmf_shard_init(world = 1) needs no NCCL; mmf_vault_search_sharded then equals mmf_vault_search"""
    if DOUBLE:
        pytest.skip("library entry points")
    n_rows, nq, k = 20000, 33, 10
    vault = synth.vault_rows(n_rows, seed=55)
    q, _, _ = synth.queries(nq, n_rows, seed=56, plant_frac=0.5, vault_seed=55)
    e2 = mmf_b200.Engine("cuda:0")
    try:
        e2.vault_load(vault, mode="fp32")
        with pytest.raises(mmf_b200.MMFError):
            e2.vault_search_sharded(q, k)                        # shard_init has not been called
        e2.shard_init(0, 1)
        assert e2.shard_info()[:2] == (0, 1)
        a = [npy(t) for t in e2.vault_search_sharded(q, k)]
        b = [npy(t) for t in e2.vault_search(q, k)]
        assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))
        g = e2.shard_all_gather(e2.vault_search_candidates(q, k), 1)
        m = [npy(t) for t in e2.topk_merge(g, k)]
        assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(m, b))
    finally:
        e2.close()


# ------------------------------------------------------------------------------ 8f rank 3 / 4 on the kernels
def test_sharded_vault_directory_into_the_kernels(tmp_path):
    """8f rank 3: a vault written as a sharded raw directory (vault_io.save_vault_dir), opened rank by rank
    (open_vault_dir) and searched shard by shard + candidate merge == the unsharded search of the same rows == the oracle;
    metadata records come back by global row id"""
    from mmf_b200 import vault_io
    n_rows, nq, k = 25_003, 21, 5
    rows = synth.vault_rows(n_rows, seed=81).astype(np.float16)          # a CUDA-built vault stores fp16 rows
    meta = [{"title": f"headline {i}", "url": f"https://example.org/{i}", "date": "2024-01-01"} for i in range(n_rows)]
    vault_io.save_vault_dir(str(tmp_path / "vault"), rows, meta, rows_per_shard=7_000)
    q, prow, _ = synth.queries(nq, n_rows, seed=82, plant_frac=0.5, vault_seed=81)
    world = 3
    e = _engine()
    try:
        e.vault_load(rows, mode="fp32")
        full = [npy(t) for t in e.vault_search(q, k)]
        packed = []
        for r in range(world):
            shard, lo, total = vault_io.open_vault_dir(str(tmp_path / "vault"), rank=r, world=world)
            assert total == n_rows and shard.dtype == np.float16
            e.vault_load(np.ascontiguousarray(shard), mode="fp32", row_offset=lo)
            packed.append(e.vault_search_candidates(q, k).clone())
        scores, rws, disc = [npy(t) for t in e.topk_merge(torch.stack(packed), k)]
        assert np.array_equal(rws, full[1]) and np.array_equal(scores, full[0]) and np.array_equal(disc, full[2])
        ri, rs, _ = oracle.vault_search_batched(rows.astype(np.float32), q, k)
        assert_topk(rws, scores, ri, rs, FP32_TOL, "sharded directory vs oracle")
        got_meta = vault_io.read_metadata(str(tmp_path / "vault"), rows=[int(x) for x in rws[:, 0]])
        assert [m["title"] for m in got_meta] == [f"headline {int(x)}" for x in rws[:, 0]]
    finally:
        e.close()


def test_search_similar_articles_dropin_on_gpu(tmp_path):
    """8f rank 4: search_similar_articles through the real kernels vs the fixtures of the reference's own function; the
    caller's engine keeps its resident vault (the database lives in a private handle)"""
    import fakes
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "similar.npz"))
    with open(os.path.join(GOLDEN, "similar_cases.json")) as fh:
        c = json.load(fh)
    k = c["top_k"]
    db = {"article_ids": c["article_ids"], "text_contents": c["text_contents"], "image_paths": c["image_paths"],
          "image_embeddings": g["image_embeddings"], "text_embeddings": g["text_embeddings"]}
    engine = _engine()
    engine.vault_load(np.eye(6, 512, dtype=np.float32), mode="fp32")
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
    assert engine.vault_rows == 6
    engine.close()
