"""Property tests of the CPU oracle: the invariants the GPU parity tests lean on --
scale invariance and symmetry of the cosine, row-permutation equivariance and shard invariance of the vault
search, the as-shipped per-query algorithm == its batched restatement, the threshold edges, softmax sanity of
the fusion judge.  hypothesis draws the shapes / seeds; the arrays come from numpy generators."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

import oracle

SET = dict(max_examples=25, deadline=None, derandomize=True, database=None)   # same examples every run


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 40), dim=st.sampled_from([8, 100, 512]),
       scale=st.floats(1e-3, 1e3, allow_nan=False))
def test_cosine_scale_invariant_symmetric_bounded(seed, n, dim, scale):
    r = np.random.default_rng(seed)
    a = r.standard_normal((n, dim)).astype(np.float32)
    b = r.standard_normal((n, dim)).astype(np.float32)
    s = oracle.cosine_pairs(a, b)
    assert np.allclose(s, oracle.cosine_pairs(b, a), atol=2e-6)
    assert np.allclose(s, oracle.cosine_pairs(a * np.float32(scale), b), atol=2e-6)
    assert np.all(np.abs(s) <= 1 + 1e-5)
    assert np.allclose(oracle.cosine_pairs(a, a), 1.0, atol=1e-5)
    assert [oracle.clip_match_label(float(x)) for x in s] == ["Match" if x >= 0.25 else "Mismatch" for x in s]


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(2, 300), nq=st.integers(1, 6), k=st.integers(1, 12))
def test_vault_search_permutation_and_shard_invariance(seed, n, nq, k):
    r = np.random.default_rng(seed)
    v = r.standard_normal((n, 32)).astype(np.float32) * r.uniform(0.1, 5, (n, 1)).astype(np.float32)
    q = r.standard_normal((nq, 32)).astype(np.float32)
    idx, sc, disc = oracle.vault_search_batched(v, q, k)
    kk = min(k, n)
    assert idx.shape == (nq, kk) and np.all(sc[:, :-1] >= sc[:, 1:])
    # permuting the rows permutes the indices (scores are distinct with probability 1); the CPU BLAS is not
    # bit-invariant under a change of matrix shape / row order, so scores agree to rounding only
    perm = r.permutation(n)
    idx_p, sc_p, _ = oracle.vault_search_batched(v[perm], q, k)
    assert np.array_equal(perm[idx_p], idx) and np.allclose(sc_p, sc, atol=2e-6)
    # any contiguous sharding + candidate merge == the unsharded search
    cuts = sorted(set(r.integers(1, n, size=min(3, n - 1)).tolist()))
    bounds = [0] + cuts + [n]
    parts = [oracle.vault_search_batched(v[lo:hi], q, k, row_offset=lo) for lo, hi in zip(bounds, bounds[1:])]
    mi, ms = oracle.merge_topk([p[0] for p in parts], [p[1] for p in parts], k)
    assert np.array_equal(mi, idx) and np.allclose(ms, sc, atol=2e-6)
    # the as-shipped per-query algorithm (misinfo_forensics.py:438-464) gives the same answer
    for i in range(nq):
        si, ss, sd = oracle.vault_search_as_shipped(v, q[i], k)
        assert np.array_equal(si, idx[i]) and np.allclose(ss, sc[i], atol=2e-6)
        assert (sd > 0) == (disc[i] > 0) or abs(float(ss[0]) - 0.85) < 1e-5


@settings(**SET)
@given(x=st.floats(-1.5, 1.5, allow_nan=False, width=32))
def test_discrepancy_rule_edges(x):
    s = np.float32(x)
    d = oracle.discrepancy_rule(np.array([s]))[0]
    assert d == (s if float(s) > 0.85 else np.float32(0.0))
    assert oracle.discrepancy_rule(np.array([np.float32(0.85)]))[0] == np.float32(0.85)        # 0.85f > 0.85 as a double
    assert oracle.discrepancy_rule(np.array([np.nextafter(np.float32(0.85), np.float32(0))]))[0] == 0
    assert oracle.discrepancy_rule(np.array([np.float32("nan")]))[0] == 0


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 50))
def test_fusion_forward_is_a_distribution_and_matches_verdict(seed, n):
    torch.manual_seed(seed % (2 ** 31))
    layer = torch.nn.Sequential(torch.nn.Linear(5, 64), torch.nn.ReLU(), torch.nn.Dropout(0.2), torch.nn.Linear(64, 32),
                                torch.nn.ReLU(), torch.nn.Linear(32, 2)).eval()
    w = {k: v.detach() for k, v in layer.state_dict().items()}
    x = np.random.default_rng(seed).uniform(-0.5, 1.5, (n, 5)).astype(np.float32)
    p = oracle.fusion_forward(w, x)
    assert p.shape == (n, 2) and np.allclose(p.sum(1), 1.0, atol=1e-6) and np.all(p >= 0)
    with torch.no_grad():
        ref = torch.softmax(layer(torch.from_numpy(x)), dim=1).numpy()
    assert np.allclose(p, ref, atol=1e-7)
    v = oracle.fusion_verdict(w, dict(zip(oracle.FUSION_ORDER, map(float, x[0]))))
    assert v["verdict"] == int(v["fake_probability"] > 0.5)
    assert abs(v["fake_probability"] - p[0, 1]) < 1e-6 and abs(v["fake_probability"] + v["real_probability"] - 1) < 1e-6
