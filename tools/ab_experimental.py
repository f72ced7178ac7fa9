#!/usr/bin/env python
"""A/B timing of the kernel variants behind environment switches (one GPU), through the Python API.
(tools/cabi_selftest does the same comparisons without torch, in seconds.)

    python tools/ab_experimental.py            # both experiments, prints one JSON line each
    python tools/ab_experimental.py hist       # C4 shard shapes: bucket pool vs histogram bound (top-100, bf16)
    python tools/ab_experimental.py screen     # C2: 3-pass fp32-exact vs screened search (top-10, fp32-exact)

CUDA events on the launching stream, 5 warm-ups, 20 timed searches per arm, arms interleaved so that clock /
power drift hits both.  Results of the two arms are compared bit for bit (hist) / against the streaming
kernel (screen) before anything is timed."""
import contextlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmf_b200  # noqa: E402


@contextlib.contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    os.environ.update({k: str(v) for k, v in kw.items()})
    try:
        yield
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)


def timed(fn, reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def ab(eng, q, k, arms, rounds=3):
    for name, kw in arms:
        with env(**kw):
            for _ in range(5):
                eng.vault_search(q, k, algo="mma")
    best = {name: float("inf") for name, _ in arms}
    for _ in range(rounds):
        for name, kw in arms:
            with env(**kw):
                best[name] = min(best[name], timed(lambda: eng.vault_search(q, k, algo="mma")))
    return best


def run_hist(eng):
    g = torch.Generator(device="cuda").manual_seed(1)
    out = {}
    for rows in (1_250_000, 2_500_000, 5_000_000):
        vault = torch.randn(rows, 512, device="cuda", generator=g)
        eng.vault_load(vault, mode="bf16")
        del vault
        q = torch.randn(4096, 512, device="cuda", generator=g)
        with env(MMF_MMA_BOUND="pool"):
            base = eng.vault_search(q, 100, algo="mma")
        with env(MMF_MMA_BOUND="hist"):
            got = eng.vault_search(q, 100, algo="mma")
        same = all(torch.equal(a, b) for a, b in zip(base, got))
        ms = ab(eng, q, 100, [("pool", {"MMF_MMA_BOUND": "pool"}), ("hist", {"MMF_MMA_BOUND": "hist"})])
        flops = 2.0 * 4096 * rows * 512
        out[str(rows)] = {"identical": same, "ms": ms, "tflops": {n: flops / (t * 1e-3) / 1e12 for n, t in ms.items()}}
    print(json.dumps({"experiment": "hist bound, 4096 queries x bf16 shard, top-100", "rows": out}))


def run_screen(eng):
    g = torch.Generator(device="cuda").manual_seed(2)
    vault = torch.randn(1_000_000, 512, device="cuda", generator=g)
    eng.vault_load(vault, mode="fp32")
    q = torch.randn(256, 512, device="cuda", generator=g)
    q[:26] = vault[torch.arange(26, device="cuda") * 37_003] + 0.4 * q[:26]
    del vault
    exact = eng.vault_search(q, 10, algo="stream")
    with env(MMF_MMA_SCREEN="1"):
        got = eng.vault_search(q, 10, algo="mma")
    same = all(torch.equal(a, b) for a, b in zip(exact, got))
    ms = ab(eng, q, 10, [("3pass", {"MMF_MMA_SCREEN": "0"}), ("screen", {"MMF_MMA_SCREEN": "1"}),
                         ("screen+12stages", {"MMF_MMA_SCREEN": "1", "MMF_MMA_STAGES": "12"}),
                         ("screen+prefetch", {"MMF_MMA_SCREEN": "1", "MMF_MMA_PREFETCH": "1"}),
                         ("screen+prefetch+12stages", {"MMF_MMA_SCREEN": "1", "MMF_MMA_PREFETCH": "1", "MMF_MMA_STAGES": "12"}),
                         ("screen+lean", {"MMF_MMA_SCREEN": "1", "MMF_MMA_LEAN": "1"}),
                         ("screen+prefetch+12stages+lean", {"MMF_MMA_SCREEN": "1", "MMF_MMA_PREFETCH": "1",
                                                            "MMF_MMA_STAGES": "12", "MMF_MMA_LEAN": "1"})])
    print(json.dumps({"experiment": "screened fp32-exact search, 256 queries x 1M rows, top-10",
                      "identical_to_stream_kernel": same, "ms": ms,
                      "queries_per_s": {n: 256 / (t * 1e-3) for n, t in ms.items()},
                      "algorithmic_GBps": {n: 2.048 / (t * 1e-3) for n, t in ms.items()}}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["hist", "screen"]
    eng = mmf_b200.Engine("cuda:0")
    if "screen" in which:
        run_screen(eng)
    if "hist" in which:
        run_hist(eng)
