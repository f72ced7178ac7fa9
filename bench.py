#!/usr/bin/env python
"""Benchmark of the scoring hot path (BASELINE.json metric: vault queries/s + roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4] [--records auto|none|c1,c3,c4,kernels,c5]
                  [--impl reference]

A step is one pass of the hot path over one batch of synthetic embeddings:
caption/image cosine -> Truth-Vault top-k + discrepancy -> score assembly + fusion judge.

The HEADLINE line (metric / value / roofline / e2e / cpu_baseline) is the workload given by --workload, default c2, so
that BENCH_rNN / SCALE_rNN stay comparable from round to round.  The same JSON line carries, under "records", the
other BASELINE.json configs measured in the same run (each with its own timing, roofline and parity gate):

  c1  1000 queries vs a 100k-row fp32-exact vault, top-10, + fusion judge on the (1000,5) scores     [N = 1]
  c2  256 queries vs a 1M x 512 fp32-exact vault, top-10, + cosine + fusion (configs[1]).  N > 1: every rank is an
      independent replica with its own 256-query batch (queries shard with no collective) -> "scaling": "weak"
  c3  batch-1 latency mode: 1 query vs the 1M fp32-exact vault, top-10, p50 / p99 over 1000 distinct queries [N = 1]
  c4  4096 queries vs a 10M x 512 bf16 vault, top-100.  N = 1: the whole vault on one GPU (the strong-scaling base);
      N > 1: ROW-SHARDED over the N ranks, one ncclAllGather of the per-shard candidates (owned by the library,
      csrc/shard.cu) + merge, with the local search, the all-gather and the merge timed separately and the result
      checked bit for bit against an unsharded search on rank 0                                        [every N]
  kernels  K1 cosine GB/s sweep up to 1M pairs, K5 fusion judge microseconds                            [N = 1]
  c5  full batched analyze: random-init RoBERTa / EfficientNet / CLIP producers (stock PyTorch) feeding the scoring
      kernels, data-parallel replicas; samples/s and the encoder / hot-path split of the step          [every N]

`--impl reference` times the reference's own CPU algorithm (the oracle port of misinfo_forensics.py:438-464: per-query
renormalisation of the whole vault in NumPy) on the host cores, rank 0 only.
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c1": dict(q=1000, rows=100_000, k=10, mode="fp32", desc="1000 queries vs 100k-row fp32-exact vault, top-10 + caption/image cosine + fusion judge"),
    "c2": dict(q=256, rows=1_000_000, k=10, mode="fp32", desc="256 queries vs 1M-row fp32-exact vault, top-10 + caption/image cosine + fusion judge"),
    "c3": dict(q=1, rows=1_000_000, k=10, mode="fp32", desc="batch-1 latency: 1 query vs 1M-row fp32-exact vault, top-10"),
    "c4": dict(q=4096, rows=10_000_000, k=100, mode="bf16", desc="4096 queries vs 10M-row bf16 vault row-sharded over the ranks, top-100, all-gather merge"),
}
# committed `ncu --set full` summaries (profiles/<name>.ncu_summary.txt) of the dominant kernel of each workload
NCU_SUMMARY = {"c2": "r02_c2_screen", "c3": "r02_c3_stream_tma", "c4": "r02_c4shard_hist", "c1": None}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def ncu_traffic(summary):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` summary of this workload / kernel variant (profiles/<summary>.ncu_summary.txt), else None.
    NOT measured in this run (a number taken under a profiler is never a bench value): the line says so."""
    if not summary:
        return None
    path = os.path.join(ROOT, "profiles", f"{summary}.ncu_summary.txt")
    if not os.path.exists(path):
        return None
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total, seen = 0.0, 0
    with open(path) as fh:
        for line in fh:
            f = line.split()
            if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[2] in mult and seen < 2:
                total += float(f[1]) * mult[f[2]]
                seen += 1
    return total if seen == 2 else None


class ClockSampler:
    """nvidia-smi clocks / power / throttle reasons UNDER THE LOAD of the timed steps.  nvidia-smi needs ~0.1 s to print its
    first line and a C2 timed region lasts ~8 ms, so: the process is started before the warm-up, `window_start()` is
    called right before the timed region, and `hold_load()` keeps launching the very same steps (untimed, after the
    closing event) until the window is 0.5 s long.  Only lines stamped inside the window are kept."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    MIN_WINDOW_S = 0.5

    def __init__(self, index: int, period_ms: int = 20):
        self.proc, self.t0, self.t1, self.wall0, self.held = None, None, None, None, False
        self.spawned = time.perf_counter()
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def window_start(self):
        self.t0, self.wall0 = datetime.datetime.now(), time.perf_counter()

    def hold_load(self, step, sync):
        """keep the load of the timed region going until the sampling window is MIN_WINDOW_S long"""
        if self.proc is None or self.wall0 is None:
            return
        self.held = True
        # (nvidia-smi prints its first line some tenths of a second after it was spawned: the window also lasts until then)
        while time.perf_counter() - self.wall0 < self.MIN_WINDOW_S or time.perf_counter() - self.spawned < 2 * self.MIN_WINDOW_S:
            for _ in range(8):
                step()
            sync()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.t1 = datetime.datetime.now()
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        rows = []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((ts, float(f[1]), float(f[2]), float(f[3]), f[4:8]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        window = ("timed region + the same steps continued untimed to %.1f s" % self.MIN_WINDOW_S) if self.held else "timed region"
        if not inside:                      # clock skew / no line in the window: say so instead of pretending
            inside, window = rows, "whole sampler lifetime (no line was stamped inside the load window)"
        sm, mx, pw, reasons = [r[1] for r in inside], [r[2] for r in inside], [r[3] for r in inside], set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None, "samples": len(sm),
                "window": window, "reasons": sorted(reasons)}


def host_info():
    """CPU model / thread count / BLAS of the box the CPU baseline ran on; best effort, never raises."""
    info = {}
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    info["cpu_model"] = line.split(":", 1)[1].strip()
                    break
    except Exception:
        pass
    try:
        info["torch_threads"] = int(torch.get_num_threads())
        info["blas"] = "mkl" if torch.backends.mkl.is_available() else "not mkl"
        info["numpy"] = np.__version__
        info["torch"] = torch.__version__
    except Exception:
        pass
    return info


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def stats(ms_list):
    """median / min / max / mean of per-step device times (the timed region of a short step is a few ms in all)"""
    s = sorted(ms_list)
    return {"median": s[len(s) // 2], "min": s[0], "max": s[-1], "mean": sum(s) / len(s), "n": len(s)}


# ----------------------------------------------------------------------------- CPU side (oracle)
def cpu_reference_step(vault_host, q_host, text_host, img_host, head, fusion_w, n_sample, k):
    """The reference's own algorithm on `n_sample` samples of the batch: cosine (:399-404),
    as-shipped vault search (:438-464, whole-vault renormalisation per query) and fusion_verdict."""
    import oracle
    for i in range(n_sample):
        sim = oracle.cosine_pairs(text_host[i:i + 1], img_host[i:i + 1], scalar_loop=True)[0]
        _, _, disc = oracle.vault_search_as_shipped(vault_host, q_host[i], k)
        oracle.fusion_verdict(fusion_w, {"ai_score": head[i, 0], "misinfo_score": head[i, 1], "deepfake_score": head[i, 2],
                                         "clip_similarity": float(sim), "vault_discrepancy": disc})


def cpu_batched_step(vault_norm_t, q_host, k):
    """The 'batched torch restatement' of BASELINE.md 4(2): normalise once, Qn @ Vn.T, torch.topk."""
    q = torch.from_numpy(q_host)
    qn = q / q.norm(dim=-1, keepdim=True)
    s = qn @ vault_norm_t.T
    return torch.topk(s, k, dim=1)


def cpu_baseline_for(vault_host, mode, qh, th, hh, K, budget_s):
    """The as-shipped reference algorithm (oracle port) on a bounded sample of the step's queries + the batched
    restatement on a bounded sample of the vault rows, on this box's host cores."""
    import oracle
    from mmf_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    if mode == "bf16":
        vault_host = torch.from_numpy(vault_host).bfloat16().float().numpy()
    fw = synth.fusion_state_dict()
    Q, n_local = qh.shape[0], vault_host.shape[0]
    n_sample = 0
    t0 = time.perf_counter()
    while True:
        cpu_reference_step(vault_host, qh[n_sample:n_sample + 1], th[n_sample:n_sample + 1], qh[n_sample:n_sample + 1],
                           hh[n_sample:n_sample + 1], fw, 1, K)
        n_sample += 1
        if time.perf_counter() - t0 > budget_s or n_sample >= Q:
            break
    dt = time.perf_counter() - t0
    cpu = {"value": n_sample / dt, "unit": "queries/s", "cores": os.cpu_count() or 1, "kind": "port",
           "sample": f"{n_sample} of the {Q} queries of one step, each vs the full {n_local}-row vault, reference "
                     "algorithm as shipped (NumPy, whole-vault renormalisation per query)"}
    sub = min(n_local, 200_000)
    vn = torch.from_numpy(oracle.vault_normalise(vault_host[:sub]).astype(np.float32))
    cpu_batched_step(vn, qh, min(K, sub))
    t0 = time.perf_counter()
    cpu_batched_step(vn, qh, min(K, sub))
    dtb = (time.perf_counter() - t0) * (n_local / sub)
    cpu["batched_restatement"] = {"value": Q / dtb, "unit": "queries/s", "threads": torch.get_num_threads(),
                                  "sample": f"all {Q} queries vs {sub} vault rows, time scaled to {n_local} rows; "
                                            "normalise once + Qn@Vn.T + torch.topk"}
    cpu["host"] = host_info()
    return cpu


def run_reference_arm(args, wl):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from mmf_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    rows, k = wl["rows"], wl["k"]
    g = np.random.default_rng(synth.VAULT_SEED)
    vault = g.standard_normal((rows, 512), dtype=np.float32)
    if wl["mode"] == "bf16":
        vault = torch.from_numpy(vault).bfloat16().float().numpy()
    n_sample = 1
    q = np.random.default_rng(synth.QUERY_SEED).standard_normal((max(n_sample, 1), 512), dtype=np.float32)
    a, b = synth.caption_image_pairs(max(n_sample, 8))
    head = synth.head_scores(max(n_sample, 8))
    fw = synth.fusion_state_dict()
    for _ in range(args.warmup):
        cpu_reference_step(vault, q, a, b, head, fw, n_sample, k)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(vault, q, a, b, head, fw, n_sample, k)
    dt = (time.perf_counter() - t0) / args.steps
    value = n_sample / dt
    cores = os.cpu_count() or 1
    sample = f"{n_sample} query per step vs the full {rows}-row vault, as-shipped per-query renormalisation (NumPy)"
    print(json.dumps({
        "impl": "reference", "metric": "vault queries/s", "value": value, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "queries_per_step": n_sample, "vault_rows_total": rows, "vault_rows_per_gpu": rows,
                   "dim": 512, "top_k": k, "vault_mode": wl["mode"], "algo": "reference as shipped (NumPy, per-query renormalisation)",
                   "parallelism": "host CPU, %d cores (BLAS threads)" % cores,
                   "sample": "each step = %d query of the workload's %d-query batch against the full vault" % (n_sample, wl["q"])},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample, "host": host_info()},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------- GPU side
class Ctx:
    """what every workload needs: ranks, device, peaks, the barrier"""

    def __init__(self, args):
        import torch.distributed as dist
        self.args = args
        self.rank, self.local_rank, self.world = dist_env()
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = dist
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.hbm_peak, self.tf_peak, self.tf_sustained, self.peak_kind = peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, values):
        if self.world == 1:
            return list(values)
        t = torch.tensor(list(values), device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()


def gen_vault_rows(dev, total_rows, lo, hi):
    """Rows [lo, hi) of the synthetic vault, generated on the device in 1M-row slabs seeded by the GLOBAL slab id, so a
    shard never depends on the others; un-normalised on purpose: vault_load normalises."""
    from mmf_b200 import synth
    slab = 1_000_000
    parts = []
    for s0 in range(lo - lo % slab, hi, slab):
        g = torch.Generator(device=dev).manual_seed(synth.VAULT_SEED + s0 // slab)
        blk = torch.randn(min(slab, total_rows - s0), 512, device=dev, generator=g)
        parts.append(blk[max(lo, s0) - s0:min(hi, s0 + slab) - s0])
    return torch.cat(parts) if len(parts) > 1 else parts[0]


def make_batch(vault_rows, n_local, Q, seed, dev):
    """Q queries: N(0,1)*3 with every 10th replaced by a planted near-duplicate of a uniformly drawn row of
    `vault_rows` at the cosines of synth.PLANT_COSINES (both sides of the 0.85 rule); caption embeddings and head
    scores from the seeded generators of mmf_b200.synth.  Host tensors are PINNED (the e2e leg copies from them)."""
    from mmf_b200 import synth
    gq = np.random.default_rng(seed)
    q_host = torch.from_numpy(gq.standard_normal((Q, 512), dtype=np.float32) * 3.0)
    n_plant = max(1, Q // 10)
    pick = torch.from_numpy(gq.integers(0, n_local, n_plant))
    planted = vault_rows[pick.to(dev)].cpu()
    noise = torch.from_numpy(gq.standard_normal((n_plant, 512), dtype=np.float32))
    cosv = torch.tensor([synth.PLANT_COSINES[i % 5] for i in range(n_plant)])
    pn = planted / planted.norm(dim=1, keepdim=True)
    noise = noise - (noise * pn).sum(1, keepdim=True) * pn
    noise = noise / noise.norm(dim=1, keepdim=True)
    q_host[:n_plant] = (cosv[:, None] * pn + torch.sqrt(1 - cosv ** 2)[:, None] * noise) * 2.0
    a_np, _ = synth.caption_image_pairs(Q, seed=seed + 1)
    return {"q": q_host, "text": torch.from_numpy(a_np), "head": torch.from_numpy(synth.head_scores(Q, seed=seed + 2)),
            "pick": pick, "cosv": cosv, "n_plant": n_plant}


def check_planted(rows0, scores0, batch, row_offset, tol, what):
    n = batch["n_plant"]
    assert torch.equal(rows0[:n].cpu(), batch["pick"] + row_offset), f"{what}: planted rows not recovered"
    assert torch.allclose(scores0[:n].cpu(), batch["cosv"], atol=tol), f"{what}: planted cosines off"


def hbm_roofline(ctx, wl_name, n_rows, Q, K, mode, search_ms, screened, rows_overridden):
    """HBM roofline of the vault search.  `achieved` counts the bytes the search has to STREAM: N*D*elem (SURVEY.md 8d's
    per-query figure) -- except for the screened fp32-exact search, which by construction reads only the fp16 hi planes
    (N*D*2) plus a few dozen rows per query for the exact re-scoring; crediting it the full N*D*4 would report 1.2x the
    measured HBM peak for bytes that never move, so that figure is kept aside as `frac_vs_survey_bytes`."""
    elem = 2 if mode == "bf16" else 4
    survey_bytes = float(n_rows) * 512 * elem
    nbytes = float(n_rows) * 512 * 2 if screened else survey_bytes
    achieved = nbytes / (search_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": ctx.hbm_peak, "unit": "GB/s", "frac": achieved / ctx.hbm_peak,
            "traffic": None, "kernel": "vault search (query prep + search + top-k select)",
            "algorithmic_bytes_per_launch": nbytes, "survey_bytes_per_launch": survey_bytes,
            "frac_vs_survey_bytes": survey_bytes / (search_ms * 1e-3) / 1e9 / ctx.hbm_peak,
            "tensor_tflops_algorithmic": 2.0 * Q * n_rows * 512 / (search_ms * 1e-3) / 1e12}
    if Q >= 16:
        # fp32-exact tcgen05 path.  top_k <= 16 (default): SCREENED search -- one f16 pass over the hi planes (half of the
        # stored bytes, a third of the MMA work), then exact fp32 re-scoring of the rows inside the proven error band
        # (DESIGN.md 9).  Otherwise: 3 f16 passes.  Only 2*Q*N*D flop are credited either way.
        passes = 3 if (mode == "fp32" and not screened) else 1
        issued = passes * roof["tensor_tflops_algorithmic"]
        roof.update({"mma_passes": passes, "tensor_tflops_issued": issued, "tensor_frac_issued": issued / ctx.tf_peak,
                     "tensor_frac_sustained": issued / ctx.tf_sustained,
                     "variant": "screened (hi-plane pass + exact re-scoring)" if screened else "%d-pass" % passes,
                     "bytes_streamed_per_launch": nbytes, "hbm_frac_streamed": roof["frac"],
                     "note": ("the screened search streams only the fp16 hi planes (N*D*2 bytes, what `achieved` counts) and re-scores "
                              "the few rows inside the error band from hi+lo; at 256 queries it sits ON the ridge (256 flop/B against "
                              "bf16 peak / HBM peak ~ 250): the MMAs of one pass over the hi planes take longer at the measured tensor "
                              "peak than the bytes take at the measured HBM peak, and under tensor load the SM clock is ~1.4-1.5 GHz "
                              "(clock64 / globaltimer inside the kernel, profiles/r02_screen_triage_clock.log), so `tensor_frac_issued` / "
                              "`tensor_frac_sustained` bound this search as tightly as `frac` does") if screened else
                             "tensor-bound once the MMA passes are counted; the HBM fraction is the algorithmic roofline"})
    summary = NCU_SUMMARY.get(wl_name)
    traffic = ncu_traffic(summary) if ctx.world == 1 and not rows_overridden else None
    roof["traffic"] = traffic
    roof["traffic_source"] = (f"profiles/{summary}.ncu_summary.txt (committed ncu --set full capture of this kernel; NOT measured "
                              "in this run)") if traffic is not None else None
    roof["peak_source"] = f"MEASURED_PEAKS.json ({ctx.peak_kind})"
    roof["kernel_ms"] = search_ms
    return roof


def run_replica(ctx, wl_name, wl, headline, algo="auto", rows_overridden=False, keep=None):
    """c1 / c2 / c3: every rank is an independent replica (own vault copy, own query batch, no collective).
    Returns (line-parts dict, objects to reuse: engine, vault, rows)."""
    import mmf_b200
    from mmf_b200 import synth
    args, dev, rank, world = ctx.args, ctx.dev, ctx.rank, ctx.world
    Q, K, n_rows, mode = wl["q"], wl["k"], wl["rows"], wl["mode"]
    if keep is not None:
        eng, vault, vault_rows = keep
    else:
        eng = mmf_b200.Engine(dev)
        vault_rows = gen_vault_rows(dev, n_rows, 0, n_rows)
        vault = mmf_b200.TruthVault(eng, vault_rows, None, mode=mode)
        eng.fusion_load(synth.fusion_state_dict())
    batch = make_batch(vault_rows, n_rows, Q, synth.QUERY_SEED + rank + (0 if wl_name == "c2" else 1000), dev)
    text_host, img_host, head_host = batch["text"].pin_memory(), batch["q"].pin_memory(), batch["head"].pin_memory()
    text_dev, img_dev, head_dev = text_host.to(dev), img_host.to(dev), head_host.to(dev)
    tol = 1e-2 if mode == "bf16" else 1e-5

    # parity gate of the timed path before timing it: planted rows at their cosine, one-call path == three-call path
    out = eng.score_batch(text_dev, img_dev, head_dev, None, K, algo=algo)
    torch.cuda.synchronize()
    check_planted(out["vault_rows"][:, 0], out["vault_scores"][:, 0], batch, 0, tol, wl_name)

    def step_device():
        # three library calls, no torch arithmetic; the events bracket the vault search alone (dominant kernel)
        sim = eng.cosine_pairs(text_dev, img_dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vs, vr, disc = vault.search(img_dev, K, mmf_b200.VAULT_THRESHOLD, algo)
        e1.record()
        x, probs, verdict, conf = eng.verdict_assemble(head_dev, None, sim, disc)
        return (e0, e1), (vr, vs, probs)

    _, (vr3, vs3, probs3) = step_device()
    torch.cuda.synchronize()
    assert torch.equal(vr3, out["vault_rows"]) and torch.equal(vs3, out["vault_scores"]) and torch.equal(probs3, out["probs"]), \
        f"{wl_name}: mmf_score_batch differs from cosine + search + verdict_assemble"

    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step_device()
    ctx.barrier()
    launches0 = eng.launch_count
    if sampler:
        sampler.window_start()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    search_ev = []
    marks[0].record()
    for i in range(args.steps):
        ev, _ = step_device()
        search_ev.append(ev)
        marks[i + 1].record()
    ctx.barrier()
    launches = eng.launch_count - launches0
    if sampler:
        sampler.hold_load(step_device, torch.cuda.synchronize)
    step_ms = marks[0].elapsed_time(marks[-1]) / args.steps
    per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
    per_search = [a.elapsed_time(b) for a, b in search_ev]
    search_ms = statistics.mean(per_search)
    clocks = sampler.stop() if sampler else None

    # e2e: the same work through the public host-buffer API (Engine.score_batch_submit / _collect -> mmf_score_batch_submit):
    # every step copies its inputs from PINNED host memory and reads its results back into host memory, inside the timed
    # region; two batches are kept in flight, so the copies of one overlap the kernels of the other
    def e2e_stream(n):
        eng.score_batch_submit(0, text_host, img_host, head_host, None, K, algo=algo)
        res = None
        for i in range(1, n):
            eng.score_batch_submit(i & 1, text_host, img_host, head_host, None, K, algo=algo)
            res = eng.score_batch_collect((i - 1) & 1)
        return eng.score_batch_collect((n - 1) & 1) if n > 0 else res

    e2e_stream(max(4, args.warmup))
    ctx.barrier()
    n_e2e = max(args.steps, 20)
    t0 = time.perf_counter()
    res = e2e_stream(n_e2e)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
    assert np.array_equal(res["vault_rows"], out["vault_rows"].cpu().numpy()), f"{wl_name}: e2e result differs from the device path"
    for _ in range(3):
        eng.score_batch_host(text_host, img_host, head_host, None, K, algo=algo)
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        eng.score_batch_host(text_host, img_host, head_host, None, K, algo=algo)
    sync_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
    h2d = (text_host.numel() + img_host.numel() + head_host.numel()) * 4
    d2h = sum(v.nbytes for v in res.values())

    step_ms, search_ms, e2e_ms, sync_ms = ctx.max_over_ranks([step_ms, search_ms, e2e_ms, sync_ms])
    queries_per_step = Q * world
    screened = mode == "fp32" and K <= 16 and Q >= 16 and eng.get_option("screen") != 0 and algo != "stream"
    roof = hbm_roofline(ctx, wl_name, n_rows, Q, K, mode, search_ms, screened, rows_overridden)
    part = {
        "value": queries_per_step / (step_ms * 1e-3), "ms_per_step": step_ms, "step_ms_stats": stats(per_step),
        "search_ms_stats": stats(per_search), "timed_region_ms": step_ms * args.steps,
        "config": {"workload": wl["desc"], "queries_per_step": queries_per_step, "vault_rows_total": n_rows,
                   "vault_rows_per_gpu": n_rows, "dim": 512, "top_k": K, "vault_mode": mode, "algo": algo,
                   "parallelism": "replica x%d, queries sharded, no collective" % world,
                   "l2": f"vault shard {n_rows * 512 * (2 if mode == 'bf16' else 4) / 1e6:.0f} MB streamed per step "
                         + ("(> 126 MB L2), no flush needed" if n_rows * 512 * 2 > 126e6 else "(the hi planes fit the 126 MB L2: a resident-vault number)")},
        "roofline": roof,
        "e2e": {"value": queries_per_step / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "Engine.score_batch_submit / score_batch_collect (mmf_score_batch_submit): pinned host buffers in, host arrays "
                       "out, 2 batches in flight (the copies of one overlap the kernels of the other)",
                "sync_call": {"ms_per_step": sync_ms, "value": queries_per_step / (sync_ms * 1e-3),
                              "api": "Engine.score_batch_host (mmf_score_batch_host): one blocking call per batch"}},
        "gpu_launches": int(launches), "clocks": clocks,
        "parity": "planted rows recovered at their cosines (tol %g); one-call path == three-call path == e2e path bit for bit" % tol,
    }
    return part, (eng, vault, vault_rows, batch)


def run_c3_latency(ctx, eng, vault, K):
    """batch-1 latency mode: distribution over 1000 DISTINCT queries, one search each, device-timed,
    plus the host-buffer call (H2D + search + D2H + sync) wall-clock."""
    import mmf_b200
    from mmf_b200 import synth
    dev = ctx.dev
    gl = torch.Generator(device=dev).manual_seed(synth.QUERY_SEED + 99)
    qs = torch.randn(1000, 512, device=dev, generator=gl)
    for i in range(5):
        vault.search(qs[i:i + 1], K, mmf_b200.VAULT_THRESHOLD, "auto")
    evs = []
    for i in range(qs.shape[0]):
        a_ev, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_ev.record()
        vault.search(qs[i:i + 1], K, mmf_b200.VAULT_THRESHOLD, "auto")
        b_ev.record()
        evs.append((a_ev, b_ev))
    torch.cuda.synchronize()
    lat = sorted(a_ev.elapsed_time(b_ev) for a_ev, b_ev in evs)
    qh = qs[:200].cpu().numpy()
    for i in range(5):
        eng.vault_search_host(qh[i:i + 1], K)
    wall = []
    for i in range(qh.shape[0]):
        t0 = time.perf_counter()
        eng.vault_search_host(qh[i:i + 1], K)
        wall.append((time.perf_counter() - t0) * 1e3)
    wall.sort()
    return {"queries": len(lat), "unit": "ms", "p50": lat[len(lat) // 2], "p90": lat[int(len(lat) * 0.9)],
            "p99": lat[int(len(lat) * 0.99)], "max": lat[-1], "mean": sum(lat) / len(lat),
            "what": "vault search of ONE query (prep + streaming kernel + in-kernel merge), CUDA events",
            "host_call": {"p50": wall[len(wall) // 2], "p99": wall[int(len(wall) * 0.99)], "queries": len(wall),
                          "what": "Engine.vault_search_host: H2D + search + D2H + sync, wall clock"}}


def run_c4(ctx, wl, steps, warmup, rows_overridden=False):
    """C4: 4096 queries vs the 10M-row bf16 vault, top-100.  world == 1: one GPU holds everything.  world > 1: row-sharded,
    TruthVault(exchange="nccl") = local search + ncclAllGather (library-owned communicator) + merge in ONE library call;
    the three phases are also timed separately through the phase entry points."""
    import mmf_b200
    from mmf_b200 import synth
    dev, rank, world, dist = ctx.dev, ctx.rank, ctx.world, ctx.dist
    Q, K, total_rows, mode = wl["q"], wl["k"], wl["rows"], wl["mode"]
    plan = mmf_b200.ShardPlan(total_rows, world)
    lo, hi = plan.bounds(rank)
    n_local = hi - lo
    eng = mmf_b200.Engine(dev)
    rows = gen_vault_rows(dev, total_rows, lo, hi)
    vault = mmf_b200.TruthVault(eng, rows, None, mode=mode, rank=rank, world=world, n_total=total_rows, row_offset=lo,
                                exchange="nccl" if world > 1 else None)
    # one batch for all ranks: rank 0's (its planted rows live in shard 0)
    batch = make_batch(rows, n_local, Q, synth.QUERY_SEED + 4, dev)
    del rows
    q_dev = batch["q"].to(dev)
    if world > 1:
        dist.broadcast(q_dev, 0)
        pd = batch["pick"].to(dev)
        dist.broadcast(pd, 0)
        batch["pick"] = pd.cpu()
    q_host = q_dev.cpu().pin_memory()

    scores, rws, disc = vault.search(q_dev, K)
    torch.cuda.synchronize()
    check_planted(rws[:, 0], scores[:, 0], batch, 0, 1e-2, "c4")
    exact = None
    if world > 1:
        # bit-exact gate: the sharded result on every rank == an UNSHARDED search of the whole vault (rank 0 loads all
        # 10M rows into a second handle once; outside every timed region)
        ref = [None]
        if rank == 0:
            eng_full = mmf_b200.Engine(dev)
            eng_full.vault_load(gen_vault_rows(dev, total_rows, 0, total_rows), mode=mode)
            fs, fr, fd = eng_full.vault_search(q_dev, K)
            torch.cuda.synchronize()
            ref[0] = (fs, fr, fd)
        flags = torch.zeros(1, device=dev, dtype=torch.int32)
        for t_i in range(3):
            mine = (scores, rws, disc)[t_i]
            buf = mine.clone() if rank != 0 else ref[0][t_i].clone()
            dist.broadcast(buf, 0)                                  # rank 0's unsharded result to everybody
            if not torch.equal(buf.view(torch.int32) if buf.dtype == torch.float32 else buf,
                               mine.view(torch.int32) if mine.dtype == torch.float32 else mine):
                flags += 1
        dist.all_reduce(flags)
        exact = int(flags.item()) == 0
        if rank == 0:
            eng_full.close()
            del eng_full, ref
            torch.cuda.empty_cache()
        assert exact, "c4: the row-sharded search differs from the unsharded one"

    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    for _ in range(warmup):
        vault.search(q_dev, K)
    ctx.barrier()
    if sampler:
        sampler.window_start()
    l0, c0 = eng.launch_count, eng.collective_count
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    marks[0].record()
    for i in range(steps):
        vault.search(q_dev, K)
        marks[i + 1].record()
    ctx.barrier()
    launches, collectives = eng.launch_count - l0, eng.collective_count - c0
    step_ms = marks[0].elapsed_time(marks[-1]) / steps
    per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
    # keep the same load going until the clock sampler's window is 0.5 s long; the step holds a collective, so every rank
    # runs the same number of extra steps (computed from the max-over-ranks step time)
    step_all = ctx.max_over_ranks([step_ms])[0]
    extra = int(min(400, max(0.0, (ClockSampler.MIN_WINDOW_S * 1e3 - step_all * steps) / max(step_all, 1e-3))))
    for _ in range(extra):
        vault.search(q_dev, K)
    torch.cuda.synchronize()
    if sampler:
        sampler.held = extra > 0
    clocks = sampler.stop() if sampler else None

    # the phases, each bracketed by events on the launching stream: local search -> all-gather -> merge
    ph = {"search": [], "all_gather": [], "merge": []}
    for i in range(warmup + steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        packed = eng.vault_search_candidates(q_dev, K)
        ev[1].record()
        gathered = eng.shard_all_gather(packed, world) if world > 1 else packed[None]
        ev[2].record()
        ms_, mr_, md_ = eng.topk_merge(gathered, K)
        ev[3].record()
        if i >= warmup:
            ph["search"].append((ev[0], ev[1])); ph["all_gather"].append((ev[1], ev[2])); ph["merge"].append((ev[2], ev[3]))
    ctx.barrier()
    assert torch.equal(mr_, rws) and torch.equal(ms_, scores), "c4: phase entry points differ from the one-call search"
    ph_ms = {k: statistics.mean(a.elapsed_time(b) for a, b in v) for k, v in ph.items()}

    # e2e: host queries in, host results out, every step
    for _ in range(2):
        s_, r_, d_ = vault.search(q_host.to(dev, non_blocking=True), K)
        r_.cpu()
    ctx.barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, steps // 2)
    for _ in range(n_e2e):
        s_, r_, d_ = vault.search(q_host.to(dev, non_blocking=True), K)
        res = (s_.cpu(), r_.cpu(), d_.cpu())
    e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e

    step_ms, search_ms, ag_ms, merge_ms, e2e_ms = ctx.max_over_ranks([step_ms, ph_ms["search"], ph_ms["all_gather"], ph_ms["merge"], e2e_ms])
    flops_rank = 2.0 * Q * n_local * 512
    tf_search = flops_rank / (search_ms * 1e-3) / 1e12
    tf_step = flops_rank / (step_ms * 1e-3) / 1e12
    traffic = ncu_traffic(NCU_SUMMARY["c4"]) if (world == 8 or n_local == 1_250_000) and not rows_overridden else None
    return {
        "metric": "vault queries/s", "value": Q / (step_ms * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": step_ms, "step_ms_stats": stats(per_step), "scaling": "strong", "dtype": "bf16",
        "config": {"workload": wl["desc"], "queries_per_step": Q, "vault_rows_total": total_rows, "vault_rows_per_gpu": n_local, "dim": 512,
                   "top_k": K, "vault_mode": mode,
                   "parallelism": ("vault row-sharded x%d + ncclAllGather of the packed candidates (library-owned communicator, "
                                   "csrc/shard.cu) + merge" % world) if world > 1 else "one GPU holds the whole vault (strong-scaling base)",
                   "l2": f"vault shard {n_local * 1024 / 1e6:.0f} MB per rank (> 126 MB L2), no flush needed"},
        "phases_ms": {"local_search": search_ms, "all_gather": ag_ms, "merge": merge_ms,
                      "all_gather_us": ag_ms * 1e3, "merge_us": merge_ms * 1e3,
                      "all_gather_bytes_received_per_rank": world * Q * K * 8 if world > 1 else 0,
                      "what": "phase entry points (mmf_vault_search_candidates / mmf_shard_all_gather / mmf_topk_merge), CUDA events "
                              "on the launching stream, max over ranks; the timed step above is the ONE-call mmf_vault_search_sharded"},
        "roofline": {"bound": "tensor", "achieved": tf_search, "peak": ctx.tf_peak, "unit": "TFLOP/s", "frac": tf_search / ctx.tf_peak,
                     "frac_of_sustained_peak": tf_search / ctx.tf_sustained, "kernel": "local vault search (query prep + tcgen05 search + "
                     "top-k select), per rank", "algorithmic_flops_per_launch": flops_rank, "kernel_ms": search_ms,
                     "whole_step_tflops_per_rank": tf_step, "whole_step_frac": tf_step / ctx.tf_peak,
                     "whole_step_frac_of_sustained_peak": tf_step / ctx.tf_sustained,
                     "traffic": traffic, "traffic_source": ("profiles/%s.ncu_summary.txt (committed capture of the 1.25M-row shard shape; "
                                                            "NOT measured in this run)" % NCU_SUMMARY["c4"]) if traffic else None,
                     "peak_source": f"MEASURED_PEAKS.json ({ctx.peak_kind})"},
        "e2e": {"value": Q / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": Q * 512 * 4,
                "d2h_bytes_per_step": Q * K * 12 + Q * 4, "api": "TruthVault.search on pinned host queries + .cpu() of the results"},
        "gpu_launches": int(launches), "collectives": int(collectives), "clocks": clocks,
        "bit_exact_vs_unsharded": exact,
        "parity": "planted rows recovered (tol 1e-2); " + ("sharded == unsharded search of all %d rows on rank 0, bit for bit (scores, rows, discrepancy)" % total_rows
                                                           if world > 1 else "unsharded (nothing to compare against)"),
    }


def run_kernels(ctx, eng):
    """K1 (cosine) large-B sweep against the HBM roof, K5 (fusion judge) microseconds."""
    dev = ctx.dev
    out = {"cosine_pairs": [], "fusion_judge": []}

    def timed(fn, reps):
        """device time of one call.  The small sizes take microseconds, less than it takes Python to issue a call, so 20
        calls are captured in a CUDA graph and the replay is timed (what a C caller sees back to back)."""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        cg = torch.cuda.CUDAGraph()
        per = 20
        with torch.cuda.stream(side):
            with torch.cuda.graph(cg, stream=side):
                for _ in range(per):
                    fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        cg.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            cg.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / (reps * per)
    g = torch.Generator(device=dev).manual_seed(7)
    big_a = torch.randn(1_000_000, 512, device=dev, generator=g)
    big_b = torch.randn(1_000_000, 512, device=dev, generator=g)
    for B in (256, 4096, 65536, 1_000_000):
        xa, xb = big_a[:B], big_b[:B]
        ms = timed(lambda: eng.cosine_pairs(xa, xb), 10 if B <= 65536 else 2)
        gbs = B * 4100.0 / (ms * 1e-3) / 1e9
        out["cosine_pairs"].append({"pairs": B, "us": ms * 1e3, "GBps_algorithmic": gbs, "frac_of_hbm_peak": gbs / ctx.hbm_peak,
                                    "note": "operands %s" % ("fit L2 (a resident number)" if B * 4096 < 100e6 else "exceed the 126 MB L2")})
    del big_a, big_b
    x = torch.rand(1_000_000, 5, device=dev, generator=g)
    for B in (256, 1000, 65536, 1_000_000):
        xs = x[:B]
        ms = timed(lambda: eng.fusion_forward(xs), 10)
        out["fusion_judge"].append({"samples": B, "us": ms * 1e3, "Msamples_per_s": B / (ms * 1e-3) / 1e6,
                                    "GBps_algorithmic": B * 40.0 / (ms * 1e-3) / 1e9})
    out["what"] = ("CUDA events around replays of a CUDA graph of 20 calls, 3 warm-ups; 4 100 B / pair (cosine), 28 B in + 12 B out / "
                   "sample (fusion), SURVEY.md 8(d)")
    return out


def run_c5(ctx, batch=256, rows=1_000_000, top_k=5, steps=8, warmup=3, keep=None):
    """C5 (BASELINE.json configs[4]): full batched analyze -- random-init PyTorch producers (RoBERTa-base heads,
    EfficientNet-B0, CLIP ViT-B/32; bf16 autocast) feeding the scoring kernels through ONE library call, data-parallel
    replicas (vault replicated per GPU, no collective).  Reports samples/s and how the step splits between the encoders
    (stock PyTorch, out of scope) and this repo's kernels."""
    import importlib.util
    import mmf_b200
    from mmf_b200 import synth
    spec = importlib.util.spec_from_file_location("bench_c5", os.path.join(ROOT, "tools", "bench_c5.py"))
    c5 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(c5)
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    prod = tuple(m.to(dev) for m in c5.build_producers())
    if keep is not None and keep[1].n_total == rows and keep[1].mode == "fp32":
        eng, vault = keep[0], keep[1]
        own = False
    else:
        eng = mmf_b200.Engine(dev)
        vault = mmf_b200.TruthVault(eng, gen_vault_rows(dev, rows, 0, rows), None, mode="fp32")
        eng.fusion_load(synth.fusion_state_dict())
        own = True
    px, cid, rid = c5.synthetic_inputs(batch, dev, 100 + rank)

    def step(ev=None):
        if ev:
            ev[0].record()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            t, i, h = c5.encode(prod, px, cid, rid)
        if ev:
            ev[1].record()
        out = eng.score_batch(t, i, h, None, top_k)
        if ev:
            ev[2].record()
        return out
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    for _ in range(warmup):
        step()
    ctx.barrier()
    if sampler:
        sampler.window_start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for e in evs:
        out = step(e)
    verdict = out["verdict"].cpu()
    ctx.barrier()
    enc_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / steps
    hot_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / steps
    step_ms = evs[0][0].elapsed_time(evs[-1][2]) / steps
    clocks = sampler.stop() if sampler else None
    step_ms, enc_ms, hot_ms = ctx.max_over_ranks([step_ms, enc_ms, hot_ms])
    if own:
        eng.close()
    return {"metric": "analyze samples/s", "value": batch * world / (step_ms * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": step_ms, "scaling": "weak", "dtype": "bf16 autocast encoders + f32 hot path", "data": "synthetic",
            "config": {"workload": "C5: batched analyze, random-init RoBERTa-base + EfficientNet-B0 + CLIP ViT-B/32 producers -> cosine + "
                                   "vault top-%d + fusion judge" % top_k, "batch_per_gpu": batch, "vault_rows_per_gpu": rows,
                       "parallelism": "replica x%d, vault replicated, no collective" % world},
            "split": {"encoders_ms": enc_ms, "scoring_hot_path_ms": hot_ms, "hot_path_fraction": hot_ms / max(step_ms, 1e-9),
                      "what": "CUDA events around the encoder forwards (stock PyTorch) and around mmf_score_batch (this repo), max over ranks"},
            "fake_verdicts_rank0": int(verdict.sum()), "clocks": clocks}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="mmf_b200", choices=["mmf_b200", "reference"])
    ap.add_argument("--algo", default="auto", choices=["auto", "stream", "mma"])
    ap.add_argument("--rows", type=int, default=0, help="override vault rows of every workload (debug only; the line then says so)")
    ap.add_argument("--records", default="auto", help="auto (N = 1: c1,c3,c4,kernels,c5; N > 1: c4,c5) | none | comma list")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wls = {k: dict(v) for k, v in WORKLOADS.items()}
    if args.rows:
        for v in wls.values():
            v["rows"] = args.rows
            v["desc"] += f" [rows overridden to {args.rows}]"
    wl = wls[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args, wl)
    args.warmup = max(args.warmup, 3)
    ctx = Ctx(args)
    rank, world = ctx.rank, ctx.world
    records_arg = args.records.lower()
    if records_arg == "auto":
        want = ["c1", "c3", "c4", "kernels", "c5"] if world == 1 else ["c4", "c5"]
    elif records_arg == "none":
        want = []
    else:
        want = [r for r in records_arg.split(",") if r]
    want = [r for r in want if r != args.workload]

    records, line = {}, None
    if args.workload == "c4":
        line = run_c4(ctx, wl, args.steps, args.warmup, bool(args.rows))
        line.update({"higher_is_better": True, "vs_baseline": None, "data": "synthetic", "cpu_baseline": None})
        keep = None
    else:
        part, keep = run_replica(ctx, args.workload, wl, True, args.algo, bool(args.rows))
        line = {"metric": "vault queries/s", "value": part.pop("value"), "unit": "queries/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": part.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        line.update(part)
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            eng, vault, vault_rows, batch = keep
            cpu = cpu_baseline_for(vault_rows.cpu().numpy(), wl["mode"], batch["q"].numpy(), batch["text"].numpy(),
                                   batch["head"].numpy(), wl["k"], 12.0)
        line["cpu_baseline"] = cpu
        if args.workload == "c3" and world == 1:
            line["latency"] = run_c3_latency(ctx, keep[0], keep[1], wl["k"])

    for name in want:
        t0 = time.perf_counter()
        try:
            if name == "c4":
                rec = run_c4(ctx, wls["c4"], min(args.steps, 10), 3, bool(args.rows))
            elif name == "c5":
                rec = run_c5(ctx, rows=wls["c2"]["rows"], keep=keep if args.workload == "c2" else None)
            elif name == "kernels":
                if world > 1:
                    continue
                import mmf_b200
                from mmf_b200 import synth
                eng_k = keep[0] if keep else mmf_b200.Engine(ctx.dev)
                if not keep:
                    eng_k.fusion_load(synth.fusion_state_dict())
                rec = run_kernels(ctx, eng_k)
            elif name in ("c1", "c2", "c3"):
                if world > 1:
                    continue
                # c3 searches the c2 vault (same 1M fp32-exact rows): reuse the resident copy when the headline is c2
                reuse = keep[:3] if (keep and name == "c3" and args.workload == "c2") else None
                rec, kept = run_replica(ctx, name, wls[name], False, "auto", bool(args.rows), keep=reuse)
                rec.update({"metric": "vault queries/s", "unit": "queries/s", "steps": args.steps, "warmup": args.warmup})
                if name == "c3":
                    rec["latency"] = run_c3_latency(ctx, kept[0], kept[1], wls[name]["k"])
                if name == "c1" and rank == 0 and not args.no_cpu_baseline:
                    rec["cpu_baseline"] = cpu_baseline_for(kept[2].cpu().numpy(), "fp32", kept[3]["q"].numpy(), kept[3]["text"].numpy(),
                                                           kept[3]["head"].numpy(), wls[name]["k"], 4.0)
                if reuse is None:
                    kept[0].close()
                del kept
            else:
                continue
            rec["record_wall_s"] = time.perf_counter() - t0
            records[name] = rec
        except AssertionError:
            raise                                   # a parity gate failed: no line at all
        except Exception as e:                      # a record must never cost the headline
            records[name] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()

    if rank == 0:
        line["records"] = records
        print(json.dumps(line))
    if world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
