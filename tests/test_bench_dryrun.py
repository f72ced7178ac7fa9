"""bench.py is what the driver runs; a Python-level slip in it costs the round's measurement.  This dry run executes
bench.main() on the CPU with the CUDA-facing pieces swapped for stand-ins (the engine is the oracle-backed test double
of tests/cpu_engine.py, torch.cuda is a small fake, vault rows are cut down with --rows), and checks the JSON line it
prints against the contract: keys, units, the roofline / cpu_baseline / e2e objects.  No number it produces means
anything -- only that every code path of the harness runs and the line is well-formed."""
import importlib.util
import json
import os
import sys
import time
import types

import pytest
import torch

from cpu_engine import OracleEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Event:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


class _CountingEngine(OracleEngine):
    """the double plus the launch counter bench.py reads"""

    def __init__(self, device=None):
        super().__init__()
        self._n = 0

    @property
    def launch_count(self):
        self._n += 3
        return self._n

    @launch_count.setter
    def launch_count(self, v):
        pass



def _load_bench(monkeypatch):
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    cpu = torch.device("cpu")
    fake_cuda = types.SimpleNamespace(set_device=lambda *a, **k: None, synchronize=lambda *a, **k: None, Event=_Event,
                                      is_available=lambda: False, empty_cache=lambda: None)
    real_randn, real_generator = torch.randn, torch.Generator

    class _TorchProxy:
        """torch, with 'cuda' devices mapped to the CPU"""
        cuda = fake_cuda

        def __getattr__(self, name):
            return getattr(torch, name)

        @staticmethod
        def device(*a, **k):
            return cpu

        @staticmethod
        def Generator(device=None):
            return real_generator()

        @staticmethod
        def randn(*a, device=None, **k):
            return real_randn(*a, **k)

    monkeypatch.setattr(bench, "torch", _TorchProxy())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    import mmf_b200
    monkeypatch.setattr(mmf_b200, "Engine", _CountingEngine)
    return bench


def _run(bench, monkeypatch, capsys, argv):
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    bench.main()
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "records"}


def _check_roofline(r, bound):
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "traffic_source", "kernel_ms"} <= set(r)
    assert r["bound"] == bound and r["unit"] == ("TFLOP/s" if bound == "tensor" else "GB/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] is None        # --rows override: no ncu traffic claimed


@pytest.mark.parametrize("workload", ["c2", "c3", "c4", "c1"])
def test_bench_line_is_well_formed(monkeypatch, capsys, workload):
    bench = _load_bench(monkeypatch)
    rows = {"c1": 2000, "c2": 3000, "c3": 3000, "c4": 2500}[workload]
    d = _run(bench, monkeypatch, capsys, ["--workload", workload, "--rows", str(rows), "--steps", "2", "--warmup", "3", "--records", "none"])
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "vault queries/s" and d["unit"] == "queries/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 2 and d["warmup"] == 3 and d["value"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["scaling"] == ("strong" if workload == "c4" else "weak") and d["dtype"] == ("bf16" if workload == "c4" else "f32")
    assert set(d["config"]) >= {"workload", "queries_per_step", "vault_rows_total", "vault_rows_per_gpu", "top_k", "parallelism", "l2"}
    _check_roofline(d["roofline"], "tensor" if workload == "c4" else "hbm")
    assert d["records"] == {} and {"median", "min", "max"} <= set(d["step_ms_stats"])
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["unit"] == "queries/s"
    assert d["gpu_launches"] > 0 and "reasons" in d["clocks"]
    if workload == "c4":
        assert d["cpu_baseline"] is None and {"local_search", "all_gather_us", "merge_us"} <= set(d["phases_ms"])
        assert d["bit_exact_vs_unsharded"] is None                       # one rank: nothing to compare against
    else:
        c = d["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c and "host" in c and "batched_restatement" in c
        assert "sync_call" in e and "score_batch_submit" in e["api"]
    if workload == "c2":
        r = d["roofline"]
        assert r["variant"].startswith("screened") and r["mma_passes"] == 1 and r["bytes_streamed_per_launch"] == rows * 512 * 2
        assert e["d2h_bytes_per_step"] == 256 * (4 + 2 * 4 + 10 * 4 + 10 * 8 + 4 + 4 + 5 * 4 + 4)      # verdict, probs, top-10 scores / rows, sim, disc, x, conf
    if workload == "c3":
        lat = d["latency"]
        assert lat["queries"] == 1000 and lat["p50"] <= lat["p90"] <= lat["p99"] <= lat["max"] and "host_call" in lat


def test_bench_default_run_carries_the_records(monkeypatch, capsys):
    """what the driver runs (`bench.py --gpus 1 --steps K --warmup W`): the c2 headline plus the c1 / c3 / c4 / kernels records"""
    bench = _load_bench(monkeypatch)
    monkeypatch.setattr(bench, "run_kernels", lambda ctx, eng: {"cosine_pairs": [], "fusion_judge": []})   # 1M-pair sweeps: GPU-sized
    monkeypatch.setattr(bench, "run_c5", lambda ctx, **kw: {"metric": "analyze samples/s", "value": 1.0})      # three encoders: GPU-sized
    d = _run(bench, monkeypatch, capsys, ["--rows", "2000", "--steps", "2", "--warmup", "3", "--no-cpu-baseline"])
    assert d["cpu_baseline"] is None and set(d["records"]) == {"c1", "c3", "c4", "kernels", "c5"}
    for name, rec in d["records"].items():
        assert "error" not in rec, (name, rec)
    assert d["records"]["c4"]["config"]["top_k"] == 100 and d["records"]["c4"]["roofline"]["bound"] == "tensor"
    assert d["records"]["c3"]["latency"]["queries"] == 1000 and d["records"]["c1"]["config"]["queries_per_step"] == 1000


def test_bench_reference_arm_line(monkeypatch, capsys):
    bench = _load_bench(monkeypatch)
    d = _run(bench, monkeypatch, capsys, ["--impl", "reference", "--rows", "4000", "--steps", "2", "--warmup", "1"])
    assert d["impl"] == "reference" and d["metric"] == "vault queries/s" and d["unit"] == "queries/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(d["config"]) >= {"workload", "queries_per_step", "vault_rows_total", "top_k", "parallelism"}


def test_graft_entry_smoke_logic(monkeypatch, capsys):
    """__graft_entry__.smoke() -- the driver's first GPU call -- with the engine swapped for the test double: its own
    Python (inputs, the oracle comparisons, the final print) must run; the real thing needs the B200"""
    spec = importlib.util.spec_from_file_location("graft_entry_under_test", os.path.join(ROOT, "__graft_entry__.py"))
    ge = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ge)
    import mmf_b200
    monkeypatch.setattr(mmf_b200, "Engine", _CountingEngine)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    ge.smoke()
    assert "smoke ok" in capsys.readouterr().out


def test_clock_sampler_keeps_only_lines_inside_the_load_window(monkeypatch):
    """ClockSampler.stop(): nvidia-smi lines stamped before window_start() or after stop() do not count; with no line in the
    window the result says so instead of passing idle clocks off as clocks under load"""
    import datetime
    bench = _load_bench(monkeypatch)
    t0 = datetime.datetime(2026, 1, 1, 12, 0, 0)

    def stamp(ms):
        return (t0 + datetime.timedelta(milliseconds=ms)).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]

    lines = [f"{stamp(-300)}, 1965, 1965, 150.0, Not Active, Not Active, Not Active, Not Active",
             f"{stamp(100)}, 1420, 1965, 600.0, Not Active, Not Active, Not Active, Active",
             f"{stamp(200)}, 1400, 1965, 640.0, Not Active, Not Active, Not Active, Active",
             f"{stamp(900)}, 1965, 1965, 200.0, Not Active, Not Active, Not Active, Not Active"]

    class _Proc:
        def terminate(self):
            pass

        def communicate(self, timeout=None):
            return "\n".join(lines) + "\n", ""

    class _Now(datetime.datetime):
        @classmethod
        def now(cls, tz=None):
            return t0 + datetime.timedelta(milliseconds=500)

    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.proc, s.t0, s.t1, s.wall0, s.held, s.spawned = _Proc(), t0, None, 0.0, True, 0.0
    monkeypatch.setattr(bench.datetime, "datetime", _Now)
    out = s.stop()
    assert out["samples"] == 2 and out["sm_mhz"] == 1410.0 and out["sm_min_mhz"] == 1400.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["power_w_max"] == 640.0 and out["window"].startswith("timed region +")
    s2 = bench.ClockSampler.__new__(bench.ClockSampler)
    s2.proc, s2.t0, s2.t1, s2.wall0, s2.held, s2.spawned = _Proc(), t0 + datetime.timedelta(milliseconds=300), None, 0.0, False, 0.0
    lines[:] = [lines[0], lines[3]]
    out2 = s2.stop()
    assert out2["samples"] == 2 and out2["window"].startswith("whole sampler lifetime")
