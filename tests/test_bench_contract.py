"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE stdout line (JSON with the contract's
keys), library chatter on stdout is diverted, and the clock sampler degrades gracefully without NVML / nvidia-smi."""
import importlib.util
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "patches/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("3D patches/sec") and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps",
                        "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_clock_sampler_without_gpu_tools():
    bench = _load_bench()
    s = bench.ClockSampler(0, period=0.01)
    s.start()
    s.active.set()
    time.sleep(0.05)
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples", "source"}
    assert isinstance(out["reasons"], list)


def test_native_arm_refuses_to_run_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr


def test_cpu_arm_does_not_import_the_product():
    """The CPU arm (`--impl reference`, `cpu_baseline`) is the oracle port only: building its step must not pull the
    product package in."""
    code = ("import sys; sys.path.insert(0, %r); import importlib.util as u; "
            "spec = u.spec_from_file_location('b', %r); m = u.module_from_spec(spec); sys.argv = ['bench.py']; "
            "spec.loader.exec_module(m); step = m.cpu_oracle_step_factory((32, 64, 64)); l = step(); "
            "assert l == l; assert not any(k.startswith('multitalent_b200') for k in sys.modules), "
            "[k for k in sys.modules if k.startswith('multitalent_b200')]" % (ROOT, os.path.join(ROOT, "bench.py")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]


def test_resenc_cpu_arm_runs():
    bench = _load_bench()
    step = bench.cpu_oracle_step_factory((32, 64, 64), "resenc")
    l0 = step()
    assert l0 == l0 and abs(l0) < 1e6


def test_last_kernel_symbol_reports_the_dispatch():
    from multitalent_b200 import _lib as L
    k = L.lib().mtb200_last_kernel()
    assert isinstance(k, bytes)
