"""bench.py's host-side contract, checked without a GPU: the reference arm prints exactly ONE JSON line on stdout with
the contract's keys (native-library output on file descriptor 1 is diverted to stderr), and the CPU-baseline sample
sizing grows by doubling, stops at the budget and reports the best operating point."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--config", "cfg1", "--cpu-batch", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[:500]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "train images/sec (fwd+bwd)" and line["unit"] == "images/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("cfg1")


def test_cpu_sample_grows_by_doubling_and_stops_at_the_cliff(monkeypatch):
    bench = _bench_module()
    # a fake reference whose cost per image is 1 ms up to 8 images and falls off a cliff above (like the 64 x 64-tap
    # CPU convolution did on a GPU box): the sizing must try 2, 4, 8, 16 and report the 8-image operating point
    clock = {"t": 0.0}
    sizes = []

    def fake_step_fn(cfg, B):
        sizes.append(B)

        def step():
            clock["t"] += B * 1e-3 if B <= 8 else B * 5.0
        return step

    monkeypatch.setattr(bench, "_reference_step_fn", fake_step_fn)
    monkeypatch.setattr(bench.time, "perf_counter", lambda: clock["t"])
    ips, s_per_step, kind, threads, B = bench.cpu_reference_images_per_s(bench.PRESETS["cfg2"], 16, 2, 1, budget_s=8.0)
    assert kind == "reference" and threads >= 1
    assert sizes == [2, 4, 8, 16]
    assert B == 8 and abs(ips - 1000.0) < 1e-6 and abs(s_per_step - 8e-3) < 1e-9
    # a fake reference that is linear in the batch stops at the requested batch
    sizes.clear()
    monkeypatch.setattr(bench, "_reference_step_fn", lambda cfg, B: (sizes.append(B), (lambda: clock.__setitem__("t", clock["t"] + B * 1e-3)))[1])
    ips, _, _, _, B = bench.cpu_reference_images_per_s(bench.PRESETS["cfg2"], 12, 1, 1, budget_s=8.0)
    assert sizes == [2, 4, 8, 12] and B == 12 and abs(ips - 1000.0) < 1e-6
