"""bench.py contract pieces that need no GPU: the reference arm (CPU port of the reference path) prints one JSON line with the
agreed keys, and the peak table is read tolerantly from whatever the driver wrote."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--queries", "8",
                          "--batches", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "egonets/s" and d["higher_is_better"] is True and d["value"] > 0
    from oracle import build_ref
    # the unmodified reference (byte-compiled into oracle/_ref by oracle/build_ref.py) when it has been built, else the oracle port
    assert d["cpu_baseline"]["kind"] == ("reference" if build_ref.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert ("cpu_port" in d) == build_ref.available()
    assert d["e2e"] == {"value": d["value"], "unit": "egonets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["metric"].startswith("egonets/s (fwd+bwd) PGAT d=250") and "workload" in d["config"]


def test_reference_arm_on_non_zero_ranks_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_measured_peaks_are_read_tolerantly(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.load_peaks()[2].startswith("fallback")
    for payload, want in (({"hbm_gbs": 6451.2, "bf16_tflops_sustained": 1401.7}, 6451.2),
                          ({"hbm": {"copy_GBps": 6451.2}, "bf16": {"burst_tflops": 1687.0, "sustained_tflops": 1401.7}}, 6451.2),
                          ({"hbm_bandwidth_TBps": 6.45}, 6450.0)):
        (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps(payload))
        hbm, tf, src = bench.load_peaks()
        assert abs(hbm - want) < 1e-6 and src.startswith("measured") and tf > 100
    (tmp_path / "MEASURED_PEAKS.json").write_text("not json")
    assert bench.load_peaks()[2].startswith("fallback")
