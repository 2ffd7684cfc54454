"""bench.py's reference arm runs on CPU and prints one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cells", "6",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "dofs/s" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["one_core"]["cores"] == 1 and d["cpu_baseline"]["extrapolated"] in (True, False)
    assert d["cpu_baseline"]["setup"]["patch_lists_s_full_estimate"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_product_arm_fails_loudly_without_gpu():
    import ssc_b200
    if ssc_b200.device_count() > 0:
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--cells", "4", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr


def test_reference_arm_under_a_launcher_prints_once_and_checks_the_rank_count():
    """--gpus N without a launcher starts N ranks itself (the contract's torchrun line); rank 0 alone prints; a launcher
    that started a different number of ranks is an error, not a silently mislabelled line."""
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cells", "5", "--steps", "1", "--warmup", "0"]
    env = {k: v for k, v in os.environ.items() if k not in ("WORLD_SIZE", "RANK", "LOCAL_RANK")}
    out = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["n_gpus"] == 2
    bad = subprocess.run(cmd + ["--gpus", "1"], capture_output=True, text=True, timeout=600, env=dict(env, WORLD_SIZE="2", RANK="0"))
    assert bad.returncode != 0 and "launcher started 2" in bad.stderr


def test_reference_arm_does_not_load_the_product_library():
    """The reference arm stands on oracle/ and the synthetic FE generator alone: libssc_b200.so (the product) must not be
    mapped into that process (the arm used the product's host builder for its dof lists)."""
    code = ("import sys, json; sys.argv=['bench.py','--impl','reference','--cells','4','--steps','1','--warmup','0'];"
            "sys.path.insert(0, %r); import runpy\n"
            "try:\n runpy.run_path(%r, run_name='__main__')\nexcept SystemExit: pass\n"
            "maps=open('/proc/self/maps').read(); print('PRODUCT_LOADED' if 'libssc_b200.so' in maps else 'PRODUCT_NOT_LOADED')") % (ROOT, os.path.join(ROOT, "bench.py"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "PRODUCT_NOT_LOADED" in out.stdout, out.stdout[-500:]
