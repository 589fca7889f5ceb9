"""CPU: the bench.py contract pieces that do not need a GPU -- the reference arm's JSON line (keys the driver reads), the
bounded adaptive CPU sample, the non-zero ranks of a multi-process reference launch exiting without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                          cwd=ROOT, timeout=600)


def test_reference_arm_json_line():
    r = _run(["--impl", "reference", "--config", "C", "--steps", "2", "--warmup", "1", "--cpu-budget", "4"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images/sec per introspective E+D step" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("CIFAR-10-shaped")
    cb = d["cpu_baseline"]
    # kind "reference" = the unmodified reference trainer from oracle/_ref (built by __graft_entry__.build() / oracle/build_ref.py
    # wherever /root/reference exists); "port" = the oracle's restatement, only when those scripts are absent
    have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "soft_intro_vae", "train_soft_intro_vae.py")) or \
        os.path.isdir("/root/reference/soft_intro_vae")
    assert cb["kind"] == ("reference" if have_ref else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "batch" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["wall_s"] < 120


def test_reference_arm_other_ranks_exit_without_work():
    r = _run(["--impl", "reference", "--config", "C"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_port_fallback():
    """without the reference scripts the arm falls back to the oracle port and says so"""
    r = _run(["--impl", "reference", "--config", "C", "--steps", "1", "--warmup", "0", "--cpu-budget", "2"], env={"SIVAE_CPU_BASELINE": "port"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
