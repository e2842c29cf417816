"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`, the reference's own CPU code from
oracle/_ref, or the oracle port) prints exactly ONE JSON line on stdout with the contract's keys, whatever the libraries below write to stdout;
under torchrun rank 0 alone prints and every rank exits 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e")


def _check(stdout, n_gpus):
    lines = [l for l in stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines[:3]
    d = json.loads(lines[0])
    for k in KEYS:
        assert k in d, k
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    _check(r.stdout, 1)


def test_reference_arm_under_torchrun_rank0_only():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29733",
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    _check(r.stdout, 2)


def test_stdout_is_claimed_at_the_descriptor_level():
    code = "import bench, os; bench._claim_stdout(); print('library noise'); os.system('echo child noise'); bench.emit({'ok': 1})"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 0 and r.stdout == '{"ok": 1}\n' and "library noise" in r.stderr and "child noise" in r.stderr
