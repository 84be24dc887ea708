"""The reference arm of bench.py (`--impl reference`: the oracle port on the host cores) prints one JSON line with the keys the
driver reads.  Runs on CPU only; the GPU arm shares the same line layout (checked on the GPU box by the driver itself)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-masks", "2", "--ref-targets", "8"], capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mask x target CDS comparisons/sec" and d["unit"] == "comparisons/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "comparisons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "2 masks x 8 targets" in d["config"]["workload"]      # the arm says what it actually timed
    assert d["gpu_library_mapped"] is False        # the CPU arm takes its inputs from the host-only generator: libcdsgpu.so is not mapped


def test_reference_arm_under_torchrun_prints_once():
    """Launched like the driver launches N > 1 (torchrun, one process per GPU): rank 0 alone runs and prints the reference line,
    the other ranks exit 0 without work."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
           "--ref-masks", "2", "--ref-targets", "8"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
