"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`: the oracle port of the reference's
CPU path, timed on the host cores) prints exactly ONE JSON line on stdout with the keys the driver reads, on the same
metric / unit / config as the GPU arm; under torchrun only rank 0 works and prints."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(cmd, env=None):
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    return [l for l in res.stdout.splitlines() if l.strip()]


def _check_reference_line(line, n_gpus):
    j = json.loads(line)
    assert j["impl"] == "reference" and j["metric"] == "maxsim_query_doc_pairs_per_s" and j["unit"] == "pairs/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == n_gpus and j["steps"] == 1 and j["warmup"] == 0
    assert j["value"] > 0 and j["ms_per_step"] > 0 and j["vs_baseline"] is None and j["data"] == "synthetic"
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "pairs" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    return j


def test_reference_arm_prints_one_json_line_on_our_config():
    sys.path.insert(0, str(ROOT))
    import bench
    lines = _run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert len(lines) == 1
    j = _check_reference_line(lines[0], 1)
    assert j["config"] == bench.workload_config(1)                      # the same `config` dict as the GPU arm
    assert j["config"]["workload"].startswith("ColPali-v1.3 MaxSim: 25000 queries")


def test_reference_arm_under_torchrun_only_rank0_prints():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    lines = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                  "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), "bench.py", "--impl", "reference",
                  "--gpus", "2", "--steps", "1", "--warmup", "0"], env=env)
    json_lines = [l for l in lines if l.startswith("{")]
    assert len(json_lines) == 1
    j = _check_reference_line(json_lines[0], 2)
    assert j["config"]["parallelism"] == "doc-shard x2"
