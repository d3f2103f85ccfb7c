"""On-GPU multi-rank correctness (NCCL): needs >= 2 GPUs, skipped otherwise.  Launches tests/nccl_worker.py with
torch.distributed.run, one process per GPU -- the same way bench.py --gpus N is launched."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2])
def test_nccl_shard_invariance_broadcast_and_hash_merge(tmp_path, world):
    """`world` NCCL ranks x B/world envs through SelfplayRunner == one rank x B (bit for bit, 3 consecutive steps, DeepSea with an even
    and Subleq with an uneven split); dist.broadcast_params and dist.merge_hash_sets on device tensors."""
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = tmp_path / "report.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    rep = json.load(open(out))
    assert rep.get("ok"), rep
    assert rep["world"] == world and rep["deepsea"]["bits_set"] > 0 and rep["subleq"]["bits_set"] > 0
