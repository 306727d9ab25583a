"""Multi-GPU parity as a -m gpu test: one torchrun rank per GPU runs tools/mgpu_check.py — sharded matcher
byte-identical to the single-GPU matcher, sharded window solve bitwise identical on every rank and equal to the
single-GPU solve.  Skips on a box with fewer than 2 GPUs (the 1-GPU boxes of the round-end run); logs of the 2/4/8-GPU
runs are committed under profiles/."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("cfg", ["C2", "C3"])
def test_sharded_pass_matches_single_gpu(world, cfg):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + (os.getpid() + 7 * world) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "mgpu_check.py"), cfg]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stderr[-4000:]
    assert r.stdout.count("bitwise_identical_across_ranks=True") == world
