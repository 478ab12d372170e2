"""Data-parallel invariance on real GPUs (needs >= 2): W ranks x B/W with the NCCL gradient allreduce captured in the
step's CUDA graph must match one rank x B.  Runs scripts/dp_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("mode", ["nccl", "peer"])
@pytest.mark.parametrize("world", [2])
def test_dp_invariance(world, mode):
    """nccl: allreduce captured in the step's CUDA graph; peer: in-kernel NVLink peer-memory exchange of the fused step."""
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    extra = ["--peer"] if mode == "peer" else []
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29517" if mode == "nccl" else "29518",
                        os.path.join(ROOT, "scripts", "dp_check.py")] + extra,
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "DP_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
