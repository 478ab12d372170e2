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


@pytest.mark.parametrize("mode", ["nccl", "peer", "wide_nccl", "wide_peer"])
@pytest.mark.parametrize("world", [2])
def test_dp_invariance(world, mode):
    """nccl: allreduce captured in the step's CUDA graph; peer: in-kernel NVLink peer-memory exchange of the persistent step;
    wide_*: the tcgen05 kernel plan with the NCCL allreduce / the two-phase peer-memory exchange fused with the optimizer."""
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    extra = {"nccl": [], "peer": ["--peer"], "wide_nccl": ["--wide"], "wide_peer": ["--wide", "--peer"]}[mode]
    port = {"nccl": "29517", "peer": "29518", "wide_nccl": "29519", "wide_peer": "29520"}[mode]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", port,
                        os.path.join(ROOT, "scripts", "dp_check.py")] + extra,
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "DP_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
