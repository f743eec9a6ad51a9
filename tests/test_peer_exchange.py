"""Two ranks (one process per GPU, torch.distributed.run): the per-round exchange over peer memory fused with the update (cgrt_peer_*) and the
ncclAllReduce path both reproduce the one-GPU render — same accepted photons and radii bit for bit, flux to the accumulators' precision."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("how", ["peer", "nccl"])
@pytest.mark.parametrize("accum", [0, 1])
def test_two_rank_render_equals_one_gpu(how, accum):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box (gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tests", "peer_worker.py"), str(accum), how],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"PEER_OK {how} {accum}" in r.stdout
