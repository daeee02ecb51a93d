"""PartitionedDNAMap across GPUs (needs >= 2 devices; skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_pmap_matches_oracle(gpu, world):
    if gpu < world:
        pytest.skip("needs %d GPUs, box has %d" % (world, gpu))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(29740 + world), os.path.join(ROOT, "tests", "nccl_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]
    assert "PMAP OK world %d" % world in r.stdout
