"""PartitionedDNAMap across GPUs (needs >= 2 devices; skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_world(world, extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(29740 + world), os.path.join(ROOT, "tests", "nccl_worker.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]
    assert "PMAP OK world %d" % world in r.stdout


@pytest.mark.parametrize("world", [2, 4, 8])
def test_pmap_matches_oracle(gpu, world):
    if gpu < world:
        pytest.skip("needs %d GPUs, box has %d" % (world, gpu))
    run_world(world)
    run_world(world, {"GENOME_B200_TUNE": "single_pass_min=1"})


@pytest.mark.parametrize("env", [{"GENOME_B200_TUNE": "route=2"}, {"GENOME_B200_TUNE": "a2a=1"}, {"GENOME_B200_TUNE": "batches=5,slice_bits=1"},
                                 {"GENOME_B200_TUNE": "single_pass_min=1"}, {"GENOME_B200_TUNE": "single_pass_min=1,batches=5,slice_bits=2"}],
                         ids=["two-level", "nccl-staged", "many-batches", "single-pass", "single-pass-many-batches"])
def test_pmap_routing_variants(gpu, env):
    """The same sharded run through the other routing paths: receiver-side re-bucketing, NCCL send/recv staging instead of
    peer stores, more batches than buffer sets; and the single-pass form (slabs in the owners' inboxes, no count pass: what large
    batches take by default) forced onto these small inputs, incl. the overflow lists of the all-reads-alike case."""
    if gpu < 2:
        pytest.skip("needs 2 GPUs, box has %d" % gpu)
    run_world(2, env)


@pytest.mark.parametrize("world", [1, 2, 8])
def test_pmap_sharded_graph_build(gpu, world):
    """Graph.buildGraph over the shards WITHOUT a replica (csrc/sgraph.cuh over the NCCL + CUDA-IPC fabric of comm.cu): the
    same worker, same oracle comparisons, with pgraph_sharded = 1 (gb_tune).  world = 1 runs the NCCL fabric (single-rank
    collectives, no IPC mapping to open) on a one-GPU box."""
    if gpu < world:
        pytest.skip("needs %d GPUs, box has %d" % (world, gpu))
    run_world(world, {"GENOME_B200_TUNE": "pgraph_sharded=1"})
    if world > 1:
        run_world(world, {"GENOME_B200_TUNE": "pgraph_sharded=0"})  # the replicated build (all-gather of the shards) beside it
