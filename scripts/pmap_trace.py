"""Sharded insert under torchrun with GENOME_B200_TRACE=1 (debugging aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bench import make_workload
from genome_b200.dnamap import Communicator, PartitionedDNAMap, torch_broadcast
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
b, n_reads, windows, G, distinct = make_workload("C2", rank, world, 1.0)
d = torch.zeros(b.size + 16, dtype=torch.uint8, device="cuda"); d[:b.size].copy_(torch.from_numpy(b))
comm = Communicator(rank, world, local, torch_broadcast)
cap = int(distinct / world * 1.15)
m = PartitionedDNAMap(31, comm, cap)
for rep in range(4):
    m.clear(cap)
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    m.insert_reads_device(d.data_ptr(), b.size, n_reads)
    torch.cuda.synchronize()
    if rank == 0: print("rep %d insert %.3f ms" % (rep, (time.perf_counter() - t) * 1e3), file=sys.stderr)
m.close(); comm.close(); dist.destroy_process_group()
