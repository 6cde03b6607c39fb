"""Bare host<->device copy ceiling, one process per GPU, all at once (VERDICT r1 #3b).

  python tools/pcie_probe.py                                   # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
         tools/pcie_probe.py [--no-bind]                       # --all-gpus: every rank copies at the same time

Each rank moves what one e2e step of bench.py moves (2.77 GB image D2H + 0.69 GB PCM H2D, pinned host
memory, two streams) with no kernels, with and without pinning the process to the CPUs / NUMA node of
its GPU before the pinned buffers are allocated and first touched.  Rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


F, B, S = 1349969, 2049, 172800000
res = {"n_gpus": world, "topology": bench.gpu_numa_cpus(local), "host_cpus": os.cpu_count()}
res["unbound"] = bench.pcie_ceiling(dev, world, barrier, F * B, S * 4, dist)
if "--no-bind" not in sys.argv:
    res["bind"] = bench.numa_bind(local, world)
    res["bound"] = bench.pcie_ceiling(dev, world, barrier, F * B, S * 4, dist)
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
