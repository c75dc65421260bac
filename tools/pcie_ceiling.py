"""Per-rank host->device copy ceiling with N ranks uploading at once (torchrun, one rank per GPU) -- the denominator of the
end-to-end scaling numbers: config 3's e2e step is one 375-MB upload per rank per step, so its N-GPU efficiency cannot exceed
(per-rank H2D bandwidth with N concurrent uploads) / (bandwidth of one upload alone).

Each rank: pinned source (375 MB, the scene of config 3), device destination, 20 back-to-back cudaMemcpyAsync, all ranks
released together by a barrier; device-timed.  Run twice: pinned buffer allocated before and after binding the process to
the GPU's NUMA node (splatter360_b200.io.bind_to_gpu_numa_node).  Rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from splatter360_b200.io import bind_to_gpu_numa_node, gpu_numa_node  # noqa: E402


def measure(dev, world, nbytes, iters=20):
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    src.fill_(1)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        dst.copy_(src, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    gbs = torch.tensor([nbytes * iters / (a.elapsed_time(b) * 1e-3) / 1e9], device=dev)
    if world > 1:
        all_ = [torch.zeros_like(gbs) for _ in range(world)]
        dist.all_gather(all_, gbs)
        return [float(x.item()) for x in all_]
    return [float(gbs.item())]


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 375390272
    node = gpu_numa_node(dev)
    cpus_before = len(os.sched_getaffinity(0))
    unbound = measure(dev, world, nbytes)
    bind = bind_to_gpu_numa_node(dev)
    bound = measure(dev, world, nbytes)
    nodes = [None] * world
    if world > 1:
        dist.all_gather_object(nodes, node)
    else:
        nodes = [node]
    if rank == 0:
        print(json.dumps({"n_gpus": world, "bytes_per_copy": nbytes, "gpu_numa_nodes": nodes, "cpus_before_bind": cpus_before,
                          "bind_rank0": bind,
                          "h2d_GBps_per_rank_unbound": [round(x, 2) for x in unbound], "aggregate_unbound": round(sum(unbound), 1),
                          "h2d_GBps_per_rank_bound": [round(x, 2) for x in bound], "aggregate_bound": round(sum(bound), 1)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
