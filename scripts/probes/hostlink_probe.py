#!/usr/bin/env python3
"""Host-link probe for the end-to-end multi-GPU numbers (VERDICT r1, weak point 4): what pinned copies reach per rank
when 1, 2, 4 or 8 ranks copy at once, H2D only / D2H only / both ways, for three ways of getting the pinned buffer:
  torch      torch.empty(pin_memory=True) from wherever the process happens to run
  numa       the process bound to the CPUs of the GPU's NUMA node (sysfs) before allocating and touching the buffer
  threads    one process, one thread per GPU (rank 0 only, world size 1 run): cudaHostAlloc per thread
Launch:  torchrun --nproc-per-node N scripts/probes/hostlink_probe.py     (prints one JSON line per configuration on rank 0)
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def gpu_numa_cpus(index: int):
    """CPUs of the NUMA node the GPU hangs off (None when sysfs does not say)."""
    try:
        props = torch.cuda.get_device_properties(index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None, node
        txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus, node
    except Exception:
        return None, -1


def measure(dev, h_up, h_dn, nbytes, mode, reps=6):
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def step():
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s_up):
                d_a.copy_(h_up, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s_dn):
                h_dn.copy_(d_b, non_blocking=True)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return nbytes * reps / float(t.item()) * 1e-9          # GB/s per rank and direction, slowest rank


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 1 << 30
    cpus, node = gpu_numa_cpus(local)
    info = {"rank": rank, "gpu": local, "numa_node": node, "numa_cpus": len(cpus) if cpus else None,
            "affinity_before": len(os.sched_getaffinity(0)), "nodes_online": open("/sys/devices/system/node/online").read().strip()
            if os.path.exists("/sys/devices/system/node/online") else "?"}
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, info)
    else:
        gathered = [info]
    results = {}
    for alloc in ("torch", "numa"):
        if alloc == "numa":
            if cpus:
                try:
                    os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
                except Exception:
                    pass
        h_up = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h_dn = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h_up.fill_(1); h_dn.fill_(2)                       # first touch from the (possibly re-bound) process
        for mode in ("h2d", "d2h", "both"):
            results[f"{alloc}.{mode}"] = measure(dev, h_up, h_dn, nbytes, mode)
        del h_up, h_dn
    if rank == 0:
        print(json.dumps({"n_ranks": world, "gbs_per_rank_each_way": results, "aggregate_both_gbs":
                          {k: v * world for k, v in results.items() if k.endswith("both")}, "ranks": gathered}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
