#!/usr/bin/env python
"""Host -> device bandwidth per rank when N ranks upload at the same time (the e2e curve).

    python tools/h2d_probe.py                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_probe.py

Every rank copies a 1 GiB page-locked host buffer to its GPU `--reps` times (cudaMemcpyAsync on its own stream),
all ranks starting together; reported: GB/s per rank and in aggregate, (a) with the scheduler's CPU placement and
(b) with the process bound to the CPUs of the GPU's NUMA node when sysfs tells which one that is.  Rank 0 prints
one JSON line."""
import argparse
import json
import os
import time

import torch


def numa_of_gpu(idx: int):
    try:
        bus = torch.cuda.get_device_properties(idx).pci_bus_id if hasattr(torch.cuda.get_device_properties(idx), "pci_bus_id") else None
    except Exception:
        bus = None
    try:
        import subprocess
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(idx)],
                             capture_output=True, text=True).stdout.strip().lower()
        bus = bus[4:] if len(bus) > 12 else bus            # 00000000:1B:00.0 -> 0000:1b:00.0
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            return int(f.read().strip())
    except Exception:
        return None


def cpus_of_node(node: int):
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            out = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                out.update(range(int(a), int(b or a) + 1))
            return out
    except Exception:
        return None


def run(dev, reps, gib, dist_on):
    import torch.distributed as dist
    host = torch.empty(int(gib * (1 << 30)) // 4, dtype=torch.float32).pin_memory()
    host.fill_(1.0)
    devbuf = torch.empty_like(host, device=dev)
    st = torch.cuda.Stream(dev)
    with torch.cuda.stream(st):
        devbuf.copy_(host, non_blocking=True)
    st.synchronize()
    if dist_on:
        dist.barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(st):
        for _ in range(reps):
            devbuf.copy_(host, non_blocking=True)
    st.synchronize()
    dt = time.perf_counter() - t0
    if dist_on:
        dist.barrier()
    return reps * host.numel() * 4 / dt / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--gib", type=float, default=1.0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    res = {}
    res["default_affinity"] = run(dev, args.reps, args.gib, dist_on)
    node = numa_of_gpu(local)
    cpus = cpus_of_node(node) if node is not None and node >= 0 else None
    if cpus:
        try:
            os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
            res["numa_local_affinity"] = run(dev, args.reps, args.gib, dist_on)
        except OSError:
            pass
    mine = dict(rank=rank, gpu=local, numa_node=node, ncpu_node=len(cpus) if cpus else None, **res)
    if dist_on:
        import torch.distributed as dist
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    if rank == 0:
        line = dict(probe="h2d", n_gpus=world, gib_per_copy=args.gib, reps=args.reps, os_cpu_count=os.cpu_count(), ranks=allr,
                    aggregate_default=sum(r["default_affinity"] for r in allr),
                    aggregate_numa_local=sum(r.get("numa_local_affinity", 0.0) for r in allr) or None)
        print(json.dumps(line), flush=True)
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
