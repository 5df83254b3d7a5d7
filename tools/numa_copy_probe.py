#!/usr/bin/env python3
"""GPU box, under torchrun: does it matter where the pinned host buffers live?  Every rank times plain pinned H2D copies, all
ranks at once -- first with the process where the launcher put it, then bound to the CPUs next to its GPU (sysfs
local_cpulist) with a freshly allocated (first-touched) buffer.  Prints the aggregate GB/s of both and the topology facts.

    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/numa_copy_probe.py
"""
import os, time
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
N = int(os.environ.get("PROBE_GB", "2")) << 30
d = torch.empty(N, dtype=torch.uint8, device=dev)


def measure(tag):
    h = torch.empty(N, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)                                            # first touch by this thread
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(4):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    x = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(x, op=dist.ReduceOp.MAX)
    agg = world * 4 * N / float(x.item()) / 1e9
    mine = 4 * N / dt / 1e9
    print(f"[rank {rank}] {tag}: {mine:.1f} GB/s alone-clock, aggregate {agg:.1f} GB/s", flush=True)
    del h


p = torch.cuda.get_device_properties(lr)
bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
base = "/sys/bus/pci/devices/" + bus
try:
    cpus = open(base + "/local_cpulist").read().strip(); node = open(base + "/numa_node").read().strip()
except OSError as e:
    cpus, node = "", f"? ({e})"
print(f"[rank {rank}] gpu {lr} pci {bus} numa_node {node} local_cpulist {cpus} allowed {sorted(os.sched_getaffinity(0))[:4]}...({len(os.sched_getaffinity(0))})", flush=True)
measure("as launched")
want = set()
for part in cpus.split(","):
    if part:
        a, _, b = part.partition("-")
        want |= set(range(int(a), int(b or a) + 1))
want &= os.sched_getaffinity(0)
if want:
    os.sched_setaffinity(0, want)
    measure(f"bound to {len(want)} cpus next to the gpu")
else:
    print(f"[rank {rank}] nothing to bind to", flush=True)
if world > 1:
    dist.destroy_process_group()
