"""Concurrent host<->device copy rates of several GPUs of one box (one process, one stream and one pinned buffer per
GPU): what the host side can absorb when every rank of a multi-GPU e2e run copies at once.  Prints the PCIe /
NUMA topology first.  Usage: python tools/pcie_probe_multi.py [size_MiB]"""
import subprocess
import sys
import time

import torch

for cmd in ("nvidia-smi topo -m", "lscpu | grep -i -E 'model name|socket|numa|^CPU\\(s\\)'", "grep -i huge /proc/meminfo"):
    print("$", cmd)
    print(subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout)

size = (int(sys.argv[1]) if len(sys.argv) > 1 else 22) << 20
n = torch.cuda.device_count()
dev, host, streams = [], [], []
for i in range(n):
    dev.append(torch.empty(size, dtype=torch.uint8, device=f"cuda:{i}"))
    host.append(torch.empty(size, dtype=torch.uint8).pin_memory())
    streams.append(torch.cuda.Stream(device=i))


def run(gpus, direction, reps=48):
    for i in gpus:
        torch.cuda.synchronize(i)
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in gpus:
            with torch.cuda.stream(streams[i]):
                if direction == "d2h":
                    host[i].copy_(dev[i], non_blocking=True)
                else:
                    dev[i].copy_(host[i], non_blocking=True)
    for i in gpus:
        streams[i].synchronize()
    return size * reps * len(gpus) / (time.perf_counter() - t0) / 1e9


subsets = [[i] for i in range(n)] + [list(range(k)) for k in (2, 4, 8) if k <= n]
if n >= 8:
    subsets += [[0, 4], [0, 2], [0, 1, 4, 5], [0, 2, 4, 6]]
for direction in ("d2h", "h2d"):
    for g in subsets:
        run(g, direction, 4)
        print(f"{direction} GPUs {g}: {run(g, direction):6.1f} GB/s aggregate")
