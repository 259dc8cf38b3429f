"""Where does the e2e figure come from?  Pinned H2D / D2H bandwidth of this box, one direction at a time and both
together, with the process's CPU/NUMA placement (the e2e step of bench.py moves 956 MB each way)."""
import os
import subprocess
import time

import torch


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return f"<{e}>"


print("cpus allowed:", sorted(os.sched_getaffinity(0)))
print(sh("nvidia-smi topo -m | head -8"))
bus = sh("nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i 0").lower()
print("bus", bus, "numa_node", sh(f"cat /sys/bus/pci/devices/{bus[4:]}/numa_node"))
print(sh("lscpu | grep -i 'numa\\|model name\\|socket'"))
print(sh("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv -i 0"))

n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * n / dt / 1e9


for _ in range(2):
    print("H2D only %.1f GB/s | D2H only %.1f GB/s | both: %.1f GB/s each way" % (run(1, 0), run(0, 1), run(1, 1)))

# is the bidirectional rate a property of the pinned pages a process happens to get?  fresh buffers each round
for r in range(6):
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    run(1, 1, reps=2)
    print("fresh pinned buffers, round %d: both %.1f GB/s each way (H2D only %.1f, D2H only %.1f)" % (r, run(1, 1), run(1, 0), run(0, 1)))
