"""Where do the pinned buffers live relative to the GPU?  Prints the NUMA layout the process sees, the GPU's node, and the
H2D / D2H bandwidth of a 3.5 MB pinned buffer allocated under each memory policy (set_mempolicy MPOL_BIND per node)."""
import ctypes, glob, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def read(p):
    try:
        return open(p).read().strip()
    except Exception as e:
        return f"<{type(e).__name__}>"

print("allowed cpus", sorted(os.sched_getaffinity(0)))
for n in sorted(glob.glob("/sys/devices/system/node/node*")):
    print(os.path.basename(n), "cpus", read(n + "/cpulist"), "mem", read(n + "/meminfo").splitlines()[0] if os.path.exists(n + "/meminfo") else "")
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
bus = pynvml.nvmlDeviceGetPciInfo(h).busId
bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
short = bus[-12:] if len(bus) > 12 else bus
print("gpu0 pci", bus, "numa_node", read(f"/sys/bus/pci/devices/{short}/numa_node"), "local_cpulist", read(f"/sys/bus/pci/devices/{short}/local_cpulist"))
libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy = 238   # x86_64

def bind(node):
    if node is None:
        r = libc.syscall(SYS_set_mempolicy, 0, None, 0)
    else:
        mask = ctypes.c_ulong(1 << node)
        r = libc.syscall(SYS_set_mempolicy, 2, ctypes.byref(mask), 64)   # MPOL_BIND
    return r, ctypes.get_errno()

dev = torch.device("cuda:0")
d = torch.empty(887040 + 1584, dtype=torch.float32, device=dev)
nodes = [None] + [int(os.path.basename(n)[4:]) for n in sorted(glob.glob("/sys/devices/system/node/node*"))]
for node in nodes:
    r = bind(node)
    h_ = torch.empty(d.numel(), dtype=torch.float32).pin_memory()
    h_.fill_(1.0)
    bind(None)
    for name, fn in (("H2D", lambda: d.copy_(h_, non_blocking=True)), ("D2H", lambda: h_.copy_(d, non_blocking=True))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 200
        print(f"policy node={node} (set_mempolicy -> {r}): {name} {us:.1f} us = {d.numel() * 4 / us / 1e3:.1f} GB/s")
