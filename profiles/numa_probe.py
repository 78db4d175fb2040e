"""Host side of the e2e path at N > 1: is the aggregate PCIe rate limited by NUMA placement of the pinned buffers?
Run under torchrun (one rank per GPU).  Every rank times, together with all the others, H2D 240 MB + D2H 320 MB on two streams
(the C2 step's traffic) with pinned buffers allocated (a) as the process happens to run, (b) after binding the process to the
CPUs NVML names as local to its GPU (first touch then lands on the local node)."""
import os, sys, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo", init_method="env://")


def barrier():
    if world > 1:
        dist.barrier()


def gpu_cpus(idx):
    """CPUs local to CUDA device idx, from NVML (matched by UUID) -- or None"""
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(idx)
        h = None
        uuid = str(getattr(props, "uuid", ""))
        for i in range(pynvml.nvmlDeviceGetCount()):
            hh = pynvml.nvmlDeviceGetHandleByIndex(i)
            u = pynvml.nvmlDeviceGetUUID(hh)
            u = u.decode() if isinstance(u, bytes) else u
            if uuid and uuid.replace("GPU-", "") in u:
                h = hh
                break
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 64
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        pci = pynvml.nvmlDeviceGetPciInfo(h).busId
        pci = pci.decode() if isinstance(pci, bytes) else pci
        node = None
        try:
            node = open("/sys/bus/pci/devices/%s/numa_node" % pci.lower()[-12:]).read().strip()
        except OSError:
            pass
        return cpus, pci, node
    except Exception as e:  # noqa: BLE001
        return None, repr(e), None


def measure(tag):
    nq = 10_000_000
    Xh = torch.empty((nq, 3), dtype=torch.float64).pin_memory(); Xh.zero_()
    Oh = torch.empty((nq, 4), dtype=torch.float64).pin_memory(); Oh.zero_()
    Xd = torch.empty((nq, 3), dtype=torch.float64, device="cuda"); Od = torch.zeros((nq, 4), dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(2):
        Xd.copy_(Xh, non_blocking=True); Oh.copy_(Od, non_blocking=True)
    torch.cuda.synchronize(); barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s1):
            Xd.copy_(Xh, non_blocking=True)
        with torch.cuda.stream(s2):
            Oh.copy_(Od, non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    barrier()
    print("rank %d %-10s H2D 240 MB + D2H 320 MB together: %.2f ms = %.1f GB/s (cpus now: %s)" % (rank, tag, dt * 1e3, 0.56 / dt, sorted(os.sched_getaffinity(0))[:4]), flush=True)
    return dt


if rank == 0:
    os.system("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)' ; nvidia-smi topo -m 2>/dev/null | head -14")
barrier()
cpus, pci, node = gpu_cpus(local)
print("rank %d gpu %d pci %s numa_node %s nvml cpus %s (%d)" % (rank, local, pci, node, (cpus or [])[:6], len(cpus or [])), flush=True)
barrier()
a = measure("default")
if cpus:
    os.sched_setaffinity(0, cpus)
    b = measure("numa-local")
if world > 1:
    dist.destroy_process_group()
