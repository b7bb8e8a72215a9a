"""H2D bandwidth from pinned host memory as a function of chunk size and of the CPU/NUMA node the pinned buffer was
allocated (first-touched) from.  Tool only (uses torch for brevity); informs the 4K e2e number, which is PCIe-bound."""
import os, time
import torch

def cpus_by_node():
    nodes = {}
    base = "/sys/devices/system/node"
    if os.path.isdir(base):
        for d in sorted(os.listdir(base)):
            if d.startswith("node") and d[4:].isdigit():
                s = open(f"{base}/{d}/cpulist").read().strip()
                cp = []
                for part in s.split(","):
                    if not part: continue
                    a, _, b = part.partition("-")
                    cp += list(range(int(a), int(b or a) + 1))
                nodes[int(d[4:])] = cp
    return nodes

def bw(nbytes_chunk, total=2 << 30, dev=0, streams=1):
    n = max(1, total // nbytes_chunk)
    host = torch.empty(nbytes_chunk * min(n, 8), dtype=torch.uint8).pin_memory()
    host.fill_(1)
    d = torch.empty(nbytes_chunk * min(n, 8), dtype=torch.uint8, device=f"cuda:{dev}")
    ss = [torch.cuda.Stream(dev) for _ in range(streams)]
    torch.cuda.synchronize(dev)
    best = 0.0
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(n):
            k = i % min(n, 8)
            with torch.cuda.stream(ss[i % streams]):
                d[k * nbytes_chunk:(k + 1) * nbytes_chunk].copy_(host[k * nbytes_chunk:(k + 1) * nbytes_chunk], non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        best = max(best, n * nbytes_chunk / dt / 1e9)
    return best

if __name__ == "__main__":
    nodes = cpus_by_node()
    print("numa nodes:", {k: (v[0], v[-1], len(v)) for k, v in nodes.items()}, "cpu_count", os.cpu_count())
    try:
        print("gpu0 numa:", open("/sys/bus/pci/devices/" + torch.cuda.get_device_properties(0).pci_bus_id.lower() + "/numa_node").read().strip())
    except Exception as e:
        import subprocess
        print(subprocess.run("nvidia-smi topo -m | head -12", shell=True, capture_output=True, text=True).stdout)
    allowed = sorted(os.sched_getaffinity(0))
    print("allowed cpus:", allowed[:4], "...", len(allowed))
    for node, cp in (nodes.items() if nodes else [(None, allowed)]):
        cp = [c for c in cp if c in allowed]
        if not cp: continue
        os.sched_setaffinity(0, cp)
        for chunk in (2073600, 16588800, 64 << 20):
            print(f"node {node} chunk {chunk/1e6:6.1f} MB: 1 stream {bw(chunk):5.1f} GB/s   2 streams {bw(chunk, streams=2):5.1f} GB/s", flush=True)
    os.sched_setaffinity(0, allowed)
