"""Host -> device ingest ceiling of the box: N processes, one per GPU, each streaming pinned 1080p-luma-sized (2 MB) and
2160p-10-bit-luma-sized (16.6 MB) chunks to its GPU at the same time, for N = 1, 2, 4, 8 (as many as are visible).
Prints per-GPU and aggregate GB/s: if the aggregate stops growing with N, the host (memory / PCIe root) is the limiter of
the multi-GPU end-to-end numbers, not the engine.  Uses the library's own pinned allocator and cudaMemcpyAsync path
(bv_submit into a PSNR-only context would add kernels; this measures the copies alone, through torch for brevity)."""
import multiprocessing as mp
import os, sys, time


def worker(rank, n, bar, q, chunk, total):
    import torch
    torch.cuda.set_device(rank)
    nb = 8
    host = torch.empty(chunk * nb, dtype=torch.uint8).pin_memory()
    host.fill_(rank + 1)
    dev = torch.empty(chunk * nb, dtype=torch.uint8, device=f"cuda:{rank}")
    st = torch.cuda.Stream(rank)
    iters = max(1, total // chunk)
    for rep in range(3):
        torch.cuda.synchronize(rank)
        bar.wait()
        t0 = time.perf_counter()
        with torch.cuda.stream(st):
            for i in range(iters):
                k = i % nb
                dev[k * chunk:(k + 1) * chunk].copy_(host[k * chunk:(k + 1) * chunk], non_blocking=True)
        torch.cuda.synchronize(rank)
        dt = time.perf_counter() - t0
        bar.wait()
    q.put((rank, iters * chunk / dt / 1e9))


def main():
    import torch
    ng = torch.cuda.device_count()
    print(f"visible GPUs: {ng}; host cpus: {os.cpu_count()}", flush=True)
    ctx = mp.get_context("spawn")
    for chunk, label in ((2073600, "1080p 8-bit luma plane (2.07 MB)"), (16588800, "2160p 10-bit luma plane (16.6 MB)")):
        for n in (1, 2, 4, 8):
            if n > ng:
                continue
            bar, q = ctx.Barrier(n), ctx.Queue()
            ps = [ctx.Process(target=worker, args=(r, n, bar, q, chunk, 4 << 30)) for r in range(n)]
            [p.start() for p in ps]
            res = sorted(q.get(timeout=300) for _ in range(n))
            [p.join() for p in ps]
            per = [round(v, 1) for _, v in res]
            print(f"{label}: {n} concurrent uploader(s): aggregate {sum(per):6.1f} GB/s, per GPU {per}", flush=True)


if __name__ == "__main__":
    main()
