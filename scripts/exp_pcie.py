"""PCIe copy bandwidth of this box (pinned host memory, cudaMemcpyAsync via torch) at the e2e leg's transfer sizes."""
import torch, time
dev = "cuda:0"
for mb in (0.26, 1.0, 4.46, 16, 64, 256):
    nbytes = int(mb * 1e6)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    for name, (dst, src) in (("d2h", (h, d)), ("h2d", (d, h))):
        for _ in range(3): dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps): dst.copy_(src, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name} {mb:7.2f} MB: {ms*1e3:8.1f} us  {nbytes/ms/1e6:6.1f} GB/s")
# both directions at once on two streams
nbytes = int(4.46e6)
d1 = torch.empty(nbytes, dtype=torch.uint8, device=dev); h1 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d2 = torch.empty(int(1.05e6), dtype=torch.uint8, device=dev); h2 = torch.empty(int(1.05e6), dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 200
print(f"duplex d2h 4.46 MB + h2d 1.05 MB: {dt*1e6:.1f} us per pair -> d2h {nbytes/dt/1e9:.1f} GB/s")
