"""Is PCIe full duplex here?  Concurrent D2H + H2D on two streams, device-timed, at the e2e leg's sizes and larger."""
import torch
dev = "cuda:0"
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for mb_out, mb_in in ((64, 64), (4.46, 1.05), (4.46, 0.0), (4.19, 0.0), (4.46, 4.46)):
    d1 = torch.empty(int(mb_out * 1e6), dtype=torch.uint8, device=dev); h1 = torch.empty(int(mb_out * 1e6), dtype=torch.uint8).pin_memory()
    d2 = torch.empty(max(1, int(mb_in * 1e6)), dtype=torch.uint8, device=dev); h2 = torch.empty(max(1, int(mb_in * 1e6)), dtype=torch.uint8).pin_memory()
    reps = 50
    def run():
        start = torch.cuda.Event(enable_timing=True); end1 = torch.cuda.Event(enable_timing=True); end2 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record()
        s1.wait_event(start); s2.wait_event(start)
        with torch.cuda.stream(s1):
            for _ in range(reps): h1.copy_(d1, non_blocking=True)
            end1.record()
        if mb_in > 0:
            with torch.cuda.stream(s2):
                for _ in range(reps): d2.copy_(h2, non_blocking=True)
                end2.record()
        torch.cuda.synchronize()
        return start.elapsed_time(end1) / reps, (start.elapsed_time(end2) / reps if mb_in > 0 else 0.0)
    run()
    t1, t2 = run()
    print(f"d2h {mb_out} MB: {t1*1e3:.1f} us/copy ({mb_out/t1:.1f} GB/s)   h2d {mb_in} MB: {t2*1e3:.1f} us/copy" + (f" ({mb_in/t2:.1f} GB/s)" if mb_in > 0 else ""))
