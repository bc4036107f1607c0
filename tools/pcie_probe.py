#!/usr/bin/env python3
"""PCIe ceiling of the box: pinned H2D alone, D2H alone, both at once (torch, two streams).
The e2e leg of bench.py moves 118 MB each way per 8K frame; this is the number it is bound by."""
import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def h2d():
    d_a.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_b, non_blocking=True)


def both():
    ev = torch.cuda.Event()
    ev.record()
    s1.wait_event(ev)
    s2.wait_event(ev)
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)


gb = n / 1e9
print("H2D alone  %.1f GB/s" % (gb / timed(h2d) * 1e3))
print("D2H alone  %.1f GB/s" % (gb / timed(d2h) * 1e3))
t = timed(both)
print("both       %.1f GB/s each way (%.1f fps of 8K frames)" % (gb / t * 1e3, gb / t * 1e3 / 0.11796))
