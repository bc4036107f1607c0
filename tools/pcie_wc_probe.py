#!/usr/bin/env python3
"""Does write-combined pinned memory move the PCIe ceiling?  H2D alone and H2D + D2H at once, with
the upload buffer allocated cudaHostAllocDefault and cudaHostAllocWriteCombined (cuda-python,
no torch).  An experiment for bench.py's e2e leg, which is bound by exactly these copies."""
import json
import time

from cuda.bindings import runtime as rt


def ck(res):
    err, *rest = res
    if int(err) != 0:
        raise RuntimeError(rt.cudaGetErrorString(err)[1].decode())
    return rest[0] if len(rest) == 1 else rest


n = 1 << 29
ck(rt.cudaSetDevice(0))
d_a, d_b = ck(rt.cudaMalloc(n)), ck(rt.cudaMalloc(n))
h_out = ck(rt.cudaHostAlloc(n, rt.cudaHostAllocDefault))
s1, s2 = ck(rt.cudaStreamCreate()), ck(rt.cudaStreamCreate())
out = {"unit": "GB/s"}
for name, flags in (("default", rt.cudaHostAllocDefault), ("write_combined", rt.cudaHostAllocWriteCombined)):
    h_in = ck(rt.cudaHostAlloc(n, flags))
    ck(rt.cudaMemset(d_a, 0, n))

    def h2d():
        ck(rt.cudaMemcpyAsync(d_a, h_in, n, rt.cudaMemcpyKind.cudaMemcpyHostToDevice, s1))

    def both():
        ck(rt.cudaMemcpyAsync(d_a, h_in, n, rt.cudaMemcpyKind.cudaMemcpyHostToDevice, s1))
        ck(rt.cudaMemcpyAsync(h_out, d_b, n, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost, s2))

    for label, fn in (("h2d_alone", h2d), ("both_each_way", both)):
        fn()
        ck(rt.cudaDeviceSynchronize())
        best = 0.0
        for _ in range(3):
            t0 = time.perf_counter()
            for _ in range(8):
                fn()
            ck(rt.cudaDeviceSynchronize())
            best = max(best, 8 * n / 1e9 / (time.perf_counter() - t0))
        out["%s_%s" % (name, label)] = round(best, 1)
    ck(rt.cudaFreeHost(h_in))
print(json.dumps(out))
