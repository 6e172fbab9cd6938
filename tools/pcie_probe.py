#!/usr/bin/env python
"""What the host link of this box delivers: pinned H2D / D2H / both at once, strided (2-D) DMA of 48-byte fields out of 424-byte records,
and the same fields read / written by a kernel through the mapped alias.  Prints GB/s; used to judge the e2e numbers of bench.py."""
import ctypes as C
import time
import torch

rt = C.CDLL("libcudart.so.12")
N = 10_000_000
REC = 424
dev = torch.device("cuda:0")
host = torch.zeros(N * REC, dtype=torch.uint8, pin_memory=True)
host2 = torch.zeros(N * REC, dtype=torch.uint8, pin_memory=True)
d = torch.zeros(N * REC, dtype=torch.uint8, device=dev)
d2 = torch.zeros(N * REC, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


gb = N * REC / 1e9
t = timed(lambda: d.copy_(host, non_blocking=True)); print(f"H2D pinned            {gb / t:6.1f} GB/s ({t * 1e3:.1f} ms)")
t = timed(lambda: host.copy_(d, non_blocking=True)); print(f"D2H pinned            {gb / t:6.1f} GB/s ({t * 1e3:.1f} ms)")


def both():
    with torch.cuda.stream(s1):
        d.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2):
        host2.copy_(d2, non_blocking=True)


t = timed(both); print(f"H2D + D2H concurrent  {2 * gb / t:6.1f} GB/s total ({t * 1e3:.1f} ms)")
# 2-D DMA: 48 bytes of every record
kindH2D, kindD2H = 1, 2
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
for w in (48, 64, 128):
    t = timed(lambda: rt.cudaMemcpy2DAsync(d.data_ptr(), REC, host.data_ptr(), REC, w, N, kindH2D, None))
    print(f"2-D DMA H2D width {w:3d}   {N * w / 1e9 / t:6.1f} GB/s of payload ({t * 1e3:.1f} ms)")
    t = timed(lambda: rt.cudaMemcpy2DAsync(host.data_ptr(), REC, d.data_ptr(), REC, w, N, kindD2H, None))
    print(f"2-D DMA D2H width {w:3d}   {N * w / 1e9 / t:6.1f} GB/s of payload ({t * 1e3:.1f} ms)")
# packed fields: one contiguous 48 N byte copy (what a host-side pack would ship)
t = timed(lambda: d[:48 * N].copy_(host[:48 * N], non_blocking=True)); print(f"contiguous 48N H2D    {48 * N / 1e9 / t:6.1f} GB/s ({t * 1e3:.1f} ms)")


def both2d():
    rt.cudaMemcpy2DAsync(d.data_ptr(), REC, host.data_ptr(), REC, 128, N, kindH2D, C.c_void_p(s1.cuda_stream))
    rt.cudaMemcpy2DAsync(C.c_void_p(host2.data_ptr() + 64), REC, C.c_void_p(d2.data_ptr() + 64), REC, 64, N, kindD2H, C.c_void_p(s2.cuda_stream))


t = timed(both2d); print(f"2-D DMA H2D width 128 + D2H width 64 concurrently: {t * 1e3:.1f} ms (alone: see above)")


def both2d_chunked(k=8):
    per = N // k
    for i in range(k):
        o = i * per * REC
        rt.cudaMemcpy2DAsync(C.c_void_p(d.data_ptr() + o), REC, C.c_void_p(host.data_ptr() + o), REC, 128, per, kindH2D, C.c_void_p(s1.cuda_stream))
        rt.cudaMemcpy2DAsync(C.c_void_p(host2.data_ptr() + o + 64), REC, C.c_void_p(d2.data_ptr() + o + 64), REC, 64, per, kindD2H, C.c_void_p(s2.cuda_stream))


t = timed(both2d_chunked); print(f"the same in 8 chunks each: {t * 1e3:.1f} ms")
