"""H2D bandwidth of one 146.7 MB copy (the c3 step's inputs): torch pinned memory vs write-combined pinned memory (cudaHostAlloc),
one copy vs two halves on two streams.  python tools/h2d_probe.py"""
import ctypes
import time

import torch
from cuda import cudart

N = 146_700_768
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
pin = torch.empty(N, dtype=torch.uint8).pin_memory()
err, wc_ptr = cudart.cudaHostAlloc(N, cudart.cudaHostAllocWriteCombined)
assert err == cudart.cudaError_t.cudaSuccess, err
ctypes.memset(wc_ptr, 1, N)
err, pl_ptr = cudart.cudaHostAlloc(N, cudart.cudaHostAllocDefault)
ctypes.memset(pl_ptr, 1, N)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
H2D = cudart.cudaMemcpyKind.cudaMemcpyHostToDevice


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return dt * 1e3, N / dt / 1e9


def torch_pinned():
    with torch.cuda.stream(s1):
        dev.copy_(pin, non_blocking=True)


def raw(ptr):
    def f():
        cudart.cudaMemcpyAsync(dev.data_ptr(), ptr, N, H2D, s1.cuda_stream)
    return f


def raw_split(ptr):
    h = N // 2 // 4096 * 4096
    def f():
        cudart.cudaMemcpyAsync(dev.data_ptr(), ptr, h, H2D, s1.cuda_stream)
        cudart.cudaMemcpyAsync(dev.data_ptr() + h, ptr + h, N - h, H2D, s2.cuda_stream)
    return f


for name, fn in (("torch pinned, one copy", torch_pinned), ("cudaHostAlloc default, one copy", raw(pl_ptr)),
                 ("cudaHostAlloc write-combined, one copy", raw(wc_ptr)), ("write-combined, two halves on two streams", raw_split(wc_ptr)),
                 ("default pinned, two halves on two streams", raw_split(pl_ptr))):
    ms, gbs = timeit(fn)
    print(f"{name:48s} {ms:7.3f} ms  {gbs:6.1f} GB/s")
