"""Probe: host-to-device copy rate of 12.8 MB of pinned memory on the default stream and on fresh streams."""
import torch, time
n = 12_800_000
src = torch.empty(n, dtype=torch.uint8).pin_memory()
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
def timed(stream):
    with torch.cuda.stream(stream):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); dst.copy_(src, non_blocking=True); b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b)
d = torch.cuda.default_stream()
for _ in range(3): timed(d)
print("default stream:", ["%.3f" % timed(d) for _ in range(5)])
for k in range(6):
    s = torch.cuda.Stream()
    for _ in range(2): timed(s)
    print("new stream %d (prio 0):" % k, ["%.3f" % timed(s) for _ in range(4)])
s = torch.cuda.Stream(priority=-1)
for _ in range(2): timed(s)
print("high-priority stream:", ["%.3f" % timed(s) for _ in range(4)])
# two halves back to back (like the library: t then q)
h = n // 2
def timed2(stream):
    with torch.cuda.stream(stream):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); dst[:h].copy_(src[:h], non_blocking=True); dst[h:].copy_(src[h:], non_blocking=True); b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b)
print("two copies, default:", ["%.3f" % timed2(d) for _ in range(4)])
import ctypes, numpy as np
import sys; sys.path.insert(0, ".")
# the same through the raw runtime API (what libfmatch.so does): own non-blocking stream, cudaMalloc'd destination
rt = ctypes.CDLL("libcudart.so.12")
vp = ctypes.c_void_p
stream, dptr, e0, e1 = vp(), vp(), vp(), vp()
assert rt.cudaStreamCreateWithFlags(ctypes.byref(stream), 1) == 0
assert rt.cudaMalloc(ctypes.byref(dptr), ctypes.c_size_t(n + 4096)) == 0
rt.cudaEventCreate(ctypes.byref(e0)); rt.cudaEventCreate(ctypes.byref(e1))
def raw(dst_ptr, pieces):
    rt.cudaEventRecord(e0, stream)
    off = 0
    for p in pieces:
        rt.cudaMemcpyAsync(vp(dst_ptr + off), vp(src.data_ptr() + off), ctypes.c_size_t(p), 1, stream)
        off += p
    rt.cudaEventRecord(e1, stream)
    rt.cudaStreamSynchronize(stream)
    ms = ctypes.c_float()
    rt.cudaEventElapsedTime(ctypes.byref(ms), e0, e1)
    return ms.value
for _ in range(2): raw(dptr.value, [n])
print("raw API, cudaMalloc dst, one copy :", ["%.3f" % raw(dptr.value, [n]) for _ in range(4)])
print("raw API, cudaMalloc dst, two copies:", ["%.3f" % raw(dptr.value, [h, h]) for _ in range(4)])
print("raw API, torch dst, two copies     :", ["%.3f" % raw(dst.data_ptr(), [h, h]) for _ in range(4)])
# destination offset by 256-byte multiples like the library's layout
print("raw API, cudaMalloc dst + 256      :", ["%.3f" % raw(dptr.value + 256, [h, h]) for _ in range(4)])
src2 = torch.empty(h, dtype=torch.uint8).pin_memory(); src3 = torch.empty(h, dtype=torch.uint8).pin_memory()
def raw2():
    rt.cudaEventRecord(e0, stream)
    rt.cudaMemcpyAsync(vp(dptr.value), vp(src2.data_ptr()), ctypes.c_size_t(h), 1, stream)
    rt.cudaMemcpyAsync(vp(dptr.value + h), vp(src3.data_ptr()), ctypes.c_size_t(h), 1, stream)
    rt.cudaEventRecord(e1, stream); rt.cudaStreamSynchronize(stream)
    ms = ctypes.c_float(); rt.cudaEventElapsedTime(ctypes.byref(ms), e0, e1); return ms.value
print("raw API, two separate pinned tensors:", ["%.3f" % raw2() for _ in range(4)])
