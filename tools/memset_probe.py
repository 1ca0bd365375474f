#!/usr/bin/env python
"""probe: do cudaMemsetAsync calls on one stream queue behind a large host->device copy on another stream?
(20 memsets of 4 KB between two events, with and without a 360 MB upload in flight)"""
import ctypes as C, torch, time
rt = C.CDLL("libcudart.so.12")
rt.cudaMemsetAsync.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
host = torch.empty(360 << 20, dtype=torch.uint8, pin_memory=True)
dev = torch.empty(360 << 20, dtype=torch.uint8, device="cuda")
small = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
def run(with_copy, what):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if with_copy:
        with torch.cuda.stream(sb): dev.copy_(host, non_blocking=True)
        time.sleep(0.001)
    with torch.cuda.stream(sa):
        e0.record()
        for i in range(20):
            if what == "memset": rt.cudaMemsetAsync(small.data_ptr() + 4096 * i, 0, 4096, sa.cuda_stream)
            elif what == "fill": small[4096 * i: 4096 * (i + 1)].zero_()
            else: small[4096 * i: 4096 * (i + 1)].copy_(small[4096 * (i + 40): 4096 * (i + 41)])
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for what in ("memset", "fill", "d2d"):
    for wc in (False, True, False, True):
        run(wc, what)
        print(f"20 x {what} 4 KB, upload in flight: {wc}: {run(wc, what):.3f} ms", flush=True)
