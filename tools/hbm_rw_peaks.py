#!/usr/bin/env python
"""Read-only and write-only HBM streams on this GPU, next to the read+write copy figure MEASURED_PEAKS.json holds:
the yardsticks for kernels that only read (match filter: the bf16 gallery) or only write (stem: 411 MB of activations).
CUDA events, best of 10, 2 GiB buffers (>> 126 MB L2)."""
import json
import torch

dev = torch.device("cuda", 0)
n = 1 << 30                                    # 2 GiB of bf16
a = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_()
b = torch.empty_like(a)


def best(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        t.append(s.elapsed_time(e))
    return min(t)

nbytes = a.numel() * 2
copy_ms = best(lambda: b.copy_(a))
write_ms = best(lambda: b.zero_())              # cudaMemsetAsync-class write stream
fill_ms = best(lambda: b.fill_(1.5))            # a kernel writing 16 B per thread
read_ms = best(lambda: a.view(torch.int32).sum())   # reduction kernel: reads everything, writes nothing
i8 = a.view(torch.int8)
read2_ms = best(lambda: torch.max(i8))
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facerecognitionpipeline_b200 import _native
ctx = _native.Context(0)
ms = C.c_float(0)
ctx.frb_debug_stream_bw(a.data_ptr(), nbytes, 0, 10, C.byref(ms))
own_read = nbytes / ms.value / 1e6
ctx.frb_debug_stream_bw(b.data_ptr(), nbytes, 1, 10, C.byref(ms))
own_write = nbytes / ms.value / 1e6
print(json.dumps(dict(read_only_gbs=own_read, write_only_gbs=own_write, note="libfrb200 stream probes (frb_debug_stream_bw): 16-byte loads / stores, 4 in flight per thread")))
print(json.dumps(dict(copy_gbs=2 * nbytes / copy_ms / 1e6, write_memset_gbs=nbytes / write_ms / 1e6, write_fill_gbs=nbytes / fill_ms / 1e6,
                      read_sum_gbs=nbytes / read_ms / 1e6, read_max_gbs=nbytes / read2_ms / 1e6,
                      note="2 GiB buffers, best of 10, CUDA events; copy counts read + write bytes")))
