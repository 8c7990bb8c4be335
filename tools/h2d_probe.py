"""Host->device copy rate by kind of host allocation (default pinned, write-combined, registered malloc) and by who wrote
the buffer last (CPU or a D2H DMA) - why the upload half of the host-buffer path runs below the download half here."""
import ctypes as C
import json
import time

import torch

rt = C.CDLL('libcudart.so.12')
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
H2D, D2H = 1, 2
n = 8 << 20
dev = torch.zeros(n, dtype=torch.uint8, device='cuda')
stream = torch.cuda.current_stream().cuda_stream


def rate(host_ptr, kind, reps=40):
    for _ in range(3):
        rt.cudaMemcpyAsync(dev.data_ptr() if kind == H2D else host_ptr, host_ptr if kind == H2D else dev.data_ptr(), n, kind, stream)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        rt.cudaMemcpyAsync(dev.data_ptr() if kind == H2D else host_ptr, host_ptr if kind == H2D else dev.data_ptr(), n, kind, stream)
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0) / 1e9


out = {}
for name, flags in (('default', 0), ('portable', 1), ('mapped', 2), ('write_combined', 4)):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), n, flags) == 0
    C.memset(p, 1, n)
    out[name] = {'h2d_after_cpu_write': rate(p.value, H2D), 'd2h': rate(p.value, D2H), 'h2d_after_dma_write': rate(p.value, H2D)}
import numpy as np
a = np.ones(n + 4096, dtype=np.uint8)
ptr = (a.ctypes.data + 4095) & ~4095
assert rt.cudaHostRegister(ptr, n, 0) == 0
out['registered_malloc'] = {'h2d': rate(ptr, H2D), 'd2h': rate(ptr, D2H)}
t = torch.ones(n, dtype=torch.uint8).pin_memory()
out['torch_pin_memory'] = {'h2d': rate(t.data_ptr(), H2D), 'd2h': rate(t.data_ptr(), D2H)}
print(json.dumps(out, indent=1))
