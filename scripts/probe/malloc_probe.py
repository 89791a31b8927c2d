import ctypes, time
rt = ctypes.CDLL("libcudart.so.12") if False else None
import torch
torch.cuda.init(); torch.zeros(1, device="cuda")
rt = ctypes.CDLL(torch.__file__.rsplit("/",1)[0] + "/../nvidia/cuda_runtime/lib/libcudart.so.12")
rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
rt.cudaFree.argtypes = [ctypes.c_void_p]
for gib in (1, 1, 4, 16, 32, 1):
    p = ctypes.c_void_p()
    t0 = time.perf_counter(); rc = rt.cudaMalloc(ctypes.byref(p), gib << 30); t1 = time.perf_counter()
    print(f"cudaMalloc {gib} GiB rc={rc}: {1e3*(t1-t0):.2f} ms")
ps = []
t0 = time.perf_counter()
for i in range(16):
    p = ctypes.c_void_p(); rt.cudaMalloc(ctypes.byref(p), 1 << 30); ps.append(p)
print(f"16 x 1 GiB: {1e3*(time.perf_counter()-t0):.2f} ms")
