"""D2H rate of strided 2-D copies (rows of a knot-major value array) vs one contiguous copy -- development tool for the direct
host-buffer path (csrc/qck_pipe.cpp)."""
import ctypes, time
import torch

rt = ctypes.CDLL("libcudart.so.12")
nnz, knots = 6674, 9999
dev = torch.empty(nnz * knots, dtype=torch.float64, device="cuda:0").normal_()
host = torch.empty(nnz * knots, dtype=torch.float64).pin_memory()
st = torch.cuda.Stream()
D2H = 2


def run(width, rows, pitch=nnz, reps=5, pieces=1):
    best = 0.0
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        per = rows // pieces
        for p in range(pieces):
            off = p * per * pitch * 8
            if width == pitch:
                rc = rt.cudaMemcpyAsync(ctypes.c_void_p(host.data_ptr() + off), ctypes.c_void_p(dev.data_ptr() + off), ctypes.c_size_t(width * per * 8), D2H,
                                        ctypes.c_void_p(st.cuda_stream))
            else:
                rc = rt.cudaMemcpy2DAsync(ctypes.c_void_p(host.data_ptr() + off), ctypes.c_size_t(pitch * 8), ctypes.c_void_p(dev.data_ptr() + off),
                                          ctypes.c_size_t(pitch * 8), ctypes.c_size_t(width * 8), ctypes.c_size_t(per), D2H, ctypes.c_void_p(st.cuda_stream))
            assert rc == 0, rc
        st.synchronize()
        best = max(best, width * per * pieces * 8 / (time.perf_counter() - t0) * 1e-9)
    return best


print(f"contiguous {nnz * knots * 8e-6:.0f} MB: {run(nnz, knots):.1f} GB/s")
for w in (8, 162, 324, 834, 1643, 3303):
    print(f"2-D copy, rows of {w:5d} doubles ({w * 8:6d} B), pitch {nnz * 8} B, {knots} rows: {run(w, knots):6.1f} GB/s   in 32 pieces: {run(w, knots, pieces=32):6.1f} GB/s   in 128 pieces: {run(w, knots, pieces=128):6.1f} GB/s")

# page-locked by cudaHostAlloc (torch pin_memory) vs. a pageable numpy array registered afterwards (cudaHostRegister)
import numpy as np, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qcknot
arr = np.empty(nnz * knots)
arr[:] = 0.0
qcknot.host_register(arr)
for name, ptr in (("cudaHostAlloc", host.data_ptr()), ("cudaHostRegister(numpy)", arr.ctypes.data)):
    best = 0.0
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rt.cudaMemcpyAsync(ctypes.c_void_p(ptr), ctypes.c_void_p(dev.data_ptr()), ctypes.c_size_t(nnz * knots * 8), D2H, ctypes.c_void_p(st.cuda_stream))
        st.synchronize()
        best = max(best, nnz * knots * 8 / (time.perf_counter() - t0) * 1e-9)
    print(f"contiguous D2H into {name}: {best:.1f} GB/s")
qcknot.host_unregister(arr)
