"""Host-to-device copy of n pcl::PointXYZ records: plain 16 B per point against a strided (2-D) DMA that skips the padding
word (12 B of every 16).  cuda-python runtime bindings."""
import time
from cuda import cudart
def ck(r):
    assert r[0] == cudart.cudaError_t.cudaSuccess, r[0]
    return r[1:] if len(r) > 2 else (r[1] if len(r) == 2 else None)
n = 10_000_000
ck(cudart.cudaSetDevice(0))
h = ck(cudart.cudaHostAlloc(16 * n, 0))
d = ck(cudart.cudaMalloc(16 * n))
s = ck(cudart.cudaStreamCreate())
e0 = ck(cudart.cudaEventCreate()); e1 = ck(cudart.cudaEventCreate())
def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        ck(cudart.cudaEventRecord(e0, s)); fn(); ck(cudart.cudaEventRecord(e1, s)); ck(cudart.cudaStreamSynchronize(s))
        best = min(best, ck(cudart.cudaEventElapsedTime(e0, e1)))
    return best
k = cudart.cudaMemcpyKind.cudaMemcpyHostToDevice
t = timed(lambda: ck(cudart.cudaMemcpyAsync(d, h, 16 * n, k, s)))
print(f"plain 16 B/pt: {t:.3f} ms = {16 * n / t / 1e6:.1f} GB/s")
for width, spitch, dpitch, rows in ((12, 16, 12, n), (12, 16, 16, n), (16 * 1024 - 4, 16 * 1024, 16 * 1024, n // 1024)):
    t = timed(lambda: ck(cudart.cudaMemcpy2DAsync(d, dpitch, h, spitch, width, rows, k, s)), reps=3)
    print(f"2-D width {width} spitch {spitch} dpitch {dpitch} rows {rows}: {t:.3f} ms = {width * rows / t / 1e6:.1f} GB/s payload")
