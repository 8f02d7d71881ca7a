"""Times the device whitening GEMM b = three . half at a given shape (development tool).
usage: python tools/whiten_bench.py n naux"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metalquicha_b200 import B200FockEngine
n, q = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(0)
three = np.empty((n * n, q), order="F")
for p in range(q):                      # symmetric slabs, cheap to make
    a = rng.standard_normal((n, n)); three[:, p] = (a + a.T).reshape(-1)
half = rng.standard_normal((q, q)); half = np.asfortranarray(half + half.T)
eng = B200FockEngine(0)
for rep in range(2):
    t0 = time.perf_counter(); eng.set_tensor_from_3c(three, half, n); wall = time.perf_counter() - t0
    ms, fl = eng.last_whiten()
print(json.dumps({"n": n, "naux": q, "whiten_gemm_ms": round(ms, 3), "tflops": round(fl / ms * 1e-9, 2),
                  "reference_gemm_flops_ratio": round(2.0 * q * q * n * n / fl, 3), "wall_s_incl_upload_and_pack": round(wall, 3)}))
t0 = time.perf_counter(); b = three @ half; cpu = time.perf_counter() - t0
print(json.dumps({"cpu_dgemm_s": round(cpu, 3), "cpu_tflops": round(2.0 * q * q * n * n / cpu * 1e-12, 3), "cores": os.cpu_count()}))
