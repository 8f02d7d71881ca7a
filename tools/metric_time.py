import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metalquicha_b200 import B200FockEngine
eng = B200FockEngine(0)
for naux in (600, 1800, 1800, 1800):
    rng = np.random.default_rng(1)
    u, _ = np.linalg.qr(rng.standard_normal((naux, naux)))
    m = np.asfortranarray((u * np.exp(rng.uniform(-2, 2, naux))[None, :]) @ u.T); m = np.asfortranarray(0.5 * (m + m.T))
    t0 = time.perf_counter(); h = eng.metric_inverse_sqrt(m); w = time.perf_counter() - t0
    print(naux, "wall %.1f ms" % (1e3 * w), "device/sweeps", eng.last_metric())
eng.close()
