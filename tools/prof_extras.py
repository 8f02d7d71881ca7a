"""Development tool: one call each of the round-2 secondary paths, for an ncu capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from metalquicha_b200 import B200FockEngine, synth
from test_gpu_device_scf import _synthetic_fragment

eng = B200FockEngine(0)
# rank-2 response operator at the c2 orbital dimensions on a 256-function auxiliary range
n, o, q = 688, 80, 256
eng.synth_tensor(n, 1800, 7, synth.default_scale(n, 1800), q_begin=0, q_count=q)
c = synth.synth_orbitals(1, n, n)
x = np.asfortranarray(1e-3 * c[:, o:2 * o]); co = np.asfortranarray(c[:, :o])
d = np.asfortranarray(x @ co.T + co @ x.T)
g = eng.response_operator_df(x, co, d)
# tensor construction: metric^(-1/2) + slab-streamed whitening, naux = 600, n = 200
n2, q2 = 200, 600
rng = np.random.default_rng(2)
u, _ = np.linalg.qr(rng.standard_normal((q2, q2)))
metric = np.asfortranarray((u * np.exp(rng.uniform(-2, 2, q2))[None, :]) @ u.T); metric = np.asfortranarray(0.5 * (metric + metric.T))
three = np.asfortranarray(rng.standard_normal((n2 * n2, q2)))
three = three.reshape(n2, n2, q2, order="F"); three = np.asfortranarray((0.5 * (three + three.transpose(1, 0, 2))).reshape(n2 * n2, q2, order="F"))
half = eng.build_df_tensor(three, metric, n2)
# gradient densities on that tensor
c2 = synth.synth_orbitals(3, n2, 40)
gamma, omega = eng.df_gradient_densities(half, np.asfortranarray(2 * c2 @ c2.T), c2, 40)
# device-resident SCF of a trimer-sized fragment
s, h, b = _synthetic_fragment(972, 72, 15, 340)
eng.set_tensor(b)
r = eng.run_scf_fragment(h, s, 30)
print("ok", float(np.abs(g).max()), float(np.abs(omega).max()), r["iterations"])
eng.close()
