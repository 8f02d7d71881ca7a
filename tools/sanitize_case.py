"""Small end-to-end exercise of every kernel, for compute-sanitizer (development tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metalquicha_b200 import B200FockEngine, synth
from metalquicha_b200.engine import metric_inverse_sqrt
from oracle import df_fock_oracle as oracle
eng = B200FockEngine(0)
worst = 0.0
for (n, o, q) in [(24, 5, 40), (70, 33, 37), (130, 17, 20)]:
    b, h, d, c = synth.synth_problem(3, n, o, q)
    eng.set_tensor(b)
    f = eng.build_fock_df(h, d, c, o); worst = max(worst, np.max(np.abs(f - oracle.build_fock_df(h, b, d, c, o))))
    e = eng.last_energy()
    eng.synth_tensor(n, q, 5, synth.default_scale(n, q)); eng.build_jk(d, c, o)
    cb = synth.synth_orbitals(9, n, max(1, o - 1))
    eng.build_jk_uhf(d, c, o, cb, max(1, o - 1))
three, metric = synth.synth_physical_like_tensor(1, 24, 30, 2)
eng.build_df_tensor(three, metric, 24)
_, h, d, c = synth.synth_problem(1, 24, 5, 30, with_tensor=False)
f = eng.build_fock_df(h, d, c, 5)
worst = max(worst, np.max(np.abs(f - oracle.build_fock_df(h, oracle.whiten(three, metric), d, c, 5))))
eng.close()
print("sanitize_case max err", worst)
assert worst < 1e-10
