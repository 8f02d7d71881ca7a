"""Development tool: a fragment's whole SCF, device-resident vs host-driven (engine Fock build + NumPy SCF step)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from metalquicha_b200 import B200FockEngine
from oracle import scf_oracle as scf
from test_gpu_device_scf import _synthetic_fragment

eng = B200FockEngine(0)
for (n, n_occ, naux) in ((72, 15, 340), (48, 10, 227), (24, 5, 113)):
    s, h, b = _synthetic_fragment(900 + n, n, n_occ, naux)
    eng.set_tensor(b)
    def fb(h_, d, c, no):
        f = eng.build_fock_df(np.asfortranarray(h_), np.asfortranarray(d), np.asfortranarray(c), no)
        return f, eng.last_energy()
    for tag, fn in (("device", lambda ce=1: eng.run_scf_fragment(h, s, 2 * n_occ, check_every=ce)),
                    ("device_q4", lambda: eng.run_scf_fragment(h, s, 2 * n_occ, check_every=4)),
                    ("host", lambda: scf.run_rhf(h, s, 2 * n_occ, fb))):
        fn(); ts = []
        for _ in range(5):
            t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
        it = r["iterations"]
        print(f"n={n} {tag:10s} iterations={it} E={r['electronic']:.10f} best {1e3*min(ts):.3f} ms  = {1e6*min(ts)/(it+1):.1f} us per build+step")
eng.close()
