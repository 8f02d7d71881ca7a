"""Development tool: the general-size device-resident SCF at a c2-like shape vs the host-driven loop."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from metalquicha_b200 import B200FockEngine, synth
from oracle import scf_oracle as scf
from test_gpu_device_scf import _synthetic_fragment

eng = B200FockEngine(0)
for (n, n_occ, naux) in ((200, 40, 400), (688, 80, 1800)):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n)) * (0.15 / np.sqrt(n)); s = np.asfortranarray(np.eye(n) + a + a.T)
    h = np.asfortranarray(synth.synth_core_hamiltonian(n, n) - 2.0 * np.diag(np.linspace(1.0, 0.0, n)))
    eng.synth_tensor(n, naux, 5, 0.3 * synth.default_scale(n, naux))
    def fb(h_, d, c, no):
        f = eng.build_fock_df(np.asfortranarray(h_), np.asfortranarray(d), np.asfortranarray(c), no)
        return f, eng.last_energy()
    t0 = time.perf_counter(); r = eng.run_scf(h, s, 2 * n_occ, max_iter=30); t_dev = time.perf_counter() - t0
    t0 = time.perf_counter(); r = eng.run_scf(h, s, 2 * n_occ, max_iter=30); t_dev = min(t_dev, time.perf_counter() - t0)
    t0 = time.perf_counter(); ref = scf.run_rhf(h, s, 2 * n_occ, fb, max_iter=30); t_host = time.perf_counter() - t0
    print(f"n={n} naux={naux}: device-resident {r['iterations']} it E={r['electronic']:.9f} {1e3*t_dev:.1f} ms ({1e3*t_dev/(r['iterations']+1):.1f} ms/it, launches {eng.last_launches()});"
          f" host-driven {ref['iterations']} it E={ref['electronic']:.9f} {1e3*t_host:.1f} ms ({1e3*t_host/(ref['iterations']+1):.1f} ms/it)")
eng.close()
