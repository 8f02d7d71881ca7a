import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metalquicha_b200 import B200FockEngine, synth
from oracle import gto_integrals as gto, scf_oracle as scf
eng = B200FockEngine(0)
for n in (1, 2, 3, 4, 5, 6, 7, 8, 9, 33):
    b = 0.1 * synth.synth_tensor(1, n, 5)
    eng.set_tensor(b)
    rng = np.random.default_rng(n); a = rng.standard_normal((n, n)) * 0.05; s = np.eye(n) + a + a.T
    h = synth.synth_core_hamiltonian(1, n)
    try:
        r = eng.run_scf_fragment(h, s, 2 * max(1, n // 3), max_iter=30)
        print(n, "ok n_mo", r["orbitals"].shape, r["iterations"], r["electronic"])
    except Exception as ex:
        print(n, "ERR", ex)
sym, xyz, nel, eref = gto.H2O_STO3G
s, h, eri, enuc = gto.molecule_integrals(sym, xyz)
eng.set_tensor(scf.exact_fit_tensor(eri))
for ne in (2, 4, 10):
    try:
        r = eng.run_scf_fragment(h, s, ne, e_nuc=enuc); print("water", ne, r["energy"], r["iterations"], r["orbitals"].shape)
    except Exception as ex:
        print("water", ne, "ERR", ex)
print(np.linalg.eigvalsh(s))
