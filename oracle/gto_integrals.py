"""Gaussian integrals over contracted Cartesian s/p/d shells -- TEST INFRASTRUCTURE ONLY.

Why this exists: the reference (JorgeG94/metalquicha) holds NO J/K/F element-level vectors,
and its named basis sets (the BSE bundle) are not in the tree -- but two of its validation
programs write their basis out inline and assert total energies against PySCF to 1e-9 Eh:

    validation/check_rhf.f90:54-92    H2  / STO-3G, R = 1.4 bohr        E = -1.1167143251
    validation/check_rhf.f90:94-153   H2O / STO-3G, standard geometry   E = -74.9658162796
    (basis data: hydrogen_sto3g :148-159, oxygen_sto3g :161-178 of the same file)

and its validation manifest holds density-fitted energies in a basis whose data is published and (for H and O)
also present in the tree in another form:

    validation/validation_tests_cpu.json:899-904  H2O / 6-31G* fitted with 6-31G*    E = -76.188111755038
    validation/validation_tests_cpu.json:905-910  CH4 / 6-31G** fitted with 6-31G**  E = -40.381603512964
    validation/validation_tests_cpu.json:719-723  H2O / 6-31G*, exact integrals      E = -76.010317945971

libcint (third party, absent: JorgeG94/libfint v0.1.1) supplies the integrals there.  This
module restates the PUBLISHED McMurchie-Davidson scheme (J. Comput. Phys. 26, 218 (1978);
Helgaker, Jorgensen, Olsen, "Molecular Electronic-Structure Theory", ch. 9) for exactly those
shells, so that the oracle's SCF restatement (oracle/scf_oracle.py) -- and the CUDA engine behind
the same loop -- can be run on REAL integrals and pinned to the reference-held energies.
The SCF energy does not depend on the normalisation or ordering of the basis functions, so no
libcint convention enters.  Pure NumPy/Python loops: small cases only.
"""
from __future__ import annotations

import itertools
import math

import numpy as np
from scipy.special import hyp1f1


def boys(n: int, x: float) -> float:
    """F_n(x) = int_0^1 t^(2n) exp(-x t^2) dt  = 1F1(n+1/2; n+3/2; -x) / (2n+1)."""
    return float(hyp1f1(n + 0.5, n + 1.5, -x)) / (2 * n + 1)


def _double_factorial(k: int) -> int:
    return 1 if k <= 0 else k * _double_factorial(k - 2)


def cartesian_components(l: int):
    """(lx, ly, lz) of a shell: s; px py pz; dxx dxy dxz dyy dyz dzz."""
    return [(lx, ly, l - lx - ly) for lx in range(l, -1, -1) for ly in range(l - lx, -1, -1)]


class BasisFunction:
    """One contracted Cartesian Gaussian, normalised to unit self-overlap."""

    def __init__(self, center, lmn, exponents, coefficients):
        self.center = np.asarray(center, dtype=float)
        self.lmn = tuple(lmn)
        self.exps = np.asarray(exponents, dtype=float)
        l, m, n = self.lmn
        L = l + m + n
        # contraction coefficients refer to NORMALISED primitives (the BSE/STO-3G convention)
        norm = np.array([(2.0 * a / math.pi) ** 0.75 * (4.0 * a) ** (L / 2.0)
                         / math.sqrt(_double_factorial(2 * l - 1) * _double_factorial(2 * m - 1)
                                     * _double_factorial(2 * n - 1)) for a in self.exps])
        self.coefs = np.asarray(coefficients, dtype=float) * norm
        s = 0.0
        for ca, a in zip(self.coefs, self.exps):
            for cb, b in zip(self.coefs, self.exps):
                s += ca * cb * _overlap_primitive(a, self.lmn, self.center, b, self.lmn, self.center)
        self.coefs = self.coefs / math.sqrt(s)


def hermite_e(i: int, j: int, t: int, qx: float, a: float, b: float) -> float:
    """Hermite expansion coefficient E_t^{ij} of the product of two 1-D Gaussians (HJO eq. 9.5.6-7)."""
    p = a + b
    q = a * b / p
    if t < 0 or t > i + j:
        return 0.0
    if i == j == t == 0:
        return math.exp(-q * qx * qx)
    if j == 0:
        return (hermite_e(i - 1, j, t - 1, qx, a, b) / (2 * p) - (q * qx / a) * hermite_e(i - 1, j, t, qx, a, b)
                + (t + 1) * hermite_e(i - 1, j, t + 1, qx, a, b))
    return (hermite_e(i, j - 1, t - 1, qx, a, b) / (2 * p) + (q * qx / b) * hermite_e(i, j - 1, t, qx, a, b)
            + (t + 1) * hermite_e(i, j - 1, t + 1, qx, a, b))


def _overlap_primitive(a, lmn1, A, b, lmn2, B) -> float:
    p = a + b
    s = 1.0
    for d in range(3):
        s *= hermite_e(lmn1[d], lmn2[d], 0, A[d] - B[d], a, b)
    return s * (math.pi / p) ** 1.5


def _kinetic_primitive(a, lmn1, A, b, lmn2, B) -> float:
    l2, m2, n2 = lmn2
    def s(dl, dm, dn):
        t = (l2 + dl, m2 + dm, n2 + dn)
        return 0.0 if min(t) < 0 else _overlap_primitive(a, lmn1, A, b, t, B)
    term0 = b * (2 * (l2 + m2 + n2) + 3) * s(0, 0, 0)
    term1 = -2.0 * b * b * (s(2, 0, 0) + s(0, 2, 0) + s(0, 0, 2))
    term2 = -0.5 * (l2 * (l2 - 1) * s(-2, 0, 0) + m2 * (m2 - 1) * s(0, -2, 0) + n2 * (n2 - 1) * s(0, 0, -2))
    return term0 + term1 + term2


def hermite_r(t_max: int, u_max: int, v_max: int, p: float, pc) -> np.ndarray:
    """Hermite Coulomb integrals R^0_{tuv}(p, PC) for t <= t_max, u <= u_max, v <= v_max (HJO 9.9.18-20)."""
    n_max = t_max + u_max + v_max
    x, y, z = pc
    r2 = x * x + y * y + z * z
    # R^n_{000} = (-2p)^n F_n(p r^2); build downwards in n
    r = np.zeros((n_max + 1, t_max + 1, u_max + 1, v_max + 1))
    for n in range(n_max + 1):
        r[n, 0, 0, 0] = (-2.0 * p) ** n * boys(n, p * r2)
    for n in range(n_max - 1, -1, -1):
        for t in range(t_max + 1):
            for u in range(u_max + 1):
                for v in range(v_max + 1):
                    if t + u + v == 0 or t + u + v > n_max - n:
                        continue
                    if t > 0:
                        val = x * r[n + 1, t - 1, u, v] + ((t - 1) * r[n + 1, t - 2, u, v] if t > 1 else 0.0)
                    elif u > 0:
                        val = y * r[n + 1, t, u - 1, v] + ((u - 1) * r[n + 1, t, u - 2, v] if u > 1 else 0.0)
                    else:
                        val = z * r[n + 1, t, u, v - 1] + ((v - 1) * r[n + 1, t, u, v - 2] if v > 1 else 0.0)
                    r[n, t, u, v] = val
    return r[0]


class _Pair:
    """Everything about a primitive pair that the Coulomb integrals reuse."""
    __slots__ = ("p", "P", "coef", "ex", "ey", "ez", "lx", "ly", "lz")

    def __init__(self, a, lmn1, A, ca, b, lmn2, B, cb):
        self.p = a + b
        self.P = (a * A + b * B) / self.p
        self.coef = ca * cb
        self.lx, self.ly, self.lz = lmn1[0] + lmn2[0], lmn1[1] + lmn2[1], lmn1[2] + lmn2[2]
        self.ex = [hermite_e(lmn1[0], lmn2[0], t, A[0] - B[0], a, b) for t in range(self.lx + 1)]
        self.ey = [hermite_e(lmn1[1], lmn2[1], t, A[1] - B[1], a, b) for t in range(self.ly + 1)]
        self.ez = [hermite_e(lmn1[2], lmn2[2], t, A[2] - B[2], a, b) for t in range(self.lz + 1)]


def _pairs(f1: BasisFunction, f2: BasisFunction):
    return [_Pair(a, f1.lmn, f1.center, ca, b, f2.lmn, f2.center, cb)
            for ca, a in zip(f1.coefs, f1.exps) for cb, b in zip(f2.coefs, f2.exps)]


def one_electron(basis, charges, centers):
    """Overlap S, kinetic T and nuclear-attraction V matrices."""
    n = len(basis)
    s, t, v = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            fi, fj = basis[i], basis[j]
            sij = tij = vij = 0.0
            for ca, a in zip(fi.coefs, fi.exps):
                for cb, b in zip(fj.coefs, fj.exps):
                    sij += ca * cb * _overlap_primitive(a, fi.lmn, fi.center, b, fj.lmn, fj.center)
                    tij += ca * cb * _kinetic_primitive(a, fi.lmn, fi.center, b, fj.lmn, fj.center)
            for pr in _pairs(fi, fj):
                for z, c in zip(charges, centers):
                    r = hermite_r(pr.lx, pr.ly, pr.lz, pr.p, pr.P - np.asarray(c, dtype=float))
                    acc = 0.0
                    for tt in range(pr.lx + 1):
                        for uu in range(pr.ly + 1):
                            for vv in range(pr.lz + 1):
                                acc += pr.ex[tt] * pr.ey[uu] * pr.ez[vv] * r[tt, uu, vv]
                    vij += -z * pr.coef * 2.0 * math.pi / pr.p * acc
            s[i, j] = s[j, i] = sij
            t[i, j] = t[j, i] = tij
            v[i, j] = v[j, i] = vij
    return s, t, v


def _eri_pairs(pairs_ab, pairs_cd) -> float:
    total = 0.0
    for pa in pairs_ab:
        for pc in pairs_cd:
            p, q = pa.p, pc.p
            alpha = p * q / (p + q)
            r = hermite_r(pa.lx + pc.lx, pa.ly + pc.ly, pa.lz + pc.lz, alpha, pa.P - pc.P)
            acc = 0.0
            for t in range(pa.lx + 1):
                for u in range(pa.ly + 1):
                    for v in range(pa.lz + 1):
                        e_ab = pa.ex[t] * pa.ey[u] * pa.ez[v]
                        if e_ab == 0.0:
                            continue
                        for tau in range(pc.lx + 1):
                            for nu in range(pc.ly + 1):
                                for phi in range(pc.lz + 1):
                                    sign = -1.0 if (tau + nu + phi) & 1 else 1.0
                                    acc += e_ab * sign * pc.ex[tau] * pc.ey[nu] * pc.ez[phi] * r[t + tau, u + nu, v + phi]
            total += pa.coef * pc.coef * 2.0 * math.pi ** 2.5 / (p * q * math.sqrt(p + q)) * acc
    return total


def electron_repulsion(basis) -> np.ndarray:
    """(ab|cd) in chemists' notation, all n^4 elements (8-fold symmetry used to fill)."""
    n = len(basis)
    pairs = {(i, j): _pairs(basis[i], basis[j]) for i in range(n) for j in range(i + 1)}
    eri = np.zeros((n, n, n, n))
    for i in range(n):
        for j in range(i + 1):
            ij = i * (i + 1) // 2 + j
            for k in range(n):
                for l in range(k + 1):
                    if k * (k + 1) // 2 + l > ij:
                        continue
                    val = _eri_pairs(pairs[(i, j)], pairs[(k, l)])
                    for a, b in ((i, j), (j, i)):
                        for c, d in ((k, l), (l, k)):
                            eri[a, b, c, d] = val
                            eri[c, d, a, b] = val
    return eri


def nuclear_repulsion(charges, centers) -> float:
    e = 0.0
    for (za, ra), (zb, rb) in itertools.combinations(zip(charges, centers), 2):
        e += za * zb / float(np.linalg.norm(np.asarray(ra, dtype=float) - np.asarray(rb, dtype=float)))
    return e


# ---- the reference's inline STO-3G data (validation/check_rhf.f90:148-178) ----------------------
STO3G = {
    "H": [(0, [3.42525091, 0.62391373, 0.16885540], [0.15432897, 0.53532814, 0.44463454])],
    "O": [(0, [130.7093200, 23.8088610, 6.4436083], [0.15432897, 0.53532814, 0.44463454]),
          (0, [5.0331513, 1.1695961, 0.3803890], [-0.09996723, 0.39951283, 0.70011547]),
          (1, [5.0331513, 1.1695961, 0.3803890], [0.15591627, 0.60768372, 0.39195739])],
}
# STO-3G in the ten digits the Basis Set Exchange distributes (the reference's named "sto-3g").  The inline table
# above is an older tabulation of the same set with seven to eight significant digits (the two agree to 3e-7
# relative: tests/test_reference_golden_energies.py); it is what check_rhf.f90 itself uses, but the twelve-digit
# energies of the validation manifest were computed with the named set and need these digits (with the inline
# table the UHF energy below comes out 2.7e-8 off, with this one 4e-12).
STO3G_BSE = {
    "H": [(0, [3.425250914, 0.6239137298, 0.1688554040], [0.1543289673, 0.5353281423, 0.4446345422])],
    "O": [(0, [130.7093214, 23.80886605, 6.443608313], [0.1543289673, 0.5353281423, 0.4446345422]),
          (0, [5.033151319, 1.169596125, 0.3803889600], [-0.09996722919, 0.3995128261, 0.7001154689]),
          (1, [5.033151319, 1.169596125, 0.3803889600], [0.1559162750, 0.6076837186, 0.3919573931])],
}
CHARGE = {"H": 1, "C": 6, "O": 8}


def build_basis(symbols, coords, table=STO3G):
    basis = []
    for sym, xyz in zip(symbols, coords):
        for l, exps, coefs in table[sym]:
            for lmn in cartesian_components(l):
                basis.append(BasisFunction(xyz, lmn, exps, coefs))
    return basis


# geometries of validation/check_rhf.f90 (bohr): run_h2 :62-65, run_water :106-110
H2_STO3G = (["H", "H"], [[0.0, 0.0, 0.0], [0.0, 0.0, 1.4]], 2, -1.1167143251)
H2O_STO3G = (["O", "H", "H"], [[0.0, 0.0, -0.1364652], [0.0, 1.4304924, 1.0826636], [0.0, -1.4304924, 1.0826636]],
             10, -74.9658162796)


# ---- 6-31G* (Pople): published data, restated ---------------------------------------------------
# The reference reads its named basis sets from the Basis Set Exchange bundle (basis_sets/PROVENANCE.md),
# which is not in the tree.  6-31G* is published data: W. J. Hehre, R. Ditchfield, J. A. Pople, J. Chem.
# Phys. 56, 2257 (1972) (H, O valence/core), P. C. Hariharan, J. A. Pople, Theor. Chim. Acta 28, 213 (1973)
# (the d polarisation function, exponent 0.8 on O), in the ten digits the Basis Set Exchange distributes.
# The same numbers are held by the reference tree itself, in GAMESS form (coefficient x primitive norm, eight
# digits), in the PROJECTION BASIS SET block of tools/efp_validation/reference/water_6-31gs_cmo.efp:1711-1743:
# tests/test_reference_golden_energies.py checks this table against that block.
# "SP" shells are written out as an s and a p shell over the same exponents; d shells are CARTESIAN (6d),
# as the reference routes them (test/test_mqc_libcint_cartesian.f90: water is 19 functions in 6-31G*).
POPLE_631GS = {
    "H": [(0, [18.73113696, 2.825394365, 0.6401216923], [0.03349460434, 0.2347269535, 0.8137573261]),
          (0, [0.1612777588], [1.0])],
    "O": [(0, [5484.671660, 825.2349460, 188.0469580, 52.96450000, 16.89757040, 5.799635340],
              [0.001831074430, 0.01395017220, 0.06844507810, 0.2327143360, 0.4701928980, 0.3585208530]),
          (0, [15.53961625, 3.599933586, 1.013761750], [-0.1107775495, -0.1480262627, 1.130767015]),
          (1, [15.53961625, 3.599933586, 1.013761750], [0.07087426823, 0.3397528391, 0.7271585773]),
          (0, [0.2700058226], [1.0]),
          (1, [0.2700058226], [1.0]),
          (2, [0.8000000000], [1.0])],
}
# 6-31G**: the same plus a p shell (exponent 1.1) on H; carbon from the same papers.  No file of the reference
# tree holds the carbon numbers: what confirms them is that the CH4 energy below comes out to 2e-12.
POPLE_631GSS = {
    "H": POPLE_631GS["H"] + [(1, [1.1000000000], [1.0])],
    "O": POPLE_631GS["O"],
    "C": [(0, [3047.524880, 457.3695180, 103.9486850, 29.21015530, 9.286662960, 3.163926960],
              [0.001834737132, 0.01403732281, 0.06884262226, 0.2321844432, 0.4679413484, 0.3623119853]),
          (0, [7.868272350, 1.881288540, 0.5442492580], [-0.1193324198, -0.1608541517, 1.143456438]),
          (1, [7.868272350, 1.881288540, 0.5442492580], [0.06899906659, 0.3164239610, 0.7443082909]),
          (0, [0.1687144782], [1.0]),
          (1, [0.1687144782], [1.0]),
          (2, [0.8000000000], [1.0])],
}
BOHR_TO_ANGSTROM = 0.529177210903           # src/core/mqc_physical_constants.F90:35 (CODATA 2018, the default)
# validation/inputs/sample_inputs/w1.xyz (Angstrom), the geometry of validation_tests_cpu.json:899-904:
# "RHF H2O 6-31g* density fitted with 6-31g*", expected_energy -76.188111755038 (manifest tolerance 1e-9)
W1_ANGSTROM = [[0.00000000000000, 0.00000000009155, 0.10077199490609],
               [0.00000000000000, 0.77250895271063, -0.46780199741728],
               [0.00000000000000, -0.77250895280218, -0.46780199748881]]
H2O_631GS_DF = (["O", "H", "H"], [[x / BOHR_TO_ANGSTROM for x in row] for row in W1_ANGSTROM], 10, -76.188111755038)
# the same molecule and basis WITHOUT density fitting: validation_tests_cpu.json:719-723 "RHF H2O 6-31g* (CPU)"
H2O_631GS_EXACT_ENERGY = -76.010317945971
# validation/inputs/sample_inputs/ch4.xyz (Angstrom); validation_tests_cpu.json:905-910
# "RHF CH4 6-31g** density fitted with 6-31g**", expected_energy -40.381603512964
_CH4_A = 0.62757974260912
CH4_ANGSTROM = [[0.0, 0.0, 0.0], [_CH4_A, _CH4_A, _CH4_A], [_CH4_A, -_CH4_A, -_CH4_A], [-_CH4_A, _CH4_A, -_CH4_A],
                [-_CH4_A, -_CH4_A, _CH4_A]]
CH4_631GSS_DF = (["C", "H", "H", "H", "H"], [[x / BOHR_TO_ANGSTROM for x in row] for row in CH4_ANGSTROM], 10,
                 -40.381603512964)
# validation/inputs/sample_inputs/oh.xyz (Angstrom); validation_tests_cpu.json:1844-1849
# "UHF OH sto-3g multiplicity 2 (CPU)", expected_energy -74.362637545612: the reference-held two-spin energy
OH_STO3G_UHF = (["O", "H"], [[0.0, 0.0, 0.0], [0.0, 0.0, 0.9697 / BOHR_TO_ANGSTROM]], 9, 2, -74.362637545612)
DF_CASES = {"h2o_631gs": (H2O_631GS_DF, POPLE_631GS), "ch4_631gss": (CH4_631GSS_DF, POPLE_631GSS)}


def _single(f: BasisFunction):
    """An auxiliary function as a 'pair' with a unit s function of exponent 0 on the same centre."""
    return [_Pair(a, f.lmn, f.center, ca, 0.0, (0, 0, 0), f.center, 1.0) for ca, a in zip(f.coefs, f.exps)]


def three_centre(basis, aux) -> np.ndarray:
    """(mu nu|P), shape (n, n, naux): what `three_centre` of the reference hands to build_df_tensor
    (mqc_libcint_integrals.F90:913-990), flattened there as three(mu + n*nu, P)."""
    n, naux = len(basis), len(aux)
    singles = [_single(f) for f in aux]
    out = np.zeros((n, n, naux))
    for i in range(n):
        for j in range(i + 1):
            pr = _pairs(basis[i], basis[j])
            for k in range(naux):
                out[i, j, k] = out[j, i, k] = _eri_pairs(pr, singles[k])
    return out


def two_centre(aux) -> np.ndarray:
    """(P|Q), the Coulomb metric of the auxiliary basis."""
    naux = len(aux)
    singles = [_single(f) for f in aux]
    out = np.zeros((naux, naux))
    for i in range(naux):
        for j in range(i + 1):
            out[i, j] = out[j, i] = _eri_pairs(singles[i], singles[j])
    return out


def df_case_integrals(name: str):
    """(S, H, three(n*n, naux) column-major flattened, metric(naux, naux), E_nuc, n_electrons, E_reference) of one
    of the reference's density-fitted validation cases: the inputs of build_df_tensor + run_libcint_rhf(aux=...).
    The auxiliary basis is the orbital basis itself, as the validation inputs ask (aux_basis == basis)."""
    (symbols, coords, n_electrons, e_ref), table = DF_CASES[name]
    basis = build_basis(symbols, coords, table)
    aux = build_basis(symbols, coords, table)
    charges = [CHARGE[s] for s in symbols]
    s, t, v = one_electron(basis, charges, coords)
    n, naux = len(basis), len(aux)
    three = np.asfortranarray(three_centre(basis, aux).reshape(n * n, naux, order="F"))   # (mu + n*nu, P), integrals.F90:985
    return s, t + v, three, two_centre(aux), nuclear_repulsion(charges, coords), n_electrons, e_ref


def molecule_integrals(symbols, coords, table=STO3G):
    """(S, H = T + V, eri, E_nuc) of a molecule (default: the inline STO-3G basis of check_rhf.f90)."""
    basis = build_basis(symbols, coords, table)
    charges = [CHARGE[s] for s in symbols]
    s, t, v = one_electron(basis, charges, coords)
    return s, t + v, electron_repulsion(basis), nuclear_repulsion(charges, coords)
