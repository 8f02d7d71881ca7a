"""bench.py's reference arm runs on CPU and keeps the driver's JSON contract: ONE line on
stdout with the agreed keys (the b200 arm needs a GPU and is exercised by the driver)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "builds/s" and d["higher_is_better"] is True
    assert d["metric"] == "DF-J/K Fock builds/sec" and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "builds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_pin_block_is_serialisable_and_lands_on_the_reference_energy(tmp_path):
    """bench.run_reference_pin with a stand-in engine backed by the oracle (the real one needs a GPU; the same
    calls run on a B200 in tests/test_gpu_reference_df_energies.py): the block must be plain JSON and must
    reproduce the reference-held density-fitted energy.  Run in a child process: importing bench.py claims
    file descriptor 1 for its one JSON line, which must not happen to the test runner."""
    out_file = tmp_path / "pin.json"
    code = f"""
import importlib.util, json, os, sys
import numpy as np
sys.path.insert(0, {ROOT!r})
from oracle import df_fock_oracle as oracle, scf_oracle as scf
spec = importlib.util.spec_from_file_location("bench_module", os.path.join({ROOT!r}, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)

class StandIn:
    def set_tensor(self, b):
        self.b = np.asarray(b)
    def build_fock_df(self, h, density, coeff, n_occ):
        assert h.flags.f_contiguous and density.flags.f_contiguous and coeff.flags.f_contiguous
        self.f = oracle.build_fock_df(h, self.b, density, coeff, n_occ)
        self.e = oracle.electronic_energy(h, self.f, density)
        return self.f
    def last_energy(self):
        return self.e
    def build_fock_df_uhf(self, h, d_a, d_b, c_a, n_a, c_b, n_b):
        return oracle.build_fock_df_uhf(h, self.b, d_a, d_b, c_a, n_a, c_b, n_b)
    def run_scf(self, h, s, n_electrons, e_nuc=0.0):
        def builder(h_, density, coeff, n_occ):
            f = oracle.build_fock_df(h_, self.b, density, coeff, n_occ)
            return f, oracle.electronic_energy(h_, f, density)
        return scf.run_rhf(h, s, n_electrons, builder, e_nuc=e_nuc)
    def close(self):
        pass

open({str(out_file)!r}, "w").write(json.dumps(bench.run_reference_pin(StandIn)))
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    block = json.loads(out_file.read_text())
    assert block["ok"] and block["abs_err"] <= 1e-9 and block["reference_held_energy"] == -76.188111755038
    assert block["two_spin"]["ok"] and block["two_spin"]["reference_held_energy"] == -74.362637545612
