"""Where the host-buffer build spends its time beyond the device-resident one (development tool).
usage: python tools/e2e_probe.py n n_occ naux [steps]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metalquicha_b200 import B200FockEngine, synth
n, o, q = (int(x) for x in sys.argv[1:4]); steps = int(sys.argv[4]) if len(sys.argv) > 4 else 30
_, h, d, c = synth.synth_problem(1, n, o, q, with_tensor=False)
eng = B200FockEngine(0); eng.synth_tensor(n, q, 1, synth.default_scale(n, q))
def pinned(a):
    t = torch.empty(a.shape[::-1], dtype=torch.float64).pin_memory()
    v = t.numpy().T
    v[...] = a
    return v
hp, dp, cp = pinned(h), pinned(d), pinned(c)
fp = pinned(np.zeros((n, n)))
dh, dd, dc = (torch.from_numpy(np.ascontiguousarray(a.T)).cuda() for a in (h, d, c)); df = torch.empty_like(dh)
out = {"shape": [n, o, q]}
for _ in range(3): eng.build_fock_device(dh, dd, dc, o, df)
t0 = time.perf_counter()
for _ in range(steps): eng.build_fock_device(dh, dd, dc, o, df, sync=False)
torch.cuda.synchronize(); out["device_async_ms"] = 1e3 * (time.perf_counter() - t0) / steps
t0 = time.perf_counter()
for _ in range(steps): eng.build_fock_device(dh, dd, dc, o, df, sync=True)
out["device_sync_each_ms"] = 1e3 * (time.perf_counter() - t0) / steps
for _ in range(3): eng.build_fock_df(hp, dp, cp, o, out=fp)
t0 = time.perf_counter()
for _ in range(steps): eng.build_fock_df(hp, dp, cp, o, out=fp)
out["host_pinned_ms"] = 1e3 * (time.perf_counter() - t0) / steps
eng.set_profiling(True); eng.last_timings()
for _ in range(steps): eng.build_fock_df(hp, dp, cp, o, out=fp)
out["host_pinned_phases_ms"] = {k: round(v / steps, 4) for k, v in eng.last_timings().items()}
eng.set_profiling(False)
# raw copies of one matrix
tp = torch.empty(n * n, dtype=torch.float64).pin_memory(); tg = torch.empty(n * n, dtype=torch.float64, device="cuda")
for name, fn in (("d2h", lambda: tp.copy_(tg, non_blocking=True)), ("h2d", lambda: tg.copy_(tp, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): fn(); torch.cuda.synchronize()
    out[f"raw_{name}_ms_{8*n*n/1e6:.1f}MB"] = 1e3 * (time.perf_counter() - t0) / 20
print(json.dumps(out))
