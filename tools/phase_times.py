"""Per-phase device times of the build at a given shape (development tool).
usage: python tools/phase_times.py n n_occ naux [steps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metalquicha_b200 import B200FockEngine, synth
n, o, q = (int(x) for x in sys.argv[1:4]); steps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
_, h, d, c = synth.synth_problem(1, n, o, q, with_tensor=False)
eng = B200FockEngine(0); eng.synth_tensor(n, q, 1, synth.default_scale(n, q))
dh, dd, dc = (torch.from_numpy(np.ascontiguousarray(a.T)).cuda() for a in (h, d, c)); df = torch.empty_like(dh)
for _ in range(3): eng.build_fock_device(dh, dd, dc, o, df)
eng.set_profiling(True); eng.last_timings()
st = torch.cuda.ExternalStream(eng.stream()); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(steps): eng.build_fock_device(dh, dd, dc, o, df, sync=False)
e1.record(st); e1.synchronize()
ph = {k: round(v / steps, 4) for k, v in eng.last_timings().items()}
ms = e0.elapsed_time(e1) / steps
npair = n * (n + 1) // 2
import hashlib
fp = hashlib.sha256(df.cpu().numpy().tobytes()).hexdigest()[:16]
print(json.dumps({"shape": [n, o, q], "tail_split": "off" if os.environ.get("MQCB200_NO_TAIL_SPLIT") else "on", "fock_sha": fp, "ktile": os.environ.get("MQCB200_KTILE", "default"), "ms_per_build": round(ms, 4), "phases_ms": ph,
                  "k1_tflops": round(2.0 * n * n * o * q / ph["k_half_transform"] * 1e-9, 2) if ph["k_half_transform"] else None,
                  "k2_tflops_alg": round(1.0 * n * n * o * q / ph["k_accumulate"] * 1e-9, 2) if ph["k_accumulate"] else None,
                  "j1_gbs": round(8.0 * npair * q / ph["j_gamma"] * 1e-6, 1) if ph["j_gamma"] else None,
                  "j2_gbs": round(8.0 * npair * q / ph["j_accumulate"] * 1e-6, 1) if ph["j_accumulate"] else None}))
