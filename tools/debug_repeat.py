import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metalquicha_b200 import B200FockEngine, synth
n, o, q = (int(x) for x in sys.argv[1:4]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
b, h, d, c = synth.synth_problem(77, n, o, q)
eng = B200FockEngine(0); eng.set_tensor(b)
ks = [eng.build_jk(d, c, o, want_j=False)[1] for _ in range(reps)]
from collections import Counter
# majority value per element = "truth"
ref = np.median(np.stack(ks), axis=0)
nbadruns = 0; tiles = Counter()
for k in ks:
    bad = np.argwhere(np.abs(k - ref) > 1e-9)
    if len(bad):
        nbadruns += 1
        for a, b_ in bad:
            if a >= b_: tiles[(a // 64, b_ // 64)] += 1
print(f"splits={os.environ.get('MQCB200_KSPLITS')} shape={n},{o},{q}: bad runs {nbadruns}/{reps}; bad lower-tri elements per 64-tile: {dict(tiles)}")
