import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from metalquicha_b200 import B200FockEngine
from test_gpu_device_scf import _synthetic_fragment
eng = B200FockEngine(0)
s, h, b = _synthetic_fragment(972, 72, 15, 340)
eng.set_tensor(b)
r = eng.run_scf_fragment(h, s, 30)
print(r["iterations"], r["electronic"])
eng.close()
