#!/bin/bash
# Single-GPU measurement pass (development tool; results -> gpurun_out/).
for wl in c2 c5 c4 c3; do
  python bench.py --workload $wl > gpurun_out/s_$wl.json 2> gpurun_out/s_$wl.err || tail -5 gpurun_out/s_$wl.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s_ref_c2.json 2>/dev/null
python - <<'PY'
import json
for name in ["c2", "c5", "c4", "c3", "ref_c2"]:
    try:
        d = json.load(open(f"gpurun_out/s_{name}.json"))
    except Exception as e:
        print(name, "FAILED", e); continue
    print(name, round(d["value"], 3), d["unit"], "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 3),
          "cpu", (d.get("cpu_baseline") or {}).get("value"), "parity", (d.get("parity") or {}).get("max_abs_err_vs_oracle"))
    s = d.get("summary")
    if s:
        print("   K", round(s["K_tflops"], 2), round(s["K_frac_of_fp64_peak"], 3), "J", round(s["J_gbs"], 1),
              round(s["J_frac_of_hbm_peak"], 3), "passes", s["J_passes_over_B"],
              {k: round(v, 4) for k, v in s["phase_ms_per_build"].items()})
PY
