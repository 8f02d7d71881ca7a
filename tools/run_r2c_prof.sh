#!/bin/bash
# Profile pass after the half-transform / SYRK schedule changes (single GPU): each program is first run plainly (exit 0), then under ncu.
mkdir -p gpurun_out
B="python bench.py --workload c2 --secondary none --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-setup-timings"
$B > gpurun_out/r02c_prof_plain_c2.json 2> gpurun_out/r02c_prof_plain_c2.err; echo "plain c2 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_c2.csv $B > gpurun_out/r02c_ncu_l.log 2>&1; echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_half_transform|k_accumulate" -c 4 \
  -o gpurun_out/r02c_prof_c2 -f $B > gpurun_out/r02c_ncu_f.log 2>&1; echo "ncu full c2 rc=$?"
B4="python bench.py --workload c4 --secondary none --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-setup-timings"
$B4 > gpurun_out/r02c_prof_plain_c4.json 2> gpurun_out/r02c_prof_plain_c4.err; echo "plain c4 rc=$?"
timeout 1500 ncu --set full --clock-control none -k regex:"k_half_transform|k_accumulate" -c 10 \
  -o gpurun_out/r02c_prof_c4 -f $B4 > gpurun_out/r02c_ncu_c4.log 2>&1; echo "ncu full c4 rc=$?"
for r in c2 c4; do
  python tools/ncu_summary.py gpurun_out/r02c_prof_$r.ncu-rep > gpurun_out/r02c_ncu_summary_$r.md 2> gpurun_out/r02c_ncu_summary_$r.err
done
python - <<'PY'
import csv, json, subprocess
out = {}
for tag in ("c2", "c4"):
    raw = subprocess.run(["ncu", "-i", f"gpurun_out/r02c_prof_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    units = rows[1]
    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    agg = {}
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").split("<")[0]
        agg.setdefault(name, []).append(to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi]))
    out[tag] = {k: sum(v) / len(v) for k, v in agg.items()}
    out[tag + "_launches_captured"] = {k: len(v) for k, v in agg.items()}
json.dump(out, open("gpurun_out/r02c_ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
PY
rm -f gpurun_out/r02c_prof_c2.ncu-rep gpurun_out/r02c_prof_c4.ncu-rep
ls -la gpurun_out | head -40
