#!/bin/bash
# Per-kernel device times of the config-3 / config-4 / small slot chains (tools/tb_cost_probe.py under the ncu launch-list
# pass: serialised and cold, so shares rather than absolutes). Output: gpurun_out/slot_kernels.txt
mkdir -p gpurun_out
python tools/tb_cost_probe.py > gpurun_out/slot_probe.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/slot_launches.csv \
  python tools/tb_cost_probe.py > gpurun_out/slot_probe_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = []
with open("gpurun_out/slot_launches.csv") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(unit, v)
        rows.append((r["Kernel Name"][:60], r["Grid Size"], v))
# the first chain of each workload
with open("gpurun_out/slot_kernels.txt", "w") as out:
    for name, grid, us in rows[:80]:
        out.write("%-62s %-18s %8.1f us\n" % (name, grid, us))
PY
cat gpurun_out/slot_probe.log
