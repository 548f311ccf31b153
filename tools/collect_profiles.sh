#!/usr/bin/env bash
# Turn the scratch results of tools/gpu_check.sh (gpurun_out/) into the tracked summaries under profiles/.
# usage: tools/collect_profiles.sh <tag>      (run here, on the CPU box; needs ncu to read the .ncu-rep files)
set -euo pipefail
TAG="$1"; SRC=gpurun_out; DST=profiles/$TAG
mkdir -p "$DST"
for f in pytest_gpu smoke bench bench_ref nvidia_smi sanitizer_memcheck sanitizer_racecheck; do
  for e in log txt; do [ -f "$SRC/${f}_$TAG.$e" ] && cp "$SRC/${f}_$TAG.$e" "$DST/$f.$e"; done
done
[ -f "$SRC/launches_$TAG.csv" ] && { cp "$SRC/launches_$TAG.csv" "$DST/ncu_launches.csv"; python tools/ncu_launch_shares.py "$SRC/launches_$TAG.csv" > "$DST/ncu_launch_shares.txt"; }
METRICS='gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|sm__throughput.avg.pct_of_peak_sustained_elapsed|l1tex__t_sector_hit_rate.pct|lts__t_sector_hit_rate.pct|smsp__inst_executed.sum|launch__occupancy_limit_registers|launch__occupancy_limit_shared_mem|launch__grid_size|launch__block_size|smsp__issue_active.avg.pct_of_peak_sustained_active|sm__cycles_active.avg|launch__shared_mem_per_block_static|l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum|l1tex__t_requests_pipe_lsu_mem_global_op_red.sum|l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum|lts__t_sectors_op_atom.sum|lts__t_sectors_op_red.sum|lts__t_sectors_srcunit_tex_op_atom.sum|l1tex__t_bytes.sum|lts__t_bytes.sum'
for k in integrate mc mcfilter alloc; do
  R="$SRC/prof_${k}_$TAG.ncu-rep"
  [ -f "$R" ] || continue
  ncu -i "$R" --page raw --csv 2>/dev/null | python tools/ncu_raw_pick.py "$METRICS" > "$DST/ncu_full_${k}.txt"
  ncu -i "$R" --page source --csv 2>/dev/null | python tools/ncu_source_summary.py 30 > "$DST/ncu_source_${k}.txt" || true
done
# DRAM traffic of the roofline kernel per launch (bench.py reports it as roofline.traffic)
if [ -f "$DST/ncu_full_integrate.txt" ]; then
  python - "$DST/ncu_full_integrate.txt" "$TAG" > profiles/integrate_traffic.json <<'PY'
import json, re, sys
txt = open(sys.argv[1]).read()
def first(name):
    m = re.search(name + r"\s+([0-9.]+)\s+(\w+)", txt)
    v, u = float(m.group(1)), m.group(2)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
print(json.dumps({"kernel": "vh::integrate_kernel", "dram_bytes_per_launch": first("dram__bytes_read.sum") + first("dram__bytes_write.sum"),
                  "source": f"profiles/{sys.argv[2]}/ncu_full_integrate.txt (ncu --set full, one launch of the bench workload, frame ~130)"}))
PY
fi
ls -la "$DST"
