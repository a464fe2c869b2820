#!/bin/bash
# Full ncu capture of the top kernel (the tcgen05 GEMM) inside one steady-state AlexNet step + raw-metric summary.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16_tcgen05 -c 12 \
  -o gpurun_out/prof_gemm -f python scripts/profile_step.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/prof_gemm_raw.csv")))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_pipe_uniform.sum"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
out = open("gpurun_out/prof_gemm_summary.txt", "w")
line = " | ".join(w for w, _ in idx); print(line); out.write(line + "\n")
for r in rows[2:]:
    line = " | ".join(r[i][:48] for _, i in idx); print(line); out.write(line + "\n")
PY
