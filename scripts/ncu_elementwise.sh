#!/bin/bash
# Full ncu capture of the streaming (non-GEMM) kernels of one AlexNet step: who is bound by what.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'maxpool|space_to_depth|crop_mirror|lrn_|relu_bias|sgd_flat' -c 24 \
  -o gpurun_out/prof_elt -f python scripts/profile_step.py > gpurun_out/ncu_elt.log 2>&1
tail -2 gpurun_out/ncu_elt.log
ncu -i gpurun_out/prof_elt.ncu-rep --page raw --csv > gpurun_out/prof_elt_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/prof_elt_raw.csv")))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__inst_executed.sum"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
out = open("gpurun_out/prof_elt_summary.txt", "w")
for r in rows[2:]:
    line = "\n".join("  %-80s %s" % (w, r[i][:60]) for w, i in idx); print(line); print(); out.write(line + "\n\n")
PY
