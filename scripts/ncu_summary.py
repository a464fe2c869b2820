#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i … --page raw --csv`) into a small text table for profiles/.
    python scripts/ncu_summary.py gpurun_out/c4_prof_gemm_tf32.ncu-rep profiles/ncu/gemm_tf32_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "smsp__inst_executed.sum", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(out, "w") as f:
        f.write("# %s  (ncu --set full --clock-control none --import-source on; one row per captured launch)\n" % rep.split("/")[-1])
        f.write(" | ".join("%s [%s]" % (w, units[i]) if units[i] else w for w, i in idx) + "\n")
        for r in rows[2:]:
            f.write(" | ".join(r[i][:70] for _, i in idx) + "\n")
    print(open(out).read()[:3000])


if __name__ == "__main__":
    main()
