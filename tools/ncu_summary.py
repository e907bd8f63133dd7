"""Condense `ncu --set full` captures (gpurun_out/*.ncu-rep) into one CSV of the judged metrics per launch.

    python tools/ncu_summary.py profiles/r01c_ncu_full_summary.csv gpurun_out/r01c_prof_step.ncu-rep gpurun_out/r01c_prof_attn.ncu-rep
"""
import csv
import io
import os
import subprocess
import sys

COLS = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
        "smsp__inst_executed.sum"]


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    rows = []
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rd = list(csv.reader(io.StringIO(txt)))
        hdr = rd[0]
        units = rd[1]
        idx = {c: hdr.index(c) for c in COLS if c in hdr}
        for r in rd[2:]:
            if len(r) != len(hdr):
                continue
            rec = {"capture": os.path.splitext(os.path.basename(rep))[0]}
            for c, i in idx.items():
                u = units[i]
                rec[c] = r[i] + (" " + u if u and c not in ("ID", "Kernel Name", "Grid Size", "Block Size") else "")
            rows.append(rec)
    with open(out_path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=["capture"] + [c for c in COLS if any(c in r for r in rows)])
        w.writeheader()
        for r in rows:
            w.writerow(r)
    print(f"{len(rows)} launches -> {out_path}")


if __name__ == "__main__":
    main()
