#!/usr/bin/env python
"""Turn an .ncu-rep (read HERE, no GPU needed) into the short markdown summary committed under profiles/.

    python bench/ncu_summary.py gpurun_out/prof_seq.ncu-rep > profiles/lstm_seq_ncu.md
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print("no data in", rep)
        return
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of `{rep}`\n")
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        print(f"## {rec.get('Kernel Name', '?')}  (id {rec.get('ID', '?')})\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in rec:
                print(f"| {k} | {rec[k]} | {units[hdr.index(k)]} |")
        print()


if __name__ == "__main__":
    main()
