#!/usr/bin/env python
"""Roofline report of one training step (runs HERE, no GPU): per-launch device times from an ncu launch list
(`bench/profile.sh launches` -> gpurun_out/launches.csv) against the driver-measured peaks in MEASURED_PEAKS.json.

    python bench/roofline.py profiles/logs/launches_r1_final.csv > profiles/roofline.md

FLOPs / bytes are analytic for BASELINE.json config 3 (2-layer-1024, T=128, B=256, bf16); every number is per launch.
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T, B, H, D = 128, 256, 1024, 1024
M = T * B
GEMM_FLOP = 2.0 * M * 4 * H * D
MB = 1e6


def model():
    """kernel-name regex -> (label, flop per launch, HBM bytes per launch, what bounds it)"""
    act, cseq, hseq, img = M * 4 * H * 2, M * H * 4, M * H * 2, M * H * 2
    return [
        (r"lstm_seq_kernel<\(bool\)0|lstm_seq_kernel<0", "LSTM forward recurrence (persistent, 128 steps)", GEMM_FLOP,
         act + act + cseq + hseq + img, "latency chain per step (see profiles/README.md 2b)"),
        (r"lstm_seq_kernel<\(bool\)1|lstm_seq_kernel<1", "LSTM backward recurrence (persistent, 129 steps)", GEMM_FLOP,
         act + 2 * cseq + hseq + act + act, "latency chain per step"),
        (r"gemm_bf16_tn_kernel", "tcgen05 GEMM (x-projection / input gradient)", GEMM_FLOP, M * D * 2 + 4 * H * D * 2 + act, "tensor pipe"),
        (r"nvjet", "cuBLAS weight-gradient GEMM", GEMM_FLOP, act + M * D * 2 + 4 * H * D * 4, "tensor pipe"),
        (r"colsum_bf16", "bias-gradient column sums", 0.0, act, "HBM"),
        (r"flat_adam", "flat Adam + bf16 shadow", 0.0, None, "HBM"),
        (r"head_xent", "fused head + softmax-xent + accuracy", 0.0, None, "latency (tiny)"),
        (r"transpose01_rows", "batch-major -> time-major feed", 0.0, 2 * M * D * 2, "HBM"),
        (r"transpose2d_b16", "weight transposes", 0.0, 2 * 4 * H * H * 2, "HBM/L2"),
    ]


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "launches.csv")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm, tf_burst, tf_sust = peaks["hbm_gbs"], peaks["bf16_tflops"], peaks["bf16_tflops_sustained"]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    names = [r[ki] for r in rows]
    vals = [float(r[vi]) for r in rows]
    adam = [i for i, n in enumerate(names) if "flat_adam" in n]
    lo, hi = (adam[-2] + 1, adam[-1] + 1) if len(adam) >= 2 else (0, len(rows))
    n_params = 2 * (4 * H * D + 4 * H * H + 4 * H) + H * 10 + 10
    mdl = model()
    agg = collections.OrderedDict()
    other = [0, 0.0]
    for n, v in zip(names[lo:hi], vals[lo:hi]):
        for rx, label, flop, byts, bound in mdl:
            if re.search(rx, n):
                a = agg.setdefault(label, [0, 0.0, flop, byts, bound])
                a[0] += 1
                a[1] += v
                break
        else:
            other[0] += 1
            other[1] += v
    total = sum(a[1] for a in agg.values()) + other[1]
    print("# Roofline of one training step (config 3, one B200)\n")
    print(f"Source: `{os.path.relpath(path, ROOT)}` (ncu `gpu__time_duration`, serialized launches of the last profiled step); peaks from "
          f"`MEASURED_PEAKS.json`: HBM copy {hbm:.0f} GB/s, cuBLAS bf16 {tf_burst:.0f} TFLOP/s burst / {tf_sust:.0f} sustained.\n")
    print("| kernel | launches | us per launch | share of step | TFLOP/s (of sustained peak) | HBM GB/s (of measured) | bound by |")
    print("|---|---|---|---|---|---|---|")
    for label, (c, v, flop, byts, bound) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        us = v / c / 1e3
        if byts is None and "Adam" in label:
            byts = n_params * (4 * 4 + 3 * 4 + 2)
        tf = f"{flop / (us * 1e-6) / 1e12:.0f} ({flop / (us * 1e-6) / 1e12 / tf_sust:.2f})" if flop else "-"
        gb = f"{byts / (us * 1e-6) / 1e9:.0f} ({byts / (us * 1e-6) / 1e9 / hbm:.2f})" if byts else "-"
        print(f"| {label} | {c} | {us:.1f} | {100 * v / total:.1f} % | {tf} | {gb} | {bound} |")
    print(f"| other (fills, small copies, loss bookkeeping) | {other[0]} | {other[1] / max(1, other[0]) / 1e3:.1f} | {100 * other[1] / total:.1f} % | - | - | launch granularity |")
    print(f"\nSum of launches: {total / 1e6:.3f} ms ({hi - lo} launches).  FLOP floor of the step (all GEMM-shaped work at the sustained cuBLAS peak): "
          f"{(4 + 4 + 3 + 1) * GEMM_FLOP / (tf_sust * 1e12) * 1e3:.2f} ms.")


if __name__ == "__main__":
    main()
