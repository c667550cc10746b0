#!/bin/bash
# Profiling recipe (run under gpurun on ONE GPU; see /opt/skills/guides/B200_PROFILING.md).
#   bench/profile.sh launches   -> per-launch device time of one training step        (gpurun_out/launches.csv)
#   bench/profile.sh seq        -> ncu --set full of the persistent LSTM kernels       (gpurun_out/prof_seq.ncu-rep)
#   bench/profile.sh gemm       -> ncu --set full of the tcgen05 GEMM                  (gpurun_out/prof_gemm.ncu-rep)
#   bench/profile.sh sanitize   -> compute-sanitizer memcheck/racecheck/synccheck over the kernel tests
set -u
mkdir -p gpurun_out
what=${1:-launches}
BENCH="python bench.py --steps 2 --warmup 1 --no_e2e"
case "$what" in
  launches)
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches.out 2>&1 ;;
  seq)
    ncu --set full --clock-control none --import-source on -k regex:lstm_seq_kernel -s 4 -c 4 -f -o gpurun_out/prof_seq $BENCH > gpurun_out/prof_seq.out 2>&1 ;;
  gemm)
    ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_kernel -s 2 -c 1 -f -o gpurun_out/prof_gemm $BENCH > gpurun_out/prof_gemm.out 2>&1 ;;
  sanitize)
    for tool in memcheck racecheck synccheck; do
      compute-sanitizer --tool $tool python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "pointwise or head_xent or adam or (tcgen05_gemm and 128-128-64) or (persistent and 3-128-64-64) or generic_shape" > gpurun_out/sanitizer_$tool.log 2>&1
      tail -3 gpurun_out/sanitizer_$tool.log
    done ;;
esac
