#!/bin/bash
# Profiling recipe (run under gpurun on ONE GPU; see /opt/skills/guides/B200_PROFILING.md).  ncu serialises kernels, and the
# layer wavefront's kernels WAIT FOR EACH OTHER on the device, so every ncu run below disables the wavefront
# (LSTM_TS_WAVEFRONT=0) and profiles the same kernel instantiations one at a time (LSTM_TS_SEQ_VARIANT=2 = two batch tiles
# per CTA, the variant the wavefront launches).  The concurrent timeline of the product path comes from bench/trace_step.py.
#   bench/profile.sh launches   -> per-launch device time of one (sequential-layers) training step   (gpurun_out/launches.csv)
#   bench/profile.sh seq        -> ncu --set full of the persistent LSTM kernels (fwd + bwd)          (gpurun_out/prof_seq.ncu-rep)
#   bench/profile.sh gemm       -> ncu --set full of the tcgen05 GEMM (x-projection, dX, dW)          (gpurun_out/prof_gemm.ncu-rep)
#   bench/profile.sh small      -> ncu --set full of head fwd/bwd, flat Adam, column sums             (gpurun_out/prof_small.ncu-rep)
#   bench/profile.sh ar         -> ncu --set full of the fused allreduce kernels, world = 1           (gpurun_out/prof_ar.ncu-rep)
#   bench/profile.sh timeline   -> torch.profiler kernel timeline of one step with the wavefront on   (gpurun_out/trace_step_n1.txt)
#   bench/profile.sh sanitize   -> compute-sanitizer memcheck/racecheck/synccheck over the kernel tests
set -u
mkdir -p gpurun_out
what=${1:-launches}
export LSTM_TS_WAVEFRONT=0
BENCH="python bench.py --steps 2 --warmup 1 --no_e2e --no_baseline --cuda_graph 0"
NCU="ncu --set full --clock-control none --import-source on -f"
case "$what" in
  launches)
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches.out 2>&1 ;;
  seq)
    LSTM_TS_SEQ_VARIANT=2 $NCU -k regex:lstm_seq_kernel -s 1 -c 2 -o gpurun_out/prof_seq $BENCH > gpurun_out/prof_seq.out 2>&1 ;;
  gemm)
    $NCU -k regex:gemm2_kernel -s 5 -c 4 -o gpurun_out/prof_gemm $BENCH > gpurun_out/prof_gemm.out 2>&1 ;;
  small)
    $NCU -k "regex:head_fwd_tc|head_bwd|flat_adam|colsum_bf16" -s 4 -c 4 -o gpurun_out/prof_small $BENCH > gpurun_out/prof_small.out 2>&1 ;;
  ar)
    $NCU -k "regex:ar_two_shot|ar_one_shot" -s 2 -c 4 -o gpurun_out/prof_ar python bench/ar_single.py > gpurun_out/prof_ar.out 2>&1 ;;
  timeline)
    LSTM_TS_WAVEFRONT=1 python bench/trace_step.py > gpurun_out/trace_step.out 2>&1 ;;
  sanitize)
    for tool in memcheck racecheck synccheck; do
      compute-sanitizer --tool $tool python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "pointwise or head_forward or adam or (gemm2 and 1-128) or (persistent and 3-128-64-64) or generic or (folded and 128-3-128)" > gpurun_out/sanitizer_$tool.log 2>&1
      tail -3 gpurun_out/sanitizer_$tool.log
    done ;;
esac
