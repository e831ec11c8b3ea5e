#!/bin/bash
# Round 2 evidence, in two GPU calls (what a call writes under gpurun_out/ must stay below 64 MiB):
#   gpurun --timeout 1200 -- 'bash profiles/r2_capture_call.sh lines'     bench line, probes, ncu launch list
#   gpurun --timeout 1800 -- 'bash profiles/r2_capture_call.sh captures'  ncu --set full of the hot kernels
set -u
mkdir -p gpurun_out
if [ "${1:-lines}" = "lines" ]; then
  timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
  echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_1gpu.err
  timeout 200 python profiles/assembly_probe.py 4000 2000 3 > gpurun_out/r2_assembly_probe.log 2>&1; echo "assembly probe rc=$?"
  MAG_TUNE=32 timeout 100 python profiles/two_level_kernels.py > gpurun_out/r2_two_level_kernels.log 2>&1; cat gpurun_out/r2_two_level_kernels.log
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/r2_launches_step_4000x2000_12it.csv python profiles/two_level_kernels.py > /dev/null 2>&1
  echo "launch list rc=$?"
  timeout 100 python profiles/example_probe.py > gpurun_out/r2_example_probe.txt 2>&1; echo "example probe rc=$?"
else
  # one launch of each kernel of a step: assembly, coarse setup, initial residual, first iteration
  timeout 900 ncu --set full --clock-control none \
      -k 'regex:gather_|eliminate_|coarse_galerkin|band_cholesky|band_inverse|coarse_restrict|coarse_apply|coarse_gather|pcg_' \
      -c 19 -o gpurun_out/r2_step python profiles/two_level_kernels.py > gpurun_out/r2_ncu_full_step.log 2>&1
  echo "step capture rc=$?"
  timeout 600 ncu --set full --clock-control none -k 'regex:rs_hist|rs_scatter|emit_incidence' -c 7 \
      -o gpurun_out/r2_sort python profiles/two_level_kernels.py > gpurun_out/r2_ncu_full_sort.log 2>&1
  echo "sort capture rc=$?"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:small_cg -c 1 -o gpurun_out/r2_small \
      python profiles/example_probe.py > gpurun_out/r2_small_ncu.log 2>&1
  echo "small capture rc=$?"
  ls -la gpurun_out/*.ncu-rep
fi
