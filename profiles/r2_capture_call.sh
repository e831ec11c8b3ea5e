#!/bin/bash
# Round 2: plain bench run, then the ncu launch list and one --set full capture of every shipped hot kernel
# (assembly, coarse setup, PCG iteration) on the 16 M-DOF plate.   gpurun --timeout 1500 -- 'bash profiles/r2_capture_call.sh'
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_1gpu.err
if timeout 100 python profiles/two_level_kernels.py > gpurun_out/r2_two_level_kernels.log 2>&1; then
  cat gpurun_out/r2_two_level_kernels.log
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/r2_launches_step_4000x2000_12it.csv python profiles/two_level_kernels.py > /dev/null 2>&1
  echo "launch list rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k 'regex:rs_hist_kernel<unsigned int>|rs_scatter_kernel<unsigned int>|gather_|eliminate_|coarse_galerkin|band_cholesky|band_inverse|coarse_restrict|coarse_apply|coarse_gather|pcg_' \
      -c 40 -o gpurun_out/r2_step python profiles/two_level_kernels.py > gpurun_out/r2_ncu_full_step.log 2>&1
  echo "full capture rc=$?"
fi
