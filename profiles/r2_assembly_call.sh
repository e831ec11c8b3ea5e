#!/bin/bash
# Round 2, fused gather assembly (mag_options.assembly = 0, the default): parity suite against every assembly mode,
# phase timers of the three modes at 16 M DOF, ncu launch list and --set full capture of the fused kernels.
#   gpurun --timeout 1200 -- 'bash profiles/r2_assembly_call.sh [quick]'
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_suite_fused_default.log 2>&1
echo "suite(default = fused gather, one pass) rc=$?"; tail -n 3 gpurun_out/r2_suite_fused_default.log
MAG_TUNE=64 timeout 400 python -m pytest tests/test_gpu_gather.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_suite_fused_twopass.log 2>&1
echo "suite(fused gather, two passes) rc=$?"; tail -n 3 gpurun_out/r2_suite_fused_twopass.log
if [ "${1:-}" != "quick" ]; then
  for mode in 1 2; do
    MAGNETITE_B200_TEST_ASSEMBLY=$mode timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_suite_assembly$mode.log 2>&1
    echo "suite(assembly=$mode) rc=$?"; tail -n 3 gpurun_out/r2_suite_assembly$mode.log
  done
fi
MAG_TUNE=64 timeout 200 python profiles/assembly_probe.py 4000 2000 3 0 > gpurun_out/r2_assembly_probe_twopass.log 2>&1; echo "two-pass probe rc=$?"; cat gpurun_out/r2_assembly_probe_twopass.log
if timeout 200 python profiles/assembly_probe.py 4000 2000 3 > gpurun_out/r2_assembly_probe_fused.log 2>&1; then
    cat gpurun_out/r2_assembly_probe_fused.log
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file gpurun_out/r2_launches_assembly_fused.csv python profiles/assembly_probe.py 4000 2000 1 0 \
        > gpurun_out/r2_ncu_launches_fused.log 2>&1
    echo "ncu launch list rc=$?"
    # the probe assembles a warm-up plate first: skip its fused launch
    timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:fused_rows' -s 1 -c 1 \
        -o gpurun_out/r2_fused python profiles/assembly_probe.py 4000 2000 1 0 > gpurun_out/r2_ncu_full_fused.log 2>&1
    echo "ncu full capture rc=$?"
else
    echo "assembly probe failed"; tail -n 20 gpurun_out/r2_assembly_probe_fused.log
fi
