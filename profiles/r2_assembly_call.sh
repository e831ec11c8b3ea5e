#!/bin/bash
# Round 2, assembly: parity suite against both assembly modes, phase timers at 16 M DOF, ncu launch list and
# --set full capture of the gather and elimination kernels.
#   gpurun --timeout 1200 -- 'bash profiles/r2_assembly_call.sh [quick]'
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_suite_default.log 2>&1
echo "suite(default = gather) rc=$?"; tail -n 3 gpurun_out/r2_suite_default.log
MAGNETITE_B200_TEST_ASSEMBLY=1 timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_suite_assembly1.log 2>&1
echo "suite(assembly=1, sorted keys) rc=$?"; tail -n 3 gpurun_out/r2_suite_assembly1.log
if timeout 200 python profiles/assembly_probe.py 4000 2000 3 > gpurun_out/r2_assembly_probe.log 2>&1; then
    cat gpurun_out/r2_assembly_probe.log
    if [ "${1:-}" != "quick" ]; then
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file gpurun_out/r2_launches_assembly.csv python profiles/assembly_probe.py 4000 2000 1 0 \
        > gpurun_out/r2_ncu_launches.log 2>&1
    echo "ncu launch list rc=$?"
    # the probe assembles a warm-up plate first: skip its launches (one of each kernel)
    timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:gather_fill|gather_count|eliminate_' -s 4 -c 4 \
        -o gpurun_out/r2_gather python profiles/assembly_probe.py 4000 2000 1 0 > gpurun_out/r2_ncu_full.log 2>&1
    echo "ncu full capture rc=$?"
    fi
else
    echo "assembly probe failed"; tail -n 20 gpurun_out/r2_assembly_probe.log
fi
