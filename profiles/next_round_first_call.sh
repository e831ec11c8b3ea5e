#!/bin/bash
# What the first GPU call of the next round should run (this round's GPU budget ended before it could):
#   1. the WHOLE parity suite with the gather assembly as the default of the Python binding — the evidence needed
#      to flip mag_options_default().assembly to 1;
#   2. the assembly probe (sorted keys vs gather, 16 M-DOF plate) as a plain run;
#   3. the ncu launch list of the same probe and one --set full capture of gather_fill_kernel / gather_count_kernel
#      (recipe: /opt/skills/guides/B200_PROFILING.md), each only after the plain run exited 0.
# Usage:  gpurun --timeout 900 -- 'bash profiles/next_round_first_call.sh'     (one GPU; ~3 minutes)
# Read the captures back here with:
#   ncu -i gpurun_out/r2_gather.ncu-rep --page raw --csv | grep -E 'dram__bytes_(read|write)\.sum|gpu__time_duration|sm__warps_active|launch__registers_per_thread|local_'
#   ncu -i gpurun_out/r2_gather.ncu-rep --page source --csv
set -u
mkdir -p gpurun_out
MAGNETITE_B200_TEST_ASSEMBLY=1 timeout 300 python -m pytest tests -q -m gpu > gpurun_out/r2_suite_gather_default.log 2>&1
echo "suite(gather default) rc=$?" | tee -a gpurun_out/r2_suite_gather_default.log
timeout 300 python -m pytest tests -q -m gpu > gpurun_out/r2_suite_default.log 2>&1
echo "suite(default) rc=$?" | tee -a gpurun_out/r2_suite_default.log
if timeout 120 python profiles/assembly_probe.py 4000 2000 3 > gpurun_out/r2_assembly_probe.log 2>&1; then
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file gpurun_out/r2_launches_assembly_probe.csv python profiles/assembly_probe.py 4000 2000 1 \
        > gpurun_out/r2_ncu_launches.log 2>&1
    echo "ncu launch list rc=$?"
    # the last gather assembly of the run: skip the warm-up launches of both kernels
    timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:gather_(fill|count)_kernel' -s 2 -c 2 \
        -o gpurun_out/r2_gather python profiles/assembly_probe.py 4000 2000 1 > gpurun_out/r2_ncu_full.log 2>&1
    echo "ncu full capture rc=$?"
else
    echo "assembly probe failed; see gpurun_out/r2_assembly_probe.log"
fi
tail -3 gpurun_out/r2_suite_gather_default.log gpurun_out/r2_suite_default.log
cat gpurun_out/r2_assembly_probe.log
