#!/bin/bash
# Round-2 ncu session: launch list of the default bench command, one --set full capture of every kernel of one e2e step.
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
python tools/run_step.py --steps 2 > gpurun_out/plain_step.log 2>&1 &&
NL=$(grep "^step 1" gpurun_out/plain_step.log | awk '{print $4}') && echo "launches per step: $NL" &&
ncu --set full --clock-control none \
    -k regex:'gemm_tc|batch_attn|layernorm|patch_im2col|conv3x3|attn_tc|proj_tc|stats_partial|fold_stats|pad_reflect|f32_to_bf16' \
    -s $NL -c $NL -o gpurun_out/r02_step python tools/run_step.py --steps 2 > gpurun_out/ncu_step.log 2>&1
echo "full capture exit $?"
# the report is > 64 MiB: export what we read here and drop it
ncu -i gpurun_out/r02_step.ncu-rep --page raw --csv > gpurun_out/r02_step_raw.csv 2> gpurun_out/ncu_export.err
ncu -i gpurun_out/r02_step.ncu-rep --page details --csv > gpurun_out/r02_step_details.csv 2>> gpurun_out/ncu_export.err
rm -f gpurun_out/r02_step.ncu-rep; ls -la gpurun_out | head -40
cat gpurun_out/plain_step.log; tail -3 gpurun_out/ncu_step.log
