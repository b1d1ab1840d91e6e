#!/bin/bash
# Round evidence on one B200: official bench line, ncu launch list (time + DRAM bytes per launch), full-set captures.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_cfg3_1gpu.json 2> gpurun_out/bench_cfg3_1gpu.err; tail -c 600 gpurun_out/bench_cfg3_1gpu.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 330 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_run.log 2>&1
# third sorted k_trace launch of a wave (bounce 3), PBR shade of bounce 1
ncu --set full --clock-control none --import-source on -k regex:k_trace --launch-skip 3 -c 1 -f -o gpurun_out/trace_full \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade --launch-skip 5 -c 1 -f -o gpurun_out/shade_pbr_full \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1
ls -la gpurun_out
