#!/bin/bash
# single-GPU records after the buddy-check culling (run under gpurun from the repo root)
set -u
python bench.py --steps 20 --warmup 5 > gpurun_out/final_cfg2.json 2> gpurun_out/final_cfg2.err
tail -1 gpurun_out/final_cfg2.err
timeout 300 python bench.py --config 1 --no-per-slab > gpurun_out/final_cfg1.json 2> gpurun_out/final_cfg1.err
timeout 400 python bench.py --config 3 --no-per-slab --no-cpu > gpurun_out/final_cfg3.json 2> gpurun_out/final_cfg3.err
timeout 600 python bench.py --config 4 --no-per-slab --no-cpu > gpurun_out/final_cfg4.json 2> gpurun_out/final_cfg4.err
timeout 600 python bench.py --config 5 --no-per-slab --no-cpu --no-e2e --steps 3 > gpurun_out/final_cfg5.json 2> gpurun_out/final_cfg5.err
timeout 120 python profiles/prof_step.py --config 2 --warmup 1 > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2_final.csv python profiles/prof_step.py --config 2 --warmup 1 > gpurun_out/ncu_ll_final.log 2>&1
ls -la gpurun_out/final_cfg[1-5].json
