#!/bin/bash
# the single-GPU records of round 2 (run under gpurun from the repo root)
set -u
python bench.py > gpurun_out/final_cfg2.json 2> gpurun_out/final_cfg2.err
python bench.py --impl reference > gpurun_out/final_cfg2_reference.json 2> gpurun_out/final_cfg2_reference.err
for c in 1 3 4 5; do
  timeout 900 python bench.py --config $c --no-per-slab > gpurun_out/final_cfg$c.json 2> gpurun_out/final_cfg$c.err
done
timeout 120 python profiles/prof_step.py --config 2 --warmup 1 > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2_final.csv python profiles/prof_step.py --config 2 --warmup 1 > gpurun_out/ncu_ll_final.log 2>&1
tail -1 gpurun_out/ncu_ll_final.log
ls -la gpurun_out/final_cfg*.json
