#!/bin/bash
set -u
python bench.py > gpurun_out/final_cfg2.json 2> gpurun_out/final_cfg2.err
tail -2 gpurun_out/final_cfg2.err
timeout 600 python bench.py --config 5 --no-per-slab --no-cpu --no-e2e --steps 3 > gpurun_out/final_cfg5.json 2> gpurun_out/final_cfg5.err
tail -2 gpurun_out/final_cfg5.err
ls -la gpurun_out/final_cfg2.json gpurun_out/final_cfg5.json
