for d in 2 3 4 6; do
  OBSGRID_BUDDY_CELL_DIV=$d timeout 300 python bench.py --config 4 --steps 5 --warmup 3 --no-per-slab --no-cpu --no-e2e --fdda-times 0 > gpurun_out/celldiv_$d.json 2> gpurun_out/celldiv_$d.err || tail -5 gpurun_out/celldiv_$d.err
done
OBSGRID_BUDDY_CELL_DIV=3 timeout 300 python bench.py --config 3 --steps 5 --warmup 3 --no-per-slab --no-cpu --no-e2e --fdda-times 0 > gpurun_out/celldiv_cfg3_3.json 2>/dev/null
OBSGRID_BUDDY_CELL_DIV=2 timeout 300 python bench.py --config 3 --steps 5 --warmup 3 --no-per-slab --no-cpu --no-e2e --fdda-times 0 > gpurun_out/celldiv_cfg3_2.json 2>/dev/null
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "buddy or qc" 2>&1 | tail -3
