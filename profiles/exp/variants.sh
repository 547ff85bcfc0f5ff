#!/bin/bash
# bench every variant library under obsgrid_b200/lib/variants (and the default build) on cfg2, resident only
cd "$(dirname "$0")/../.."
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    b=d["breakdown"]; print(f'{sys.argv[1]:40s} step {d["ms_per_step"]:.3f} ms  qc {b["qc_ms"]:.3f} oa {b["oa_ms"]:.3f} scan {b["cressman_scan_kernels_ms"]:.3f}  frac {d["roofline"]["frac"]:.4f}')
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
python bench.py --no-cpu --no-e2e --steps 5 "$@" > gpurun_out/var_default.json 2> gpurun_out/var_default.err; show gpurun_out/var_default.json
for so in obsgrid_b200/lib/variants/lib_*.so; do
  n=$(basename $so .so)
  OBSGRID_GPU_LIB=$PWD/$so python bench.py --no-cpu --no-e2e --steps 5 "$@" > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err; show gpurun_out/var_$n.json
done
