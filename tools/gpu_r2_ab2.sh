#!/bin/bash
mkdir -p gpurun_out
VARIANTS="r2_default r2_fsnolock" WL="f 524288" TAG=fslock bash tools/gpu_r2_variants.sh
for m in 1 2 4 8 16; do
  PBC_B200_PARAM_EXTRA="b200_prod_share $m" timeout 300 python bench.py --steps 3 --warmup 3 --workload prod16 --configs none --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(json.dumps({'prod_share': $m, 'outputs_per_s': d['value'], 'stage_ms': d['stage_ms'], 'e2e': d['e2e']['value'], 'same': d['e2e']['matches_device_resident_output']}))" | tee -a gpurun_out/r2_prod_share.jsonl
done
