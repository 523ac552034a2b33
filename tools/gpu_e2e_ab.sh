#!/bin/bash
# e2e (host buffers) of every default config for several builds of the library: LIBS="build/liba.so build/libb.so"
mkdir -p gpurun_out
: > gpurun_out/r2_e2e_ab.jsonl
for lib in $LIBS; do
  PBC_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > /tmp/e2e_ab.json 2> /tmp/e2e_ab.err || tail -3 /tmp/e2e_ab.err
  python - "$lib" <<'PY' | tee -a gpurun_out/r2_e2e_ab.jsonl
import json, sys
d = json.loads([l for l in open('/tmp/e2e_ab.json') if l.startswith('{')][-1])
row = {"lib": sys.argv[1], "a": [round(d["value"]), round(d["e2e"]["value"])]}
for k, v in (d.get("configs") or {}).items():
    if "error" not in v: row[k] = [round(v["value"]), round(v["e2e"]["value"])]
print(json.dumps(row))
PY
done
