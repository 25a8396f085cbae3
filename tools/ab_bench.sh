#!/bin/bash
# A/B of library builds on one GPU: tools/ab_bench.sh name1 name2 ...  ("default" = libmehhua.so, else libmehhua_<name>.so)
# prints images/s and the K2 stage time of `bench.py --steps 12 --warmup 3` per build
for v in "$@"; do
  if [ "$v" = default ]; then unset MEHHUA_LIB; else export MEHHUA_LIB=$PWD/aod_meh_hua_b200/libmehhua_$v.so; fi
  python bench.py --no-cpu-baseline --e2e-steps 0 --steps 12 --warmup 3 $AB_ARGS > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<PY
import json, sys
v = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{v}.json"))
    print(f"{v:12s} {d['value']:9.1f} img/s  K2 {d['roofline']['stage_ms_per_step']['k2_dirichlet']:.3f} ms  step {d['ms_per_step']:.3f} ms")
except Exception as e:
    print(v, "failed", e)
PY
done
