#!/bin/bash
# Bench sweep over the BASELINE.json configs (one GPU).  Output: one JSON line per run.
out=${1:-gpurun_out/r1_sweep.jsonl}
: > $out
for wl in cfg1_retina_r50_512_voc cfg2_ssd300_voc cfg3_retina_r50_800x1344_coco cfg3p_retina_r50_800x800_coco cfg5_retina_r101_1344_coco; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>/dev/null | tail -1 >> $out
done
for T in 10 50 200 500; do
  timeout 600 python bench.py --workload cfg4_ssd512_coco --samples $T --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>/dev/null | tail -1 >> $out
done
python - <<PY
import json
for l in open("$out"):
    l=l.strip()
    if not l: continue
    d=json.loads(l)
    st=d["roofline"]["stage_ms_per_step"]
    print(d["config"]["workload"], "T", d["config"]["samples"], "B", d["config"]["batch_per_gpu"], "img/s %.0f"%d["value"], "k1a frac %.3f"%d["roofline"]["frac"], "e2e %.0f"%d["e2e"]["value"], {k: round(v,3) for k,v in st.items()}, "status", d["config"]["status_bits"])
PY
