#!/bin/bash
# BASELINE.json configs 3, 4 (T = 10 / 50 / 200) and 5 on N GPUs of one box + the H2D probe:
#   tools/scale_sweep.sh N [out.jsonl]       (one JSON line per run, appended)
N=${1:-1}
out=${2:-gpurun_out/r2_scale.jsonl}
run() {
  if [ "$N" = 1 ]; then timeout 900 python "$@"
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 "$@"; fi
}
run tools/h2d_probe.py 2>/dev/null | grep '^{' >> $out
run bench.py --gpus $N --no-cpu-baseline 2>/dev/null | grep '^{' >> $out
for T in 10 50 200; do
  run bench.py --gpus $N --workload cfg4_ssd512_coco --samples $T --no-cpu-baseline --e2e-steps 3 2>/dev/null | grep '^{' >> $out
done
run bench.py --gpus $N --workload cfg5_retina_r101_1344_coco --steps 24 --warmup 3 --no-cpu-baseline --e2e-steps 3 2>/dev/null | grep '^{' >> $out
python - "$out" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    if d.get("probe") == "h2d":
        print(f"h2d N={d['n_gpus']}: per rank {[round(r['default_affinity'], 1) for r in d['ranks']]} GB/s, aggregate {d['aggregate_default']:.1f}"
              f" | numa-local {d['aggregate_numa_local']}")
    else:
        c = d["config"]
        print(f"N={d['n_gpus']} {c['workload']} T={c['samples']} B={c['batch_per_gpu']} steps={d['steps']}: {d['value']:.0f} img/s, e2e {d['e2e']['value']:.0f}, "
              f"k1 stage frac {d['roofline']['k1_stage_frac']:.3f}, K2 {d['roofline']['k2']['draws_per_s']:.3e} draws/s")
PY
