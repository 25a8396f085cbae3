python -m pytest tests/test_gpu_parity.py tests/test_pool_identity.py tests/test_dropin.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t14.log
for w in "cfg3:" "cfg4:--workload cfg4_ssd512_coco --samples 50" "cfg5:--workload cfg5_retina_r101_1344_coco --steps 12" "cfg2:--workload cfg2_ssd300_voc"; do
  tag=${w%%:*}; AB_ARGS="${w#*:}" tools/ab_bench.sh default > gpurun_out/ab_$tag.txt 2>&1
  cp gpurun_out/ab_default.json gpurun_out/ab_${tag}_default.json
done
cat gpurun_out/r2_t14.log gpurun_out/ab_cfg3.txt gpurun_out/ab_cfg4.txt gpurun_out/ab_cfg5.txt gpurun_out/ab_cfg2.txt
