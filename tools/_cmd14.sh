python -m pytest tests/test_gpu_sampler.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_t15.log
for w in "cfg3:" "cfg4:--workload cfg4_ssd512_coco --samples 50" "cfg4t10:--workload cfg4_ssd512_coco --samples 10"; do
  tag=${w%%:*}; AB_ARGS="${w#*:}" tools/ab_bench.sh default > gpurun_out/ab_$tag.txt 2>&1
  cp gpurun_out/ab_default.json gpurun_out/ab_${tag}_default.json
done
cat gpurun_out/r2_t15.log gpurun_out/ab_cfg3.txt gpurun_out/ab_cfg4.txt gpurun_out/ab_cfg4t10.txt
