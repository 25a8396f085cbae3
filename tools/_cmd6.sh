python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t9.log
tools/ab_bench.sh default > gpurun_out/ab_cfg3.txt 2>&1; cp gpurun_out/ab_default.json gpurun_out/ab3_default.json
AB_ARGS="--workload cfg4_ssd512_coco --samples 50" tools/ab_bench.sh default > gpurun_out/ab_cfg4.txt 2>&1; cp gpurun_out/ab_default.json gpurun_out/ab4_default.json
AB_ARGS="--workload cfg5_retina_r101_1344_coco --steps 12" tools/ab_bench.sh default > gpurun_out/ab_cfg5.txt 2>&1; cp gpurun_out/ab_default.json gpurun_out/ab5_default.json
cat gpurun_out/r2_t9.log gpurun_out/ab_cfg3.txt gpurun_out/ab_cfg4.txt gpurun_out/ab_cfg5.txt
