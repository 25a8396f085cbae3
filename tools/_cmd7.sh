python -m pytest tests/test_gpu_parity.py tests/test_pool_identity.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t10.log
MEHHUA_LIB=$PWD/aod_meh_hua_b200/libmehhua_x_t64.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t10_t64.log
tools/ab_bench.sh default x_t64 x_t32 > gpurun_out/ab_cfg3.txt 2>&1
for v in default x_t64 x_t32; do cp gpurun_out/ab_$v.json gpurun_out/ab3_$v.json; done
AB_ARGS="--workload cfg4_ssd512_coco --samples 50" tools/ab_bench.sh default x_t64 x_t32 > gpurun_out/ab_cfg4.txt 2>&1
for v in default x_t64 x_t32; do cp gpurun_out/ab_$v.json gpurun_out/ab4_$v.json; done
cat gpurun_out/r2_t10.log gpurun_out/r2_t10_t64.log gpurun_out/ab_cfg3.txt gpurun_out/ab_cfg4.txt
