timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:k3a_nms|k1c_rescan|k1c_gather|k1b_select' -c 4 -f -o gpurun_out/r2_cfg4_kernels python bench.py --workload cfg4_ssd512_coco --samples 50 --no-cpu-baseline --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/ncu_cfg4.log 2>&1
tail -2 gpurun_out/ncu_cfg4.log | cut -c1-200
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pool_topk" 2>&1 | tail -2
