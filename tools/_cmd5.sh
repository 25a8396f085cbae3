python tools/capture_stats.py > gpurun_out/r2_capture_stats_cfg3.json 2> gpurun_out/cs.err
python tools/capture_stats.py --workload cfg5_retina_r101_1344_coco --images 148 --chunk 37 > gpurun_out/r2_capture_stats_cfg5.json 2>> gpurun_out/cs.err
python tools/capture_stats.py --workload cfg4_ssd512_coco --images 296 > gpurun_out/r2_capture_stats_cfg4.json 2>> gpurun_out/cs.err
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:k1b_select|k3a_nms|k3b_pairs|k3c_hua|k1t_threshold|k1c_parked' -c 6 -f -o gpurun_out/r2_small_kernels python bench.py --no-cpu-baseline --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/ncu_small.log 2>&1
cat gpurun_out/r2_capture_stats_cfg3.json gpurun_out/r2_capture_stats_cfg5.json gpurun_out/r2_capture_stats_cfg4.json; tail -3 gpurun_out/cs.err; tail -5 gpurun_out/ncu_small.log
