#!/bin/bash
# One GPU call that refreshes the evidence under gpurun_out/ (copied to profiles/ afterwards):
#   GPU tests, the default bench line, launch lists of cfg 3 / cfg 4, ncu --set full of the K1 / K2 kernels, the N=1 sweep.
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_gputests_final.log
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
tools/launch_list.sh gpurun_out/r2_launch_list_final.csv > gpurun_out/r2_launch_list_final.txt 2>&1
tools/launch_list.sh gpurun_out/r2_launch_list_cfg4.csv --workload cfg4_ssd512_coco --samples 50 > gpurun_out/r2_launch_list_cfg4.txt 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k 'regex:k1t_threshold|k1a_keys|k1b_select|k1c_parked|k2_dirichlet' -c 6 -f \
    -o gpurun_out/r2_final_kernels python bench.py --no-cpu-baseline --batch 128 --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/ncu_final.log 2>&1
rm -f gpurun_out/r2_scale_n1.jsonl
tools/scale_sweep.sh 1 gpurun_out/r2_scale_n1.jsonl > gpurun_out/r2_scale_n1.txt 2>&1
cat gpurun_out/r2_gputests_final.log gpurun_out/r2_launch_list_final.txt gpurun_out/r2_scale_n1.txt; tail -2 gpurun_out/ncu_final.log | cut -c1-300
