#!/bin/bash
# compute-sanitizer over the GPU tests that touch this round's new kernels (memcheck everywhere, racecheck on the
# shared-memory heavy ones); summaries under gpurun_out/.
SEL='stagewise_parity or pool_topk or guard_banded or aggregation_matrix or many_pairs or per_image_shapes'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mi.py -m gpu -q -x -k "$SEL or mi_" > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stagewise_parity and not random or pool_topk_matches or many_pairs" > gpurun_out/r2_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_racecheck.log
MEHHUA_PARKED_BULK=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_size" > gpurun_out/r2_memcheck_bulk.log 2>&1
echo "memcheck bulk rc=$?" >> gpurun_out/r2_memcheck_bulk.log
tail -4 gpurun_out/r2_memcheck.log gpurun_out/r2_racecheck.log gpurun_out/r2_memcheck_bulk.log
