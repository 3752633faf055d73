#!/bin/bash
# One gpurun call for the tabulated structure embedding (GSE mode 5): accuracy + kernel time, a short bench with the mode on,
# then the GPU tests that touch the embedding.  Everything lands in gpurun_out/ as it is produced.
mkdir -p gpurun_out
export GEOB200_GSE_MODE=5
timeout 150 python tests/gse_table_check.py > gpurun_out/gse_table_check.txt 2>&1
echo "check rc=$?" >> gpurun_out/gse_table_check.txt
timeout 170 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gse_mode5.json 2> gpurun_out/bench_gse_mode5.err
echo "bench rc=$?" >> gpurun_out/bench_gse_mode5.err
timeout 120 python -m pytest tests/test_gpu_ops.py -x -q -k "gse" > gpurun_out/pytest_gse_ops.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_batch.py tests/test_gpu_native.py -x -q > gpurun_out/pytest_e2e_mode5.txt 2>&1
tail -3 gpurun_out/gse_table_check.txt gpurun_out/pytest_gse_ops.txt gpurun_out/pytest_e2e_mode5.txt
tail -c 600 gpurun_out/bench_gse_mode5.err
head -c 400 gpurun_out/bench_gse_mode5.json
