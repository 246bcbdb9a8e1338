#!/bin/bash
source scripts/ab.sh
mkdir -p gpurun_out/r02p10
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "strided_halo or transposed_halo or wgrad_halo or tcgen05_halo" > gpurun_out/r02p10/test.log 2>&1; tail -3 gpurun_out/r02p10/test.log
timeout 100 python scripts/resample_time.py 2>&1 | grep strided | tee gpurun_out/r02p10/resample.txt
timeout 100 python scripts/wgrad_time.py 2>&1 | tee gpurun_out/r02p10/wgrad.txt
{
run A=1
rund A=1
} 2>&1 | tee gpurun_out/r02p10/ab.txt
