#!/bin/bash
source scripts/ab.sh
mkdir -p gpurun_out/r02p6
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py -m gpu -x -q > gpurun_out/r02p6/test.log 2>&1; tail -3 gpurun_out/r02p6/test.log
{
run A=1
rund A=1
} 2>&1 | tee gpurun_out/r02p6/ab.txt
