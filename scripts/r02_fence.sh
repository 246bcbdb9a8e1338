#!/bin/bash
source scripts/ab.sh
mkdir -p gpurun_out/r02p12
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py -m gpu -x -q > gpurun_out/r02p12/test.log 2>&1; tail -2 gpurun_out/r02p12/test.log
timeout 100 python scripts/halo_timeline.py plain > gpurun_out/r02p12/halo_tl.txt 2>&1; sed -n 22,30p gpurun_out/r02p12/halo_tl.txt | cut -c1-100
{
run A=1
rund A=1
} 2>&1 | tee gpurun_out/r02p12/ab.txt
