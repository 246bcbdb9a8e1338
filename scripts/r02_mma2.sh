#!/bin/bash
source scripts/ab.sh
mkdir -p gpurun_out/r02p11
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py -m gpu -x -q > gpurun_out/r02p11/test.log 2>&1; tail -3 gpurun_out/r02p11/test.log
timeout 100 python scripts/halo_timeline.py plain > gpurun_out/r02p11/halo_tl.txt 2>&1; sed -n 1,14p gpurun_out/r02p11/halo_tl.txt | cut -c1-100
{
run A=1
rund A=1
} 2>&1 | tee gpurun_out/r02p11/ab.txt
