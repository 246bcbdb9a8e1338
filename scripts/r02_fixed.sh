#!/bin/bash
source scripts/ab.sh
mkdir -p gpurun_out/r02m
timeout 120 python -m pytest tests/test_gpu_unet.py -m gpu -x -q -k "bit_identical" > gpurun_out/r02m/test0.log 2>&1; tail -4 gpurun_out/r02m/test0.log
timeout 90 python scripts/determinism_check.py 2>&1 | grep -v Warn | tail -2
DMU_GN_FIXED_SUMS=0 timeout 90 python scripts/determinism_check.py 2>&1 | grep -v Warn | tail -1
{
run DMU_GN_FIXED_SUMS=0
run DMU_GN_FIXED_SUMS=1
rund DMU_GN_FIXED_SUMS=0
rund DMU_GN_FIXED_SUMS=1
} 2>&1 | tee gpurun_out/r02m/ab.txt | cut -c1-80
