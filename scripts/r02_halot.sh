#!/bin/bash
source scripts/ab.sh
mkdir -p gpurun_out/r02p8
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "transposed_halo or tcgen05" > gpurun_out/r02p8/test.log 2>&1; tail -5 gpurun_out/r02p8/test.log
{
run DMU_HALO_T=0
run DMU_HALO_T=1

rund DMU_HALO_T=0
rund DMU_HALO_T=1
} 2>&1 | tee gpurun_out/r02p8/ab.txt
