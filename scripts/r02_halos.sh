#!/bin/bash
source scripts/ab.sh
mkdir -p gpurun_out/r02p9
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "strided_halo or transposed_halo or tcgen05" > gpurun_out/r02p9/test.log 2>&1; tail -5 gpurun_out/r02p9/test.log
{
run DMU_HALO_S=0
run DMU_HALO_S=1
rund DMU_HALO_S=0
rund DMU_HALO_S=1
} 2>&1 | tee gpurun_out/r02p9/ab.txt
