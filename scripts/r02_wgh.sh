#!/bin/bash
# A/B of the halo weight-gradient kernel inside the training step (development aid)
source scripts/ab.sh
mkdir -p gpurun_out/r02p3
{
run DMU_WGRAD_HALO=0
run DMU_WGRAD_HALO=1






run DMU_WGRAD_HALO=0
run DMU_WGRAD_HALO=1
} 2>&1 | tee gpurun_out/r02p3/ab.txt
