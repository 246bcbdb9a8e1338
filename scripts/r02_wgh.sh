#!/bin/bash
# A/B of the halo weight-gradient kernel inside the training step (development aid)
source scripts/ab.sh
mkdir -p gpurun_out/r02p3
{
run DMU_WGRAD_HALO=0
run DMU_WGRAD_HALO=1
run DMU_WGRAD_HALO_MIN_TILES=300
run DMU_WGRAD_HALO_CTAS=74
run DMU_WGRAD_HALO_CTAS=100
run DMU_WGRAD_HALO_CTAS=111 DMU_WGRAD_HALO_MIN_TILES=300
run DMU_WGRAD_HALO_STAGES=2
run DMU_WGRAD_HALO_STAGES=4
run DMU_WGRAD_HALO=0
run DMU_WGRAD_HALO=1
} 2>&1 | tee gpurun_out/r02p3/ab.txt
