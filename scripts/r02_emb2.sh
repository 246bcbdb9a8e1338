#!/bin/bash
source scripts/ab.sh
mkdir -p gpurun_out/r02p5
{
run DMU_EARLY_EMB=0
run DMU_EARLY_EMB=0 DMU_SIDE_PRIO=1
run DMU_EARLY_EMB=1 DMU_SIDE_PRIO=1
run DMU_EARLY_EMB=1 DMU_SIDE_PRIO=1 DMU_WGRAD_HALO=0
run DMU_EARLY_EMB=0 DMU_WGRAD_HALO=0
} 2>&1 | tee gpurun_out/r02p5/ab.txt
