#!/bin/bash
source scripts/ab.sh
mkdir -p gpurun_out/r02p4
timeout 600 python -m pytest tests/test_gpu_unet.py -m gpu -x -q > gpurun_out/r02p4/test_unet.log 2>&1; tail -3 gpurun_out/r02p4/test_unet.log
{
run DMU_EARLY_EMB=0
run DMU_EARLY_EMB=1
run DMU_EARLY_EMB=1 DMU_WGRAD_HALO_STAGES=3
run DMU_EARLY_EMB=0
run DMU_EARLY_EMB=1
} 2>&1 | tee gpurun_out/r02p4/ab.txt
python scripts/step_trace.py > gpurun_out/r02p4/trace.log 2>&1; cp gpurun_out/*.csv gpurun_out/r02p4/ 2>/dev/null; ls gpurun_out/r02p4
