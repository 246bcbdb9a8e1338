# First GPU calls of the next round: the two switches that were written after round 1's GPU budget was spent.
# Run under gpurun, e.g.  gpurun --timeout 300 -- 'bash scripts/next_gpu_checks.sh 2>&1 | tee gpurun_out/next_checks.txt'
set -u
. scripts/ab.sh

echo "== 1. tensor-core weight gradient of the 3-channel boundary layers (conv_edge_tc.cu) =="
DMU_EDGE_WGRAD_TC=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k "stem or head or wgrad" 2>&1 | tail -5
DMU_EDGE_WGRAD_TC=1 timeout 200 python -m pytest tests/test_gpu_unet.py -x -q -k "train_step_golden or whole_step" 2>&1 | tail -3

echo "== 2. A/B inside the step graph =="
run DMU_EDGE_WGRAD_TC=0
run DMU_EDGE_WGRAD_TC=1
run DMU_EDGE_WGRAD_TC=1 DMU_EDGE_WGRAD_CTAS=1
run DMU_EDGE_WGRAD_TC=1 DMU_EDGE_WGRAD_CTAS=4

# 3. (needs gpurun --gpus 2) all-reduces inside the step graph:
#   for v in 0 1; do DMU_DP_GRAPH=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#       --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 | tail -1 | cut -c1-200; done
