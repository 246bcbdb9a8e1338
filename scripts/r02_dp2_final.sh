# 2-GPU bench line at the round's last commit (NCCL_MAX_CTAS default from the package)
set -u
O=gpurun_out/r02s; mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras > $O/dp2.json 2> $O/dp2.err
echo "rc=$?"; cut -c1-250 $O/dp2.json; tail -2 $O/dp2.err | cut -c1-200
