# 4-GPU bench line at the round's last commit
set -u
O=gpurun_out/r02y; mkdir -p $O
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras --no-cpu > $O/dp4.json 2> $O/dp4.err
echo "rc=$?"; cut -c1-200 $O/dp4.json
