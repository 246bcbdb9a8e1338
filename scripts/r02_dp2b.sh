set -u
O=gpurun_out/r02z; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras > $O/dp2.json 2> $O/dp2.err
echo "dp2 rc=$?"; cut -c1-260 $O/dp2.json; tail -3 $O/dp2.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload ddim --no-extras > $O/dp2_ddim.json 2> $O/dp2_ddim.err
echo "dp2 ddim rc=$?"; cut -c1-260 $O/dp2_ddim.json
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras 2>/dev/null | cut -c1-260
