set -u
O=gpurun_out/r02dp8; mkdir -p $O
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 30 --warmup 5 --no-extras > $O/dp$N.json 2> $O/dp$N.err
echo "dp$N rc=$?"; cut -c1-260 $O/dp$N.json; tail -2 $O/dp$N.err | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/ref$N.json 2> $O/ref$N.err
echo "ref$N rc=$?"; cut -c1-200 $O/ref$N.json
