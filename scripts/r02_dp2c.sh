set -u
O=gpurun_out/r02z; mkdir -p $O
for e in A=1 NCCL_PROTO=Simple; do
env $e timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras 2> $O/dp2c.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('dp2 $e', round(d['value']), d['ms_per_step'])"
done
