# 2-GPU step: range-wise Adam beside the last all-reduce; NCCL's CTA count capped
set -u
O=gpurun_out/r02j; mkdir -p $O
i=0
for v in default 16 default 16; do
  i=$((i+1))
  if [ "$v" = "default" ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$v; fi
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-cpu > $O/dp2_$i.json 2> $O/dp2_$i.err
  echo "NCCL_MAX_CTAS=$v rc=$?"; python -c "
import json,sys
d=json.loads(open('$O/dp2_$i.json').read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'])" 2>&1 | tail -1
done
