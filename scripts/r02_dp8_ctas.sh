# 8-GPU step: NCCL's CTA count capped vs default
set -u
O=gpurun_out/r02k8; mkdir -p $O
i=0
for v in 16 default; do
  i=$((i+1))
  if [ "$v" = "default" ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$v; fi
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2955$i bench.py --gpus 8 --steps 30 --warmup 5 --no-extras --no-cpu > $O/dp8_$v.json 2> $O/dp8_$v.err
  echo "NCCL_MAX_CTAS=$v rc=$?"; python -c "
import json,sys
d=json.loads(open('$O/dp8_$v.json').read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'])" 2>&1 | tail -1
done
