set -u
O=gpurun_out/r02h; mkdir -p $O
for v in 0 1; do
  DMU_DP_GRAPH=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$v bench.py --gpus 2 --steps 30 --warmup 5 --no-extras > $O/dp_graph$v.json 2> $O/dp_graph$v.err
  echo "DMU_DP_GRAPH=$v rc=$?" | tee -a $O/summary.txt; cut -c1-220 $O/dp_graph$v.json | tee -a $O/summary.txt; tail -3 $O/dp_graph$v.err | cut -c1-300
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras 2>/dev/null | cut -c1-220 | tee -a $O/summary.txt
