# Evidence for profiles/: HBM roofline of the update kernels, ncu launch list + full captures, compute-sanitizer passes
set -u
O=gpurun_out/r02k; mkdir -p $O
timeout 300 python scripts/update_kernels_hbm.py > $O/update_kernels_hbm.json 2> $O/update_kernels_hbm.err; echo "hbm rc=$?"; tail -2 $O/update_kernels_hbm.err
timeout 300 python scripts/gn_time.py > $O/gn_time.txt 2>&1
# ncu: launch list of two training steps (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > $O/ncu_list.log 2>&1
python scripts/ncu_launch_summary.py $O/launches.csv 60 > $O/launch_summary.txt 2>&1
# ncu --set full: one capture each of the halo conv, the per-tap conv with the GroupNorm epilogue, the slab GroupNorm backward / forward, wgrad
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_halo_kernel|gn_bwd_smem_kernel|gn_fwd_smem_kernel" -c 6 -o $O/full_a -f python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > $O/ncu_full_a.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kernel" -s 40 -c 8 -o $O/full_b -f python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > $O/ncu_full_b.log 2>&1
ls -la $O/*.ncu-rep
# compute-sanitizer: memcheck + racecheck over the tensor-core / GroupNorm kernel tests (small cases)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tcgen05 or halo or split_k or gn_epilogue or single_pass" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gn_epilogue or single_pass or (tcgen05 and case0)" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -5 $O/sanitizer_memcheck.log; tail -5 $O/sanitizer_racecheck.log
