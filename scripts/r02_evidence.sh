# Evidence for profiles/: ncu launch lists of the training step and of a DDIM evaluation, ncu --set full captures of the main kernels
set -u
O=gpurun_out/r02e2; mkdir -p $O
NCU="ncu --clock-control none --profile-from-start off"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_train.csv python scripts/ncu_step.py train 2 > $O/ncu_list_train.log 2>&1; echo "list train rc=$?"
python scripts/ncu_launch_summary.py $O/launches_train.csv 60 > $O/launch_summary_train.txt 2>&1; head -5 $O/launch_summary_train.txt
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_ddim.csv python scripts/ncu_step.py ddim 1 > $O/ncu_list_ddim.log 2>&1; echo "list ddim rc=$?"
python scripts/ncu_launch_summary.py $O/launches_ddim.csv 60 > $O/launch_summary_ddim.txt 2>&1; head -5 $O/launch_summary_ddim.txt
# full captures (one training step): the largest conv (halo kernel), the GroupNorm-epilogue convs, weight gradients, slab GroupNorm backward, edge wgrad, Adam
timeout 1500 $NCU --set full --import-source on -k regex:"conv3x3_halo_kernel|gn_bwd_smem_kernel|edge_wgrad_tc_kernel|adam_ema_kernel" -c 8 -o $O/full_a -f python scripts/ncu_step.py train 1 > $O/ncu_full_a.log 2>&1; echo "full a rc=$?"
timeout 1500 $NCU --set full --import-source on -k regex:"conv_tc_gn_kernel" -s 4 -c 6 -o $O/full_b -f python scripts/ncu_step.py train 1 > $O/ncu_full_b.log 2>&1; echo "full b rc=$?"
timeout 1500 $NCU --set full --import-source on -k regex:"wgrad_tc_kernel" -s 2 -c 4 -o $O/full_c -f python scripts/ncu_step.py train 1 > $O/ncu_full_c.log 2>&1; echo "full c rc=$?"
# DDIM evaluation: tensor-core attention, the 64x64 halo conv with the statistics epilogue, the apply pass
timeout 1500 $NCU --set full --import-source on -k regex:"attn_fwd_tc_kernel|conv3x3_halo_kernel|gn_apply_kernel" -c 10 -o $O/full_d -f python scripts/ncu_step.py ddim 1 > $O/ncu_full_d.log 2>&1; echo "full d rc=$?"
ls -la $O/*.ncu-rep
