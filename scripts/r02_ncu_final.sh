# ncu --set full captures of the kernels added / changed in the round's last session
set -u
O=gpurun_out/r02h; mkdir -p $O
NCU="ncu --clock-control none --profile-from-start off"
timeout 400 $NCU --set full --import-source on -k regex:"wgrad_halo_kernel|conv4x4t_halo_kernel|conv4x4s2_halo_kernel" -c 6 -o $O/full_train_new -f python scripts/ncu_step.py train 1 > $O/ncu_full_a.log 2>&1; echo "full train rc=$?"
timeout 400 $NCU --set full --import-source on -k regex:"conv3x3_halo_kernel|conv4x4t_halo_kernel|conv4x4s2_halo_kernel" -c 12 -o $O/full_ddim_new -f python scripts/ncu_step.py ddim 1 > $O/ncu_full_b.log 2>&1; echo "full ddim rc=$?"
ls -la $O/*.ncu-rep
