# ncu launch lists of the training step and of a DDIM evaluation (final state of the round)
set -u
O=gpurun_out/r02f; mkdir -p $O
NCU="ncu --clock-control none --profile-from-start off"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_train.csv python scripts/ncu_step.py train 2 > $O/ncu_list_train.log 2>&1; echo "list train rc=$?"
python scripts/ncu_launch_summary.py $O/launches_train.csv 60 > $O/launch_summary_train.txt 2>&1; head -12 $O/launch_summary_train.txt
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_ddim.csv python scripts/ncu_step.py ddim 1 > $O/ncu_list_ddim.log 2>&1; echo "list ddim rc=$?"
python scripts/ncu_launch_summary.py $O/launches_ddim.csv 60 > $O/launch_summary_ddim.txt 2>&1; head -30 $O/launch_summary_ddim.txt
