# round-2 profile set (one gpurun call): launch list + `--set full` of the conv, fused linear attention and attention kernels
mkdir -p gpurun_out
B=1024 REPS=3 python scripts/profile_forward.py > gpurun_out/plain.log 2>&1 || exit 1
B=1024 REPS=3 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_forward.py > gpurun_out/ncu1.log 2>&1
B=1024 REPS=1 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 0 -c 2 -o gpurun_out/prof_conv python scripts/profile_forward.py > gpurun_out/ncu2.log 2>&1
B=1024 REPS=1 ncu --set full --clock-control none --import-source on -k regex:linattn_fused -s 0 -c 1 -o gpurun_out/prof_laf python scripts/profile_forward.py > gpurun_out/ncu3.log 2>&1
B=1024 REPS=1 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 0 -c 1 -o gpurun_out/prof_att python scripts/profile_forward.py > gpurun_out/ncu4.log 2>&1
tail -n 2 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log gpurun_out/ncu4.log
B=1024 REPS=1 ncu --set full --clock-control none --import-source on -k regex:stem_umma -s 0 -c 1 -o gpurun_out/prof_stem python scripts/profile_forward.py > gpurun_out/ncu5.log 2>&1
tail -n 2 gpurun_out/ncu5.log
