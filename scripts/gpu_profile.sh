set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
B=${B:-1024} REPS=3 python scripts/profile_forward.py > gpurun_out/plain.log 2>&1 &&
B=${B:-1024} REPS=3 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_forward.py > gpurun_out/ncu1.log 2>&1
B=${B:-1024} REPS=2 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 60 -c 4 -o gpurun_out/prof_conv python scripts/profile_forward.py > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/plain.log gpurun_out/ncu1.log gpurun_out/ncu2.log
ls -la gpurun_out
