mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
B=${B:-1024} REPS=1 python scripts/profile_forward.py > gpurun_out/plain.log 2>&1 &&
B=${B:-1024} REPS=1 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-linattn} -s ${SKIP:-0} -c ${COUNT:-1} -o gpurun_out/prof_k python scripts/profile_forward.py > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu2.log
