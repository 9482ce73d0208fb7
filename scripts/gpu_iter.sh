# one optimisation iteration on the GPU box: parity tests -> bench -> ncu launch list (shares per op)
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -q --timeout 300 -x > gpurun_out/tests.log 2>&1
rc=$?
tail -n 5 gpurun_out/tests.log
if [ $rc -ne 0 ]; then echo "TESTS FAILED"; tail -n 60 gpurun_out/tests.log; exit 1; fi
timeout 900 python bench.py --steps ${STEPS:-2} --warmup 3 --batch ${BENCH_BATCH:-1024} > gpurun_out/bench.log 2>&1
tail -n 2 gpurun_out/bench.log
if [ "${NCU:-1}" = "1" ]; then
B=${B:-1024} REPS=3 python scripts/profile_forward.py > gpurun_out/plain.log 2>&1 &&
B=${B:-1024} REPS=3 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_forward.py > gpurun_out/ncu1.log 2>&1
fi
