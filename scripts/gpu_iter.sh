# one optimisation iteration on the GPU box: parity tests -> bench (x RUNS) -> ncu launch list (shares per op)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -q --timeout 300 -x > gpurun_out/tests.log 2>&1
rc=$?
tail -n 3 gpurun_out/tests.log
if [ $rc -ne 0 ]; then echo "TESTS FAILED"; tail -n 60 gpurun_out/tests.log; exit 1; fi
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("RUN value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "conv_us", round(d["roofline"]["us_per_launch"],1), "launches", d["gpu_launches"], d["clocks"])'
for i in $(seq 1 ${RUNS:-2}); do
timeout 900 python bench.py --steps ${STEPS:-2} --warmup 3 --batch ${BENCH_BATCH:-1024} > gpurun_out/bench.log 2> gpurun_out/bench.err || tail -n 20 gpurun_out/bench.err
python -c "$P" < gpurun_out/bench.log
done
if [ "${NCU:-1}" = "1" ]; then
B=${B:-1024} REPS=3 python scripts/profile_forward.py > gpurun_out/plain.log 2>&1 &&
B=${B:-1024} REPS=3 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_forward.py > gpurun_out/ncu1.log 2>&1
fi
