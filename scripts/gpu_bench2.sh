mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("RUN", round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],1), d["clocks"]["samples"])'
for i in 1 2 3; do
DDM_BENCH_NO_CLOCKS=1 timeout 900 python bench.py --steps 2 --warmup 3 2>/dev/null | python -c "$P"
done
for i in 1 2 3; do
timeout 900 python bench.py --steps 2 --warmup 3 2>/dev/null | python -c "$P"
done
