mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
tail -n 1 gpurun_out/bench_n$N.log | cut -c1-400
tail -n 5 gpurun_out/bench_n$N.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -n 1 gpurun_out/bench_ref.log | cut -c1-600
