set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "sampler_step" > gpurun_out/t_small.log 2>&1
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q --timeout 300 -k "self_condition" > gpurun_out/t_unet.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 --batch ${BENCH_BATCH:-1024} > gpurun_out/bench.log 2>&1
tail -n 3 gpurun_out/t_small.log gpurun_out/t_unet.log gpurun_out/smoke.log gpurun_out/bench.log
