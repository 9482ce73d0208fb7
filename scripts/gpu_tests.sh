set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "not conv and not gemm and not sample_unshuffle and not upsample and not downsample" > gpurun_out/t_small.log 2>&1
echo "small rc=$?" >> gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "conv or gemm or upsample or downsample" > gpurun_out/t_conv.log 2>&1
echo "conv rc=$?" >> gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q --timeout 300 > gpurun_out/t_unet.log 2>&1
echo "unet rc=$?" >> gpurun_out/rc.txt
tail -5 gpurun_out/t_small.log gpurun_out/t_conv.log gpurun_out/t_unet.log
cat gpurun_out/rc.txt
