python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
for d in 0 1 2 3 4 5 6 7; do DDM_CONV_DEBUG=$d python scripts/conv_bisect.py 2>&1 | grep DBG; done
