python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
export TAGS=${TAGS:-downs.0.0.block1,downs.0.0.block2,downs.0.2.to_qkv,downs.0.2.to_out,ups.3.0.res_conv}
for d in ${DBGS:-0 1 2 3 7}; do DDM_CONV_DEBUG=$d python scripts/conv_bisect.py 2>&1 | grep DBG; done
