# end-of-round check on the GPU box: build, smoke, full GPU test-suite, default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('SMOKE OK')" 2>&1 | tail -n 2
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final.log").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "steps", "warmup", "ms_per_step", "gpu_launches")}, "e2e", round(d["e2e"]["value"], 1), d["clocks"],
      "frac", round(d["roofline"]["frac"], 3), "whole", round(d["roofline"]["whole_step"]["frac"], 3), "cpu", round(d["cpu_baseline"]["value"], 2))
PY
