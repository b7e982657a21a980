#!/bin/bash
# round 2, session 2, final single-GPU call: full parity suite, smoke(), the default bench line, ncu launch list, then the generated-product-rows variant
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2t_pytest.log; cat gpurun_out/r2t_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; tail -c 600 gpurun_out/r2t_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2t_bench.json").read().strip().splitlines()[-1])
k = d["int_roofline"]["kernels"]
print("ms/step %.2f e2e %.2f prefetch %.2f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["prefetch_pipeline"]["ms_per_step"]),
      {n: round(v["ms_per_step"], 2) for n, v in k.items() if v["ms_per_step"] > 0.3}, "frac", round(d["roofline"]["frac"], 3), "step_frac", round(d["int_roofline"]["step_frac"], 3))
print("cairo", d.get("cairo_prove", {}).get("value"), d.get("cairo_prove", {}).get("stages_ms"))
c4 = d.get("c4_one_gpu") or {}
print("c4_one_gpu", {k_: c4.get(k_) for k_ in ("ms_per_step", "parity_ok", "stages_ms", "error")}, (c4.get("e2e") or {}).get("ms_per_step"))
print("cpu", d.get("cpu_baseline", {}).get("value"), "peak", d["roofline"]["peak"])
print("c3_one_gpu", (d.get("c3_one_gpu") or {}).get("ms"), "c5_one_gpu", (d.get("c5_one_gpu") or {}).get("ms"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 2 --warmup 3 --no-cairo --no-cpu-baseline --no-c4 > gpurun_out/r2t_ncu_bench.log 2>&1
