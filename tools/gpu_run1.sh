#!/bin/bash
# round 2, first GPU call: parity suite, bench line, upload-mode and FRI-tail variants, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.err
for mode in dma2d copy; do
  S252_HOST_UPLOAD=$mode python bench.py --steps 5 --warmup 3 --no-cairo --no-cpu-baseline > gpurun_out/r2a_bench_$mode.json 2>/dev/null
done
for g in 3 10; do
  S252_HOST_GROUPS=$g python bench.py --steps 5 --warmup 3 --no-cairo --no-cpu-baseline > gpurun_out/r2a_bench_groups$g.json 2>/dev/null
done
for t in 0 11 14; do
  S252_FRI_TAIL_LOG=$t python bench.py --steps 5 --warmup 3 --no-cairo --no-cpu-baseline > gpurun_out/r2a_bench_tail$t.json 2>/dev/null
done
python tools/sweep.py c5 > gpurun_out/r2a_sweep_c5.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2a_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d["int_roofline"]["kernels"]
        print(f, "ms/step %.2f e2e %.2f prefetch %.2f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["prefetch_pipeline"]["ms_per_step"]),
              {n: round(v["ms_per_step"], 2) for n, v in k.items()}, "frac", round(d["roofline"]["frac"], 3), d.get("cairo_prove", {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-cairo --no-cpu-baseline > gpurun_out/r2a_ncu_bench.log 2>&1
tail -3 gpurun_out/r2a_sweep_c5.log
