#!/bin/bash
# round 2, single-GPU tuning call: upload-group and FRI-tail sweeps, 2-D DMA microbench, ncu --set full of the NTT passes,
# compute-sanitizer on the small parity suite, and the default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_pytest.log
python tools/dma2d_bench.py > gpurun_out/r2c_dma2d.json 2> gpurun_out/r2c_dma2d.err; cat gpurun_out/r2c_dma2d.json; tail -3 gpurun_out/r2c_dma2d.err
for g in 2 3 4 8; do
  S252_HOST_GROUPS=$g python bench.py --steps 5 --warmup 3 --no-cairo --no-cpu-baseline --no-c4 > gpurun_out/r2c_bench_groups$g.json 2>/dev/null
done
for t in 9 10 12; do
  S252_FRI_TAIL_LOG=$t python bench.py --steps 5 --warmup 3 --no-cairo --no-cpu-baseline --no-c4 > gpurun_out/r2c_bench_tail$t.json 2>/dev/null
done
python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 400 gpurun_out/r2c_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d["int_roofline"]["kernels"]
        print(f, "ms/step %.2f e2e %.2f prefetch %.2f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["prefetch_pipeline"]["ms_per_step"]),
              {n: round(v["ms_per_step"], 2) for n, v in k.items()}, "frac", round(d["roofline"]["frac"], 3), d.get("cairo_prove", {}).get("value"),
              (d.get("c4_one_gpu") or {}).get("ms_per_step"), (d.get("c4_one_gpu") or {}).get("parity_ok"), (d.get("c4_one_gpu") or {}).get("error"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ntt_pass -s 12 -c 4 -o gpurun_out/r2c_ntt_full -f \
    python bench.py --steps 1 --warmup 3 --no-cairo --no-cpu-baseline --no-c4 > gpurun_out/r2c_ncu_full.log 2>&1
ncu -i gpurun_out/r2c_ntt_full.ncu-rep --page raw --csv > gpurun_out/r2c_ntt_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c_ntt_full.ncu-rep --page details --csv > gpurun_out/r2c_ntt_full_details.csv 2>/dev/null
ls -la gpurun_out/r2c_ntt_full* | head
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not large and not 18 and not 17" > gpurun_out/r2c_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2c_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fri or merkle or shared or interpolate_and_commit" > gpurun_out/r2c_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2c_racecheck.log
