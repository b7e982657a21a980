#!/bin/bash
# round 2, single-GPU tuning call: kernel variants (parity + bench each), upload-group sweep, 2-D DMA microbench,
# ncu --set full of the NTT passes, compute-sanitizer on the small parity suite, the default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
for v in "" _gen _nof2 _gennof2 _lb2; do
  echo "== variant '$v'"
  S252_LIB_SUFFIX=$v python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "field or fft or transform or interpolate or fri_commit" 2>&1 | tail -1
  S252_LIB_SUFFIX=$v python bench.py --steps 5 --warmup 3 --no-cairo --no-cpu-baseline --no-c4 > gpurun_out/r2c_bench_var$v.json 2>/dev/null
done
python tools/dma2d_bench.py > gpurun_out/r2c_dma2d.json 2> gpurun_out/r2c_dma2d.err; cat gpurun_out/r2c_dma2d.json; tail -2 gpurun_out/r2c_dma2d.err
for g in 2 3 4 8; do
  S252_HOST_GROUPS=$g python bench.py --steps 5 --warmup 3 --no-cairo --no-cpu-baseline --no-c4 > gpurun_out/r2c_bench_groups$g.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d["int_roofline"]["kernels"]
        print(f, "ms/step %.2f e2e %.2f prefetch %.2f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["prefetch_pipeline"]["ms_per_step"]),
              {n: round(v["ms_per_step"], 2) for n, v in k.items() if v["ms_per_step"] > 0.3}, "frac", round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ntt_pass -s 12 -c 4 -o gpurun_out/r2c_ntt_full -f \
    python bench.py --steps 1 --warmup 3 --no-cairo --no-cpu-baseline --no-c4 > gpurun_out/r2c_ncu_full.log 2>&1
ncu -i gpurun_out/r2c_ntt_full.ncu-rep --page raw --csv > gpurun_out/r2c_ntt_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c_ntt_full.ncu-rep --page details --csv > gpurun_out/r2c_ntt_full_details.csv 2>/dev/null
ls -la gpurun_out/r2c_ntt_full* | head
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not 18 and not 17 and not 16 and not 15 and not three_pass" > gpurun_out/r2c_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2c_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fri_commit or merkle_build or shared_transform_two" > gpurun_out/r2c_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2c_racecheck.log
