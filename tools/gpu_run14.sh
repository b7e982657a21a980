#!/bin/bash
# round 2, session 2: Montgomery reduction with the 2^27 rows as funnel shifts (-DS252_MONT_SHIFT=1, library suffix _shift) against the default
mkdir -p gpurun_out
S252_LIB_SUFFIX=_shift timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_deep.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2p_pytest_shift.log
for v in "" "_shift"; do
  S252_LIB_SUFFIX=$v python bench.py --steps 5 --warmup 3 --no-cairo --no-cpu-baseline --no-c4 > gpurun_out/r2p_bench$v.json 2> gpurun_out/r2p_bench$v.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2p_bench$v.json").read().strip().splitlines()[-1])
k = d["int_roofline"]["kernels"]
print("variant '$v': ms/step %.3f" % d["ms_per_step"], {n: round(x["ms_per_step"], 3) for n, x in k.items() if x["ms_per_step"] > 0.3}, "root", d["result"]["last_root"][:16], "nonce", d["result"]["nonce"])
PY
done
