#!/bin/bash
# round 2, session 2: warp-level barriers after the first radix-4 steps (default) against block barriers everywhere (-DS252_NTT_WARP_SYNC=0, suffix _nows)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2r_pytest_ws.log
for v in "_nows" ""; do
  S252_LIB_SUFFIX=$v python bench.py --steps 5 --warmup 3 --no-cairo --no-cpu-baseline --no-c4 > gpurun_out/r2r_bench$v.json 2> gpurun_out/r2r_bench$v.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2r_bench$v.json").read().strip().splitlines()[-1])
k = d["int_roofline"]["kernels"]
print("variant '$v': ms/step %.3f" % d["ms_per_step"], {n: round(x["ms_per_step"], 3) for n, x in k.items() if x["ms_per_step"] > 8}, "root", d["result"]["last_root"][:16], "nonce", d["result"]["nonce"])
PY
done
