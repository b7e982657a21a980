#!/bin/bash
# round 2, multi-GPU call (N = number of visible GPUs): NCCL parity tests of the sharded paths, then the sharded bench line
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
python -m pytest tests/test_distributed.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b_pytest_n$N.log; cat gpurun_out/r2b_pytest_n$N.log
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2b_bench_n$N.json 2> gpurun_out/r2b_bench_n$N.err
tail -c 1500 gpurun_out/r2b_bench_n$N.err | grep -v "NCCL INFO" | tail -20
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2b_bench_n$N.json").read().strip().splitlines()[-1])
    print("C4 sharded: %.2f ms/step (e2e %.2f) %.2f G elems/s parity %s stages %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"] / 1e9, d["parity_ok"], d["stages_ms"]))
    c = d.get("cairo_prove", {})
    for k in ("fib", "fib_large"):
        if k in c:
            print(k, c[k]["program"], "%.2f ms" % c[k]["value"], c[k].get("parity_ok"), c[k]["stages_ms"])
    if "error" in c:
        print("cairo error:", c["error"])
except Exception as e:
    print("unreadable bench line:", e)
PY
