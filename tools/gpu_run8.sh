#!/bin/bash
# round 2, N-GPU call without the pytest part: the sharded bench line (C4 commit, C3 one column, ONE Cairo proof), then C3 at 2^26
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2d_bench_n$N.json 2> gpurun_out/r2d_bench_n$N.err
grep -v "NCCL INFO" gpurun_out/r2d_bench_n$N.err | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cairo --c3-log-n 26 > gpurun_out/r2d_bench_c3_26_n$N.json 2> gpurun_out/r2d_bench_c3_26_n$N.err
grep -v "NCCL INFO" gpurun_out/r2d_bench_c3_26_n$N.err | tail -8
python - <<PY
import json
for f in ("gpurun_out/r2d_bench_n$N.json", "gpurun_out/r2d_bench_c3_26_n$N.json"):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f)
        print("  C4 sharded: %.2f ms/step (e2e %.2f) %.2f G elems/s parity %s stages %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"] / 1e9, d["parity_ok"], d["stages_ms"]))
        print("  C3:", d.get("c3_one_column"))
        c = d.get("cairo_prove", {})
        for k in ("fib", "fib_large"):
            if k in c:
                print(" ", k, c[k]["program"], "%.2f ms" % c[k]["value"], c[k].get("parity_ok"), c[k]["stages_ms"])
        if "error" in c:
            print("  cairo error:", c["error"])
    except Exception as e:
        print(f, "unreadable bench line:", e)
PY
