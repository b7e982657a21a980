#!/bin/bash
# round 2, session 2, N-GPU call: NCCL tests (block-mode DEEP tables on real ranks), then the sharded bench line with the C5 object
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 900 python -m pytest tests/test_distributed.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2g_pytest_n$N.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err
grep -v "NCCL INFO" gpurun_out/r2g_bench_n$N.err | tail -15
python - <<PY
import json
f = "gpurun_out/r2g_bench_n$N.json"
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    print("  C4 sharded: %.2f ms/step (e2e %.2f) %.2f G elems/s parity %s stages %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"] / 1e9, d["parity_ok"], d["stages_ms"]))
    print("  roofline:", {k: d["roofline"][k] for k in ("kernel", "achieved", "peak", "frac", "peak_source")})
    print("  C3:", d.get("c3_one_column"))
    print("  C5:", d.get("c5_fri"))
    c = d.get("cairo_prove", {})
    for k in ("fib", "fib_large"):
        if k in c:
            print(" ", k, c[k]["program"], "%.2f ms" % c[k]["value"], c[k].get("parity_ok"), c[k]["stages_ms"], c[k]["commit_detail_ms"])
    if "error" in c:
        print("  cairo error:", c["error"])
except Exception as e:
    print(f, "unreadable bench line:", e)
PY
