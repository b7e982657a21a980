#!/bin/bash
# round 2, session 2, N-GPU call: the full sharded bench line (C4 through the C-ABI collective call + the torch.distributed path, C3, C5, ONE Cairo proof)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2n_bench_n$N.json 2> gpurun_out/r2n_bench_n$N.err
grep -v "NCCL INFO" gpurun_out/r2n_bench_n$N.err | tail -8
grep -c "NCCL INFO" gpurun_out/r2n_bench_n$N.err
python - <<PY
import json
f = "gpurun_out/r2n_bench_n$N.json"
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    print("  C4:", d.get("call"), "| %.2f ms/step (e2e %.2f) %.2f G elems/s parity %s stages %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"] / 1e9, d["parity_ok"], d["stages_ms"]))
    print("  torch path:", {k: v for k, v in (d.get("torch_distributed_path") or {}).items() if k != "how"}, "| err:", d.get("c_abi_path_error"))
    print("  roofline:", {k: d["roofline"][k] for k in ("kernel", "achieved", "peak", "frac")})
    print("  C3:", {k: v for k, v in (d.get("c3_one_column") or {}).items() if k != "workload"})
    print("  C5:", {k: v for k, v in (d.get("c5_fri") or {}).items() if k != "workload"})
    c = d.get("cairo_prove", {})
    for k in ("fib", "fib_large"):
        if k in c:
            tp = c[k].get("torch_distributed_path", {})
            print(" ", k, c[k]["program"], "%.2f ms (C ABI)" % c[k]["value"], c[k].get("parity_ok"), c[k].get("stages_ms_rank0_host_marks"))
            print("      torch.distributed path: %.2f ms" % tp.get("value", -1), tp.get("stages_ms_synchronised"), tp.get("commit_detail_ms"))
    if "error" in c:
        print("  cairo error:", c["error"])
except Exception as e:
    print(f, "unreadable bench line:", e)
PY
