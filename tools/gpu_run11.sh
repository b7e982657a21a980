#!/bin/bash
# round 2, session 2: the C-ABI sharded commit inside the bench line (N = number of GPUs of the box), small C4 first, then the full line
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
run() {
  if [ "$N" = "1" ]; then python bench.py --mode sharded "$@"; else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; fi
}
run --steps 2 --warmup 3 --no-cairo --c3-log-n 0 --c5-log-n 16 --c4-log-n 14 > gpurun_out/r2j_small_n$N.json 2> gpurun_out/r2j_small_n$N.err
tail -3 gpurun_out/r2j_small_n$N.err
run --steps 3 --warmup 3 --no-cairo --c3-log-n 0 --c5-log-n 0 > gpurun_out/r2j_c4_n$N.json 2> gpurun_out/r2j_c4_n$N.err
tail -3 gpurun_out/r2j_c4_n$N.err
python - <<PY
import json
for f in ("gpurun_out/r2j_small_n$N.json", "gpurun_out/r2j_c4_n$N.json"):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, "\n  ", d.get("call"), "| %.2f ms/step e2e %.2f parity %s launches %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity_ok"], d["gpu_launches"]))
        print("   torch path:", d.get("torch_distributed_path"), "| err:", d.get("c_abi_path_error"))
        print("   c5:", d.get("c5_fri"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
