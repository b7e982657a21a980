for g in 1 3 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/c4_sharded.py 22 33 8 $g 2>&1 | grep "^{" | tee -a gpurun_out/c4_sharded_r1b.jsonl
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 tools/c4_sharded.py 22 33 8 4 2>&1 | grep "^{" | tee -a gpurun_out/c4_sharded_r1b.jsonl
