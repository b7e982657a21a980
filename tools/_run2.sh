python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_cairo_worker.py 100 4 3 3 1 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -25
