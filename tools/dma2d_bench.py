"""Host -> device throughput of strided 2-D DMA (cudaMemcpy2DAsync from pinned memory) as a function of the run width:
what a column group of a row-major table costs on the way to the GPU (s252_interpolate_and_commit with S252_HOST)."""
import json
import sys

import torch
from cuda.bindings import runtime as rt

rows, pitch_elems = 1 << 19, 34
host = torch.empty((rows, pitch_elems, 4), dtype=torch.int64, pin_memory=True)
host.random_()
dev = torch.empty((rows, pitch_elems, 4), dtype=torch.int64, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
out = []
for cols in (1, 2, 4, 6, 8, 11, 17, 34):
    width = cols * 32
    best = 1e9
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        (err,) = rt.cudaMemcpy2DAsync(dev.data_ptr(), width, host.data_ptr(), pitch_elems * 32, width, rows,
                                      rt.cudaMemcpyKind.cudaMemcpyHostToDevice, stream)
        assert int(err) == 0, err
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1))
    out.append({"run_bytes": width, "rows": rows, "ms": round(best, 3), "GBps": round(rows * width / best / 1e6, 2)})
    print(out[-1], file=sys.stderr)
print(json.dumps({"pitch_bytes": pitch_elems * 32, "copies": out}))
