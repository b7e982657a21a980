#!/bin/bash
# round 2, session 2: corrected integer-pipe probes (data-dependent multiplicands) + the block-mode DEEP test
mkdir -p gpurun_out
python tools/microbench.py > gpurun_out/r2f_int_peaks.json 2> gpurun_out/r2f_int_peaks.err; cat gpurun_out/r2f_int_peaks.json; tail -3 gpurun_out/r2f_int_peaks.err
timeout 600 python -m pytest tests/test_gpu_deep.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r2f_pytest_deep.log
