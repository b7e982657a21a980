#!/bin/bash
# round 2, final: ncu --set full of the NTT passes as they ship (72 wide multiplies per field multiplication, warp barriers)
mkdir -p gpurun_out
timeout 200 ncu --set full --import-source on --clock-control none -k regex:ntt_pass -s 12 -c 2 -o gpurun_out/r2u_ntt_full -f \
    python bench.py --steps 1 --warmup 3 --no-cairo --no-cpu-baseline --no-c4 > gpurun_out/r2u_ncu_full.log 2>&1
tail -2 gpurun_out/r2u_ncu_full.log | cut -c1-200
ncu -i gpurun_out/r2u_ntt_full.ncu-rep --page raw --csv > gpurun_out/r2u_ntt_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2u_ntt_full.ncu-rep --page details --csv > gpurun_out/r2u_ntt_full_details.csv 2>/dev/null
ls -la gpurun_out/r2u_*
