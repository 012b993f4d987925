#!/bin/bash
# ncu --set full capture of two steady-state pack_kernel launches of the C2 resident path
R=${1:-10000000}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"pack_kernel|parse_fused" -s 30 -c 2 -f -o gpurun_out/r02_ncu_pack \
    python tools/c2_once.py $R > gpurun_out/r02_ncu_pack.log 2>&1
