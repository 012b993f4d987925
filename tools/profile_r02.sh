#!/bin/bash
# On the GPU box: (1) launch list of the C2 resident path, (2) ncu --set full captures of the hash and pack kernels.
# Outputs land in gpurun_out/; summarise here with profiles/summarize_launches.py / ncu_metrics.py / make_ncu_hash_json.py.
R=${1:-10000000}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
    python tools/c2_once.py $R > gpurun_out/r02_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hash_kernel -s 30 -c 2 -f -o gpurun_out/r02_ncu_hash \
    python tools/c2_once.py $R > gpurun_out/r02_ncu_hash.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pack_kernel|parse_fused" -s 30 -c 2 -f -o gpurun_out/r02_ncu_pack \
    python tools/c2_once.py $R > gpurun_out/r02_ncu_pack.log 2>&1
ls -la gpurun_out/*.ncu-rep
