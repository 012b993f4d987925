#!/bin/bash
# On the GPU box: launch list of one C5 dist call through the inverted index, and ncu --set full captures of its counting
# kernel and of the tile radix sort's scatter.  Outputs land in gpurun_out/; summarise with profiles/summarize_launches.py /
# ncu_metrics.py.
mkdir -p gpurun_out
C5_ITERS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_dist_launches.csv \
    python tools/c5_once.py > gpurun_out/r02_dist_launches.log 2>&1
C5_ITERS=1 ncu --set full --clock-control none --import-source on -k regex:"dist_inverted_kernel|radix_tile_scatter_kernel|postings_runs" -c 4 -f \
    -o gpurun_out/r02_ncu_dist python tools/c5_once.py > gpurun_out/r02_ncu_dist.log 2>&1
ls -la gpurun_out/r02_ncu_dist.ncu-rep
