#!/bin/bash
# ncu evidence for the tree backward kernels: DFS-range (k3d_tree_bwd) with --set full, level-synchronous (launch times)
mkdir -p gpurun_out
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'tree_bwd|tree_fwd|k3_elem' -s 8 -c 4 -f -o $O/r02_bwd_dfs python tools/ec_probe.py --reps 2 --steps 3 > $O/r02_bwd_dfs.log 2>&1
POLEE_TREE_BWD=levels ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tree_bwd|tree_fwd|k3_elem' -s 8 -c 8 --csv --log-file $O/r02_bwd_levels_launches.csv python tools/ec_probe.py --reps 20 --steps 3 > $O/r02_bwd_levels.log 2>&1
tail -3 $O/r02_bwd_dfs.log; tail -3 $O/r02_bwd_levels.log; grep -v "^==" $O/r02_bwd_levels_launches.csv | cut -d, -f5,15- | cut -c1-160
