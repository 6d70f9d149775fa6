#!/bin/bash
# full captures of the kernels of one V-cycle at 512^3 (voxel-grid passes and the level-1 passes)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2d}
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_mg_fine|k_mg_sweep|k_mg_restrict|k_mg_prolong_add' -s 34 -c 34 -o gpurun_out/prof_cycle_$TAG -f python scratch/prof_mg.py 512 3 > gpurun_out/ncu_cycle_$TAG.log 2>&1; echo "ncu rc $?"; tail -3 gpurun_out/ncu_cycle_$TAG.log
