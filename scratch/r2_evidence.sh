#!/bin/bash
# round-2 evidence on ONE GPU: bench line (+ reference arm), launch lists with DRAM bytes, full captures of the kernels
# that changed this round.  usage: r2_evidence.sh TAG
TAG=${1:-r2a}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc $?"; cut -c1-400 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc $?"; cut -c1-300 gpurun_out/bench_ref_$TAG.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu bench launches rc $?"
PROF_WARM=0 timeout 600 ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/launches_solve3d_512_$TAG.csv python scratch/prof_mg.py 512 500000 > gpurun_out/ncu_launch3d_$TAG.log 2>&1; echo "ncu 3d launches rc $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_stencil_tma|k_cg_dir_mg|k_cg_update_mg' -s 3 -c 3 -o gpurun_out/prof_grid_$TAG -f python scratch/prof_mg.py 512 3 > gpurun_out/ncu_grid_$TAG.log 2>&1; echo "ncu grid full rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_batch2d_mg -c 1 -s 1 -o gpurun_out/prof_bmg_$TAG -f python scratch/prof_batch.py 10000 > gpurun_out/ncu_bmg_$TAG.log 2>&1; echo "ncu batch full rc $?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ls -la gpurun_out/*$TAG*
