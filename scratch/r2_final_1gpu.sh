#!/bin/bash
# final one-GPU record of the round: every GPU test, the bench line with its CPU legs, the reference arm, smoke,
# launch list of one 512^3 solve and a full capture of the batch kernel.  usage: r2_final_1gpu.sh TAG
TAG=${1:-r2d}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/gpu_tests_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc $?"; cut -c1-300 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc $?"; cut -c1-200 gpurun_out/bench_ref_$TAG.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
PROF_WARM=0 timeout 600 ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/launches_solve3d_512_$TAG.csv python scratch/prof_mg.py 512 500000 > gpurun_out/ncu_launch3d_$TAG.log 2>&1; echo "ncu 3d launches rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_batch2d_mg -c 1 -s 1 -o gpurun_out/prof_bmg_$TAG -f python scratch/prof_batch.py 10000 > gpurun_out/ncu_bmg_$TAG.log 2>&1; echo "ncu batch full rc $?"
