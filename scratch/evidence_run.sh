#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_rNN.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/gpu_tests_rNN.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_rNN.json 2> gpurun_out/bench_rNN.err; echo "bench rc $?"; cut -c1-900 gpurun_out/bench_rNN.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_rNN.json 2> gpurun_out/bench_ref_rNN.err; echo "ref rc $?"; cut -c1-200 gpurun_out/bench_ref_rNN.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_rNN.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch_rNN.log 2>&1; echo "ncu launches rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_batch2d_mg -c 1 -s 1 -o gpurun_out/prof_bmg_rNN -f python scratch/prof_batch.py 10000 > gpurun_out/ncu_bmg_rNN.log 2>&1; echo "ncu full rc $?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
