#!/bin/bash
# round 2, GPU session O: final single-GPU captures for profiles/
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err
echo "bench rc=$?"; tail -2 gpurun_out/r2o_bench_n1.err; cut -c1-600 gpurun_out/r2o_bench_n1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_link_pcg2 -s 2 -c 1 -f -o gpurun_out/prof_link_pcg2_r2o python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-small > /dev/null 2> gpurun_out/r2o_ncu.err
echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2o.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-small > /dev/null 2>&1
echo "ncu launches rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2o_ref.json 2> gpurun_out/r2o_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --config 3 --steps 20 --warmup 3 --no-cpu --no-small > gpurun_out/r2o_bench_c3.json 2>/dev/null; echo "c3 rc=$?"
