#!/bin/bash
# round 2, GPU session U: final single-GPU record: whole GPU suite, smoke, bench (both arms), launch list, config 3
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SECONDS=0
timeout 1300 python -m pytest tests -q -m gpu > gpurun_out/r2u_pytest.log 2>&1; echo "pytest -m gpu rc=$? in ${SECONDS}s: $(tail -1 gpurun_out/r2u_pytest.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err
echo "bench rc=$?"; tail -2 gpurun_out/r2u_bench_n1.err; cut -c1-400 gpurun_out/r2u_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2u_ref.json 2> gpurun_out/r2u_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2u_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2u.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-small > /dev/null 2>&1
echo "ncu launches rc=$?"
timeout 300 python bench.py --config 3 --steps 20 --warmup 3 --no-cpu --no-small > gpurun_out/r2u_bench_c3.json 2>/dev/null; echo "c3 rc=$?"; cut -c1-300 gpurun_out/r2u_bench_c3.json
