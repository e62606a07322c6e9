#!/bin/bash
# round 2, GPU session Q: link kernel warps-per-CTA A/B, small-config steady state
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in base w8 w20 w16s8; do
  if [ $v = base ]; then unset DBL_LIB; else export DBL_LIB=$GRAFT_REPO_ROOT/exp/lib_$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "chain_from_init or model_shapes or random_states or tiny or many_tiles" > gpurun_out/r2q_par_$v.log 2>&1; echo "$v parity rc=$? $(tail -1 gpurun_out/r2q_par_$v.log)"
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e --no-small > gpurun_out/r2q_bench_$v.json 2> gpurun_out/r2q_bench_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2q_bench_$v.json"))
print("$v", "it/s", round(d["value"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "pcg1", round(d["other_samplers"]["PCG-I"]["value"],1), d["state_hash"], d["phases_ms_per_sweep"]["per_rank"])
PY
done
unset DBL_LIB
for s in PCG-I PCG-II; do
timeout 300 python exp/small_steady.py $s 2>&1 | tail -4
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 14000 -c 200 --csv --log-file gpurun_out/launches_small_steady_r2q_PCG-I.csv python exp/small_steady.py PCG-I > gpurun_out/r2q_small_ncu.log 2>&1
echo "small steady ncu rc=$?"
