#!/bin/bash
# round 2, GPU session S (final) (8 GPUs): scaling series with state-hash equality, config 5 with / without device-side LPT
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2s_topo.txt 2>&1
run() { # name N extra...
  name=$1; N=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu --no-small "$@" > gpurun_out/r2s_$name.json 2> gpurun_out/r2s_$name.err
  echo "$name rc=$?"; tail -1 gpurun_out/r2s_$name.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2s_$name.json"))
    b=d.get("balance") or {}
    print("$name phases", [ {k: round(v,3) for k,v in r.items()} for r in d["phases_ms_per_sweep"]["per_rank"]][:2])
    print("$name it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "link_ms", round(d["roofline"]["kernel_ms"],3), "e2e", round(d["e2e"]["value"],2) if d.get("e2e") else None, "pcg1", round(d["other_samplers"]["PCG-I"]["value"],1) if d.get("other_samplers") else None, "hash", d["state_hash"], "bal", round(b.get("max_over_mean_link_ms",0),4), b.get("exchange"), "trace", (d.get("trace") or {}).get("max_over_mean_link_ms"))
except Exception as e: print("$name failed", e)
PY
}
run n8 8
run n4 4
run c5_n8_lpt 8 --config 5 --trace-sweeps 100 --no-e2e
