#!/bin/bash
# GPU session 3f (2 GPUs): final check after the bounded gap fill: whole GPU suite, PCG-I on 2 GPUs, N = 1 line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r3f_pytest.log 2>&1; echo "pytest -m gpu rc=$? in ${SECONDS}s: $(tail -1 gpurun_out/r3f_pytest.log)"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --sampler PCG-I --gpus 2 --steps 8 --warmup 3 --no-cpu --no-small --no-e2e > gpurun_out/r3f_n2_pcg1.json 2> gpurun_out/r3f_n2_pcg1.err; echo "n2 pcg1 rc=$?"
timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu --no-small > gpurun_out/r3f_n1.json 2> gpurun_out/r3f_n1.err; echo "n1 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r3f_n2_pcg1.json"))
print("n2 PCG-I it/s", round(d["value"],2), d["phases_ms_per_sweep"]["per_rank"], d["state_hash"])
d=json.load(open("gpurun_out/r3f_n1.json"))
print("n1 it/s", round(d["value"],2), "e2e", d["e2e"]["value"], "pcg1", d["other_samplers"]["PCG-I"]["value"], d["state_hash"])
PY
