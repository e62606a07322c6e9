#!/usr/bin/env python
"""BASELINE.json configs[4] (1M records, distortion 0.30, Zipf-1.5 split attributes -> skewed k-d-tree blocks):
convergence traces in the reference's diagnostics.csv format + block-size skew and the LPT placement it leads to.

  python profiles/config5_diagnostics.py profiles/r1c_config5 [sweeps]
writes <prefix>_diagnostics_<sampler>.csv and <prefix>_blocks.json (one B200).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import dblink_b200 as D  # noqa: E402
from dblink_b200 import synth  # noqa: E402
from dblink_b200.distributed import lpt_assign  # noqa: E402
from dblink_b200.writers import DiagnosticsWriter  # noqa: E402


def main():
    prefix = sys.argv[1]
    sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 200

    class Args:
        config, records, levels = 5, 0, -1

    enc, levels, split_attrs, name = bench.workload(Args)
    indexes, x, file, F = synth.build_encoded(enc)
    names = [a.name for a in enc["attributes"]]
    alpha = [a.alpha for a in enc["attributes"]]
    beta = [a.beta for a in enc["attributes"]]
    out = {"workload": name}
    for sampler, n, every in (("PCG-I", sweeps, 5), ("PCG-II", max(10, sweeps // 5), 2)):
        eng = D.GibbsEngine(indexes, alpha, beta, None, 2024, F)
        eng.init_state(x, file)
        part = D.KDTreePartitioner(levels, split_attrs).fit(eng.download_state()["y"])
        eng.set_partitioner(part)
        dw = DiagnosticsWriter(f"{prefix}_diagnostics_{sampler}.csv", names)
        dw.write_row(eng.summary(), eng.num_entities)
        ms = 0.0
        for it in range(n):
            eng.sweep(sampler, 1)
            ms += eng.last_sweep_ms()
            if (it + 1) % every == 0:
                dw.write_row(eng.summary(), eng.num_entities)
        dw.close()
        link, blk = eng.links()
        P = eng.num_partitions
        ent = np.bincount(blk, minlength=P).astype(np.float64)
        rec = np.bincount(blk[link], minlength=P).astype(np.float64)
        cost = ent * rec
        bal = {}
        for w in (2, 4, 8):
            owner = lpt_assign(cost, w)
            load = np.bincount(owner, weights=cost, minlength=w)
            bal[str(w)] = float(load.max() / load.mean())
        out[sampler] = {"sweeps": n, "ms_per_sweep": ms / n, "entities_per_block_min_mean_max":
                        [float(ent.min()), float(ent.mean()), float(ent.max())],
                        "cost_max_over_mean_block": float(cost.max() / cost.mean()),
                        "lpt_max_over_mean_rank_cost": bal}
    json.dump(out, open(f"{prefix}_blocks.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
