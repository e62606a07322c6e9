"""Do PCG-I and PCG-II reach the same posterior on RLdata500?  (CPU oracle; output: profiles/r2_sampler_agreement_oracle.txt; usage: python profiles/sampler_agreement_oracle.py 6000 3000)"""
import collections, gzip, csv, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O
from dblink_b200 import config
from dblink_b200.project import Project
from test_host_pipeline import GOLDEN, make_conf
from test_gpu_rldata import oracle_state

def pair_f1(link, truth):
    # pairwise F1 of one sample: pairs linked together vs truth
    def pairs(lab):
        _, inv = np.unique(lab, return_inverse=True)
        c = np.bincount(inv)
        return (c * (c - 1) // 2).sum()
    both = np.unique(np.stack([link, truth], 1), axis=0, return_counts=True)[1]
    tp = (both * (both - 1) // 2).sum()
    pp, tpairs = pairs(link), pairs(truth)
    prec = tp / pp if pp else 1.0
    rec = tp / tpairs
    return 2 * prec * rec / (prec + rec) if prec + rec else 0.0

n_sweeps, burn, thin = int(sys.argv[1]), int(sys.argv[2]), 10
for sampler in ("PCG-I", "PCG-II"):
    for seed in (319158, 1, 2, 3, 4):
        conf = make_conf(os.path.join(GOLDEN, "RLdata500.csv.gz"), "/tmp/x/", 0, "[]").replace("randomSeed : 319158", f"randomSeed : {seed}")
        proj = Project(config.parse_string(conf), base_dir="")
        st = oracle_state(O, proj, 0, [])
        _, truth = np.unique(np.array(proj.load()["ent_ids"]), return_inverse=True)
        t = time.time(); nobs, n2, f1 = [], [], []
        for it in range(n_sweeps):
            st.sweep(O.SAMPLERS[sampler])
            if it >= burn and it % thin == 0:
                link = st.link
                c = np.bincount(np.bincount(link, minlength=500))
                nobs.append(500 - c[0]); n2.append(c[2] if len(c) > 2 else 0); f1.append(pair_f1(link, truth))
        print(f"{sampler} seed={seed}: numObserved {np.mean(nobs):.2f}+-{np.std(nobs):.2f}  size2 {np.mean(n2):.2f}  sampleF1 {np.mean(f1):.3f}  ({time.time()-t:.0f}s)", flush=True)
