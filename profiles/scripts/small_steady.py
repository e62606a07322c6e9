"""RLdata10000 / 4 blocks in steady state: 300 sweeps, then a few eager sweeps for a per-kernel launch list
(run under ncu with --launch-skip), and graph vs eager rates"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from dblink_b200 import config
from dblink_b200.project import Project
from test_host_pipeline import GOLDEN, make_conf
sampler = sys.argv[1]
conf = make_conf(os.path.join(GOLDEN, "RLdata10000.csv.gz"), "/tmp/x/", 2, '["fname_c1", "lname_c1"]')
conf = conf.replace("lowDistortion : {alpha : 0.5, beta : 50.0}", "lowDistortion : {alpha : 10.0, beta : 1000.0}")
proj = Project(config.parse_string(conf), base_dir="")
eng = proj.generate_initial_state()
eng.set_graph_mode(2)
eng.sweep(sampler, 300)
s = eng.summary()
print(sampler, "after 300: isolates", s["num_isolates"], "pairs/sweep", None)
for mode in (2, 1):
    eng.set_graph_mode(mode)
    eng.sweep(sampler, 200)
    print(sampler, "mode", mode, "it/s", 1e3 * 200 / eng.last_sweep_ms(), "phases", eng.phase_ms())
eng.set_graph_mode(1)
eng.sweep(sampler, 3)   # the launches ncu lists (everything before is skipped)
