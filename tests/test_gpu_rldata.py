"""BASELINE.json configs[0] and configs[1] on the GPU: RLdata500 (1 block) through the full config -> sample ->
summarize -> evaluate pipeline, and RLdata10000 with 4 k-d-tree blocks, each also bit-checked against the oracle."""
import collections
import os

import numpy as np
import pytest

from test_host_pipeline import GOLDEN, make_conf

pytestmark = pytest.mark.gpu


def oracle_state(O, proj, levels, attr_ids):
    d = proj.load()
    names = [a.name for a in proj.matching_attributes]
    from dblink_b200.records import read_csv

    _, _, values, _ = read_csv(proj.data_path, proj.rec_id_attribute, names, None, None, proj.null_value)
    idx = []
    for a, attr in enumerate(proj.matching_attributes):
        cnt = collections.Counter(v[a] for v in values if v[a] is not None)
        idx.append(O.Index.build({k: float(c) for k, c in cnt.items()}, attr.is_constant, attr.similarity_fn.threshold,
                                 attr.similarity_fn.max_similarity, proj.expected_max_cluster_size))
    alpha = [a.alpha for a in proj.matching_attributes]
    beta = [a.beta for a in proj.matching_attributes]
    m0 = O.Model(idx, alpha, beta, None, proj.random_seed, 1)
    s0 = O.State.init(m0, d["x"], d["file"])
    tree = O.KDTree.fit(s0.y, levels, attr_ids)
    m = O.Model(idx, alpha, beta, tree, proj.random_seed, 1)
    st = O.State.from_arrays(m, d["x"], d["file"], s0.z, s0.link, s0.y, s0.theta, 0)
    st._keep = (m0, s0, tree, idx)
    return st


def test_rldata500_pipeline(oracle, tmp_path):
    from dblink_b200 import config
    from dblink_b200.project import Project

    out = str(tmp_path) + "/"
    conf = make_conf(os.path.join(GOLDEN, "RLdata500.csv.gz"), out, 0, "[]", sample_size=100, thinning=10,
                     sampler="PCG-I", cutoff=100)
    proj = Project(config.parse_string(conf), base_dir="")
    res = proj.execute(log=lambda *a: None)
    # outputs of the reference's steps exist and have its formats
    for f in ("linkage-chain.parquet", "diagnostics.csv", "cluster-size-distribution.csv", "partition-sizes.csv",
              "shared-most-probable-clusters.csv", "evaluation-results.txt"):
        assert os.path.exists(os.path.join(out, f)), f
    diag = open(os.path.join(out, "diagnostics.csv")).read().splitlines()
    assert len(diag) == 1 + 101  # initial state + 100 samples (Sampler.scala:84-89)
    assert diag[0].startswith("iteration,systemTime-ms,numObservedEntities,logLikelihood,popSize,aggDist-by")
    assert [int(r.split(",")[0]) for r in diag[1:]] == list(range(0, 1001, 10))
    # posterior quality against the ground truth (ent_id): 450 entities, 50 duplicate pairs
    pw = res["pairwise"]
    assert pw["precision"] > 0.8 and pw["recall"] > 0.75 and pw["f1score"] > 0.8, pw
    assert res["cluster"] > 0.8
    csd = open(os.path.join(out, "cluster-size-distribution.csv")).read().splitlines()
    last = dict(zip(csd[0].split(",")[1:], map(int, csd[-1].split(",")[1:])))
    assert abs(last["1"] - 400) <= 15 and abs(last["2"] - 50) <= 10
    # the chain itself is the oracle's chain, bit for bit
    eng = proj.generate_initial_state()
    st = oracle_state(oracle, proj, 0, [])
    for _ in range(4):
        eng.sweep("PCG-I", 25)
        st.sweep(oracle.PCG_I, 25)
        d = eng.download_state()
        assert np.array_equal(d["link"], st.link) and np.array_equal(d["y"], st.y) and np.array_equal(d["z"], st.z)
        assert np.array_equal(d["theta"], st.theta)


@pytest.mark.parametrize("sampler", ["PCG-I", "PCG-II"])
def test_rldata10000_four_blocks(oracle, sampler, tmp_path):
    from dblink_b200 import analysis, config, writers
    from dblink_b200.project import Project

    conf = make_conf(os.path.join(GOLDEN, "RLdata10000.csv.gz"), str(tmp_path) + "/", 2, '["fname_c1", "lname_c1"]')
    conf = conf.replace("lowDistortion : {alpha : 0.5, beta : 50.0}", "lowDistortion : {alpha : 10.0, beta : 1000.0}")
    proj = Project(config.parse_string(conf), base_dir="")
    eng = proj.generate_initial_state()
    assert eng.num_partitions == 4 and eng.num_records == 10000
    st = oracle_state(oracle, proj, 2, [3, 4])
    for _ in range(3):
        eng.sweep(sampler, 1)
        st.sweep(oracle.SAMPLERS[sampler])
        d = eng.download_state()
        for k in ("link", "y", "z", "theta", "block"):
            assert np.array_equal(d[k], getattr(st, k)), k
    eng.sweep(sampler, 150)
    link, blk = eng.links()
    rec_ids = proj.load()["rec_ids"]
    parts = writers.linkage_structure(link, blk, rec_ids)
    clusters = [frozenset(c) for cl in parts.values() for c in cl]
    pw = analysis.pairwise_metrics(clusters, proj.true_clusters())
    # one early sample (not the sMPC point estimate): true links are being found and most links are right-ish
    assert pw["TP"] >= 100 and pw["precision"] > 0.25, pw
    sizes = np.bincount(np.bincount(link, minlength=eng.num_entities))
    assert sizes[1] > 6000, sizes  # most records are singletons (8000 of 10000 in the ground truth)


def test_save_and_resume_continues_the_same_chain(tmp_path):
    """State.save / State.read (State.scala:122-193): a chain stopped and resumed equals the uninterrupted chain,
    and the writers append (Sampler.scala:71,79-81)."""
    from dblink_b200 import config, state_io
    from dblink_b200.project import Project

    data = os.path.join(GOLDEN, "RLdata500.csv.gz")
    outa, outb = str(tmp_path) + "/a/", str(tmp_path) + "/b/"
    def mk(out, n, resume):
        conf = make_conf(data, out, 0, "[]", sample_size=n, thinning=5, cutoff=0)
        return Project(config.parse_string(conf.replace("resume : false", "resume : %s" % resume)), base_dir="")

    pa = mk(outa, 8, "false")
    pa.execute(log=lambda *a: None)      # 40 sweeps in one go
    pb = mk(outb, 4, "true")
    pb.execute(log=lambda *a: None)      # 20 sweeps ...
    assert state_io.saved_state_exists(outb)
    pb2 = mk(outb, 4, "true")
    pb2.execute(log=lambda *a: None)     # ... resumed for 20 more
    sa, sb = state_io.load_state(outa), state_io.load_state(outb)
    assert sa["iteration"] == sb["iteration"] == 40
    for k in ("theta", "z", "link", "y"):
        assert np.array_equal(sa[k], sb[k]), k
    da = open(outa + "diagnostics.csv").read().splitlines()
    db = open(outb + "diagnostics.csv").read().splitlines()
    strip = lambda rows: [",".join(r.split(",")[:1] + r.split(",")[2:]) for r in rows]  # noqa: E731  (drop systemTime)
    assert strip(da) == strip(db)
