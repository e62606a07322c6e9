"""The oracle against every golden vector / known-answer test the reference holds for this path
(SURVEY.md section 8c), plus the published Philox4x32-10 known-answer vectors.  CPU only."""
import math

import numpy as np
import pytest

STATE_WEIGHTS = {  # AttributeIndexTest.scala:38-41
    "Australian Capital Territory": 0.410, "New South Wales": 7.86, "Northern Territory": 0.246, "Queensland": 4.92,
    "South Australia": 1.72, "Tasmania": 0.520, "Victoria": 6.32, "Western Australia": 2.58,
}
STATE_SIM_NORMS = {  # AttributeIndexTest.scala:48-55
    "Australian Capital Territory": 0.0027140755302269004, "New South Wales": 1.4193905286944585e-4,
    "Northern Territory": 0.00451528932619675, "Queensland": 2.2673706056780077e-4,
    "South Australia": 6.465919296781136e-4, "Tasmania": 0.00214117348291189, "Victoria": 1.7651936247903708e-4,
    "Western Australia": 4.317863538883541e-4,
}


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    u0, u1 = oracle.uniform2(123, 2, 3, 4, 5)
    assert 0.0 < u0 < 1.0 and 0.0 < u1 < 1.0


def test_similarity_fn(oracle):
    """SimilarityFnTest.scala:46-74"""
    s = lambda a, b, t=5.0: oracle.lev_similarity(a, b, t, 10.0)  # noqa: E731
    assert s("John Smith", "John Smith") == 10.0
    assert s("", "") == 10.0
    assert s("", "John Smith") == 0.0
    assert s("Jane Smith", "John Smith") == s("John Smith", "Jane Smith")
    assert s("AB", "BB") == 2.0
    assert s("AB", "BB", 0.0) == 6.0
    assert oracle.levenshtein("kitten", "sitting") == 3


@pytest.fixture(scope="module")
def state_index(oracle):
    return oracle.Index.build(STATE_WEIGHTS, False, 5.0, 10.0)


def test_attribute_index_generic(oracle, state_index):
    """AttributeIndexBehaviors.scala:7-50 for both kinds of index"""
    const = oracle.Index.build(STATE_WEIGHTS, True)
    total = sum(STATE_WEIGHTS.values())
    for ix in (const, state_index):
        assert ix.V == len(STATE_WEIGHTS)
        ids = {ix.value_id(k) for k in STATE_WEIGHTS}
        assert ids == set(range(len(STATE_WEIGHTS)))
        assert [ix.value(i) for i in range(ix.V)] == sorted(STATE_WEIGHTS)  # ids in sorted-string order
        for k, w in STATE_WEIGHTS.items():
            assert ix.phi[ix.value_id(k)] == pytest.approx(w / total, abs=1e-4)
        with pytest.raises(IndexError):
            ix.exp_sim_of(ix.V + 1, 0)
        with pytest.raises(IndexError):
            ix.exp_sim_of(0, ix.V + 1)
        with pytest.raises(IndexError):
            ix.sim_values_of(ix.V + 1)
    assert np.all(const.norm == 1.0)                       # AttributeIndexTest.scala:62-64
    assert const.nnz == 0                                  # :66-68
    assert all(const.exp_sim_of(i, j) == 1.0 for i in range(8) for j in range(8))  # :70-74


def test_attribute_index_golden(state_index):
    """AttributeIndexTest.scala:78-99, to full double precision (the reference asserts 1e-4)"""
    ix = state_index
    for k, n in STATE_SIM_NORMS.items():
        assert ix.norm[ix.value_id(k)] == pytest.approx(n, rel=1e-13)
    sa, wa = ix.value_id("South Australia"), ix.value_id("Western Australia")
    sv = ix.sim_values_of(sa)
    assert set(sv) == {7, 4}
    assert sv[7] == pytest.approx(39.813678188084864, rel=1e-13)
    assert sv[4] == pytest.approx(22026.465794806718, rel=1e-13)
    assert ix.exp_sim_of(sa, wa) == pytest.approx(39.813678188084864, rel=1e-13)
    assert ix.exp_sim_of(ix.value_id("Victoria"), ix.value_id("Tasmania")) == 1.0


def test_distortion_probs_init(oracle):
    """DistortionProbsTest.scala:24-33: theta starts at the prior mean alpha/(alpha+beta) for every file"""
    ix = oracle.Index.build({"a": 1.0, "b": 2.0}, True)
    m = oracle.Model([ix], [3.0], [3.0], None, 1, F=2)
    st = oracle.State.init(m, np.array([[0], [1], [1]], np.int32), np.array([0, 1, 1], np.int32))
    assert st.theta.tolist() == [[0.5, 0.5]]


def test_alias_sampler_conventions(oracle):
    """AliasSamplerTest.scala:46-62 (rejects negative / NaN / inf) and DiscreteDistBehavior.scala:51-53"""
    for bad in ([0.1, -0.2, 0.3], [0.1, float("nan")], [0.1, float("inf")]):
        with pytest.raises(ValueError):
            oracle.alias_build(bad)
    with pytest.raises(ValueError):
        oracle.alias_build([0.0, 0.0])
    w = np.array([0.5, 0.0, 2.5, 1.0, 0.0])
    prob, alias = oracle.alias_build(w)
    us = (np.arange(20000) + 0.5) / 20000
    draws = np.array([oracle.alias_sample(prob, alias, u) for u in us])
    freq = np.bincount(draws, minlength=5) / len(us)
    assert freq[1] == 0 and freq[4] == 0  # never samples zero-probability values
    np.testing.assert_allclose(freq, w / w.sum(), atol=2e-3)


def test_draw_index_protocol(oracle):
    """the fixed-order inverse-CDF draw: exact distribution, zero-weight exclusion, error convention"""
    rng = np.random.default_rng(0)

    def enumeration_order(n):
        """chunk-major, then lane-major, then step: the order in which the protocol lays out the mass"""
        ntiles = (n + 127) // 128
        nsteps = 4 * ntiles
        spc = 4 * max(1, (ntiles + 31) // 32)
        order = []
        for c0 in range(0, nsteps, spc):
            for lane in range(32):
                for s in range(c0, min(c0 + spc, nsteps)):
                    if s * 32 + lane < n:
                        order.append(s * 32 + lane)
        return np.array(order)

    for n in (1, 5, 32, 33, 100, 1024, 1025, 5000):
        w = rng.random(n) * (rng.random(n) < 0.4)
        if w.sum() == 0:
            w[n // 2] = 1.0
        order = enumeration_order(n)
        assert sorted(order) == list(range(n))
        cdf = np.cumsum(w[order]) / w.sum()
        pos = np.empty(n, int)
        pos[order] = np.arange(n)
        for u in (1e-9, 0.1, 0.37, 0.5, 0.731, 0.999999):
            j, st = oracle.draw_index(w, u)
            assert st == 0 and w[j] > 0
            k = pos[j]
            lo = cdf[k - 1] if k > 0 else 0.0
            assert lo - 1e-12 <= u <= cdf[k] + 1e-12  # the exact inverse CDF over that enumeration
    j, st = oracle.draw_index(np.zeros(40), 0.3)
    assert st == 1  # "zero probability mass" (IndexNonUniformDiscreteDist.scala:78-79)
    j, st = oracle.draw_index(np.array([1.0, float("inf")]), 0.3)
    assert st == 1
    # 0/1 weights: exactly uniform over the positive entries (integer sums are exact in any order)
    w = np.zeros(70)
    w[[3, 40, 41, 69]] = 1.0
    picks = [oracle.draw_index(w, u)[0] for u in (0.1, 0.3, 0.6, 0.9)]
    assert sorted(picks) == [3, 40, 41, 69]


def test_theta_draw_moments(oracle):
    """updateDistProbs GU:305-320: Beta(alpha + n_dist, beta + N - n_dist)"""
    ix = oracle.Index.build({"a": 1.0, "b": 2.0}, True)
    m = oracle.Model([ix, ix], [0.5, 10.0], [50.0, 1000.0], None, 99, F=2)
    agg = np.array([[3, 40], [0, 500]], np.int64)
    fs = np.array([200, 3000], np.int64)
    draws = np.stack([oracle.draw_theta(m, agg, fs, it) for it in range(1, 3001)])
    assert np.all((draws > 0) & (draws < 1))
    for a, (al, be) in enumerate([(0.5, 50.0), (10.0, 1000.0)]):
        for f in range(2):
            s1, s2 = al + agg[a, f], be + fs[f] - agg[a, f]
            mean, var = s1 / (s1 + s2), s1 * s2 / ((s1 + s2) ** 2 * (s1 + s2 + 1))
            assert draws[:, a, f].mean() == pytest.approx(mean, abs=5 * math.sqrt(var / 3000))
            assert draws[:, a, f].var() == pytest.approx(var, rel=0.15)
    # deterministic in (seed, iteration)
    np.testing.assert_array_equal(oracle.draw_theta(m, agg, fs, 7), oracle.draw_theta(m, agg, fs, 7))


def test_theta_draw_is_beta_distributed(oracle):
    """the protocol's Beta draw (Marsaglia-Tsang gammas, polar normals, protocol log / exp) against the exact Beta cdf,
    including the shape < 1 branch (alpha = 0.5 with no distorted value: RLdata500's prior)"""
    from scipy import stats

    ix = oracle.Index.build({"a": 1.0, "b": 2.0}, True)
    m = oracle.Model([ix, ix, ix], [0.5, 0.5, 10.0], [50.0, 2.0, 1000.0], None, 7, F=1)
    agg = np.array([[0], [0], [37]], np.int64)
    fs = np.array([500], np.int64)
    draws = np.stack([oracle.draw_theta(m, agg, fs, it) for it in range(1, 8001)])[:, :, 0]
    for a, (al, be) in enumerate([(0.5, 50.0), (0.5, 2.0), (10.0, 1000.0)]):
        s1, s2 = al + agg[a, 0], be + fs[0] - agg[a, 0]
        ks = stats.kstest(draws[:, a], stats.beta(s1, s2).cdf)
        assert ks.pvalue > 1e-3, (a, ks)


def test_protocol_log_exp_accuracy(oracle):
    """log / exp built from individually rounded IEEE operations (bit-reproducible on CPU and GPU);
    they only have to be GOOD functions: within a few ulp of libm over the ranges the theta draw uses"""
    L = oracle.lib()
    rng = np.random.default_rng(0)
    xs = np.r_[rng.uniform(0, 1, 4000), 10.0 ** rng.uniform(-300, 300, 4000), [1.0, 2.0, 0.5, 1e-310, 5e-324, 1.7976931348623157e308]]
    for x in xs:
        got, ref = L.orc_det_log(float(x)), math.log(x)
        assert abs(got - ref) <= 4 * np.spacing(abs(ref)) + 1e-320, x
    assert L.orc_det_log(1.0) == 0.0
    for x in np.r_[rng.uniform(-745, 709, 6000), rng.uniform(-1, 1, 2000), [0.0]]:
        got, ref = L.orc_det_exp(float(x)), math.exp(x)
        assert abs(got - ref) <= 4 * np.spacing(ref), x
    assert L.orc_det_exp(0.0) == 1.0 and L.orc_det_exp(-800.0) == 0.0
