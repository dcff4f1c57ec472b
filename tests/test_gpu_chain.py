"""Chained SMC sets through abcb200_chain against the oracle's restatement of the reference's per-set sequence
(AbcSmc.cpp:634-664, 1041-1066; AbcLog.cpp:81-124), and the config-1 stand-in: the dice game of examples/reference.json
(2 integer parameters with DiscreteUniform priors on [1, 1000], metrics `sum` (integer) and `sd`, sets of 300 / 500 / 500 particles,
predictive prior = half of each set, MULTIVARIATE noise) replayed as three chained sets: rank -> weights -> MVN proposals -> next set.
The simulator is examples/include/dice.h restated in numpy (its own random stream: the reference's gsl_rng is not available)."""
import numpy as np
import pytest

from abcsmc_b200 import synth

pytestmark = pytest.mark.gpu


def _nrmse(post_mets, observed):                 # ABC::calculate_nrmse, src/AbcUtil.cpp:326-345
    sim = post_mets.mean(axis=0)
    expected = (np.abs(observed) + np.abs(sim)) / 2.0
    expected[sim == observed] = 1.0
    return np.sqrt(np.mean(((sim - observed) / expected) ** 2))


def _check_set(oracle, out, met, par, target, n_pp, prev, priors, filtering, tie_report=None):
    """One set of the chain against the oracle. prev = (theta, w, dv) of the previous set or None. Returns this set's (theta, w, dv)."""
    ref = oracle.particle_ranking_PLS(met, par, target, 0.5) if filtering == 0 else oracle.particle_ranking_simple(met, target)
    order = out["order"].astype(np.int64)
    want = ref["order"][:n_pp].astype(np.int64)
    # the GPU order is a valid ascending order of the oracle's distances (exact ties may be permuted: the reference's std::sort leaves
    # them unspecified, here they come out in ascending particle index)
    assert np.array_equal(ref["dist"][order], ref["dist"][want])
    if tie_report is not None:
        tie_report.append(int(np.sum(order != want)))
    else:
        assert np.array_equal(order, want)
    if filtering == 0:
        assert out["ncomp_used"] == ref["ncomp_used"]
    sel = par[order, :]
    smet = met[order, :]
    np.testing.assert_allclose(out["doubled_variance"], oracle.calculate_doubled_variance(sel), rtol=1e-10)
    if prev is None:
        assert np.all(out["weights"] == 1.0 / n_pp)
    else:
        numer = np.ones(n_pp)
        for p in range(par.shape[1]):
            numer *= np.array([oracle.prior_likelihood(int(priors[0][p]), float(priors[1][p]), float(priors[2][p]), float(v)) for v in sel[:, p]])
        np.testing.assert_allclose(out["weights"], oracle.weight_predictive_prior(numer, sel, prev[0], prev[1], prev[2]), rtol=1e-10)
    np.testing.assert_allclose(out["nrmse"], _nrmse(smet, target), rtol=1e-12)
    np.testing.assert_allclose(out["mean_par"], sel.mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(out["mean_met"], smet.mean(axis=0), rtol=1e-12, atol=1e-13)
    assert np.array_equal(out["median_par"], np.median(sel, axis=0))
    assert np.array_equal(out["median_met"], np.median(smet, axis=0))
    return sel, out["weights"].copy(), out["doubled_variance"].copy()


@pytest.mark.parametrize("filtering", [0, 1])
def test_chain_three_sets_match_oracle(oracle, filtering):
    from abcsmc_b200 import api
    P, K, N, n_pp = 10, 20, 4000, 400
    priors = (np.zeros(P, dtype=np.int32), np.zeros(P), np.ones(P))           # ContinuousUniformPrior(0, 1) for every parameter
    chain = api.SmcChain(P)
    prev, states = None, []
    try:
        for t in range(3):
            par, met, target = synth.make_set(N, P, K, 0xD1CE + 17 * t)
            out = chain.process_set(met, par, target, n_pp, filtering=filtering, priors=priors)
            prev = _check_set(oracle, out, met, par, target, n_pp, prev, priors, filtering)
            assert chain.sets == t + 1
            states.append(chain.state())
            np.testing.assert_array_equal(states[-1][0], prev[0]); np.testing.assert_array_equal(states[-1][1], prev[1])
        # a chain re-seeded from the persisted state of set 1 gives set 2 bit for bit (nothing of the earlier sets is replayed)
        chain2 = api.SmcChain(P)
        try:
            chain2.restore(*states[1], sets_done=2)
            par, met, target = synth.make_set(N, P, K, 0xD1CE + 17 * 2)
            out2 = chain2.process_set(met, par, target, n_pp, filtering=filtering, priors=priors)
            assert np.array_equal(out2["weights"], prev[1]) and np.array_equal(out2["order"], out["order"])
            # host-supplied numerators (custom Parameter classes) take the same path
            chain2.restore(*states[1], sets_done=2)
            out3 = chain2.process_set(met, par, target, n_pp, filtering=filtering, numer_all=np.ones(N))
            np.testing.assert_allclose(out3["weights"], prev[1], rtol=1e-14)
        finally:
            chain2.close()
    finally:
        chain.close()


def _dice(theta, rng):
    """examples/include/dice.h: roll `ndice` dice with `sides` faces; metrics = (sum, sample sd of the rolls; 0 for one die)."""
    met = np.empty((theta.shape[0], 2))
    for i, (nd, sides) in enumerate(theta.astype(np.int64)):
        rolls = rng.integers(1, sides + 1, size=nd)
        met[i, 0] = rolls.sum()
        met[i, 1] = 0.0 if nd == 1 else rolls.std(ddof=1)
    return np.asfortranarray(met)


@pytest.mark.parametrize("stdsort", [False, True])
def test_dice_game_three_chained_sets(oracle, stdsort):
    """Config-1 stand-in (BASELINE.json configs[0]): shapes, priors, noise kind and set sizes of examples/reference.json.
    stdsort: abcb200_set_tie_order(1) - exact ties placed as libstdc++'s std::sort leaves them in PLS::ordered, so every rank position
    of every set is the oracle's (= the reference's), tie groups included."""
    from abcsmc_b200 import api
    ctx = api.get_context(0)
    ctx.set_tie_order(api.TIES_STDSORT if stdsort else api.TIES_BY_INDEX)
    resorts0 = ctx.tie_resorts
    try:
        _dice_game(oracle, api, stdsort)
        assert (ctx.tie_resorts - resorts0 >= 2) if stdsort else (ctx.tie_resorts == resorts0)      # sets 1 and 2 carry duplicates
    finally:
        ctx.set_tie_order(api.TIES_BY_INDEX)


def _dice_game(oracle, api, stdsort):
    rng = np.random.default_rng(20261018)
    sizes, P = [300, 500, 500], 2
    target = np.array([44.0, 2.39925])
    priors = (np.full(P, api.PRIOR_DISCRETE_UNIFORM, dtype=np.int32), np.ones(P), np.full(P, 1000.0))
    lo, hi = np.ones(P), np.full(P, 1000.0)
    chain = api.SmcChain(P)
    prev, ties = None, []
    try:
        theta = np.asfortranarray(rng.integers(1, 1001, size=(sizes[0], P)).astype(np.float64))         # set 0: draws from the prior
        for t, N in enumerate(sizes):
            if t == 1:      # exact ties on purpose: a tenth of the particles repeat another particle's parameters AND metrics
                pass
            met = _dice(theta, rng)
            if t >= 1:
                dup = rng.choice(N, size=N // 10, replace=False); src = rng.choice(N, size=N // 10)
                theta[dup, :] = theta[src, :]; met[dup, :] = met[src, :]
            n_pp = N // 2                                                                                 # predictive_prior_fraction 0.5
            out = chain.process_set(met, theta, target, n_pp, filtering=api.FILTER_PLS, priors=priors)
            prev = _check_set(oracle, out, met, theta, target, n_pp, prev, priors, 0, tie_report=None if stdsort else ties)
            if t + 1 < len(sizes):   # next set: NOISE::MULTIVARIATE proposals from the predictive prior just built (AbcSmc.cpp:491-503)
                L = api.setup_mvn_sampler(prev[0])
                np.testing.assert_allclose(L, oracle.setup_mvn_sampler(prev[0]), rtol=1e-10, atol=1e-12)
                theta = api.sample_mvn_predictive_priors(1000 + t, sizes[t + 1], prev[1], prev[0], L, lo, hi, integral=np.ones(P, dtype=np.int32))
                assert theta.min() >= 1 and theta.max() <= 1000 and np.array_equal(theta, np.round(theta))
                theta = np.asfortranarray(theta)
        # the posterior contracts towards the truth (13 dice, 8 sides) as the sets go on
        assert chain.sets == 3
        print(f"dice game: rank positions that differ from libstdc++ std::sort inside exact-tie groups, per set: {ties} of {[s // 2 for s in sizes]}")
    finally:
        chain.close()


def test_tie_order_stdsort_host_and_device_entry_points(oracle):
    """abcb200_set_tie_order(1) through abcb200_rank_pls (host buffers), abcb200_rank_pls_dev / abcb200_rank_simple_dev (device buffers):
    duplicated particles, full order and a cut that falls inside a tie group; continuous data is left alone."""
    import torch
    from abcsmc_b200 import api, device as dev
    ctx = api.get_context(0)
    cfg = synth.make_config("C2", scale=0.02)
    met, par, target, N = cfg["metrics"].copy(), cfg["params"].copy(), cfg["target"], cfg["N"]
    rng = np.random.default_rng(5)
    base = oracle.particle_ranking_PLS(met, par, target, 0.5)
    # copies of particles that sit around rank 100 (hold-out rows overwritten by hold-out rows: the fit's training rows are untouched)
    n_tr = int(round(N * 0.5))
    near = [int(i) for i in base["order"][60:140] if i >= n_tr][:12]
    spare = [int(i) for i in base["order"][-400:] if i >= n_tr][:36]
    for j, dst in enumerate(spare):
        met[dst, :] = met[near[j % len(near)], :]; par[dst, :] = par[near[j % len(near)], :]
    met, par = np.asfortranarray(met), np.asfortranarray(par)
    ref = oracle.particle_ranking_PLS(met, par, target, 0.5)
    groups = np.flatnonzero(np.diff(ref["dist"][ref["order"].astype(np.int64)]) == 0)
    assert groups.size >= 30
    cut = int(groups[groups.size // 2]) + 1                                    # top_n ends between two equal distances
    dv0 = torch.device("cuda", 0)
    dev.use_torch_stream(ctx)
    d_met, d_par, d_t = dev.host_to_colmajor_tensor(met, dv0), dev.host_to_colmajor_tensor(par, dv0), torch.from_numpy(np.ascontiguousarray(target)).to(dv0)
    ctx.set_tie_order(api.TIES_STDSORT)
    try:
        r0, expected = ctx.tie_resorts, 0
        ds = ref["dist"][ref["order"].astype(np.int64)]
        for top_n in (N, cut, 40):
            expected += 2 * int(np.any(np.diff(ds[:top_n]) == 0) or (top_n < N and ds[top_n] == ds[top_n - 1]))      # host + device entry point
            want = ref["order"][:top_n].astype(np.int64)
            got = api.particle_ranking_PLS(met, par, target, 0.5, top_n=top_n).astype(np.int64)
            assert np.array_equal(got, want), top_n
            o_dev = dev.rank_pls(ctx, d_met, d_par, d_t, 0.5, top_n=top_n)[0]
            assert np.array_equal(o_dev.cpu().numpy().astype(np.int64), want), top_n
        assert ctx.tie_resorts - r0 == expected and expected >= 4              # N and cut at least; a host pass only where ties reach the output
        rs = oracle.particle_ranking_simple(met, target)
        assert np.array_equal(api.particle_ranking_simple(met, par, target).astype(np.int64), rs["order"].astype(np.int64))
        assert np.array_equal(dev.rank_simple(ctx, d_met, d_t, top_n=cut)[0].cpu().numpy().astype(np.int64), rs["order"][:cut].astype(np.int64))
        v = np.sqrt(rng.integers(0, 40, 3000).astype(np.float64))               # PLS::ordered itself (abcb200_ordered / _top)
        assert np.array_equal(api.ordered(v).astype(np.int64), oracle.ordered(v).astype(np.int64))
        assert np.array_equal(api.ordered(v, top_n=300).astype(np.int64), oracle.ordered(v)[:300].astype(np.int64))
        # continuous data: nothing to re-derive, same answer as the default mode
        r1 = ctx.tie_resorts
        a = api.particle_ranking_PLS(cfg["metrics"], cfg["params"], target, 0.5, top_n=cfg["N_pp"])
        assert ctx.tie_resorts == r1 and np.array_equal(a.astype(np.int64), base["order"][:cfg["N_pp"]].astype(np.int64))
    finally:
        ctx.set_tie_order(api.TIES_BY_INDEX)
    # default mode on the tied data: same distances rank by rank, ties in ascending particle index
    got = api.particle_ranking_PLS(met, par, target, 0.5, top_n=N).astype(np.int64)
    assert np.array_equal(ref["dist"][got], ref["dist"][ref["order"].astype(np.int64)])
    d = ref["dist"][got]
    same = np.flatnonzero(np.diff(d) == 0)
    assert np.all(got[same] < got[same + 1])
