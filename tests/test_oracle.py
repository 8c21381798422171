"""Pin the numpy oracle (no GPU needed): known-answer vectors, closed forms,
scipy cross-checks and the reference's one in-tree formula.  The reference has
no tests for this path, so these are the pins there are."""
import numpy as np
import pytest
import scipy.linalg

from oracle import gp_numpy as o
from oracle import philox, picks
from tests.util import problem, rel_err


def test_philox_known_answer_vectors():
    """Random123 kat_vectors for philox4x32-10."""
    def h(t):
        return ['%08x' % int(x) for x in t]
    assert h(philox.philox4x32_10(0, 0, 0, 0, 0, 0)) == ['6627e8d5', 'e169c58d', 'bc57ac4c', '9b00dbd8']
    f = 0xffffffff
    assert h(philox.philox4x32_10(f, f, f, f, f, f)) == ['408f276d', '41c83b0e', 'a20bc7c6', '6d5451fd']
    assert h(philox.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344,
                                  0xa4093822, 0x299f31d0)) == ['d16cfe09', '94fdcceb', '5001e420', '24126ea1']


def test_philox_normals_are_trial_indexed_and_standard():
    a = philox.normals(100826730, 7, 512, 64)
    b = philox.normals(100826730, 7, 512, 64)
    c = philox.normals(100826730, 8, 512, 64)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(a.mean()) < 0.02 and abs(a.std() - 1.0) < 0.02
    # prefix property: the first rows do not depend on m
    assert np.array_equal(philox.normals(1, 2, 10, 64), philox.normals(1, 2, 300, 64)[:10])


def test_prior_posterior_closed_forms():
    D = 2
    kern = o.make_kernel('se', D, 2.0, [0.3, 0.4])
    Xs = np.random.RandomState(0).uniform(size=(20, D))
    g0 = o.make_gp(np.zeros((0, D)), [], 'se', 2.0, [0.3, 0.4], 0.05, 0.7)
    mu, cov = g0.eval(Xs, include_covar=True)
    assert np.allclose(mu, 0.7) and np.allclose(cov, kern(Xs, Xs))
    x0 = np.asarray([[0.4, 0.6]])
    g1 = o.make_gp(x0, [1.3], 'se', 2.0, [0.3, 0.4], 0.05, 0.7)
    mu, cov = g1.eval(Xs, include_covar=True)
    k = kern(Xs, x0).ravel()
    assert np.allclose(mu, 0.7 + k * 0.6 / 2.05, rtol=1e-13)
    assert np.allclose(cov, kern(Xs, Xs) - np.outer(k, k) / 2.05, atol=1e-13)
    # test point == training point: var -> s * noise / (s + noise)
    _, c = g1.eval(x0, include_covar=True)
    assert abs(c[0, 0] - 2.0 * 0.05 / 2.05) < 1e-13


def test_large_noise_recovers_prior():
    X, y, hp = problem(3, 30, 3)
    g = o.make_gp(X, y, 'se', 1.0, hp['bandwidths'], 1e12, 0.2)
    Xs = np.random.RandomState(1).uniform(size=(10, 3))
    mu, cov = g.eval(Xs, include_covar=True)
    kern = o.make_kernel('se', 3, 1.0, hp['bandwidths'])
    assert np.allclose(mu, 0.2, atol=1e-9) and np.allclose(cov, kern(Xs, Xs), atol=1e-9)


@pytest.mark.parametrize('kernel_type,nu', [('se', 2.5), ('matern', 0.5), ('matern', 1.5), ('matern', 2.5)])
def test_posterior_against_scipy_cho_solve(kernel_type, nu):
    X, y, hp = problem(5, 60, 4, bw=0.4)
    g = o.make_gp(X, y, kernel_type, 1.3, hp['bandwidths'], 0.02, 0.1, matern_nu=nu)
    Xs = np.random.RandomState(2).uniform(size=(25, 4))
    mu, cov = g.eval(Xs, include_covar=True)
    kern = g.gp_core.kernel
    K = kern(X, X) + 0.02 * np.eye(60)
    cf = scipy.linalg.cho_factor(K, lower=True)
    Ksx = kern(Xs, X)
    assert rel_err(mu, 0.1 + Ksx @ scipy.linalg.cho_solve(cf, y - 0.1)) < 1e-10
    assert np.max(np.abs(cov - (kern(Xs, Xs) - Ksx @ scipy.linalg.cho_solve(cf, Ksx.T)))) < 1e-10
    lml = g.gp_core.compute_log_marginal_likelihood()
    r = y - 0.1
    ref = -0.5 * r @ scipy.linalg.cho_solve(cf, r) - np.log(np.diag(cf[0])).sum() - 30 * np.log(2 * np.pi)
    assert abs(lml - ref) < 1e-9
    # direct kernel formula
    d = (X[:, None, :] - X[None, :, :]) / hp['bandwidths']
    r2 = (d ** 2).sum(-1)
    if kernel_type == 'se':
        direct = 1.3 * np.exp(-0.5 * r2)
        assert np.max(np.abs(kern(X, X) - direct)) < 1e-12


@pytest.mark.parametrize('kernel_type,nu', [('se', 2.5), ('matern', 0.5), ('matern', 1.5), ('matern', 2.5)])
def test_posterior_and_lml_against_scikit_learn(kernel_type, nu):
    """An independent third-party implementation of the same GP arithmetic (scikit-learn's
    GaussianProcessRegressor, ARD length scales, fixed hyper-parameters): posterior mean, full
    posterior covariance and log marginal likelihood.  Not a pin to Dragonfly (absent), but it
    rules out a restatement that is self-consistent and wrong."""
    gpr = pytest.importorskip('sklearn.gaussian_process')
    from sklearn.gaussian_process.kernels import RBF, Matern, ConstantKernel, WhiteKernel
    rs = np.random.RandomState(5)
    X = rs.uniform(size=(60, 4))
    y = np.sin(X.sum(1)) + 0.1 * rs.normal(size=60)
    bw = np.array([0.4, 0.3, 0.5, 0.6])
    scale, noise, mu0 = 1.3, 0.02, 0.1
    g = o.make_gp(X, y, kernel_type, scale, bw, noise, mu0, matern_nu=nu)
    base = RBF(length_scale=bw) if kernel_type == 'se' else Matern(length_scale=bw, nu=nu)
    sk = gpr.GaussianProcessRegressor(kernel=ConstantKernel(scale) * base + WhiteKernel(noise),
                                      optimizer=None, alpha=0.0).fit(X, y - mu0)
    Xs = rs.uniform(size=(25, 4))
    mu, cov = g.eval(Xs, include_covar=True)
    mu_sk, cov_sk = sk.predict(Xs, return_cov=True)
    cov_sk = cov_sk - noise * np.eye(25)      # the WhiteKernel term is observation noise, eval() returns f's covariance
    # nu = 1/2 is not differentiable at r = 0: the expanded-form distance (Appendix A) leaves
    # r ~ 1e-8 on the diagonal of K(X, X), which exp(-r) passes on at first order
    tol = 1e-6 if (kernel_type, nu) == ('matern', 0.5) else 1e-11
    assert np.max(np.abs(mu - (mu_sk + mu0))) < tol
    assert np.max(np.abs(cov - cov_sk)) < tol
    assert abs(g.gp_core.compute_log_marginal_likelihood() - sk.log_marginal_likelihood_value_) < 10 * tol


def test_pt_relations_is_eval_covariance():
    """The in-tree formula (gp_interface.py:112-119) pins eval's covariance."""
    X, y, hp = problem(7, 40, 2)
    g = o.make_gp(X, y, 'se', 1.0, hp['bandwidths'], 0.01, 0.0)
    A = np.random.RandomState(3).uniform(size=(15, 2))
    assert np.allclose(g.get_pt_relations(A, A), g.eval(A, include_covar=True)[1], atol=1e-13)


def test_stable_cholesky_ladder():
    v = np.random.RandomState(0).normal(size=(30, 3))
    M = v @ v.T  # rank 3: plain Cholesky fails
    with pytest.raises(np.linalg.LinAlgError):
        np.linalg.cholesky(M)
    L, p = o.stable_cholesky(M, return_jitter=True)
    assert p is not None and -11 <= p <= 4
    assert np.allclose(L @ L.T, M + 10.0 ** p * np.max(np.diag(M)) * np.eye(30), atol=1e-9)
    L2, p2 = o.stable_cholesky(M + np.eye(30), return_jitter=True)
    assert p2 is None
    with pytest.raises(ValueError):
        o.stable_cholesky(-1e9 * np.eye(4))


def test_draw_samples_shape_and_injection():
    X, y, hp = problem(9, 20, 2)
    g = o.make_gp(X, y, 'se', 1.0, hp['bandwidths'], 0.01, 0.0)
    Xs = np.random.RandomState(4).uniform(size=(12, 2))
    o.NORMAL_SOURCE = lambda size: np.zeros(size)
    try:
        s = g.draw_sample(samp_pts=Xs)
        assert s.shape == (12,) and np.allclose(s, g.eval(Xs)[0])
    finally:
        o.NORMAL_SOURCE = None
    np.random.seed(1)
    s1 = g.gp_core.draw_samples(3, X_test=Xs)
    assert s1.shape == (3, 12)


def test_pick_rules_and_ties():
    # MTS (mts.py:37-43): strict > keeps the first task; argmax first index
    samples = [np.asarray([0.0, 1.0, 3.0, 3.0]), np.asarray([0.5, 1.5, 3.5, 3.5])]
    assert picks.mts_pick(samples, [2, 2]) == (0, 0, 2.0)
    # joint MTS (joint_mts.py:43-57)
    sampled = np.asarray([1.0, 2.0, 0.5, 0.1, 5.0, 5.0, 2.0, 2.4, 0.0])  # 3 old (2+1), 2 tasks x 3
    t, a, g = picks.joint_mts_pick(sampled, [2, 1], 2)
    assert (t, a) == (0, 1) and g == 3.0
    assert picks.grid_only_pick(np.asarray([1.0, 4.0, 4.0, 2.0]), 2) == (0, 1, 4.0)
    assert picks.cmts_gain(np.asarray([1.0, 3.0, 2.0]), np.asarray([0.1, 0.0, 0.5])) == (1, 1.0)
    assert picks.cmts_pm_gain(np.asarray([1.0, 3.0, 2.0]), np.asarray([0.1, 0.0, 0.5]), [0.2, 0.3]) == (1, 2.7)
    assert picks.profile_pick([1.0, 2.0, 2.0]) == (1, 2.0)


def test_sample_grid_layout_is_task_major():
    g = picks.sample_grid([[0.0], [0.5], [1.0]], [[0, 1]], 2, uniforms=np.zeros((6, 1)))
    assert g[:, 0].tolist() == [0, 0, 0.5, 0.5, 1, 1]


def test_hp_candidates_and_fitter_criteria():
    from argparse import Namespace
    X, y, hp = problem(11, 25, 2)
    np.random.seed(3)
    c = o.hp_candidates(X, y, 8)
    lo, hi = o.hp_bounds(X, y)
    assert c.shape == (8, 5) and np.all(c[:, 0] > 0) and np.all(c[:, 3:] > 0)
    opts = Namespace(kernel_type='se', hp_tune_criterion='ml-post_sampling', hp_samples=3,
                     hp_search_candidates=8, use_additive_gp=False)
    o.EuclideanGPFitter._fit_counter = 0
    f = o.OracleGPFitter(list(X), list(y), opts)
    f.fit_gp()
    assert len(f.gpf_core.hp_list) == 1  # 'ml'
    best = f.gpf_core.hp_list[0]
    assert np.array_equal(best, f.gpf_core.last_candidates[np.argmax(f.gpf_core.last_lmls)])
    f.fit_gp()
    assert len(f.gpf_core.hp_list) == 3  # 'post_sampling'
    gp = f.get_next_gp()
    assert gp.gp_core.alpha is not None


def test_multi_round_hp_search_is_defined_alike_in_product_and_oracle():
    """``hp_search_rounds`` > 1: 'ml' zooms on the best candidate, other criteria draw more
    uniform candidates.  The product's host logic (ocbo_b200.gp.hp_search.search, LML injected
    from the oracle so that no GPU is needed) and the oracle's fitter must consume the numpy
    stream alike and arrive at the same candidates and the same choice."""
    from argparse import Namespace
    from ocbo_b200.gp import hp_search
    X, y, hp = problem(13, 30, 3)
    lml_of = lambda ths: np.asarray([o.log_marginal_likelihood(X, y, 'se', th) for th in ths])
    for crit, nsel in (('ml', 1), ('post_sampling', 3)):
        opts = Namespace(kernel_type='se', hp_tune_criterion=crit, hp_samples=3,
                         hp_search_candidates=16, hp_search_rounds=3, use_additive_gp=False)
        o.EuclideanGPFitter._fit_counter = 0
        np.random.seed(7)
        f = o.OracleGPFitter(list(X), list(y), opts)
        f.fit_gp()
        np.random.seed(7)
        cands, lmls = hp_search.search(X, y, 0, 16, 3, crit, lml_fn=lml_of)
        idx = hp_search.select(lmls, crit, 3)
        assert cands.shape == (16 + 2 * 15, 6)
        assert np.array_equal(cands, f.gpf_core.last_candidates)
        assert np.array_equal(lmls, f.gpf_core.last_lmls)
        assert len(idx) == nsel == len(f.gpf_core.hp_list)
        for i, th in zip(idx, f.gpf_core.hp_list):
            assert np.array_equal(cands[i], th)
        lo, hi = o.hp_bounds(X, y)
        box = np.log(np.delete(cands, 2, axis=1))
        assert (box >= np.delete(lo, 2) - 1e-12).all() and (box <= np.delete(hi, 2) + 1e-12).all()
        if crit == 'ml':  # a zoom never loses the incumbent, and the later rounds stay near it
            assert lmls.max() >= lmls[:16].max()
            best_t = np.log(np.delete(cands[int(np.argmax(lmls[:16]))], 2))
            r1 = np.log(np.delete(cands[16:31], 2, axis=1))
            assert (np.abs(r1 - best_t) <= 0.25 * (np.delete(hi, 2) - np.delete(lo, 2)) + 1e-12).all()


def test_knowledge_gradient_restatement_is_pinned_by_the_reference_function():
    """oracle.kg.knowledge_gradient == the reference's util.misc_util.knowledge_gradient, bit for
    bit, on the committed vectors (tests/golden/make_kg_golden.py ran the reference's function)."""
    import json
    import os
    from oracle import kg
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'kg_vectors.json')) as fh:
        vecs = json.load(fh)
    assert len(vecs) >= 10
    for v in vecs:
        assert kg.knowledge_gradient(np.asarray(v['a']), np.asarray(v['b'])) == v['kg']
    # closed form for two lines: E[max(0, d + s Z)] with d = a1 - a0 <= 0 shifted, s = |b1 - b0|
    got = kg.knowledge_gradient(np.asarray([0.0, 0.0]), np.asarray([-1.0, 1.0]))
    assert abs(got - 2.0 / np.sqrt(2.0 * np.pi)) <= 1e-15
