"""GPU parity tests: the CUDA path (through the C ABI) against the numpy oracle
on identical seeded inputs and injected normals.

Tolerances (BASELINE.json north_star): posterior mean / variance / samples
within 1e-8 relative in fp64 on well-conditioned fixtures; indices bit-exact.
Ill-conditioned fixtures (jitter ladder) use the cond-scaled tolerance stated
in each test (SURVEY section 7 "hard parts")."""
import ctypes

import numpy as np
import pytest

from tests.util import problem, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-8


@pytest.fixture(scope='module')
def mods():
    import torch
    from oracle import gp_numpy, philox, picks
    from ocbo_b200 import engine, rng, _lib
    from ocbo_b200.gp import gp_interface, gp_core
    return dict(torch=torch, o=gp_numpy, philox=philox, picks=picks, engine=engine, rng=rng,
                lib=_lib, gi=gp_interface, gc=gp_core)


def both_gps(mods, seed, n, D, kernel_type='se', nu=2.5, **kw):
    X, y, hp = problem(seed, n, D, **kw)
    og = mods['o'].make_gp(X, y, kernel_type, hp['scale'], hp['bandwidths'], hp['noise_var'],
                           hp['mu0'], matern_nu=nu)
    bg = mods['gi'].make_gp(X, y, kernel_type, hp['scale'], hp['bandwidths'], hp['noise_var'],
                            hp['mu0'], matern_nu=nu)
    return X, y, og, bg


@pytest.mark.parametrize('kernel_type,nu', [('se', 2.5), ('matern', 0.5), ('matern', 1.5),
                                            ('matern', 2.5)])
@pytest.mark.parametrize('n1,n2,D', [(5, 7, 1), (130, 64, 2), (257, 300, 6)])
def test_gram_matches_oracle(mods, kernel_type, nu, n1, n2, D):
    rs = np.random.RandomState(n1 * 31 + n2)
    A, B = rs.uniform(size=(n1, D)), rs.uniform(size=(n2, D))
    bw = rs.uniform(0.2, 0.8, size=D)
    ko = mods['o'].make_kernel(kernel_type, D, 1.7, bw, nu)
    kb = mods['gc'].make_kernel(kernel_type, D, 1.7, bw, nu)
    ref, got = ko(A, B), kb(A, B)
    assert got.shape == ref.shape
    # the oracle uses Dragonfly's expanded-form distances (|a|^2+|b|^2-2ab, clipped),
    # the kernel the direct form; they agree to ~1e-15 absolute in d^2.  Matern-1/2
    # has an infinite slope at r = 0, hence the looser bound there.
    tol = 1e-7 if (kernel_type == 'matern' and nu == 0.5) else 1e-12
    assert np.max(np.abs(got - ref)) <= tol * 1.7
    sym = kb(A, A)
    assert np.array_equal(sym, sym.T)


@pytest.mark.parametrize('n', [1, 100, 128, 300, 640])
def test_potrf_and_solve(mods, n):
    torch, engine = mods['torch'], mods['engine']
    X, y, og, bg = both_gps(mods, 7 + n, n, 3, noise=1e-2, bw=0.3)
    Lo, Lb = og.gp_core.L, bg.gp_core.L
    assert rel_err(Lb, Lo) <= 1e-10
    assert rel_err(bg.gp_core.alpha, og.gp_core.alpha) <= TOL
    lml_o = og.gp_core.compute_log_marginal_likelihood()
    lml_b = bg.gp_core.compute_log_marginal_likelihood()
    assert abs(lml_b - lml_o) <= 1e-9 * max(1.0, abs(lml_o))
    # inverted diagonal blocks
    post = bg.gp_core._post
    inv = post.invd[0].cpu().numpy()
    Lfull = post.L[0].cpu().numpy()
    for j in range(post.n_pad // 128):
        blk = np.tril(Lfull[j * 128:(j + 1) * 128, j * 128:(j + 1) * 128])
        assert np.max(np.abs(inv[j] @ blk - np.eye(128))) <= 1e-9
        assert np.all(np.triu(Lfull[j * 128:(j + 1) * 128, j * 128:(j + 1) * 128], 1) == 0)


def test_potrf_reports_first_bad_pivot(mods):
    """LAPACK dpotrf convention: info = column (1-based) of the first pivot <= 0."""
    torch, engine = mods['torch'], mods['engine']
    n = 256
    rs = np.random.RandomState(3)
    Q = rs.normal(size=(n, n))
    A = Q @ Q.T + n * np.eye(n)
    A[150, 150] = -1.0  # makes the leading 151 x 151 minor indefinite
    import scipy.linalg
    _, info_ref = scipy.linalg.lapack.dpotrf(A, lower=1)
    Ad = engine._h2d(A[None])
    invd = torch.empty((1, 2, 128, 128), dtype=torch.float64, device=Ad.device)
    info = torch.zeros((1,), dtype=torch.int32, device=Ad.device)
    engine.potrf(Ad, invd, info)
    assert int(info.item()) == info_ref == 151


@pytest.mark.parametrize('n,m,D', [(40, 90, 2), (200, 333, 6), (300, 128, 1)])
def test_eval_mean_and_covariance(mods, n, m, D):
    X, y, og, bg = both_gps(mods, n + m, n, D)
    Xs = np.random.RandomState(5).uniform(size=(m, D))
    mo, co = og.eval(Xs, include_covar=True)
    mb, cb = bg.eval(Xs, include_covar=True)
    assert rel_err(mb, mo) <= TOL
    assert np.max(np.abs(cb - co)) <= TOL * np.max(np.abs(co))
    assert rel_err(np.diag(cb), np.diag(co)) <= TOL
    assert np.array_equal(cb, cb.T)
    mb2, none = bg.eval(Xs)
    assert none is None and np.array_equal(mb2, mb)


def test_prior_and_single_point_posteriors(mods):
    """Closed forms: n = 0 gives the prior; n = 1 the analytic posterior."""
    gi = mods['gi']
    D = 2
    Xs = np.random.RandomState(1).uniform(size=(50, D))
    g0 = gi.make_gp(np.zeros((0, D)), [], 'se', 2.0, [0.3, 0.4], 0.05, 0.7)
    mu, cov = g0.eval(Xs, include_covar=True)
    kern = mods['o'].make_kernel('se', D, 2.0, [0.3, 0.4])
    assert np.allclose(mu, 0.7) and np.max(np.abs(cov - kern(Xs, Xs))) <= 1e-12
    x0, y0 = np.asarray([[0.4, 0.6]]), [1.3]
    g1 = gi.make_gp(x0, y0, 'se', 2.0, [0.3, 0.4], 0.05, 0.7)
    mu, cov = g1.eval(Xs, include_covar=True)
    k = kern(Xs, x0).ravel()
    mu_ref = 0.7 + k * (1.3 - 0.7) / (2.0 + 0.05)
    cov_ref = kern(Xs, Xs) - np.outer(k, k) / (2.0 + 0.05)
    assert rel_err(mu, mu_ref) <= 1e-12 and np.max(np.abs(cov - cov_ref)) <= 1e-12
    # growing an empty (pre-tuned) GP one point at a time == building it at once
    g0.add_data_single(x0[0], y0[0])
    mu2, cov2 = g0.eval(Xs, include_covar=True)
    assert np.array_equal(mu2, mu) and np.array_equal(cov2, cov)


def test_pt_relations_matches_in_tree_formula(mods):
    """get_pt_relations (gp_interface.py:112-119) == eval(A,'covar') on A,A."""
    X, y, og, bg = both_gps(mods, 11, 150, 3)
    rs = np.random.RandomState(2)
    A, B = rs.uniform(size=(70, 3)), rs.uniform(size=(140, 3))
    ref = og.get_pt_relations(A, B)
    got = bg.get_pt_relations(A, B)
    assert np.max(np.abs(got - ref)) <= TOL * np.max(np.abs(ref))
    _, cov = bg.eval(A, include_covar=True)
    assert np.max(np.abs(bg.get_pt_relations(A, A) - cov)) <= 1e-12


@pytest.mark.parametrize('n,m,D,bw', [(30, 100, 2, 0.05), (120, 260, 6, 0.15)])
def test_draw_sample_with_injected_normals(mods, n, m, D, bw):
    """Well-conditioned fixture (short length-scale, noise 1e-2)."""
    X, y, og, bg = both_gps(mods, 100 + n, n, D, noise=1e-2, bw=bw)
    Xs = np.random.RandomState(9).uniform(size=(m, D))
    z = np.random.RandomState(10).normal(size=(m, 1))
    mods['o'].NORMAL_SOURCE = lambda size: z.reshape(size)
    mods['rng'].NORMAL_SOURCE = lambda size: z.reshape(size)
    try:
        so = og.draw_sample(samp_pts=Xs)
        sb = bg.draw_sample(samp_pts=Xs)
        assert og.gp_core is not None and sb.shape == so.shape == (m,)
        cond = np.linalg.cond(og.eval(Xs, True)[1])
        assert rel_err(sb, so) <= max(TOL, 10 * cond * 2.2e-16), cond
        assert int(np.argmax(sb)) == int(np.argmax(so))
        # (means, covar) form, as the continuous strategies call it (profile_cts.py:100-102)
        mu, cov = og.eval(Xs, include_covar=True)
        so2 = og.draw_sample(means=mu, covar=cov)
        sb2 = bg.draw_sample(means=mu, covar=cov)
        assert rel_err(sb2, so2) <= max(TOL, 10 * cond * 2.2e-16)
    finally:
        mods['o'].NORMAL_SOURCE = None
        mods['rng'].NORMAL_SOURCE = None


def test_jitter_ladder_on_singular_covariance(mods):
    """Duplicate candidates make Sigma exactly singular: both paths must walk the
    ladder; samples then agree to sqrt(jitter)-scaled tolerance and the equal
    rows tie (first index wins, like np.argmax)."""
    X, y, og, bg = both_gps(mods, 21, 40, 2, noise=1e-3, bw=0.4)
    rs = np.random.RandomState(4)
    base = rs.uniform(size=(60, 2))
    Xs = np.vstack([base, base[:20], X[:10]])  # duplicates + training points
    z = rs.normal(size=(len(Xs), 1))
    mods['o'].NORMAL_SOURCE = lambda size: z.reshape(size)
    mods['rng'].NORMAL_SOURCE = lambda size: z.reshape(size)
    try:
        mu, cov = og.eval(Xs, include_covar=True)
        _, p_ref = mods['o'].stable_cholesky(cov, return_jitter=True)
        so = og.draw_sample(samp_pts=Xs)
        sb = bg.draw_sample(samp_pts=Xs)
        p_got = bg.gp_core.last_sample_jitter
        assert p_ref is not None and p_got is not None
        # knife-edge ladder decisions may differ by one rung (documented)
        assert abs(p_got - p_ref) <= 1
        jit = 10.0 ** max(p_got, p_ref) * np.max(np.diag(cov))
        assert np.max(np.abs(sb - so)) <= 50 * np.sqrt(jit) * np.max(np.abs(z))
    finally:
        mods['o'].NORMAL_SOURCE = None
        mods['rng'].NORMAL_SOURCE = None


def test_device_ladder_walks_each_problem_to_its_own_rung(mods):
    """ocbo_ladder_* on a ragged batch: a PD problem (no jitter, parked through every rung), an
    exactly singular one (needs a small rung), one that needs a LARGE rung (beyond the
    speculative ones: host-continued) and an identity-padded small one.  Rung and factor must
    equal numpy's stable_cholesky of the same matrices; an indefinite matrix raises
    ValueError like the reference."""
    torch, engine, o = mods['torch'], mods['engine'], mods['o']
    rs = np.random.RandomState(7)
    m_pad = 256

    def psd(m, rank):
        G = rs.normal(size=(m, rank))
        return G.dot(G.T) / rank

    mats = [psd(200, 400) + 1e-3 * np.eye(200),   # PD
            psd(256, 100),                        # singular: ladder
            psd(150, 50) - 1e-6 * np.eye(150),    # slightly indefinite: needs a large rung
            psd(40, 10)]                          # small + singular, padded with identity
    mv = [len(M) for M in mats]
    ref = [o.stable_cholesky(M, return_jitter=True) for M in mats]
    assert ref[0][1] is None and ref[1][1] is not None and ref[2][1] > -10
    Sh = np.tile(np.eye(m_pad), (len(mats), 1, 1))
    for b, M in enumerate(mats):
        Sh[b, :mv[b], :mv[b]] = M
    Sd = engine._h2d(Sh)
    mv_dev = engine._h2d(np.asarray(mv, dtype=np.int32), engine.I32)
    pend = engine.PendingCholesky(Sd, mv_dev)
    jit = pend.resolve()
    assert (pend.info.cpu().numpy() == 0).all()
    Lh = Sd.cpu().numpy()
    for b, (Lref, pref) in enumerate(ref):
        if pref is None:
            assert jit[b] is None
        else:
            # knife-edge ladder decisions may differ by one rung (documented)
            assert jit[b] is not None and abs(jit[b] - pref) <= 1
        if jit[b] == pref:
            Lb = np.tril(Lh[b, :mv[b], :mv[b]])
            scale = 1.0 if pref is None else 10.0 ** pref
            A = mats[b] + (0 if pref is None else scale * np.max(np.diag(mats[b])) * np.eye(mv[b]))
            assert rel_err(Lb.dot(Lb.T), A) <= 1e-12
        # identity padding survives the ladder untouched
        assert np.array_equal(np.tril(Lh[b, mv[b]:, mv[b]:]), np.eye(m_pad - mv[b]))
    bad = np.tile(np.eye(128), (1, 1, 1))
    bad[0, :4, :4] = -np.eye(4) * 1e6 + 1.0
    bad[0, 0, 0] = 1.0   # max(diag) = 1 but a -1e6 pivot: no rung up to 1e4 fixes it
    Bd = engine._h2d(bad)
    with pytest.raises(ValueError):
        engine.PendingCholesky(Bd, engine._h2d(np.asarray([4], dtype=np.int32), engine.I32)).resolve()


def test_concurrent_first_rung_equals_the_sequential_walk(mods):
    """The speculative rung -11 factored on a second stream next to the plain attempt
    (ocbo_ladder_adopt) must leave exactly what the sequential ladder leaves: factor, inverted
    diagonal blocks, rung and status, for PD, singular and needs-a-larger-rung problems."""
    torch, engine = mods['torch'], mods['engine']
    rs = np.random.RandomState(12)
    m_pad = 384

    def psd(m, rank):
        G = rs.normal(size=(m, rank))
        return G.dot(G.T) / rank

    mats = [psd(300, 600) + 1e-3 * np.eye(300), psd(384, 120), psd(200, 60) - 1e-6 * np.eye(200),
            psd(30, 5), psd(129, 129) + 1e-6 * np.eye(129)]
    mv = [len(M) for M in mats]
    Sh = np.tile(np.eye(m_pad), (len(mats), 1, 1))
    for b, M in enumerate(mats):
        Sh[b, :mv[b], :mv[b]] = M
    outs = []
    for conc in (False, True):
        Sd = engine._h2d(Sh)
        mv_dev = engine._h2d(np.asarray(mv, dtype=np.int32), engine.I32)
        pend = engine.PendingCholesky(Sd, mv_dev, spec=2, concurrent_first=conc)
        jit = pend.resolve()
        outs.append((np.tril(Sd.cpu().numpy()), pend.invd.cpu().numpy(), pend.info.cpu().numpy(), jit))
    assert outs[0][3] == outs[1][3] and outs[0][3][0] is None and outs[0][3][1] is not None
    assert outs[0][3][2] > -10
    for k in range(3):
        assert np.array_equal(outs[0][k], outs[1][k])


def test_potrf_skips_identity_padding_bit_exactly(mods):
    """ocbo_potrf_nv == ocbo_potrf on identity-padded problems, bit for bit (L and the
    inverted diagonal blocks), for valid sizes on and off the 16-column stage boundaries."""
    torch, engine = mods['torch'], mods['engine']
    rs = np.random.RandomState(3)
    mv = [1, 16, 50, 128, 129, 200, 255, 256]
    m_pad = 256
    Ah = np.tile(np.eye(m_pad), (len(mv), 1, 1))
    for b, m in enumerate(mv):
        G = rs.normal(size=(m, m + 5))
        Ah[b, :m, :m] = G.dot(G.T) / m + 0.1 * np.eye(m)
    outs = []
    for use_nv in (False, True):
        Ad = engine._h2d(Ah)
        invd = torch.zeros((len(mv), m_pad // 128, 128, 128), dtype=torch.float64, device=Ad.device)
        info = torch.zeros((len(mv),), dtype=torch.int32, device=Ad.device)
        nv = engine._h2d(np.asarray(mv, dtype=np.int32), engine.I32) if use_nv else None
        engine.potrf(Ad, invd, info, nv)
        assert (info.cpu().numpy() == 0).all()
        outs.append((np.tril(Ad.cpu().numpy()), invd.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])
    for b, m in enumerate(mv):
        L = outs[1][0][b, :m, :m]
        assert rel_err(L.dot(L.T), Ah[b, :m, :m]) <= 1e-13


def test_diagonal_block_kernels_are_bit_identical(tmp_path):
    """The blocked diagonal-block kernel (default) regroups the column steps of the
    column-per-barrier register kernel without changing the arithmetic or its order: L and
    inv(L_jj) must agree bit for bit (and the first bad pivot must be reported alike).  The
    variant is chosen per process (OCBO_DIAG_KERNEL), hence the subprocesses."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, numpy as np, torch
from ocbo_b200 import engine
rs = np.random.RandomState(11)
mv = [128, 100, 256, 300, 384, 17]
m_pad = 384
Ah = np.tile(np.eye(m_pad), (len(mv), 1, 1))
for b, m in enumerate(mv):
    G = rs.normal(size=(m, m + 3))
    Ah[b, :m, :m] = G.dot(G.T) / m + 0.05 * np.eye(m)
Ah[3, 150, 150] = -1.0          # problem 3 fails at column 151
out = {}
for tag, use_nv in (('full', False), ('nv', True)):
    Ad = engine._h2d(Ah)
    invd = torch.zeros((len(mv), m_pad // 128, 128, 128), dtype=torch.float64, device=Ad.device)
    info = torch.zeros((len(mv),), dtype=torch.int32, device=Ad.device)
    nv = engine._h2d(np.asarray(mv, dtype=np.int32), engine.I32) if use_nv else None
    engine.potrf(Ad, invd, info, nv)
    out['L_' + tag] = np.tril(Ad.cpu().numpy())
    out['inv_' + tag] = invd.cpu().numpy()
    out['info_' + tag] = info.cpu().numpy()
np.savez(sys.argv[1], **out)
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for variant in ('reg', 'blk'):
        path = str(tmp_path / (variant + '.npz'))
        env = dict(os.environ, OCBO_DIAG_KERNEL=variant, PYTHONPATH=root)
        subprocess.run([sys.executable, '-c', code, path], check=True, env=env, cwd=root, timeout=300)
        res[variant] = np.load(path)
    for variant in ('blk',):
        for key in res['reg'].files:
            ok = res['reg'][key]
            got = res[variant][key]
            if key.startswith('info'):
                assert list(got) == [0, 0, 0, 151, 0, 0] and np.array_equal(ok, got)
            else:
                good = [b for b in range(6) if b != 3]
                assert np.array_equal(ok[good], got[good]), (variant, key)


def test_kg_kernel_matches_the_reference_vectors(mods):
    """ocbo_kg against tests/golden/kg_vectors.json, produced by the reference's OWN
    knowledge_gradient (make_kg_golden.py): sort, envelope and segment sum follow it operation
    for operation, only erf / erfc / exp come from another libm (a few ulp)."""
    import json
    import os
    torch, engine = mods['torch'], mods['engine']
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'kg_vectors.json')) as fh:
        vecs = json.load(fh)
    assert len(vecs) >= 10
    one = engine._h2d(np.ones(1))  # cand_var = 1, noise = 0: slopes are divided by exactly 1
    for v in vecs:
        a, b = np.asarray(v['a']), np.asarray(v['b'])
        means = engine._h2d(a[None, :])
        inter = engine._h2d(b[None, :])
        kg = engine.knowledge_gradients(means, len(a), inter, one, 0.0, 1, 1, len(a))
        got = float(kg.cpu().numpy()[0, 0])
        assert abs(got - v['kg']) <= 1e-13 * max(1.0, abs(v['kg'])), (len(a), got, v['kg'])


def test_revi_improvements_match_the_oracle(mods):
    """The whole REVI scoring pass (V for both point sets, judgement means, posterior
    cross-covariance, candidate variances, ocbo_kg, sum over groups) against the numpy
    oracle + the pinned knowledge-gradient restatement."""
    from oracle import kg
    from ocbo_b200 import batched
    X, y, og, bg = both_gps(mods, 17, 60, 3, noise=1e-2, bw=0.3)
    rs = np.random.RandomState(2)
    G, E, C = 5, 40, 37
    cand = rs.uniform(size=(C, 3))
    judge = rs.uniform(size=(G * E, 3))
    got = batched.revi_improvements(bg, cand, judge, G, E)
    means = og.eval(judge)[0].reshape(G, E)
    inter = og.get_pt_relations(cand, judge)
    cand_vars = og.eval(cand, include_covar=True)[1].diagonal()
    ref = kg.revi_improvements(means, inter, cand_vars, og.get_estimated_noise())
    assert got.shape == ref.shape == (C,)
    assert rel_err(got, ref) <= TOL
    assert int(np.argmax(got)) == int(np.argmax(ref))


def test_dmma_sampling_path_many_draws(mods):
    """S = 256 draws go through the DMMA L*Z kernel; compare with numpy."""
    torch, engine = mods['torch'], mods['engine']
    X, y, og, bg = both_gps(mods, 31, 200, 6, noise=1e-2, bw=0.25)
    m, S = 384, 256
    Xs = np.random.RandomState(12).uniform(size=(m, 6))
    Z = np.random.RandomState(13).normal(size=(m, S))
    mods['o'].NORMAL_SOURCE = lambda size: Z.reshape(size)
    mods['rng'].NORMAL_SOURCE = lambda size: Z.reshape(size)
    try:
        ref = og.gp_core.draw_samples(S, X_test=Xs)
        got = bg.gp_core.draw_samples(S, X_test=Xs)
        assert got.shape == ref.shape == (S, m)
        cond = np.linalg.cond(og.eval(Xs, True)[1])
        assert rel_err(got, ref) <= max(TOL, 10 * cond * 2.2e-16), cond
        assert np.array_equal(np.argmax(got, axis=1), np.argmax(ref, axis=1))
    finally:
        mods['o'].NORMAL_SOURCE = None
        mods['rng'].NORMAL_SOURCE = None


def test_philox_normals_match_restatement(mods):
    engine = mods['engine']
    m, S, seed = 300, 7, 100826730
    Zd = engine.philox_normals(2, 384, S, seed, [5, 9], engine.device())
    for b, trial in enumerate([5, 9]):
        ref = mods['philox'].normals(seed, trial, 384, S)
        got = Zd[b].cpu().numpy()
        assert np.max(np.abs(got - ref)) <= 1e-12
    assert abs(got.mean()) < 0.1 and abs(got.std() - 1) < 0.1


def test_segment_argmax_and_joint_pick(mods):
    torch, engine, picks = mods['torch'], mods['engine'], mods['picks']
    rs = np.random.RandomState(8)
    T, A, S = 6, 50, 5
    num_qs = [3, 7, 5, 9, 4, 6]
    n_old = sum(num_qs)
    m = n_old + T * A
    Fh = rs.normal(size=(2, m, S))
    Fh[0, n_old + 10, 0] = Fh[0, n_old + 20, 0] = 9.0  # tie inside task 0: first index wins
    Fh[1, n_old + 2 * A + 1, 1] = Fh[1, n_old + 4 * A + 3, 1] = 7.0  # tie across tasks
    Fd = engine._h2d(Fh)
    divs = np.concatenate([[0], np.cumsum(num_qs)])
    seg = np.concatenate([divs, n_old + A * np.arange(1, T + 1)]).astype(np.int32)
    seg = np.stack([seg, seg])
    mx, ag = engine.segment_argmax(Fd, S, seg)
    task, act, gain = engine.joint_pick(mx, ag, T, T)
    task, act, gain = task.cpu().numpy(), act.cpu().numpy(), gain.cpu().numpy()
    for b in range(2):
        for s in range(S):
            t_ref, a_ref, g_ref = picks.joint_mts_pick(Fh[b, :, s], num_qs, T)
            assert (task[b, s], act[b, s]) == (t_ref, a_ref)
            assert gain[b, s] == g_ref
    # grid-only variant (sweep): old best = 0
    seg2 = np.stack([(A * np.arange(0, T + 1)).astype(np.int32)] * 2)
    G = engine._h2d(Fh[:, n_old:, :])
    mx2, ag2 = engine.segment_argmax(G, S, seg2)
    t2, a2, g2 = [x.cpu().numpy() for x in engine.joint_pick(mx2, ag2, 0, T)]
    for b in range(2):
        for s in range(S):
            assert (t2[b, s], a2[b, s], g2[b, s]) == picks.grid_only_pick(Fh[b, n_old:, s], T)


@pytest.mark.parametrize('n', [256, 1024, 3072])
def test_tile_core_dgemm_matches_cublas(mods, n):
    """The DMMA tile core with its TMA-fed, free-running operand pipeline (gemm_core.cuh) on a plain
    C = A B^T against cuBLAS: many CTAs per SM slot, many k-tiles per CTA -- a refill that raced with
    the fragment loads of a stage (tma_mainloop's "drained" barrier) shows up here at once."""
    import ctypes
    torch = mods['torch']
    from ocbo_b200._lib import call
    g = torch.Generator(device='cuda').manual_seed(n)
    a = torch.randn(n, n, dtype=torch.float64, device='cuda', generator=g)
    b = torch.randn(n, n, dtype=torch.float64, device='cuda', generator=g)
    c = torch.zeros(n, n, dtype=torch.float64, device='cuda')
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ref = a @ b.t()
    for _ in range(3):
        c.zero_()
        call('ocbo_probe_dgemm_nt', ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()),
             ctypes.c_void_p(c.data_ptr()), n, st)
        torch.cuda.synchronize()
        assert float((c - ref).abs().max() / ref.abs().max()) <= 1e-13


def test_fused_sampling_tail_equals_sample_then_argmax(mods):
    """ocbo_sample_tile_argmax (F never written) against ocbo_sample + np.max / np.argmax on the
    materialised F: bit-identical maxima, first-index arg-max on ties (duplicate rows inside a
    tile, equal maxima of two draws), NaN propagated like numpy, a failed problem left untouched."""
    import ctypes
    torch, engine = mods['torch'], mods['engine']
    from ocbo_b200._lib import call
    from ocbo_b200.engine import _p
    rs = np.random.RandomState(21)
    B, m, S = 3, 512, 64
    L = np.tril(rs.normal(size=(B, m, m))) / np.sqrt(m)
    Z = rs.normal(size=(B, m, S))
    mean = rs.normal(size=(B, m))
    # ties: row 200 repeats row 137 of problem 0 (same tile 1: 128..255) -- zero tail beyond column 137
    L[0, 200, :] = 0.0
    L[0, 200, :138] = L[0, 137, :138]
    mean[0, 200] = mean[0, 137] = 50.0       # and make the pair the tile's maximum in every draw
    Z[1, :, 5] = 0.0                          # problem 1, draw 5: F = mean, with duplicated maxima
    mean[1, 300] = mean[1, 260] = mean[1, 383] = 9.0
    mean[1, 17] = np.nan                      # NaN in tile 0 of problem 1: first NaN wins
    mean[1, 90] = np.nan
    Ld, Zd, md = engine._h2d(L), engine._h2d(Z), engine._h2d(mean)
    info = torch.zeros(B, dtype=torch.int32, device=Ld.device)
    info[2] = 7                               # problem 2 "failed its factorisation": skipped
    F = torch.zeros((B, m, S), dtype=torch.float64, device=Ld.device)
    tmax = torch.full((B, m // 128, S), -1.0, dtype=torch.float64, device=Ld.device)
    targ = torch.full((B, m // 128, S), -5, dtype=torch.int32, device=Ld.device)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    call('ocbo_sample', _p(md), m, _p(Ld), m, m, m * m, _p(Zd), S, S, m * S, _p(F), S, m * S, _p(info), B, st)
    call('ocbo_sample_tile_argmax', _p(md), m, _p(Ld), m, m, m * m, _p(Zd), S, S, m * S, _p(tmax),
         _p(targ), _p(info), B, st)
    Fh, tmax, targ = F.cpu().numpy(), tmax.cpu().numpy(), targ.cpu().numpy()
    for b in range(2):
        tiles = Fh[b].reshape(m // 128, 128, S)
        ref_arg = np.argmax(tiles, axis=1)
        ref_max = np.max(tiles, axis=1)
        assert np.array_equal(targ[b], ref_arg)
        assert np.array_equal(tmax[b], ref_max, equal_nan=True)
    assert (targ[0, 1] == 137 - 128).all()    # the earlier of the two equal rows
    assert targ[1, 2, 5] == 260 - 256 and np.isnan(tmax[1, 0]).all() and (targ[1, 0] == 17).all()
    assert (tmax[2] == -1.0).all() and (targ[2] == -5).all()
    # the unfused reduction gives the same answers (shared comparator)
    seg = np.stack([(128 * np.arange(m // 128 + 1)).astype(np.int32)] * B)
    mx, ag = engine.segment_argmax(F[:2].contiguous(), S, seg[:2])
    assert np.array_equal(mx.cpu().numpy(), tmax[:2], equal_nan=True)
    assert np.array_equal(ag.cpu().numpy(), targ[:2])


def test_batched_lml_matches_oracle(mods):
    from ocbo_b200.gp import hp_search
    X, y, hp = problem(17, 90, 3)
    np.random.seed(5)
    cands = hp_search.hp_candidates(X, y, 24)
    np.random.seed(5)
    cands_o = mods['o'].hp_candidates(X, y, 24)
    assert np.array_equal(cands, cands_o)
    lml, _ = hp_search.batched_lml(X, y, 0, cands)
    ref = np.asarray([mods['o'].log_marginal_likelihood(X, y, 'se', th) for th in cands])
    assert np.max(np.abs(lml - ref) / np.maximum(1.0, np.abs(ref))) <= 1e-9
    assert int(np.argmax(lml)) == int(np.argmax(ref))


def test_bad_arguments_raise_value_error(mods):
    """C ABI returns -k for argument k; the binding raises ValueError like the
    reference's own error convention (gp_util.py:30)."""
    lib = mods['lib']
    with pytest.raises(ValueError):
        lib.call('ocbo_potrf', None, 128, 128, 0, None, None, 1, None)
    with pytest.raises(ValueError):
        mods['gi'].make_gp(np.zeros((3, 20)), [0, 0, 0], 'se', 1.0, np.ones(20), 0.1, 0.0)
    from ocbo_b200.gp.gp_util import get_gp_fitter
    with pytest.raises(ValueError):
        get_gp_fitter('dragonfly', [[0.0]], [0.0], None)


@pytest.mark.parametrize('n,T', [(1024, 64), (2048, 128)])
def test_full_size_cholesky_backward_error(mods, n, T):
    """BASELINE sweep size (n = 1024, m = 8192) and twice it (n = 2048, m = 16384: 2 GiB of
    Sigma, 64-bit offsets everywhere): size-independent properties instead of a CPU
    comparison -- backward error of the factor and of the triangular solve."""
    torch, engine = mods['torch'], mods['engine']
    m, D = T * 128, 6
    X, y, hp = problem(1, n, D, noise=1e-2, bw=0.25)
    hp['bandwidths'] = np.asarray([0.25] * 4 + [0.05] * 2)  # sweep HPs: cond(Sigma) ~ 1e5
    bg = mods['gi'].make_gp(X, y, 'se', 1.0, hp['bandwidths'], 1e-2, 0.0)
    core = bg.gp_core
    rs = np.random.RandomState(2)
    locs = rs.uniform(size=(T, 4))
    Xs = np.hstack([np.repeat(locs, 128, axis=0), rs.uniform(size=(m, 2))])
    Xd, mv_dev, mv, mean, V, Sigma = core._posterior_at(Xs, True, False)
    Sfull = Sigma[0].clone()
    Lsig = Sigma.clone()
    invd, info, jit = engine.factor_covariance(None, Lsig, mv_dev, mv, lambda a, b: None)
    assert int(info.item()) == 0 and (jit[0] is None or m > 8192)
    if jit[0] is not None:  # (denser candidates: the ladder's rung is part of what was factored)
        Sfull += torch.eye(m, dtype=Sfull.dtype, device=Sfull.device) * (10.0 ** jit[0]) * Sfull.diagonal().max()
    L = torch.tril(Lsig[0])
    probe = torch.randn(m, 8, dtype=torch.float64, device=L.device)
    resid = L @ (L.t() @ probe) - Sfull @ probe
    assert float(resid.abs().max() / (Sfull @ probe).abs().max()) <= 1e-12
    # trsm: L_K V = K(X, Xs)
    post = core._post
    Kxs = torch.from_numpy(mods['o'].make_kernel('se', D, 1.0, hp['bandwidths'])(X, Xs)).to(L.device)
    LK = torch.tril(post.L[0])
    assert float((LK @ V[0] - Kxs).abs().max()) <= 1e-11
    # Sigma symmetric and PSD-consistent with the oracle on a 256-point subset
    og = mods['o'].make_gp(X, y, 'se', 1.0, hp['bandwidths'], 1e-2, 0.0)
    idx = rs.choice(m, 256, replace=False)
    _, cov_ref = og.eval(Xs[idx], include_covar=True)
    got = Sfull[idx][:, idx].cpu().numpy()
    assert np.max(np.abs(got - cov_ref)) <= TOL * np.max(np.abs(cov_ref))


def test_std_and_expected_improvement_without_forming_sigma(mods):
    """eval(x, 'std') and the EI-family reduction use the fused diagonal-variance kernel;
    compare with the oracle's full-covariance route (misc_util.py:368-385)."""
    from scipy.stats import norm
    from ocbo_b200 import batched
    X, y, og, bg = both_gps(mods, 41, 90, 3, noise=1e-2, bw=0.2)
    rs = np.random.RandomState(3)
    pts = rs.uniform(size=(150, 3))
    mu_o, sd_o = og.gp_core.eval(pts, 'std')
    mu_b, sd_b = bg.gp_core.eval(pts, 'std')
    assert rel_err(mu_b, mu_o) <= TOL and rel_err(sd_b, sd_o) <= TOL
    sets = [rs.uniform(size=(100, 3)) for _ in range(4)]
    bests = [0.1, 0.5, -0.2, 0.9]
    ei_max, ei_arg = batched.shared_ei_step(bg, sets, best_list=bests)
    for p, best, em, ea in zip(sets, bests, ei_max, ei_arg):
        mu, cov = og.eval(p, include_covar=True)
        sd = np.sqrt(np.diag(cov))
        z = (mu - best) / sd
        ei = sd * (z * norm.cdf(z) + norm.pdf(z))
        assert int(ea) == int(np.argmax(ei))
        assert abs(em - ei.max()) <= 1e-8 * max(ei.max(), 1e-12) + 1e-14
    cap = 0.3
    ei_max2, ei_arg2 = batched.shared_ei_step(bg, sets, cap=cap)
    for p, em, ea in zip(sets, ei_max2, ei_arg2):
        mu, cov = og.eval(p, include_covar=True)
        sd = np.sqrt(np.diag(cov))
        z = (mu - min(mu.max(), cap)) / sd
        ei = sd * (z * norm.cdf(z) + norm.pdf(z))
        assert int(ea) == int(np.argmax(ei))
