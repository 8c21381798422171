"""CUDA-graph step plans (engine.StepPlan): a replayed step must give exactly what the
eagerly launched step gives, for data that changes under a fixed shape key (the BO loop
adds one observation per step), and must fall back when the ladder needs the host."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gps(make_gp, rs, T, d, sizes):
    out = []
    for f in range(T):
        X = rs.uniform(size=(sizes[f], d))
        y = np.sin(3 * X.sum(1)) + 0.01 * rs.normal(size=len(X))
        v = float(np.var(y))
        out.append(make_gp(X, y, 'se', v, 0.3 * np.ones(d), 0.01 * v, float(np.median(y)), build=False))
    return out


def test_graph_replay_equals_eager_launches(monkeypatch):
    from ocbo_b200 import batched, engine
    from ocbo_b200.gp.gp_interface import make_gp
    rs = np.random.RandomState(5)
    T, d = 4, 3
    gps = _gps(make_gp, rs, T, d, [20, 31, 7, 44])
    engine.StepPlan.clear()
    replays = 0
    for step in range(4):
        pts = [np.vstack([g.gp_core.X_array(), rs.uniform(size=(60, d))]) for g in gps]
        Z = [rs.normal(size=(len(p), 1)) for p in pts]
        n_old = [g.gp_core.num_tr_data for g in gps]
        monkeypatch.setattr(engine, 'graphs_enabled', lambda: True)
        got = batched.mts_step(gps, pts, n_old, Z)
        jit_g = [g.gp_core.last_sample_jitter for g in gps]
        monkeypatch.setattr(engine, 'graphs_enabled', lambda: False)
        ref = batched.mts_step(gps, pts, n_old, Z)
        jit_e = [g.gp_core.last_sample_jitter for g in gps]
        assert got == ref          # floats and indices, bit for bit
        assert jit_g == jit_e
        for g in gps:              # the loop's add_data_single: sizes change, the key does not
            x = rs.uniform(size=d)
            g.add_data_single(x, float(np.sin(3 * x.sum())))
    plans = list(engine.StepPlan._plans.values())
    assert len(plans) == 1 and plans[0].replays == 4


def test_profile_and_joint_plans_equal_eager(monkeypatch):
    from ocbo_b200 import batched, engine, rng
    from ocbo_b200.gp.gp_interface import make_gp
    rs = np.random.RandomState(9)
    X = rs.uniform(size=(150, 4))
    y = np.cos(2 * X.sum(1))
    gp = make_gp(X, y, 'se', 1.0, 0.4 * np.ones(4), 1e-2, 0.0)
    acts = [np.hstack([np.tile(c, (40, 1)), rs.uniform(size=(40, 2))]) for c in rs.uniform(size=(6, 2))]
    Z = [rs.normal(size=(40, 1)) for _ in acts]
    engine.StepPlan.clear()
    outs = []
    for on in (True, True, False):
        monkeypatch.setattr(engine, 'graphs_enabled', lambda on=on: on)
        outs.append(batched.profile_step(gp, acts, Z))
    for a, b in zip(outs[0], outs[2]):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    for a, b in zip(outs[1], outs[2]):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    # joint pick: segments = 3 tasks' old points, then 3 grids
    Xj = rs.uniform(size=(30, 2))
    gj = make_gp(Xj, np.sin(4 * Xj[:, 0]) + Xj[:, 1], 'se', 1.0, 0.3 * np.ones(2), 1e-2, 0.0)
    to_samp = np.vstack([Xj, rs.uniform(size=(90, 2))])
    seg = np.asarray([0, 10, 20, 30, 60, 90, 120], dtype=np.int32)
    z = rs.normal(size=(len(to_samp), 1))
    rng.NORMAL_SOURCE = lambda size: z.reshape(size)
    try:
        picks = []
        for on in (True, True, False):
            monkeypatch.setattr(engine, 'graphs_enabled', lambda on=on: on)
            picks.append(gj.gp_core.sample_and_pick(to_samp, seg, 3))
        assert picks[0] == picks[2] and picks[1] == picks[2]
    finally:
        rng.NORMAL_SOURCE = None


def test_plan_falls_back_when_the_ladder_needs_the_host(monkeypatch):
    """Exact duplicates among the candidates AND a tiny noise make Sigma need more rungs than
    the speculative two: the planned step must hand over to the host-continued ladder and
    still agree with the eager step."""
    from ocbo_b200 import batched, engine
    from ocbo_b200.gp.gp_interface import make_gp
    rs = np.random.RandomState(2)
    X = rs.uniform(size=(40, 2))
    y = np.sin(5 * X[:, 0])
    gps = [make_gp(X, y, 'se', 1.0, np.asarray([2.0, 2.0]), 1e-10, 0.0, build=False)]
    base = rs.uniform(size=(50, 2))
    pts = [np.vstack([X, base, base])]
    Z = [rs.normal(size=(len(pts[0]), 1))]
    engine.StepPlan.clear()
    monkeypatch.setattr(engine, 'graphs_enabled', lambda: True)
    got = batched.mts_step(gps, pts, [40], Z)
    jit = gps[0].gp_core.last_sample_jitter
    monkeypatch.setattr(engine, 'graphs_enabled', lambda: False)
    ref = batched.mts_step(gps, pts, [40], Z)
    assert got == ref and jit == gps[0].gp_core.last_sample_jitter
    assert jit is not None


def test_lml_plan_equals_eager_and_handles_the_ladder(monkeypatch):
    """Batched LML over hyper-parameter candidates: replayed graph == eager launches, also for
    candidates whose Gram matrix only factors with jitter (duplicated inputs, vanishing noise)."""
    from ocbo_b200 import engine
    from ocbo_b200.gp import hp_search
    from oracle import gp_numpy as o
    rs = np.random.RandomState(4)
    X = rs.uniform(size=(70, 3))
    X[10:20] = X[:10]                       # duplicates: K is singular without the noise term
    y = np.sin(4 * X.sum(1))
    cands = hp_search.hp_candidates(X, y, 16, uniforms=rs.uniform(size=(15, 6)))
    cands[3, 1] = 1e-18                     # noise far below the ladder's first rung
    cands[7, 1] = 1e-18
    engine.StepPlan.clear()
    monkeypatch.setattr(engine, 'graphs_enabled', lambda: True)
    a1 = hp_search.batched_lml(X, y, 0, cands)[0]
    a2 = hp_search.batched_lml(X, y, 0, cands)[0]
    monkeypatch.setattr(engine, 'graphs_enabled', lambda: False)
    b = hp_search.batched_lml(X, y, 0, cands)[0]
    assert np.array_equal(a1, a2) and np.array_equal(a1, b)
    ref = np.asarray([o.log_marginal_likelihood(X, y, 'se', th) for th in cands])
    ok = [i for i in range(16) if i not in (3, 7)]
    assert np.max(np.abs(a1[ok] - ref[ok]) / np.maximum(1.0, np.abs(ref[ok]))) <= 1e-8
    assert np.all(np.isfinite(a1))          # the jittered candidates got a factor too
