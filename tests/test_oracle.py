"""CPU tests of the oracle (test infrastructure): known answers, derivative self-consistency, regression against
the committed golden vectors. The reference ships no tests or golden vectors, so the oracle is pinned by
closed-form values computed independently from the model files and by these checks."""
import os

import numpy as np
import pytest

from helpers import ROOT, grav_comp_guess, make_oracle, standing_state


def rand_state(rng, unit=True):
    x = np.zeros(51)
    x[:3] = rng.uniform(-1, 1, 3); x[2] += 1.0
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    if not unit:
        q *= rng.uniform(0.9, 1.1)
    x[3:7] = q
    x[7:26] = rng.uniform(-0.6, 0.6, 19)
    x[26:] = rng.uniform(-1, 1, 25)
    return x


def test_known_answers_appendix_e(oracle):
    x = standing_state()
    assert np.allclose(oracle.dyn_com(x), [0.0162336940, 0.000967503071, 1.00406373], atol=5e-9)
    assert np.allclose(oracle.dyn_body_pos(x, 5), [0.039468, 0.20286, 0.069], atol=1e-12)
    assert np.allclose(oracle.dyn_body_pos(x, 10), [0.039468, -0.20286, 0.069], atol=1e-12)
    _, g, _ = oracle.cost_term(oracle.TERM_COM, x, [0, 0, 0], 1.0, mode=2)  # gradient of |com|^2 w.r.t. p_b = 2 com
    assert np.allclose(g[:3] / 2, [0.0164049953, 0.000968419856, 1.00733595], atol=5e-9)
    assert abs(sum(oracle.dynamics_model().mass) - 51.649896) < 1e-9
    assert abs(sum(oracle.cost_model().mass) - 51.601) < 1e-9
    b = oracle.dyn_bias(x)
    assert abs(b[2] - 51.649896 * 1.0) < 1e-9  # weight under the shipped gravity (0,0,-1)


@pytest.mark.parametrize("term", range(6))
def test_cost_term_analytic_equals_ad(oracle, term):
    rng = np.random.default_rng(100 + term)
    for trial in range(3):
        x = rand_state(rng, unit=False)
        tgt, w, ee = rng.uniform(-0.5, 0.5, 3), rng.uniform(1, 100), trial % 2
        _, g1, H1 = oracle.cost_term(term, x, tgt, w, ee, 1)
        _, g2, H2 = oracle.cost_term(term, x, tgt, w, ee, 2)
        assert np.abs(g1 - g2).max() <= 1e-12 * max(np.abs(g1).max(), 1e-300)
        assert np.abs(H1 - H2).max() <= 1e-12 * max(np.abs(H1).max(), 1e-300)
        assert np.abs(H2 - H2.T).max() <= 1e-10 * np.abs(H2).max()


@pytest.mark.parametrize("term", range(6))
def test_cost_term_gradient_matches_central_differences(oracle, term):
    rng = np.random.default_rng(200 + term)
    x = rand_state(rng, unit=False)
    tgt, w = rng.uniform(-0.5, 0.5, 3), 7.0
    _, g, H = oracle.cost_term(term, x, tgt, w, 0, 2)
    perm = list(range(51)); perm[3], perm[4], perm[5], perm[6] = 4, 5, 6, 3  # Pinocchio index -> MuJoCo index
    gfd = np.zeros(51)
    for i in range(51):
        e = np.zeros(51); e[perm[i]] = 1e-6
        gfd[i] = (oracle.cost_term(term, x + e, tgt, w, 0, 0)[0] - oracle.cost_term(term, x - e, tgt, w, 0, 0)[0]) / 2e-6
    assert np.abs(gfd - g).max() <= 1e-7 * np.abs(g).max()
    # Hessian column by central differences of the analytic gradient
    for i in (3, 8, 30):
        e = np.zeros(51); e[perm[i]] = 1e-6
        col = (oracle.cost_term(term, x + e, tgt, w, 0, 2)[1] - oracle.cost_term(term, x - e, tgt, w, 0, 2)[1]) / 2e-6
        assert np.abs(col - H[:, i]).max() <= 1e-6 * max(np.abs(H).max(), 1e-12)


def test_dynamics_properties(oracle):
    rng = np.random.default_rng(3)
    x = rand_state(rng); x[2] = 1.05
    u = rng.uniform(-30, 30, 19)
    xn = oracle.dyn_step(x, u)[0]
    assert abs(np.linalg.norm(xn[3:7]) - 1.0) < 1e-14            # quaternion renormalised
    xs = x.copy(); xs[3:7] *= 1.07
    assert np.abs(oracle.dyn_step(xs, u)[0] - xn).max() < 1e-12   # input quaternion scale is irrelevant
    xt = x.copy(); xt[0] += 3.0; xt[1] -= 2.0                     # horizontal translation invariance
    xnt = oracle.dyn_step(xt, u)[0]
    assert np.abs(xnt[2:] - xn[2:]).max() < 1e-12 and abs(xnt[0] - xn[0] - 3.0) < 1e-12
    uc = u.copy(); uc[3] = 1e4; ud = u.copy(); ud[3] = 300.0     # torque clamped to ctrlrange (knee: 300)
    assert np.abs(oracle.dyn_step(x, uc)[0] - oracle.dyn_step(x, ud)[0]).max() == 0.0


def test_linearization_ad_vs_fd(oracle):
    rng = np.random.default_rng(4)
    x = rand_state(rng); x[2] = 1.04
    u = rng.uniform(-30, 30, 19)
    A, B = oracle.dyn_linearize_ad(x, u)
    Af, Bf = oracle.dyn_linearize(x, u, 1e-6)
    assert np.abs(A - Af).max() <= 1e-4 * np.abs(A).max()
    assert np.abs(B - Bf).max() <= 1e-4 * np.abs(B).max()
    u[3] = 400.0
    _, Bc = oracle.dyn_linearize_ad(x, u)
    assert np.abs(np.asarray(Bc)[:, 3]).max() == 0.0  # clamped actuator: zero sensitivity (Q11)


def test_solve_monotone_and_first_accept(oracle):
    s, w, win = make_oracle("walking")
    x0 = standing_state(); ug = grav_comp_guess(x0)
    s.initialize(x0, False, ug)
    c = s.solve(x0)
    ct, at = s.trace()
    it = s.iters()
    assert 1 <= it <= 10 and np.isfinite(c)
    assert (np.diff(ct[:it]) <= 1e-12).all()          # accepted steps never increase the cost
    assert 1e-6 <= s.get_lambda() <= 1e-3              # lambda stays inside the reference's bounds
    # quirk Q1: the line-search cost omits the CoM / foot terms; with them switched off nothing changes
    from mpc_ilqr_mujoco_b200 import Config
    cfg = Config(); cfg.mpc.costs.W_com = 0.0; cfg.mpc.costs.W_foot = 0.0; cfg.mpc.costs.W_foot_vel = 0.0
    s2, _, _ = make_oracle("walking", cfg=cfg)
    s.mpc_reset(); s.initialize(x0, False, ug); s2.initialize(x0, False, ug)
    assert abs(s.total_cost() - s2.total_cost()) == 0.0


def test_golden_vectors_regression(oracle):
    g = np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"))
    assert np.abs(oracle.dyn_step(g["dyn_x"], g["dyn_u"]) - g["dyn_xnext"]).max() < 1e-12
    for tag in ("standing", "walking"):
        s, w, win = make_oracle(tag)
        x0 = standing_state(); ug = g[f"{tag}_u_guess"]
        assert np.abs(ug - grav_comp_guess(x0)).max() < 1e-12
        s.initialize(x0, False, ug)
        c = s.solve(x0)
        ct, at = s.trace()
        assert (at == g[f"{tag}_solve_alpha"]).all()
        assert abs(c - g[f"{tag}_solve_cost"][0]) <= 1e-9 * abs(c)
        assert np.abs(s.get("xbar") - g[f"{tag}_solve_xbar"]).max() < 1e-8


def test_com_velocity_matches_finite_difference_of_com(oracle):
    """dyn_com_vel (the J_subtreeCom * qvel target of loadReferences, robot_utils.cpp:388-397) = d/dt of dyn_com along
    the exact configuration flow (world-frame base velocity, body-frame angular velocity, hinge rates)."""
    rng = np.random.default_rng(0)
    x = np.zeros(51)
    x[2] = 1.0
    q = rng.normal(size=4); x[3:7] = q / np.linalg.norm(q)
    x[7:26] = rng.uniform(-0.5, 0.5, 19); x[26:] = rng.uniform(-1, 1, 25)

    def flow(x, eps):
        y = x.copy(); y[0:3] += eps * x[26:29]
        w = x[29:32]; ang = np.linalg.norm(w) * eps
        b = np.array([np.cos(ang / 2), *(np.sin(ang / 2) * w / np.linalg.norm(w))]); a = x[3:7]
        y[3:7] = [a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                  a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]]
        y[7:26] += eps * x[32:]
        return y
    e = 1e-6
    fd = (oracle.dyn_com(flow(x, e)) - oracle.dyn_com(flow(x, -e))) / (2 * e)
    assert np.abs(fd - oracle.dyn_com_vel(x)).max() < 1e-9


def test_decision_margins_and_compare_solve(oracle):
    """The oracle's decision margins (test diagnostics) are consistent with its trace, and compare_solve accepts an
    identical solve, flags a forked one as a real mismatch when the margin is wide, and as a near-tie when it is not."""
    from helpers import compare_solve, grav_comp_guess, make_oracle, standing_state
    so, w, win = make_oracle("walking")
    x0 = standing_state(); ug = grav_comp_guess(x0)
    so.initialize(x0, False, ug); c = so.solve(x0)
    ct, at = so.trace(); lm, sm = so.margins(); it = so.iters()
    for k in range(it):
        assert lm[k][0] >= 0 and (lm[k][1] >= 0) == (at[k][1] != -2)
    assert (lm[it:] == -1).all() and (sm[it:] == -1).all()
    x, u = so.get("xbar"), so.get("ubar")
    assert compare_solve(so, c, it, ct, at, x, u) == "match"
    bad = at.copy(); bad[1][0] = (bad[1][0] + 1) % 8
    with pytest.raises(AssertionError):
        compare_solve(so, c, it, ct, bad, x, u)     # wide margin: a fork there is a real mismatch
    with pytest.raises(AssertionError):
        compare_solve(so, c * (1 + 1e-5), it, ct, at, x, u)
