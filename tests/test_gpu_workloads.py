"""GPU parity on what bench.py actually runs, and on the branches round 1 never exercised on the GPU.

Every test compares the CUDA path (C ABI, include/h1ilqr.h) with the CPU oracle on the same inputs: accept / reject
decisions and iteration counts identical, per-iteration cost and final x / u within 1e-6 relative (BASELINE.json),
cost derivatives within 1e-9. Instances whose solve forks at a decision the oracle itself marks as a near-tie
(helpers.compare_solve) are reported and counted, never absorbed by a looser tolerance.

  test_bench_workload_parity      config 5 instances exactly as bench.py builds them (workloads.walking_instances), solved
                                  inside a 1024-instance batch so that AUTO selects the BATCHED kernel family
  test_config3_parity             config 3: 1024 perturbed standing instances, 16 of them against the oracle
  test_config2_closed_loop        config 2: walking closed loop, 30 MPC steps from the standing pose (what main does) and
                                  35 steps across the end of the reference table (window clamp, robot_utils.cpp:430-441)
  test_horizon_200                config 4 at N = 200: a start the oracle converges from, and the diverging cold start
  test_com_velocity_term          W_com_vel > 0 with the real CoM-velocity targets (ilqr.cpp:159-161, 675-695), Q13
  test_llt_fallback               indefinite Quu -> +1e-4 I once, pivoted LDL^T on an indefinite matrix (Q9, ilqr.cpp:275-281)
  test_aerial_phase               (0,0) stance rows: balance skipped, both feet tracked (ilqr.cpp:769-775)
  test_warm_start_baseline        line-search baseline after a double failure in iteration 0 of a warm start (ilqr.cpp:318)
"""
import os

import numpy as np
import pytest

from helpers import (ROOT, compare_solve, grav_comp_guess, make_oracle, oracle_kinematics, oracle_solves, po,
                     reference_set, rel_err, standing_state)
from mpc_ilqr_mujoco_b200 import Config
from mpc_ilqr_mujoco_b200 import workloads as wl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from mpc_ilqr_mujoco_b200 import gpu as g
    g.lib()  # raises if the CUDA extension is missing: no fallback
    return g


def _check_sample(sg, solvers, sample, x_g, u_g, label, max_tie_frac=0.25):
    cost = sg.last_cost
    st, iters = sg.get_status()
    ct, at = sg.solve_trace()
    ties = []
    for k, i in enumerate(sample):
        r = compare_solve(solvers[k], cost[i], iters[i], ct[i], at[i], x_g[i], u_g[i], label=f"{label}[{i}]")
        if r == "near_tie":
            ties.append(int(i))
    print(f"{label}: {len(sample) - len(ties)} of {len(sample)} sampled instances match the oracle decision for decision; "
          f"near-ties (forked at a decision within 1e-6 x cost of its threshold): {ties}")
    assert len(ties) <= max_tie_frac * len(sample), ties
    return ties


def test_bench_workload_parity(gpu, oracle):
    """The instances bench.py times (BASELINE config 5: window row t0 = i mod 374, perturbed x_ref[t0], per-instance
    windows, cold start), 1024 of them in one batch (AUTO -> BATCHED family, as in the bench), 32 compared with the
    oracle over the full solve."""
    w = Config().build_weights()
    B = 1024
    ids = np.arange(0, 8192, 8)                 # spread over the whole 8192-instance shard of rank 0
    sg = gpu.H1IlqrBatch(w, N=25, batch=B)
    win, x0, t0 = wl.walking_instances(ids, sg.reference_kinematics)
    assert len(np.unique(t0)) > 180
    sg.set_reference_window(*win, shared=False)
    ug = grav_comp_guess(standing_state())
    sg.initialize(x0, None, ug)
    sg.last_cost, _, _ = sg.solve(x0)
    kernels = sg.stage_times()
    assert kernels["launches"] > 0
    xg, ugp = sg.get_trajectory()
    sample = np.arange(0, B, 32)
    # the oracle gets its windows from ITS OWN kinematics (independent FK), not from the GPU's
    owin, ox0, _ = wl.walking_instances(ids[sample], oracle_kinematics)
    assert np.abs(ox0 - x0[sample]).max() == 0.0
    for a, b in zip(owin, win):
        assert np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b[sample], dtype=np.float64)).max() < 1e-12
    solvers = oracle_solves(w, owin, ox0, ug)
    _check_sample(sg, solvers, sample, xg, ugp, "config5")


def test_config3_parity(gpu, oracle):
    """BASELINE config 3: 1024 perturbed standing instances on one GPU; 16 of them against the oracle."""
    w = Config().build_weights()
    B = 1024
    sg = gpu.H1IlqrBatch(w, N=25, batch=B)
    jr = np.array(gpu.default_dynamics_model().jnt_range)
    win, x0 = wl.standing_instances(np.arange(B), sg.reference_kinematics, jnt_range=jr)
    sg.set_reference_window(*win, shared=True)
    ug = grav_comp_guess(standing_state())
    sg.initialize(x0, None, ug)
    sg.last_cost, _, status = sg.solve(x0)
    assert (status == 0).all()
    xg, ugp = sg.get_trajectory()
    sample = np.arange(5, B, 64)
    owin, _ = wl.standing_instances(sample, oracle_kinematics, jnt_range=jr)
    owins = tuple(np.stack([a] * len(sample)) for a in owin)
    solvers = oracle_solves(w, owins, x0[sample], ug)
    _check_sample(sg, solvers, sample, xg, ugp, "config3")


def _closed_loop(gpu, oracle, tag, t_begin, steps, x_start, noise=None, resync=False):
    """MPC closed loop on both sides (plant = f_D of each side). Returns the number of steps compared before a
    near-tie fork (== steps when none occurred). resync: after every step the GPU side takes over the ORACLE's plant
    state, previous solution and lambda, so that every MPC step of the loop is compared on identical inputs
    (measured state, window, warm start) instead of on inputs that carry the accumulated differences of an
    ill-conditioned optimisation (see test_config2_closed_loop)."""
    so, w, _ = make_oracle(tag)
    sg = gpu.H1IlqrBatch(w, N=25, batch=1)
    refs = reference_set(tag)
    xo = x_start.copy(); xg = x_start.copy()
    ug = grav_comp_guess(standing_state())
    rng = np.random.default_rng(7)
    for k in range(steps):
        win = refs.window(t_begin + k, 25)
        so.set_reference_window(*win); sg.set_reference_window(*win, shared=True)
        if noise and k > 0:       # the same measurement disturbance on both sides: x_measured leaves the prediction
            d = np.zeros(51)
            d[7:26] = rng.uniform(-noise, noise, 19); d[26:] = rng.uniform(-10 * noise, 10 * noise, 25)
            xo = xo + d; xg = xg + d
        uo, co = so.mpc_step(xo, ug)
        ugp, cg = sg.mpc_step(xg[None], ug)
        ct, at = sg.solve_trace()
        st, it = sg.get_status()
        xt, ut = sg.get_trajectory()
        r = compare_solve(so, cg[0], it[0], ct[0], at[0], xt[0], ut[0], label=f"{tag} closed loop step {k} (t_idx {t_begin + k})")
        if r == "near_tie":
            print(f"{tag} closed loop: near-tie fork at step {k}; {k} steps compared")
            return k
        assert np.abs(ugp[0] - uo).max() <= 1e-6 * max(np.abs(uo).max(), 1.0), k
        xo = oracle.dyn_step(xo, uo)[0]
        xg = sg.dynamics_step(xg[None], ugp)[0]
        assert np.abs(xg - xo).max() <= 1e-6 * max(np.abs(xo).max(), 1.0), k
        if resync:
            xg = xo.copy()
            sg.set_previous_solution(so.get("xbar")[None], so.get("ubar")[None])
            sg.set_regularization(so.get_lambda())
    return steps


def test_config2_closed_loop(gpu, oracle):
    """BASELINE config 2: walking reference + contact_walking schedule, N = 25, warm-started MPC steps.
    (a) 30 steps from the standing pose at t_idx 0, as main/humanoid_mpc.cpp runs it;
    (b) 35 steps from x_ref[360]: from t_idx 375 on the window rows clamp at the last reference row while the
        schedule / foot-target lookups stay horizon-local (robot_utils.cpp:430-441, quirk Q6). This start is a hard
        one: cost 2e4, most line searches fail, feedback gains up to 2e5 (Quu is close to singular). Every stage
        agrees with the oracle to 1e-11 or better on identical inputs at every step of it (tools/diag_closed_loop.py,
        profiles/r02_diag_closed_loop.txt), but the optimisation itself amplifies input differences by up to 100x per
        MPC step, so two free-running fp64 implementations drift apart (3e-11 after 3 steps, 2e-8 after 12, 2e-6 after
        13). The segment is therefore compared step by step on identical inputs (resync)."""
    refs = reference_set("walking")
    assert refs.T == 400
    n = _closed_loop(gpu, oracle, "walking", 0, 30, standing_state())
    assert n >= 15
    n = _closed_loop(gpu, oracle, "walking", 360, 35, refs.x_ref_full[360].copy(), resync=True)
    assert n >= 20


def test_horizon_200(gpu, oracle):
    """BASELINE config 4 at its longest horizon. (a) converging: the guess is a 200-step trajectory that stays upright
    (tests/golden/n200_guess.npz, tools/make_n200_guess.py); (b) diverging: the reference's constant cold-start guess lets
    the robot fall within the 4 s horizon, the capture-point term takes the square root of a negative CoM height and the
    initial cost is NaN — every line search fails and the solve stops after three iterations on both sides."""
    N = 200
    g = np.load(os.path.join(ROOT, "tests", "golden", "n200_guess.npz"))
    so, w, win = make_oracle("walking", N=N)
    sg = gpu.H1IlqrBatch(w, N=N, batch=1)
    sg.set_reference_window(*win, shared=True)
    x0 = g["x0"]; U = g["U"]
    so.set("ubar", U); so.rollout_nominal(x0)
    assert rel_err(so.get("xbar"), g["X"]) < 1e-9
    sg.set_trajectory(xbar=so.get("xbar")[None], ubar=U[None])
    co = so.solve(x0)
    assert np.isfinite(co) and so.iters() >= 2
    cg, it, st = sg.solve(x0[None])
    ct, at = sg.solve_trace()
    xg, ugp = sg.get_trajectory()
    # Conditioning of this problem: 200 knots of closed-loop candidate rollouts amplify an input change by ~1e8. The
    # oracle's OWN final trajectory moves by `probe` when its initial controls move by 1e-13 relative (the size of the
    # f_D agreement between two fp64 implementations); decisions and per-iteration costs must still agree to 1e-6, the
    # first 50 knots of the trajectory too, the whole trajectory to within that envelope.
    s2, _, _ = make_oracle("walking", N=N)
    rng = np.random.default_rng(1)
    s2.set("ubar", U * (1 + 1e-13 * rng.standard_normal(U.shape))); s2.rollout_nominal(x0); s2.solve(x0)
    probe = max(rel_err(s2.get("xbar"), so.get("xbar")), rel_err(s2.get("ubar"), so.get("ubar")))
    r = compare_solve(so, cg[0], it[0], ct[0], at[0], xg[0], ugp[0], label="N=200 converging", tol_xu=max(1e-6, probe))
    assert rel_err(xg[0, :51], so.get("xbar")[:51]) < 1e-6 and rel_err(ugp[0, :50], so.get("ubar")[:50]) < 1e-6
    print("N=200 converging:", r, "iters", so.iters(), "cost", co, "x rel", rel_err(xg[0], so.get("xbar")), "u rel",
          rel_err(ugp[0], so.get("ubar")), "oracle self-response to a 1e-13 input change", probe)
    # (b) the diverging cold start
    ug = grav_comp_guess(standing_state())
    xs = wl.perturb(reference_set("walking").x_ref_full[0], 0)
    so.mpc_reset(); so.initialize(xs, False, ug)
    co = so.solve(xs)
    sg.mpc_reset(); sg.initialize(xs[None], None, ug)
    cg, it, st = sg.solve(xs[None])
    ct_o, at_o = so.trace(); ct, at = sg.solve_trace()
    assert not np.isfinite(co) and not np.isfinite(cg[0])
    assert it[0] == so.iters() == 3 and (at[0] == at_o).all()


def test_com_velocity_term(gpu, oracle):
    """W_com_vel > 0 (0 as shipped, so round 1 never ran this branch): CoM-velocity targets computed per reference row
    on the GPU (h1ilqr_reference_com_velocity) vs the oracle's, cost quadratics with the term on every knot but the
    terminal one (Q13), and a full solve."""
    cfg = Config()
    cfg.mpc.costs.W_com_vel = 35.0
    w = cfg.build_weights()
    assert w.w_com_vel == 35.0
    sg = gpu.H1IlqrBatch(w, N=25, batch=1)
    d = np.load(os.path.join(ROOT, "data", "h1_refs.npz"))
    xr = np.hstack([d["walking_q"], d["walking_v"]])
    cv_g = sg.reference_com_velocity(xr)
    cv_o = np.array([po.dyn_com_vel(r) for r in xr])
    assert np.abs(cv_o).max() > 0.05 and np.abs(cv_g - cv_o).max() < 1e-12
    refs = wl.reference_set("walking", sg.reference_kinematics, sg.reference_com_velocity)
    refs.require_com_velocity(w)
    win = refs.window(40, 25)
    assert np.abs(win[5]).max() > 0.01
    so = po.OracleSolver(w, 25)
    so.set_reference_window(*win); sg.set_reference_window(*win, shared=True)
    rng = np.random.default_rng(21)
    xb = win[0] + rng.normal(size=win[0].shape) * 0.05
    xb[:, 7:26] += rng.uniform(-0.6, 0.6, (26, 19))      # some joints beyond the inner 80 % of their range, also at the terminal knot
    ub = rng.uniform(-250, 250, (25, 19))
    so.set("xbar", xb); so.set("ubar", ub); so.cost_quadratics()
    sg.set_trajectory(xbar=xb[None], ubar=ub[None]); sg.cost_quadratics()
    lx, lu, lxx, luu = sg.get_cost_quadratics()
    for t in range(26):
        assert rel_err(lx[0, t], so.get("lx")[t]) < 1e-9 and rel_err(lxx[0, t], so.get("lxx")[t]) < 1e-9, t
    # the term is really there: without the weight the derivatives differ
    w0 = Config().build_weights()
    s0 = po.OracleSolver(w0, 25); s0.set_reference_window(*win); s0.set("xbar", xb); s0.set("ubar", ub); s0.cost_quadratics()
    assert rel_err(s0.get("lx")[:25], so.get("lx")[:25]) > 1e-3
    assert rel_err(s0.get("lx")[25], so.get("lx")[25]) == 0.0     # Q13: not at the terminal knot
    x0 = win[0][0].copy(); ug = grav_comp_guess(standing_state())
    so.initialize(x0, False, ug); so.solve(x0)
    sg.initialize(x0[None], None, ug)
    cg, it, st = sg.solve(x0[None])
    ct, at = sg.solve_trace(); xg, ugp = sg.get_trajectory()
    print("com-vel solve:", compare_solve(so, cg[0], it[0], ct[0], at[0], xg[0], ugp[0], label="W_com_vel"))


def test_llt_fallback(gpu, oracle):
    """Q9 (ilqr.cpp:275-281): when LLT of Quu + lambda I fails the reference adds 1e-4 I ONCE, without re-testing, and
    solves with the pivoted LDL^T whatever the signs of the pivots. The kernel infers "LLT fails" from the pivots of its
    own LDL^T; here Quu is made clearly indefinite at some knots through the cost Hessians (luu), and K / kff must match
    the oracle (which runs a real Cholesky test) to 1e-8."""
    so, w, win = make_oracle("walking")
    sg = gpu.H1IlqrBatch(w, N=25, batch=2)
    sg.set_reference_window(*win, shared=True)
    x0 = standing_state(); ug = grav_comp_guess(x0)
    so.initialize(x0, False, ug); so.rollout_nominal(x0); so.linearize(); so.cost_quadratics()
    luu = so.get("luu").copy()
    for t, j, v in ((24, 2, -1000.0), (24, 7, -30.0), (17, 0, -500.0), (9, 11, -2000.0), (3, 18, -50.0)):
        luu[t, j, j] = v
    so.set("luu", luu)
    so.backward_pass()
    assert np.isfinite(so.get("K")).all()
    # reference run without the negative curvature: the fallback must have changed the gains
    s2, _, _ = make_oracle("walking"); s2.initialize(x0, False, ug); s2.rollout_nominal(x0); s2.linearize(); s2.cost_quadratics(); s2.backward_pass()
    assert rel_err(so.get("K"), s2.get("K")) > 1e-2
    two = lambda a: np.stack([a, a])
    sg.set_trajectory(xbar=two(so.get("xbar")), ubar=two(so.get("ubar")))
    sg.set_linearization(two(so.get("A")), two(so.get("B")))
    luu2 = np.stack([luu, s2.get("luu")])        # instance 1 keeps the positive-definite Hessians
    sg.set_cost_quadratics(two(so.get("lx")), two(so.get("lu")), two(so.get("lxx")), luu2)
    sg.set_regularization(so.get_lambda())
    for policy in (1, 2):
        sg.set_kernel_policy(policy)
        sg.backward_pass()
        K, kff = sg.get_gains()
        assert rel_err(K[0], so.get("K")) < 1e-8 and rel_err(kff[0], so.get("kff")) < 1e-8, policy
        assert rel_err(K[1], s2.get("K")) < 1e-8 and rel_err(kff[1], s2.get("kff")) < 1e-8, policy


def test_aerial_phase(gpu, oracle):
    """Rows with both feet off the ground ((0,0) in the schedule): the balance term is skipped in the cost and in the
    derivatives (ilqr.cpp:769-775, 403-437), both feet get position targets and neither a velocity term."""
    so, w, win = make_oracle("walking")
    win = list(win)
    stance = win[4].copy()
    stance[5:12] = 0            # aerial
    stance[12:15] = (1, 0)
    stance[25] = 0              # aerial terminal knot
    win[4] = stance
    so.set_reference_window(*win)
    sg = gpu.H1IlqrBatch(w, N=25, batch=1)
    sg.set_reference_window(*win, shared=True)
    rng = np.random.default_rng(3)
    xb = win[0] + rng.normal(size=win[0].shape) * 0.03
    ub = rng.uniform(-40, 40, (25, 19))
    so.set("xbar", xb); so.set("ubar", ub); so.cost_quadratics()
    sg.set_trajectory(xbar=xb[None], ubar=ub[None]); sg.cost_quadratics()
    lx, lu, lxx, luu = sg.get_cost_quadratics()
    for t in range(26):
        assert rel_err(lx[0, t], so.get("lx")[t]) < 1e-9 and rel_err(lxx[0, t], so.get("lxx")[t]) < 1e-9, t
    assert abs(sg.total_cost()[0] - so.total_cost()) <= 1e-12 * abs(so.total_cost())
    for policy in (1, 2):
        x0 = win[0][0].copy(); ug = grav_comp_guess(standing_state())
        so.mpc_reset(); so.initialize(x0, False, ug); so.solve(x0)
        sg.set_kernel_policy(policy)
        sg.mpc_reset(); sg.initialize(x0[None], None, ug)
        cg, it, st = sg.solve(x0[None])
        ct, at = sg.solve_trace(); xg, ugp = sg.get_trajectory()
        print("aerial solve, policy", policy, compare_solve(so, cg[0], it[0], ct[0], at[0], xg[0], ugp[0], label="aerial"))


def test_warm_start_baseline(gpu, oracle):
    """Regression test for the line-search baseline of a warm start (ilqr.cpp:318: baseline = computeTotalCost of the
    trajectory AFTER forwardRolloutNominal). With a measured state that left the prediction, the shifted warm-start
    trajectory is not dynamics-consistent, so its cost differs from the cost after the re-rollout from x0. All alphas
    are made tiny here, so every candidate costs (almost exactly) the baseline and must be REJECTED (improvement below
    the 1e-6 margin): a solver that compares against the pre-rollout cost instead accepts one as soon as that cost is
    the higher of the two. 8 disturbed instances; the oracle's pre- and post-rollout costs are checked to differ."""
    w = Config().build_weights()
    B = 8
    opt = gpu.default_options()
    for i in range(8):
        opt.alphas[i] = 1e-13
    oopt = po.default_options()
    for i in range(8):
        oopt.alphas[i] = 1e-13
    refs = reference_set("standing")
    rng = np.random.default_rng(11)
    for policy in (1, 2):
        sg = gpu.H1IlqrBatch(w, N=25, batch=B, options=opt)
        sg.set_kernel_policy(policy)
        sg.set_reference_window(*refs.window(0, 25), shared=True)
        x0 = np.stack([standing_state()] * B); ug = grav_comp_guess(standing_state())
        ua, c0 = sg.mpc_step(x0, ug)                      # step 0: cold (all candidates rejected, guess kept)
        x1 = sg.dynamics_step(x0, ua)
        x1[:, 7:26] += rng.uniform(-2e-3, 2e-3, (B, 19)); x1[:, 26:] += rng.uniform(-2e-2, 2e-2, (B, 25))
        sg.set_reference_window(*refs.window(1, 25), shared=True)
        ua1, c1 = sg.mpc_step(x1, ug)                     # step 1: warm start from a disturbed measurement
        ct, at = sg.solve_trace()
        st, it = sg.get_status()
        higher = 0
        for i in range(B):
            so = po.OracleSolver(w, 25, options=oopt)
            so.set_reference_window(*refs.window(0, 25))
            so.mpc_step(x0[i], ug)
            so.set_reference_window(*refs.window(1, 25))
            so.initialize(x1[i], True, ug)
            pre = so.total_cost()
            so.rollout_nominal(x1[i])
            post = so.total_cost()
            higher += pre > post + 1e-4
            so.initialize(x1[i], True, ug)
            co = so.solve(x1[i])
            ct_o, at_o = so.trace()
            assert (at_o[:3] == -1).all() and so.iters() == 3
            assert (at[i] == at_o).all(), (policy, i, at[i][:3].tolist(), pre, post)
            assert it[i] == 3 and abs(c1[i] - co) <= 1e-9 * abs(co)
        assert higher >= 1, "no instance exercises the case pre-rollout cost > post-rollout cost"


def test_full_weight_matrices(gpu, oracle):
    """Whole symmetric Q / R / Qf (RobotUtils::setCostWeights keeps matrices and iLQR multiplies them: lx = Q (x - x_ref),
    lxx = Q, lu = R (u - u_ref), luu = R, 0.5 e'Qe in the line-search cost; ilqr.cpp:145-150, 372-373, 441): cost quadratics,
    total cost, line search and a full solve against the oracle on both kernel families; asymmetric matrices are refused;
    NULL restores the diagonal weights."""
    w = Config().build_weights()
    rng = np.random.default_rng(5)

    def spd(diag, scale):
        n = len(diag)
        M = rng.normal(size=(n, n)) * scale
        M = 0.5 * (M + M.T)
        np.fill_diagonal(M, 0.0)
        return M + np.diag(np.asarray(diag) + np.abs(M).sum(axis=1))      # diagonally dominant -> positive definite
    Q, R, Qf = spd(np.array(w.Qdiag), 1.5), spd(np.array(w.Rdiag), 2e-4), spd(np.array(w.Qfdiag), 3.0)
    so, _, win = make_oracle("walking")
    so.set_weight_matrices(Q, R, Qf)
    sg = gpu.H1IlqrBatch(w, N=25, batch=2)
    sg.set_reference_window(*win, shared=True)
    with pytest.raises(gpu.H1IlqrError):
        sg.set_weight_matrices(Q + np.triu(np.ones((51, 51)), 1), R, Qf)
    sg.set_weight_matrices(Q, R, Qf)
    xb = win[0] + rng.normal(size=win[0].shape) * 0.05
    ub = rng.uniform(-60, 60, (25, 19))
    so.set("xbar", xb); so.set("ubar", ub); so.cost_quadratics()
    two = lambda a: np.stack([a, a])
    sg.set_trajectory(xbar=two(xb), ubar=two(ub)); sg.cost_quadratics()
    lx, lu, lxx, luu = sg.get_cost_quadratics()
    for t in range(26):
        assert rel_err(lx[0, t], so.get("lx")[t]) < 1e-9 and rel_err(lxx[0, t], so.get("lxx")[t]) < 1e-9, t
    assert rel_err(lu[0], so.get("lu")) < 1e-9 and rel_err(luu[0], so.get("luu")) < 1e-9
    assert np.abs(luu[0, 3] - luu[0, 3].T).max() == 0.0 and np.abs(luu[0, 3, 0, 1]) > 0
    x0 = win[0][0].copy(); ug = grav_comp_guess(standing_state())
    for policy in (1, 2):
        sg.set_kernel_policy(policy)
        sg.set_trajectory(xbar=two(xb), ubar=two(ub))
        so.set("xbar", xb); so.set("ubar", ub)
        assert abs(sg.total_cost()[0] - so.total_cost()) <= 1e-12 * abs(so.total_cost()), policy
        so.mpc_reset(); so.initialize(x0, False, ug); so.solve(x0)
        sg.mpc_reset(); sg.initialize(two(x0), None, ug)
        cg, it, st = sg.solve(two(x0))
        ct, at = sg.solve_trace(); xg, ugp = sg.get_trajectory()
        print("full Q/R/Qf solve, policy", policy, compare_solve(so, cg[0], it[0], ct[0], at[0], xg[0], ugp[0], label="full weights"))
        assert cg[0] == cg[1]
    # the off-diagonal parts matter, and NULL brings the diagonal weights back
    s0, _, _ = make_oracle("walking"); s0.set("xbar", xb); s0.set("ubar", ub)
    assert abs(s0.total_cost() - so.total_cost()) > 1e-3 * abs(so.total_cost())
    sg.set_weight_matrices(None, None, None); sg.set_weights(w)
    sg.set_trajectory(xbar=two(xb), ubar=two(ub))
    assert abs(sg.total_cost()[0] - s0.total_cost()) <= 1e-12 * abs(s0.total_cost())
