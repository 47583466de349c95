"""GPU parity tests: the CUDA path (through the C ABI, include/h1ilqr.h) against the CPU oracle on the same
seeded inputs. Tolerances follow BASELINE.json: 1e-9 relative on cost derivatives and (analytic) A/B, 1e-6 relative
on per-iteration cost and final trajectories; fp64 throughout.
"""
import numpy as np
import pytest

from helpers import grav_comp_guess, make_oracle, reference_set, standing_state
from mpc_ilqr_mujoco_b200 import Config
from mpc_ilqr_mujoco_b200.references import perturbed_states

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def gpu():
    from mpc_ilqr_mujoco_b200 import gpu as g
    g.lib()  # raises if the CUDA extension is missing: no fallback
    return g


def random_states(n, seed, spread=1.0):
    rng = np.random.default_rng(seed)
    x = np.zeros((n, 51))
    x[:, :3] = rng.uniform(-1, 1, (n, 3)) * spread
    x[:, 2] = 1.02 + rng.uniform(-0.03, 0.08, n)
    q = rng.normal(size=(n, 4)) * 0.1
    q[:, 0] += 1
    x[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    x[:, 7:26] = rng.uniform(-0.5, 0.5, (n, 19))
    x[:, 26:] = rng.uniform(-1, 1, (n, 25))
    u = rng.uniform(-60, 60, (n, 19))
    return x, u


def test_dynamics_step_parity(gpu, oracle):
    x, u = random_states(256, 11)
    x[::7, 3:7] *= 1.03  # un-normalised quaternions are renormalised inside f_D
    u[::5] *= 8.0        # beyond ctrlrange: clamped inside f_D
    s = gpu.H1IlqrBatch(Config().build_weights(), N=25, batch=1)
    xn = s.dynamics_step(x, u)
    xo = oracle.dyn_step(x, u)
    assert np.isfinite(xn).all()
    assert np.abs(xn - xo).max() < 1e-11


def test_bias_and_reference_kinematics(gpu, oracle):
    x, _ = random_states(64, 12)
    s = gpu.H1IlqrBatch(Config().build_weights(), N=25, batch=1)
    b = s.bias_forces(x)
    com, ee = s.reference_kinematics(x)
    for i in range(x.shape[0]):
        assert np.abs(b[i] - oracle.dyn_bias(x[i])).max() < 1e-10
        assert np.abs(com[i] - oracle.dyn_com(x[i])).max() < 1e-13
        assert np.abs(ee[i, 0] - oracle.dyn_body_pos(x[i], 5)).max() < 1e-13
        assert np.abs(ee[i, 1] - oracle.dyn_body_pos(x[i], 10)).max() < 1e-13


POLICIES = [1, 2]  # KERNELS_COOPERATIVE (warp per unit), KERNELS_BATCHED (thread per unit)


def _setup_pair(gpu, tag, N=25, batch=1, t0=0, lin=0, policy=0):
    so, w, win = make_oracle(tag, N=N, t0=t0, linearization=lin)
    opt = gpu.default_options()
    opt.linearization = lin
    sg = gpu.H1IlqrBatch(w, N=N, batch=batch, options=opt)
    sg.set_kernel_policy(policy)
    sg.set_reference_window(*win, shared=True)
    return so, sg, win


def test_rollout_and_total_cost_parity(gpu, oracle):
    so, sg, win = _setup_pair(gpu, "walking")
    x0 = win[0][0].copy()
    rng = np.random.default_rng(5)
    ub = rng.uniform(-20, 20, (25, 19))
    so.set("ubar", ub)
    so.rollout_nominal(x0)
    sg.set_trajectory(ubar=ub[None])
    sg.rollout_nominal(x0[None])
    xg, _ = sg.get_trajectory()
    assert rel(xg[0], so.get("xbar")) < 1e-10
    assert abs(sg.total_cost()[0] - so.total_cost()) / so.total_cost() < 1e-12


def test_linearize_fd_parity(gpu, oracle):
    """Forward-difference mode (the reference's method, eps = 1e-5). Two independent fp64 implementations of f_D
    agree to ~1e-13; forward differencing multiplies that by 1/eps = 1e5, so FD-vs-FD parity is bounded near
    1e-8 absolute. The 1e-9 claim is carried by the analytic mode (test_linearize_analytic_parity)."""
    so, sg, win = _setup_pair(gpu, "walking", lin=1)
    x0 = win[0][0].copy()
    rng = np.random.default_rng(6)
    ub = rng.uniform(-20, 20, (25, 19))
    so.set("ubar", ub); so.rollout_nominal(x0); so.linearize()
    sg.set_trajectory(xbar=so.get("xbar")[None], ubar=ub[None])
    sg.linearize()
    A, B = sg.get_linearization()
    Ao, Bo = so.get("A"), so.get("B")
    scale = max(np.abs(Ao).max(), np.abs(Bo).max())
    assert np.abs(A[0] - Ao).max() / scale < 1e-9
    assert np.abs(B[0] - Bo).max() / scale < 1e-9
    assert np.abs(A[0] - Ao).max() < 2e-7 and np.abs(B[0] - Bo).max() < 2e-7


@pytest.mark.parametrize("policy", POLICIES)
def test_linearize_analytic_parity(gpu, oracle, policy):
    """Analytic mode (default): exact A_k = df_D/dx, B_k = df_D/du. GPU tangent propagation with a shared
    factorisation vs the oracle's forward-mode AD through its own dense f_D: 1e-9 relative on A and on B
    separately (BASELINE.json), per knot. Includes clamped torques and un-normalised quaternions."""
    so, sg, win = _setup_pair(gpu, "walking", lin=0, policy=policy)
    x0 = win[0][3].copy()
    x0[3:7] *= 1.02
    rng = np.random.default_rng(16)
    ub = rng.uniform(-30, 30, (25, 19))
    ub[::4, 3] = 400.0  # beyond ctrlrange: zero column in B
    so.set("ubar", ub); so.rollout_nominal(x0); so.linearize()
    sg.set_trajectory(xbar=so.get("xbar")[None], ubar=ub[None])
    sg.linearize()
    A, B = sg.get_linearization()
    Ao, Bo = so.get("A"), so.get("B")
    for t in range(25):
        assert rel(A[0, t], Ao[t]) < 1e-9 and rel(B[0, t], Bo[t]) < 1e-9, t
    assert np.abs(B[0, 0, 3]).max() == 0.0
    # cross-check against the reference's forward-difference definition: agreement at the FD truncation level
    sf = gpu.H1IlqrBatch(Config().build_weights(), N=25, batch=1, options=_fd_options(gpu))
    sf.set_trajectory(xbar=so.get("xbar")[None], ubar=ub[None]); sf.linearize()
    Af, Bf = sf.get_linearization()
    assert rel(Af[0], A[0]) < 5e-3 and rel(Bf[0], B[0]) < 5e-3


def _fd_options(gpu):
    o = gpu.default_options()
    o.linearization = 1
    return o


def test_cost_quadratics_parity(gpu, oracle):
    for tag in ("standing", "walking"):
        so, sg, win = _setup_pair(gpu, tag)
        x_ref = win[0]
        rng = np.random.default_rng(7)
        xb = x_ref + rng.normal(size=x_ref.shape) * 0.05
        xb[:, 7:26] += rng.uniform(-0.6, 0.6, (26, 19))
        ub = rng.uniform(-250, 250, (25, 19))
        so.set("xbar", xb); so.set("ubar", ub); so.cost_quadratics()
        sg.set_trajectory(xbar=xb[None], ubar=ub[None]); sg.cost_quadratics()
        lx, lu, lxx, luu = sg.get_cost_quadratics()
        for t in range(26):
            assert rel(lx[0, t], so.get("lx")[t]) < 1e-9
            assert rel(lxx[0, t], so.get("lxx")[t]) < 1e-9
        assert rel(lu[0], so.get("lu")) < 1e-9
        assert rel(luu[0], so.get("luu")) < 1e-9
        # AD path of the oracle (CasADi-style exact derivatives) agrees as well
        so.use_ad(True); so.cost_quadratics(); so.use_ad(False)
        assert rel(lxx[0], so.get("lxx")) < 1e-9 and rel(lx[0], so.get("lx")) < 1e-9


def _prepare_iteration(so, x0, ug):
    so.initialize(x0, False, ug)
    so.rollout_nominal(x0); so.linearize(); so.cost_quadratics()


def test_backward_pass_parity(gpu, oracle):
    so, sg, win = _setup_pair(gpu, "walking")
    x0 = standing_state()
    _prepare_iteration(so, x0, grav_comp_guess(x0))
    so.backward_pass()
    sg.set_trajectory(xbar=so.get("xbar")[None], ubar=so.get("ubar")[None])
    sg.set_linearization(so.get("A")[None], so.get("B")[None])
    sg.set_cost_quadratics(so.get("lx")[None], so.get("lu")[None], so.get("lxx")[None], so.get("luu")[None])
    sg.set_regularization(so.get_lambda())
    sg.backward_pass()
    K, kff = sg.get_gains()
    assert rel(K[0], so.get("K")) < 1e-8
    assert rel(kff[0], so.get("kff")) < 1e-8


@pytest.mark.parametrize("policy", [0, 1, 2])
def test_riccati_structured_contractions(gpu, oracle, policy):
    """After an analytic linearization the backward pass contracts over the 29 rows of [A|B] that carry information
    (h1_riccati.cuh, STRUCT): (i) every analytic linearization kernel (AUTO at batch 1: thread per column; cooperative;
    batched) writes position rows that equal unit entry + dt * velocity row to the last bit or two; (ii) the structured
    pass, the dense pass on the same [A|B] (caller-supplied linearization -> dense) and the oracle agree."""
    so, sg, win = _setup_pair(gpu, "walking", policy=policy)
    x0 = win[0][3].copy()
    x0[3:7] *= 1.01
    ug = grav_comp_guess(standing_state())
    _prepare_iteration(so, x0, ug)
    so.backward_pass()
    sg.initialize(x0[None], None, ug)
    sg.rollout_nominal(x0[None]); sg.linearize(); sg.cost_quadratics()
    A, B = sg.get_linearization()
    dt = Config().mpc.dt
    AB = np.concatenate([A[0], B[0]], axis=1)            # [knot][column][row]
    for r in list(range(3)) + list(range(7, 26)):
        v = 26 + (r if r < 3 else r - 1)
        unit = np.zeros(70); unit[r] = 1.0
        want = unit + dt * AB[:, :, v]
        assert (np.abs(AB[:, :, r] - want) <= 4e-16 * np.maximum(1.0, np.abs(want))).all(), r
    sg.backward_pass()
    K1, k1 = sg.get_gains()
    sg.set_linearization(A, B)
    sg.backward_pass()
    K2, k2 = sg.get_gains()
    assert not (K1 == K2).all()                           # two different kernels ran
    assert rel(K1, K2) < 1e-8 and rel(k1, k2) < 1e-8
    assert rel(K1[0], so.get("K")) < 1e-8 and rel(k1[0], so.get("kff")) < 1e-8
    assert rel(K2[0], so.get("K")) < 1e-8 and rel(k2[0], so.get("kff")) < 1e-8


def test_line_search_parity(gpu, oracle):
    so, sg, win = _setup_pair(gpu, "walking")
    x0 = standing_state()
    _prepare_iteration(so, x0, grav_comp_guess(x0))
    so.backward_pass()
    sg.set_trajectory(xbar=so.get("xbar")[None], ubar=so.get("ubar")[None])
    sg.set_gains(so.get("K")[None], so.get("kff")[None])
    ok_o, cost_o, ai_o = so.line_search(x0)
    ok, nc, ai = sg.line_search(x0[None])
    assert bool(ok[0]) == ok_o and ai[0] == ai_o
    assert abs(nc[0] - cost_o) / abs(cost_o) < 1e-9
    xg, ug = sg.get_trajectory()
    assert rel(xg[0], so.get("xbar")) < 1e-9 and rel(ug[0], so.get("ubar")) < 1e-9


@pytest.mark.parametrize("policy", POLICIES)
@pytest.mark.parametrize("tag,lin", [("standing", 0), ("walking", 0), ("standing", 1)])
def test_solve_parity(gpu, oracle, tag, lin, policy):
    """Full iLQR::solve: identical accept/reject decisions, per-iteration cost and final x/u within 1e-6
    relative. lin=1 runs the reference's forward-difference linearization on both sides; its FD noise
    (see test_linearize_fd_parity) propagates to ~1e-5 on the controls, hence the looser bound there."""
    so, sg, win = _setup_pair(gpu, tag, lin=lin, policy=policy)
    x0 = standing_state()
    ug = grav_comp_guess(x0)
    so.initialize(x0, False, ug)
    co = so.solve(x0)
    sg.initialize(x0[None], None, ug)
    cg, iters, status = sg.solve(x0[None])
    ct_o, at_o = so.trace()
    ct_g, at_g = sg.solve_trace()
    assert status[0] == 0
    assert (at_g[0] == at_o).all(), (at_g[0].tolist(), at_o.tolist())
    assert iters[0] == so.iters()
    assert rel(ct_g[0], ct_o) < 1e-6
    assert abs(cg[0] - co) / abs(co) < 1e-6
    xg, ugp = sg.get_trajectory()
    tol_u = 1e-6 if lin == 0 else 5e-5
    assert rel(xg[0], so.get("xbar")) < 1e-6 and rel(ugp[0], so.get("ubar")) < tol_u
    assert abs(sg.get_regularization()[0] - so.get_lambda()) < 1e-18


def test_mpc_closed_loop_parity(gpu, oracle):
    """15 MPC steps (BASELINE config 1): warm starts, persistent lambda, plant = f_D."""
    so, sg, win = _setup_pair(gpu, "standing")
    refs = reference_set("standing")
    xo = standing_state(); xg = xo.copy()
    ug = grav_comp_guess(xo)
    for k in range(15):
        w = refs.window(k, 25)
        so.set_reference_window(*w); sg.set_reference_window(*w, shared=True)
        uo, co = so.mpc_step(xo, ug)
        ugp, cg = sg.mpc_step(xg[None], ug)
        assert abs(cg[0] - co) / abs(co) < 1e-6, k
        assert np.abs(ugp[0] - uo).max() / max(np.abs(uo).max(), 1.0) < 1e-6, k
        xo = oracle.dyn_step(xo, uo)[0]
        xg = sg.dynamics_step(xg[None], ugp)[0]
    assert np.abs(xg - xo).max() < 1e-6


@pytest.mark.parametrize("policy", POLICIES)
def test_batch_equals_looped_single(gpu, oracle, policy):
    """Independent instances: a batch must reproduce the per-instance single solves bit for bit (same kernel
    family on both sides; AUTO would pick the family by batch size)."""
    cfg = Config(); w = cfg.build_weights()
    refs = reference_set("walking")
    win = refs.window(0, 25)
    B = 6
    jr = np.array(gpu.default_dynamics_model().jnt_range)
    x0 = perturbed_states(standing_state(), B, seed=0, jnt_range=jr)
    ug = grav_comp_guess(standing_state())
    sb = gpu.H1IlqrBatch(w, N=25, batch=B)
    sb.set_kernel_policy(policy)
    sb.set_reference_window(*win, shared=True)
    sb.initialize(x0, None, ug)
    cb, ib, _ = sb.solve(x0)
    xb, ub = sb.get_trajectory()
    s1 = gpu.H1IlqrBatch(w, N=25, batch=1)
    s1.set_kernel_policy(policy)
    s1.set_reference_window(*win, shared=True)
    for i in range(B):
        s1.set_regularization(1e-6)
        s1.initialize(x0[i:i + 1], None, ug)
        c1, i1, _ = s1.solve(x0[i:i + 1])
        x1, u1 = s1.get_trajectory()
        assert c1[0] == cb[i] and i1[0] == ib[i]
        assert (x1[0] == xb[i]).all() and (u1[0] == ub[i]).all()
    # and the first instance agrees with the oracle
    so, _, _ = make_oracle("walking")
    so.initialize(x0[0], False, ug)
    co = so.solve(x0[0])
    assert abs(cb[0] - co) / abs(co) < 1e-6


def test_kernel_families_agree(gpu, oracle):
    """COOPERATIVE and BATCHED kernels compute the same numbers (different operation order only): stage by stage
    on a 37-instance walking batch with per-instance windows."""
    w = Config().build_weights()
    refs = reference_set("walking")
    B = 37
    wins = [refs.window(3 * i, 25) for i in range(B)]
    win = tuple(np.stack([wi[k] for wi in wins]) for k in range(6))
    jr = np.array(gpu.default_dynamics_model().jnt_range)
    x0 = np.stack([wi[0][0] for wi in wins]) + (perturbed_states(standing_state(), B, seed=3, jnt_range=jr) - standing_state())
    x0[:, 3:7] /= np.linalg.norm(x0[:, 3:7], axis=1, keepdims=True)
    ug = grav_comp_guess(standing_state())
    out = []
    for policy in POLICIES:
        s = gpu.H1IlqrBatch(w, N=25, batch=B)
        s.set_kernel_policy(policy)
        s.set_reference_window(*win, shared=False)
        s.initialize(x0, None, ug)
        s.rollout_nominal(x0); s.linearize(); s.cost_quadratics(); s.backward_pass()
        ok, c, a = s.line_search(x0)
        out.append((s.get_linearization(), s.get_cost_quadratics(), s.get_gains(), s.get_trajectory(), c, a))
    (A1, B1), cq1, (K1, k1), (x1, u1), c1, a1 = out[0]
    (A2, B2), cq2, (K2, k2), (x2, u2), c2, a2 = out[1]
    for i in range(B):
        for t in range(25):
            assert rel(A2[i, t], A1[i, t]) < 1e-10 and rel(B2[i, t], B1[i, t]) < 1e-10, (i, t)
    for p, q in zip(cq1, cq2):
        assert rel(q, p) < 1e-10
    assert rel(K2, K1) < 1e-7 and rel(k2, k1) < 1e-7
    assert (a1 == a2).all() and rel(c2, c1) < 1e-9 and rel(x2, x1) < 1e-8 and rel(u2, u1) < 1e-8


def test_horizon_sweep(gpu, oracle):
    """BASELINE config 4: N in {50, 100} (25 is covered above, 200 in the bench) with 8 concurrent alphas."""
    for N in (50, 100):
        so, sg, win = _setup_pair(gpu, "walking", N=N)
        x0 = standing_state(); ug = grav_comp_guess(x0)
        so.initialize(x0, False, ug); sg.initialize(x0[None], None, ug)
        opt_iters = 2  # two iterations are enough to cover every stage at this horizon
        for _ in range(opt_iters):
            so.rollout_nominal(x0); so.linearize(); so.cost_quadratics(); so.backward_pass()
            ok_o, c_o, a_o = so.line_search(x0)
            sg.rollout_nominal(x0[None]); sg.linearize(); sg.cost_quadratics(); sg.backward_pass()
            ok, c, a = sg.line_search(x0[None])
            assert a[0] == a_o and abs(c[0] - c_o) / abs(c_o) < 1e-6


def test_against_committed_golden_vectors(gpu):
    """GPU vs tests/golden/oracle_golden.npz (generated by tools/make_golden.py from the oracle in the build
    container) — fixed targets that do not depend on the oracle being rebuilt on the GPU box."""
    import os
    from helpers import ROOT
    g = np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"))
    w = Config().build_weights()
    s = gpu.H1IlqrBatch(w, N=25, batch=1)
    assert np.abs(s.dynamics_step(g["dyn_x"], g["dyn_u"]) - g["dyn_xnext"]).max() < 1e-11
    assert np.abs(s.bias_forces(g["dyn_x"]) - g["dyn_bias"]).max() < 1e-10
    com, _ = s.reference_kinematics(g["dyn_x"])
    assert np.abs(com - g["dyn_com"]).max() < 1e-13
    for tag in ("standing", "walking"):
        refs = reference_set(tag)
        s.set_reference_window(*refs.window(0, 25), shared=True)
        x0 = standing_state(); ug = g[f"{tag}_u_guess"]
        s.mpc_reset(); s.initialize(x0[None], None, ug)
        s.rollout_nominal(x0[None]); s.linearize(); s.cost_quadratics(); s.backward_pass()
        A, B = s.get_linearization(); lx, lu, lxx, luu = s.get_cost_quadratics(); K, kff = s.get_gains()
        for t in range(25):
            assert rel(A[0, t], g[f"{tag}_iter0_A"][t]) < 1e-9 and rel(B[0, t], g[f"{tag}_iter0_B"][t]) < 1e-9
            assert rel(lxx[0, t], g[f"{tag}_iter0_lxx"][t]) < 1e-9 and rel(lx[0, t], g[f"{tag}_iter0_lx"][t]) < 1e-9
        assert rel(K[0], g[f"{tag}_iter0_K"]) < 1e-8 and rel(kff[0], g[f"{tag}_iter0_kff"]) < 1e-8
        s.mpc_reset(); s.initialize(x0[None], None, ug)
        c, it, st = s.solve(x0[None])
        ct, at = s.solve_trace()
        assert (at[0] == g[f"{tag}_solve_alpha"]).all()
        assert rel(ct[0], g[f"{tag}_solve_trace"]) < 1e-6
        xg, ugp = s.get_trajectory()
        assert rel(xg[0], g[f"{tag}_solve_xbar"]) < 1e-6 and rel(ugp[0], g[f"{tag}_solve_ubar"]) < 1e-6


def test_full_size_batch_properties(gpu):
    """BASELINE-size batch (1024 standing instances, config 3): size-independent properties — every instance's
    accepted costs are non-increasing, every trajectory is dynamically consistent (a fresh rollout of the
    returned controls reproduces the returned states: bit for bit with the warp-cooperative kernels, to rounding
    with the thread-sequential ones, whose rollout and line-search kernels inline the same f_D source but may
    contract multiply-adds differently), identical instances give identical results, nothing is non-finite."""
    w = Config().build_weights()
    B = 1024
    s = gpu.H1IlqrBatch(w, N=25, batch=B)
    refs = reference_set("standing")
    s.set_reference_window(*refs.window(0, 25), shared=True)
    jr = np.array(gpu.default_dynamics_model().jnt_range)
    x0 = perturbed_states(standing_state(), B, seed=0, jnt_range=jr)
    x0[1] = x0[0]  # two identical instances
    ug = grav_comp_guess(standing_state())
    s.initialize(x0, None, ug)
    cost, iters, status = s.solve(x0)
    assert (status == 0).all() and np.isfinite(cost).all() and (iters >= 1).all() and (iters <= 10).all()
    ct, at = s.solve_trace()
    for i in range(0, B, 37):
        c = ct[i][:iters[i]]
        assert (np.diff(c) <= 1e-9 * np.abs(c[:-1])).all()
    xg, ugp = s.get_trajectory()
    assert (xg[0] == xg[1]).all() and cost[0] == cost[1]
    s.rollout_nominal(x0)
    xr, _ = s.get_trajectory()
    assert np.abs(xr - xg).max() < 1e-12
    sc = gpu.H1IlqrBatch(w, N=25, batch=64)
    sc.set_kernel_policy(1)
    sc.set_reference_window(*refs.window(0, 25), shared=True)
    sc.initialize(x0[:64], None, ug)
    sc.solve(x0[:64])
    xc, _ = sc.get_trajectory()
    sc.rollout_nominal(x0[:64])
    assert (sc.get_trajectory()[0] == xc).all()
    assert np.abs(np.linalg.norm(xg[:, 1:, 3:7], axis=2) - 1).max() < 1e-12


def test_contact_schedule_generation(gpu, tmp_path):
    """contact schedule from the sole points of f_D (replaces get_contacts.py's MuJoCo collision
    query). Known answers: flat standing pose -> all eight points at the same height under the ankles; the
    standing reference is double support throughout; on the walking reference the schedule agrees with the one the
    reference ships (generated with MuJoCo mesh-hull contacts) on most rows; the CSV round-trips."""
    import os
    from helpers import ROOT
    from mpc_ilqr_mujoco_b200.references import contact_schedule_from_states, load_contact_csv, write_contact_csv
    s = gpu.H1IlqrBatch(Config().build_weights(), N=25, batch=1)
    x0 = standing_state()
    pts = s.sole_points(x0[None])[0]
    _, ee = s.reference_kinematics(x0[None])
    m = gpu.default_dynamics_model()
    for f in range(2):
        for c in range(4):
            want = ee[0, f] + np.array(m.foot_pts[f][c])     # identity orientation, zero joint angles
            assert np.abs(pts[4 * f + c] - want).max() < 1e-12
    d = np.load(os.path.join(ROOT, "data", "h1_refs.npz"))
    cs = contact_schedule_from_states(d["standing_q"], s.sole_points)
    assert cs.shape == (d["standing_q"].shape[0], 2) and (cs == 1).all()
    cw = contact_schedule_from_states(d["walking_q"], s.sole_points)    # threshold 1e-3 like get_contacts.py:141
    shipped = d["walking_contact"][:cw.shape[0]]
    agree = (cw[:shipped.shape[0]] == shipped).mean()
    print("walking schedule agreement with the shipped CSV:", agree, "stance fractions", cw.mean(axis=0), shipped.mean(axis=0))
    assert agree > 0.8
    p = tmp_path / "contact.csv"
    write_contact_csv(str(p), cw)
    assert (load_contact_csv(str(p)) == cw).all()


def test_registered_host_buffers_give_identical_results(gpu):
    """h1ilqr_host_register (page-locked caller buffers) changes how the copies run, not what they carry."""
    w = Config().build_weights()
    refs = reference_set("walking")
    B = 4
    wins = [refs.window(5 * i, 25) for i in range(B)]
    win = tuple(np.ascontiguousarray(np.stack([wi[k] for wi in wins])) for k in range(6))
    x0 = np.ascontiguousarray(np.stack([wi[0][0] for wi in wins]))
    ug = grav_comp_guess(standing_state())
    out = []
    for pin in (False, True):
        s = gpu.H1IlqrBatch(w, N=25, batch=B)
        a = tuple(s.pin_host(*win)) if pin else win
        xx = s.pin_host(x0) if pin else x0
        s.set_reference_window(*a, shared=False)
        ua, c = s.mpc_step(xx, ug)
        out.append((ua.copy(), c.copy()))
        s.close()
    assert (out[0][0] == out[1][0]).all() and (out[0][1] == out[1][1]).all()


def test_backward_pass_survives_non_finite_inputs(gpu):
    """A diverged instance (non-finite cost quadratics -> NaN on the diagonal of Quu) must neither fault nor disturb its
    neighbours: the reference carries non-finite gains on with a warning (ilqr.cpp:290-293). Regression test for the pivot
    ranking of the LDL^T (a NaN diagonal used to leave the permutation uninitialised -> out-of-bounds shared-memory reads,
    found with compute-sanitizer at N = 200)."""
    _, sg, win = _setup_pair(gpu, "walking", batch=2)
    x0 = np.stack([win[0][0], win[0][0]])
    ug = grav_comp_guess(standing_state())
    sg.initialize(x0, None, ug)
    sg.rollout_nominal(x0); sg.linearize(); sg.cost_quadratics()
    sg.backward_pass()
    K_ref, k_ref = sg.get_gains()
    lx, lu, lxx, luu = sg.get_cost_quadratics()
    luu[1, 20, 3, 3] = np.nan
    sg.set_cost_quadratics(lx, lu, lxx, luu)
    sg.backward_pass()
    K, k = sg.get_gains()
    assert (K[0] == K_ref[0]).all() and (k[0] == k_ref[0]).all()          # the healthy instance is bit-identical
    assert not np.isfinite(K[1]).all()                                     # the poisoned one reports it through its gains
    # and the long-horizon configuration that exposed it runs through the whole solve
    w = Config().build_weights()
    s = gpu.H1IlqrBatch(w, N=200, batch=1)
    refs = reference_set("walking")
    s.set_reference_window(*refs.window(0, 200), shared=True)
    xs = perturbed_states(refs.x_ref_full[0], 1, seed=0)
    s.upload_inputs(xs, ug)
    assert s.run_resident_steps(1, True) > 0.0
