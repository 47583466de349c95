#!/usr/bin/env python3
"""bench.py — batched H1 MPC solves/s on N B200s (BASELINE.json metric), one process per GPU.

A "step" is one MPC step of every instance resident on a GPU: cold-start guess (gravity compensation) + the full
multi-iteration iLQR solve (<= 10 iterations: rollout, linearization, cost quadratics, Riccati backward pass,
8-alpha line search) + first control.

  --workload config5         (default) BASELINE config 5 sharded by instance, WEAK scaling: `--batch` (8192) walking
                             instances per GPU — 8 GPUs = the 65,536-instance configuration; instance i tracks the walking
                             reference from window row t0 = i mod 374 from a perturbed state
  --workload config5_strong  the same 65,536 instances in total, split over the GPUs (STRONG scaling; 1 GPU holds them all)
  --workload config3         BASELINE config 3: 1024 perturbed standing instances on one GPU

  value        : solves/s with inputs resident in HBM (CUDA events on the solver's stream, max over ranks)
  e2e          : the same through the public C-ABI calls h1ilqr_set_reference_window + h1ilqr_mpc_step with HOST buffers
                 (page-locked once with h1ilqr_host_register; H2D of x_measured + reference windows and D2H of u_apply +
                 cost inside the timed region), over all --steps
  roofline     : the dominant kernel (k_backward, Riccati pass on the fp64 tensor cores): algorithmic flops per launch /
                 launch duration measured live with CUDA events (h1ilqr_time_stage), against the fp64 tensor peak measured
                 live; `kernels` holds the same per-launch measurement for every stage kernel next to its bound. DRAM
                 traffic / executed-flop figures come from the committed ncu capture (profiles/r02_kernel_metrics.json)
  warm_closed_loop : the workload an MPC actually runs — device-resident closed loop (h1ilqr_run_closed_loop, warm starts)
  single_instance  : BASELINE metric part 1, H1 iLQR solve ms per MPC step (N = 25, one instance): device-resident with /
                     without the CUDA-graph replay, and end to end through MPC::stepOnce of the C++ host shim (configs 1, 2)
  cpu_baseline : the CPU oracle (a port: the reference itself cannot be built here) on this box's host cores
`--impl reference` times that CPU oracle as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched_h1_mpc_solves_per_sec"
UNIT = "solves/s"
N_HORIZON = 25
TOTAL_CONFIG5 = 65536


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("H1_BENCH_WORKLOAD", "config5"), choices=["config5", "config5_strong", "config3"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("H1_BENCH_BATCH", "8192")), help="instances per GPU (config5)")
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("H1_BENCH_CPU_SAMPLE", "48")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip warm_closed_loop / single_instance / per-kernel legs")
    return ap.parse_args()


def workload(name, batch, rank, world, kinematics, jnt_range=None):
    """(window tuple, shared flag, x0, t0, description) of this rank's instances — exactly what
    tests/test_gpu_workloads.py checks against the oracle (workloads.py is shared by both)."""
    from mpc_ilqr_mujoco_b200 import workloads as wl
    if name == "config3":
        win, x0 = wl.standing_instances(np.arange(rank * batch, (rank + 1) * batch), kinematics, N=N_HORIZON, jnt_range=jnt_range)
        return win, True, x0, np.zeros(batch, dtype=np.int32), "standing"
    win, x0, t0 = wl.walking_instances(np.arange(rank * batch, (rank + 1) * batch), kinematics, N=N_HORIZON)
    return win, False, x0, t0.astype(np.int32), "walking"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def oracle_kinematics(x):
    """CoM / ankle positions of reference rows on the dynamics model, from the CPU oracle (CPU legs only)."""
    from oracle import pyoracle as po
    x = np.atleast_2d(x)
    com = np.array([po.dyn_com(r) for r in x])
    ee = np.array([[po.dyn_body_pos(r, 5), po.dyn_body_pos(r, 10)] for r in x])
    return com, ee


def _standing():
    from mpc_ilqr_mujoco_b200.references import standing_state
    return standing_state()


def cpu_oracle_rate_parallel(sample, name="config5", threads=None, linearization=0):
    """`sample` instances of the workload, one cold MPC step each, all in flight over the host threads (one oracle
    handle per instance, driven from Python threads; the C++ side releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from mpc_ilqr_mujoco_b200 import Config
    from oracle import pyoracle as po
    po.build()
    threads = threads or po.lib().orc_max_threads()
    w = Config().build_weights()
    opt = po.default_options()
    opt.linearization = linearization
    jr = np.array(po.dynamics_model().jnt_range)
    win, shared, x0, _, _ = workload(name, sample, 0, 1, oracle_kinematics, jnt_range=jr)
    ug = np.zeros(19)
    ug[:18] = po.dyn_bias(_standing())[7:25]
    solvers = []
    for i in range(sample):
        s = po.OracleSolver(w, N_HORIZON, batch=1, options=opt)
        s.set_reference_window(*(win if shared else tuple(a[i] for a in win)))
        solvers.append(s)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda k: solvers[k].mpc_step_batch(x0[k:k + 1], ug, 1), range(sample)))
    dt = time.perf_counter() - t0
    return sample / dt, threads, dt


def cpu_single_thread_ms(name, linearization, instances=2):
    """Single-thread CPU oracle, ms per cold MPC step (the reference is single-threaded): mean over `instances`."""
    v, _, dt = cpu_oracle_rate_parallel(instances, name, threads=1, linearization=linearization)
    return 1e3 * dt / instances


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation cannot be built here (MuJoCo / Pinocchio / CasADi /
    Eigen / yaml-cpp absent), so this times the CPU oracle port on all host threads."""
    if rank != 0:
        return
    sample = args.cpu_sample
    name = "config3" if args.workload == "config3" else "config5"
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_rate_parallel(max(8, sample // 6), name)
    vals, secs = [], 0.0
    for _ in range(args.steps):
        v, threads, dt = cpu_oracle_rate_parallel(sample, name)
        vals.append(v)
        secs += dt
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "config5_strong" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"H1 MPC step, N={N_HORIZON}, cold-start iLQR solve per instance (BASELINE {name} instances, analytic linearization)",
                       "instances_per_step": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} instances per step, all host threads, CPU oracle (oracle/) - the reference binary needs MuJoCo/Pinocchio/CasADi which are absent"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def host_single_instance(tag, steps):
    """BASELINE configs 1 / 2 through the reference-facing C++ API: the host shim's demo binary runs `steps` closed-loop
    MPC steps (RobotUtils / MPC::stepOnce over the C ABI, host buffers in and out, warm starts) and reports the wall
    time of every stepOnce call."""
    exe = os.path.join(ROOT, "mpc-ilqr-mujoco_b200", "host", "bin", "humanoid_mpc_demo")
    if not os.path.exists(exe):
        return {"unavailable": "host demo binary not built (make -C mpc-ilqr-mujoco_b200/host)"}
    from mpc_ilqr_mujoco_b200 import Config, dump_config_yaml
    d = np.load(os.path.join(ROOT, "data", "h1_refs.npz"))
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "data")); os.makedirs(os.path.join(td, "results"))
        md = np.load(os.path.join(ROOT, "data", "h1_models.npz"))     # the model files RobotUtils::loadModel / iLQR parse at run time
        for key, rel in (("mjcf_scene_xml", "robots/h1_description/mjcf/scene.xml"), ("mjcf_h1_xml", "robots/h1_description/mjcf/h1.xml"),
                         ("urdf_h1_urdf", "robots/h1_description/urdf/h1.urdf")):
            os.makedirs(os.path.dirname(os.path.join(td, rel)), exist_ok=True)
            open(os.path.join(td, rel), "wb").write(md[key].tobytes())
        np.savetxt(os.path.join(td, "data", "q.csv"), d[f"{tag}_q"], delimiter=",", fmt="%.17g")
        np.savetxt(os.path.join(td, "data", "v.csv"), d[f"{tag}_v"], delimiter=",", fmt="%.17g")
        with open(os.path.join(td, "data", "c.csv"), "w") as f:
            f.write("left_foot,right_foot\n")
            for r in d[f"{tag}_contact"]:
                f.write(f"{r[0]},{r[1]}\n")
        cfg = Config()
        cfg.q_ref_path, cfg.v_ref_path, cfg.contact_schedule_path = "data/q.csv", "data/v.csv", "data/c.csv"
        cfg.mpc.sim_steps = steps
        cfg.save_trajectories = False
        with open(os.path.join(td, "config.yaml"), "w") as f:
            f.write(dump_config_yaml(cfg))
        out = subprocess.run([exe, "config.yaml", str(steps)], cwd=td, capture_output=True, text=True, timeout=600)
    js = [ln for ln in out.stdout.splitlines() if ln.startswith("BENCH_JSON ")]
    if out.returncode != 0 or not js:
        return {"unavailable": (out.stderr or out.stdout)[-300:]}
    r = json.loads(js[-1][len("BENCH_JSON "):])
    ms = np.array(r["step_ms"])
    warm = ms[1:] if len(ms) > 1 else ms
    return {"steps": int(len(ms)), "first_step_ms_cold": float(ms[0]), "mean_ms": float(warm.mean()), "min_ms": float(warm.min()),
            "max_ms": float(warm.max()), "all_finite": bool(np.isfinite(r["cost"]).all()),
            "api": "MPC::stepOnce of libh1host.so (C++ shim over the C ABI), warm-started steps 2..n"}


def load_kernel_metrics():
    """ncu-derived per-knot figures written by tools/ncu_kernel_metrics.py (profiles/): DRAM traffic and executed flops."""
    for name in ("r02_kernel_metrics.json",):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            try:
                return json.load(open(p)), name
            except Exception:
                pass
    return {}, None


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    if world > 1:
        torch.cuda.set_device(local)
        # NCCL prints its version banner to stdout when the first communicator comes up: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    from mpc_ilqr_mujoco_b200 import Config, gpu
    from mpc_ilqr_mujoco_b200 import workloads as wl
    w = Config().build_weights()
    if args.workload == "config3":
        B = 1024
    elif args.workload == "config5_strong":
        B = TOTAL_CONFIG5 // world
    else:
        B = args.batch
    solver = gpu.H1IlqrBatch(w, N=N_HORIZON, batch=B, device=local)
    jr = np.array(gpu.default_dynamics_model().jnt_range)
    win, shared, x0, t0, tag = workload(args.workload, B, rank, world, solver.reference_kinematics, jnt_range=jr)
    solver.set_reference_window(*win, shared=shared)
    ug = np.zeros(19)
    ug[:18] = solver.bias_forces(_standing()[None])[0][7:25]
    solver.upload_inputs(x0, ug)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        solver.run_resident_steps(1, True)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ms = solver.run_resident_steps(args.steps, True)
    launches = solver.stage_times()["launches"]
    barrier()
    clocks = sampler.stop()
    ms_max = max_over_ranks(ms)
    value = world * B * args.steps / (ms_max * 1e-3)

    # per-instance statistics, reduced over NCCL (the only inter-GPU traffic of the path)
    ct, at = solver.solve_trace()
    xg, ugp = solver.get_trajectory()
    status, iters = solver.get_status()
    # a diverging instance is a legitimate outcome of the reference algorithm (NaN cost -> every line search fails);
    # what must hold is that every instance that reports status 0 carries a finite trajectory
    ok_inst = status == 0
    finite_ok = bool(np.isfinite(xg[ok_inst]).all() and np.isfinite(ugp[ok_inst]).all())
    stats = torch.tensor([float(iters.sum()), float(finite_ok), float(B), float(ok_inst.sum())], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    mean_iters = float(stats[0].item() / stats[2].item())
    all_finite = bool(stats[1].item() == world)
    n_ok = int(stats[3].item())

    # ---- end to end through the C-ABI with host buffers, over all --steps ----
    # the host arrays handed over every step are page-locked once (h1ilqr_host_register), as a host MPC loop would do
    # with its reference buffers: the copies inside the timed region are DMA transfers from pinned memory
    win = tuple(solver.pin_host(*win))
    x0, ug = solver.pin_host(x0, ug)
    solver.mpc_reset()
    solver.set_reference_window(*win, shared=shared)
    solver.mpc_step(x0, ug)  # warm-up of the host path
    barrier()
    tt = time.perf_counter()
    for _ in range(args.steps):
        solver.mpc_reset()
        solver.set_reference_window(*win, shared=shared)   # per-step host inputs: reference windows + x_measured
        ua, cost = solver.mpc_step(x0, ug)
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - tt)
    e2e_value = world * B * args.steps / dt
    h2d = sum(a.nbytes for a in win) + x0.nbytes + ug.nbytes
    d2h = ua.nbytes + cost.nbytes
    e2e_finite = bool(np.isfinite(ua[ok_inst]).all())

    # ---- warm-started closed loop on the device (the workload an MPC runs): every rank, max over ranks ----
    warm = None
    if not args.no_extras:
        refs = wl.reference_set(tag, solver.reference_kinematics)
        solver.set_reference_table(refs)
        solver.mpc_reset()
        cl_steps = 4
        solver.run_closed_loop(2, t_idx0=t0, x_start=x0, u_init=ug, graph=True, logs=False)       # cold step + one warm step
        barrier()
        out = solver.run_closed_loop(cl_steps, graph=True, logs=True)
        cl_ms = max_over_ranks(out["ms"])
        it_sum = torch.tensor([float(out["iters"].sum()), float(out["iters"].size)], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(it_sum, op=dist.ReduceOp.SUM)
        warm = {"value": world * B * cl_steps / (cl_ms * 1e-3), "unit": UNIT, "steps": cl_steps, "ms_per_step": cl_ms / cl_steps,
                "mean_ilqr_iterations": float(it_sum[0].item() / it_sum[1].item()),
                "api": "h1ilqr_set_reference_table + h1ilqr_run_closed_loop (window extraction, warm start, solve, first control, plant "
                       "step on the device; CUDA-graph replay; steps 3.. of the loop)"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json (of measured)" if peaks else "6650 GB/s (of fallback)"
        km, km_file = load_kernel_metrics()
        # ---- stage split of one solve (stages timed one after the other) ----
        solver.set_reference_window(*win, shared=shared)
        solver.enable_stage_timing(True)
        solver.mpc_reset(); solver.initialize(x0, None, ug)
        solver.solve(x0)
        tm = solver.stage_times()
        solver.enable_stage_timing(False)
        _, at_s = solver.solve_trace()
        first_passes = float((at_s[:, :, 0] != -2).sum())    # instance-iterations (linearization, first backward pass / line search)
        second_passes = float((at_s[:, :, 1] != -2).sum())   # second attempts after a failed line search
        stage = {k: tm[k] for k in ("rollout_ms", "linearize_ms", "cost_quadratics_ms", "backward_ms", "line_search_ms")}
        fp64_peak = solver.measure_fp64_peak()
        dmma_peak = solver.measure_fp64_mma_peak()
        # ---- per-launch device time of every stage kernel on the full batch (first iteration of a cold step) ----
        solver.mpc_reset(); solver.initialize(x0, None, ug)
        solver.rollout_nominal(x0); solver.linearize(); solver.cost_quadratics(); solver.backward_pass()
        kern = {}
        knots = B * N_HORIZON
        # stages interleaved as they are in a solve (a back-to-back train of one kernel is not what the step runs: on a warm
        # box the DMMA-heavy backward pass alone read 10 % slower that way), median of 5 rounds after one warm-up round
        stages = ("factor", "linearize", "cost_quadratics", "backward", "line_search")
        samples = {st: [] for st in stages}
        for rnd in range(6):
            for st in stages:
                ms_st = solver.time_stage(st, 1)
                if rnd:
                    samples[st].append(ms_st)
        kern = {st: float(np.median(samples[st])) for st in stages}
        bwd_s = kern["backward"] * 1e-3
        bwd_tflops = knots * BWD_FLOPS_PER_KNOT / bwd_s / 1e12
        kernels = {
            "k_backward": {"ms_per_launch": kern["backward"], "bound": "tensor (fp64 DMMA)", "achieved_tflops": bwd_tflops,
                           "frac": bwd_tflops / max(dmma_peak, 1e-9)},
            "k_linearize_tangents<3> + k_linearize_finish": {
                "ms_per_launch": kern["linearize"], "bound": "hbm",
                "achieved_gbs": knots * LIN_ALG_BYTES_PER_KNOT / (kern["linearize"] * 1e-3) / 1e9,
                "frac": knots * LIN_ALG_BYTES_PER_KNOT / (kern["linearize"] * 1e-3) / 1e9 / hbm_peak},
            "k_cost_quadratics": {"ms_per_launch": kern["cost_quadratics"], "bound": "hbm",
                                  "achieved_gbs": B * (N_HORIZON + 1) * CQ_ALG_BYTES_PER_KNOT / (kern["cost_quadratics"] * 1e-3) / 1e9,
                                  "frac": B * (N_HORIZON + 1) * CQ_ALG_BYTES_PER_KNOT / (kern["cost_quadratics"] * 1e-3) / 1e9 / hbm_peak},
            "k_line_search_quad": {"ms_per_launch": kern["line_search"], "bound": "fp64",
                                   "achieved_tflops": B * 8 * N_HORIZON * km.get("line_search_flops_per_eval", LS_FLOPS_PER_EVAL) / (kern["line_search"] * 1e-3) / 1e12,
                                   "frac": B * 8 * N_HORIZON * km.get("line_search_flops_per_eval", LS_FLOPS_PER_EVAL) / (kern["line_search"] * 1e-3) / 1e12 / max(fp64_peak, 1e-9),
                                   "flops_per_eval": km.get("line_search_flops_per_eval", LS_FLOPS_PER_EVAL)},
            "k_primal_factor_seq": {"ms_per_launch": kern["factor"]},
        }
        # ---- single-instance latency (BASELINE metric part 1) ----
        single = None
        if not args.no_extras:
            s1 = gpu.H1IlqrBatch(w, N=N_HORIZON, batch=1, device=local)
            s1.set_reference_window(*(win if shared else tuple(a[0] for a in win)), shared=True)
            s1.upload_inputs(x0[:1], ug)
            res = {}
            for graph in (False, True):
                for _ in range(3):
                    s1.run_resident_steps(1, True, graph=graph)
                res["cold_step_graph_ms" if graph else "cold_step_ms"] = s1.run_resident_steps(10, True, graph=graph) / 10
            res["launches_per_step"] = s1.stage_times()["launches"] // 10
            s1.close()
            res["host_config1_standing_15_steps"] = host_single_instance("standing", 15)
            res["host_config2_walking_100_steps"] = host_single_instance("walking", 100)
            single = res
        cpu = None
        if not args.no_cpu_baseline and world == 1:   # (rank 0 at N = 1 only)
            name = "config3" if args.workload == "config3" else "config5"
            v, threads, secs = cpu_oracle_rate_parallel(args.cpu_sample, name)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_sample} instances of the same workload, one cold MPC step each, all host threads ({secs:.1f} s)",
                   "single_thread_ms_per_mpc_step": {"analytic": cpu_single_thread_ms(name, 0), "fd": cpu_single_thread_ms(name, 1),
                                                     "note": "one host thread, mean of 2 instances; fd = the reference's forward-difference linearization"}}
        scaling = "strong" if args.workload == "config5_strong" else "weak"
        desc = {"config5": f"{B} H1 walking MPC instances per GPU (BASELINE config 5 sharded by instance: 8192/GPU x 8 = 65536)",
                "config5_strong": f"BASELINE config 5: 65536 H1 walking MPC instances in total, {B} per GPU",
                "config3": "BASELINE config 3: 1024 perturbed H1 standing MPC instances on one GPU"}[args.workload]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc + f", N={N_HORIZON}, one cold-start MPC step = full iLQR solve (<=10 iterations, analytic linearization)",
                       "name": args.workload, "instances_per_gpu": B, "horizon": N_HORIZON, "mean_ilqr_iterations": mean_iters,
                       "l2": "working set per GPU (%.1f GB of solver state) is far larger than the 126 MB L2" % (B * 1.7e-3)},
            "all_finite": all_finite and e2e_finite, "instances_status_ok": n_ok,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.steps,
                    "api": "h1ilqr_set_reference_window + h1ilqr_mpc_step (host buffers registered with h1ilqr_host_register)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            # dominant kernel of the step: the Riccati backward pass (one CTA per instance, contractions on the fp64 tensor cores)
            "roofline": {"kernel": "k_backward (Riccati backward pass, five contractions per knot as mma.sync m8n8k4 f64 = SASS DMMA)",
                         "bound": "tensor", "achieved": bwd_tflops, "peak": dmma_peak, "unit": "TFLOP/s",
                         "frac": bwd_tflops / max(dmma_peak, 1e-9),
                         "traffic": km.get("backward_dram_bytes_per_knot", BWD_TRAFFIC_BYTES_PER_KNOT) * knots,
                         "traffic_note": f"ncu --set full dram read+write per knot ({km_file or 'profiles/r01i_ncu_top_kernels.txt'}) x the {knots} knots of one launch; "
                                         f"algorithmic {BWD_ALG_BYTES_PER_KNOT / 1e3:.1f} KB per knot",
                         "launch_ms": kern["backward"], "knots_per_launch": knots,
                         "peak_source": "fp64 tensor-core peak measured live in this run (mma.sync m8n8k4 f64 probe kernel); MEASURED_PEAKS.json "
                                        "carries HBM and bf16 figures only, and the bf16 tcgen05 peak does not apply to an fp64 path",
                         "flops_per_knot": BWD_FLOPS_PER_KNOT,
                         "share_of_step": tm["backward_ms"] / max(tm["total_ms"], 1e-9),
                         "whole_solve": {"knot_passes": (first_passes + second_passes) * N_HORIZON, "stage_ms": tm["backward_ms"],
                                         "frac": (first_passes + second_passes) * N_HORIZON * BWD_FLOPS_PER_KNOT / (tm["backward_ms"] * 1e-3) / 1e12 / max(dmma_peak, 1e-9)},
                         "hbm": {"achieved": knots * BWD_ALG_BYTES_PER_KNOT / bwd_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": knots * BWD_ALG_BYTES_PER_KNOT / bwd_s / 1e9 / hbm_peak, "peak_source": hbm_src},
                         "note": "achieved = ALGORITHMIC flops (SURVEY 8(d): 1.153 MFLOP per knot, no symmetry credit) x knots of one full-batch launch / its "
                                 "duration (CUDA events, h1ilqr_time_stage, median of 5 with the stages interleaved); the kernel executes 1428 DMMA = 0.73 MFLOP per knot "
                                 "(lower triangles of Qxx / Vxx only; contractions over the 29 information-carrying rows of [A|B], h1_riccati.cuh)"},
            "kernels": kernels, "fp64_fma_peak_tflops": fp64_peak, "kernel_metrics_file": km_file,
            "stage_ms_per_solve": stage,
            "warm_closed_loop": warm,
            "single_instance": single,
            "single_instance_ms_per_mpc_step": single["cold_step_graph_ms"] if single else None,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not (all_finite and e2e_finite):
        print("bench: an instance with status 0 carries a non-finite trajectory / control", file=sys.stderr)
        sys.exit(3)


# ---- per-knot work figures ----
# Riccati backward pass: algorithmic flops of one knot (the five contractions with shared products, the
# factorisation and the solves) and its algorithmic bytes (A, B, lx, lu, lxx, luu read; K, kff written).
BWD_FLOPS_PER_KNOT = 1.153e6
BWD_ALG_BYTES_PER_KNOT = 52816.0 - (51 * 51 - 51 * 52 // 2) * 8.0 + 7904.0   # (lxx: lower triangle only, 1326 of 2601 entries)
BWD_TRAFFIC_BYTES_PER_KNOT = (5.500e9 + 0.817e9) / (4096 * 25)      # fallback when profiles/r02_kernel_metrics.json is absent (r01i capture)
# Linearization: algorithmic bytes per knot — A_k, B_k written; x_k, u_k and the factor (L, D, a) read; the parked tangents
# written and read once
LIN_ALG_BYTES_PER_KNOT = (51 * 51 + 51 * 19 + 51 + 19) * 8.0 + (25 * 11 + 25 + 25) * 8.0 + 2 * 48 * 25 * 8.0
# Cost quadratics: lx, lu, lxx (lower triangle), luu written; x, u read (per knot, incl. the terminal one)
CQ_ALG_BYTES_PER_KNOT = (51 + 19 + 51 * 52 // 2 + 19 * 19 + 51 + 19) * 8.0
# Line search: fp64 operations executed per f_D evaluation of the quad kernel (fallback; the ncu figure replaces it)
LS_FLOPS_PER_EVAL = 3.2e4

if __name__ == "__main__":
    main()
