#!/usr/bin/env python3
"""bench.py — batched H1 MPC solves/s on N B200s (BASELINE.json metric), one process per GPU.

A "step" is one MPC step of every instance resident on a GPU: cold-start guess (gravity compensation) + the full
multi-iteration iLQR solve (<= 10 iterations: rollout, linearization, cost quadratics, Riccati backward pass,
8-alpha line search) + first control. Workload = BASELINE config 5 sharded by instance: `--batch` walking-MPC
instances per GPU (weak scaling; 8192/GPU x 8 GPUs = the 65,536-instance configuration), instance i tracking
the walking reference from window row t0_i = i mod 374 with a perturbed initial state.

  value        : solves/s with inputs resident in HBM (CUDA events on the solver's stream, max over ranks)
  e2e          : the same through the public C-ABI call h1ilqr_mpc_step with HOST buffers (pinned staging,
                 H2D of x_measured + reference windows and D2H of u_apply + cost inside the timed region)
  roofline     : dominant stage (analytic linearization, k_linearize_dirs<0|1|2>) against the measured HBM peak and
                 the live-measured fp64 FMA peak; flop counts are EXECUTED fp64 operations per knot taken from ncu
                 (profiles/), not an estimate
  cpu_baseline : the CPU oracle (a port: the reference itself cannot be built here) on this box's host cores
`--impl reference` times that CPU oracle as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched_h1_mpc_solves_per_sec"
UNIT = "solves/s"
N_HORIZON = 25


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("H1_BENCH_BATCH", "8192")), help="instances per GPU")
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("H1_BENCH_CPU_SAMPLE", "48")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(batch, rank, kinematics):
    """Instances [rank*batch, (rank+1)*batch) of the sharded walking workload (deterministic, seed 0)."""
    from mpc_ilqr_mujoco_b200.references import ReferenceSet
    d = np.load(os.path.join(ROOT, "data", "h1_refs.npz"))
    refs = ReferenceSet(d["walking_q"], d["walking_v"], d["walking_contact"], kinematics)
    ids = np.arange(rank * batch, (rank + 1) * batch)
    t0 = ids % (refs.T - (N_HORIZON + 1))
    wins = [refs.window(int(t), N_HORIZON) for t in np.unique(t0)]
    lut = {int(t): w for t, w in zip(np.unique(t0), wins)}
    stack = lambda k: np.ascontiguousarray(np.stack([lut[int(t)][k] for t in t0]))
    win = tuple(stack(k) for k in range(6))
    x_nom = refs.x_ref_full[t0]
    x0 = np.vstack([_perturb_one(x_nom[j], int(i)) for j, i in enumerate(ids)])  # keyed by the GLOBAL instance id
    return win, x0


def _perturb_one(x_nom, gid):
    from mpc_ilqr_mujoco_b200.references import NQ, NV
    rng = np.random.Generator(np.random.Philox(key=0, counter=[gid, 0, 0, 0]))
    x = np.array(x_nom, dtype=np.float64)
    x[0:3] += rng.uniform(-0.02, 0.02, 3)
    rv = rng.uniform(-0.05, 0.05, 3)
    ang = np.linalg.norm(rv)
    dq = np.array([np.cos(ang / 2), *(np.sin(ang / 2) / ang * rv)])
    w0, x0, y0, z0 = x[3:7] / np.linalg.norm(x[3:7])
    w1, x1, y1, z1 = dq
    qn = np.array([w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                   w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1, w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1])
    x[3:7] = qn / np.linalg.norm(qn)
    x[7:NQ] += rng.uniform(-0.05, 0.05, NQ - 7)
    x[NQ:] += rng.uniform(-0.1, 0.1, NV)
    return x


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


def cpu_oracle_rate_parallel(sample, threads=None):
    """Same sample, but all instances in flight at once over all host threads (one oracle handle per instance
    group would serialise the groups; here groups are driven from Python threads, the C++ side releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from mpc_ilqr_mujoco_b200 import Config
    from oracle import pyoracle as po
    po.build()
    threads = threads or po.lib().orc_max_threads()
    w = Config().build_weights()
    win, x0 = workload(sample, 0, oracle_kinematics)
    ug = np.zeros(19)
    ug[:18] = po.dyn_bias(_standing())[7:25]
    solvers = []
    for i in range(sample):
        s = po.OracleSolver(w, N_HORIZON, batch=1)
        s.set_reference_window(*(a[i] for a in win))
        solvers.append(s)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda k: solvers[k].mpc_step_batch(x0[k:k + 1], ug, 1), range(sample)))
    dt = time.perf_counter() - t0
    return sample / dt, threads, dt


def oracle_kinematics(x):
    """CoM / ankle positions of reference rows on the dynamics model, from the CPU oracle (reference arm only)."""
    from oracle import pyoracle as po
    x = np.atleast_2d(x)
    com = np.array([po.dyn_com(r) for r in x])
    ee = np.array([[po.dyn_body_pos(r, 5), po.dyn_body_pos(r, 10)] for r in x])
    return com, ee


def _standing():
    from mpc_ilqr_mujoco_b200.references import standing_state
    return standing_state()


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation cannot be built here (MuJoCo / Pinocchio / CasADi /
    Eigen / yaml-cpp absent), so this times the CPU oracle port on all host threads."""
    if rank != 0:
        return
    sample = args.cpu_sample
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_rate_parallel(max(8, sample // 6))
    vals, secs = [], 0.0
    for _ in range(args.steps):
        v, threads, dt = cpu_oracle_rate_parallel(sample)
        vals.append(v)
        secs += dt
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"H1 walking MPC step, N={N_HORIZON}, cold-start iLQR solve per instance (BASELINE config 5 instances)",
                       "instances_per_step": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} instances per step, all host threads, CPU oracle (oracle/) - the reference binary needs MuJoCo/Pinocchio/CasADi which are absent"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mpc_ilqr_mujoco_b200 import Config, gpu
    w = Config().build_weights()
    B = args.batch
    solver = gpu.H1IlqrBatch(w, N=N_HORIZON, batch=B, device=local)
    win, x0 = workload(B, rank, solver.reference_kinematics)
    solver.set_reference_window(*win, shared=False)
    ug = np.zeros(19)
    ug[:18] = solver.bias_forces(_standing()[None])[0][7:25]
    solver.upload_inputs(x0, ug)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # per-instance statistics, gathered over NCCL (the only inter-GPU traffic of the path)
    ct, at = solver.solve_trace()
    xg, ugp = solver.get_trajectory()
    iters_local = (at[:, :, 0] != -2).sum(axis=1).astype(np.float64)
    stats = torch.tensor([iters_local.sum(), float(np.isfinite(xg).all()), B], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    mean_iters = float(stats[0].item() / stats[2].item())

    # ---- end to end through the C-ABI with host buffers ----
    e2e_steps = max(1, min(args.steps, 3))
    # the host arrays handed over every step are page-locked once (h1ilqr_host_register), as a host MPC loop would do
    # with its reference buffers: the copies inside the timed region are DMA transfers from pinned memory
    win = tuple(solver.pin_host(*win))
    x0, ug = solver.pin_host(x0, ug)
    solver.mpc_reset()
    solver.set_reference_window(*win, shared=False)
    solver.mpc_step(x0, ug)  # warm-up of the host path
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        solver.mpc_reset()
        solver.set_reference_window(*win, shared=False)   # per-step host inputs: reference windows + x_measured
        ua, cost = solver.mpc_step(x0, ug)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(t.item())
    h2d = sum(a.nbytes for a in win) + x0.nbytes + ug.nbytes
    d2h = ua.nbytes + cost.nbytes

    line = None
    if rank == 0:
        # ---- stage split + roofline of the dominant kernel (measured outside the timed region) ----
        solver.enable_stage_timing(True)
        solver.mpc_reset(); solver.initialize(x0, None, ug)
        _, it_s, _ = solver.solve(x0)
        tm = solver.stage_times()
        solver.enable_stage_timing(False)
        fp64_peak = solver.measure_fp64_peak()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        knots = float(it_s.sum()) * N_HORIZON               # linearized knots in that solve
        lin_s = tm["linearize_ms"] * 1e-3
        # algorithmic bytes per linearized knot: A_k, B_k written; x_k, u_k and the Mhat factor (L, D, a) read
        alg_bytes = knots * ((51 * 51 + 51 * 19 + 51 + 19) * 8.0 + FACTOR_BYTES_PER_KNOT)
        alg_flops = knots * FLOPS_PER_LINEARIZED_KNOT
        achieved = alg_bytes / lin_s / 1e9
        bwd_passes = float(it_s.sum()) * N_HORIZON          # lower bound: second attempts add passes
        dmma_peak = solver.measure_fp64_mma_peak()
        stage = {k: tm[k] for k in ("rollout_ms", "linearize_ms", "cost_quadratics_ms", "backward_ms", "line_search_ms")}
        # single-instance latency (BASELINE metric part 1): H1 iLQR solve ms per MPC step, N=25, one instance
        s1 = gpu.H1IlqrBatch(w, N=N_HORIZON, batch=1, device=local)
        s1.set_reference_window(*(a[0] for a in win), shared=True)
        s1.upload_inputs(x0[:1], ug)
        for _ in range(3):
            s1.run_resident_steps(1, True)
        single_ms = s1.run_resident_steps(5, True) / 5
        s1.close()
        cpu = None
        if not args.no_cpu_baseline:
            v, threads, secs = cpu_oracle_rate_parallel(args.cpu_sample)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_sample} instances of the same workload, one cold MPC step each, all host threads ({secs:.1f} s)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{B} H1 walking MPC instances per GPU (BASELINE config 5 sharded by instance: 8192/GPU x 8 = 65536), "
                                   f"N={N_HORIZON}, one cold-start MPC step = full iLQR solve (<=10 iterations, analytic linearization)",
                       "instances_per_gpu": B, "horizon": N_HORIZON, "mean_ilqr_iterations": mean_iters,
                       "l2": "working set per GPU (%.1f GB of solver state) is far larger than the 126 MB L2" % (B * 1.7e-3)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "h1ilqr_set_reference_window + h1ilqr_mpc_step (host buffers registered with h1ilqr_host_register)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "k_linearize_dirs<0|1|2> (analytic linearization, one thread per column of [A|B])",
                         "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": TRAFFIC_BYTES_PER_LINEARIZED_KNOT * B * N_HORIZON,
                         "traffic_note": "ncu --set full dram read+write of the three launches of one iteration, scaled to this batch (profiles/)",
                         "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback",
                         "share_of_step": tm["linearize_ms"] / max(tm["total_ms"], 1e-9),
                         "note": "the stage is fp64-pipe bound, not HBM bound: see fp64",
                         "fp64": {"achieved_tflops": alg_flops / lin_s / 1e12, "peak_tflops": fp64_peak,
                                  "frac": alg_flops / lin_s / 1e12 / max(fp64_peak, 1e-9),
                                  "flops_per_knot": FLOPS_PER_LINEARIZED_KNOT,
                                  "peak_source": "measured live (DFMA kernel)"}},
            "roofline_backward": {"kernel": "k_backward (Riccati, DMMA m8n8k4)", "bound": "fp64 tensor",
                                  "achieved_tflops": bwd_passes * 1.153e6 / (tm["backward_ms"] * 1e-3) / 1e12,
                                  "peak_tflops": dmma_peak, "peak_source": "measured live (mma.sync m8n8k4 f64 kernel)",
                                  "note": "achieved is a lower bound (second attempts after a failed line search add passes)"},
            "stage_ms_per_solve": stage,
            "single_instance_ms_per_mpc_step": single_ms,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# fp64 operations EXECUTED per linearized knot by k_linearize_dirs<0|1|2> (2 per DFMA, 1 per DADD / DMUL), from the
# smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on counters of profiles/r01f_ncu_top_kernels.txt:
# q columns 693 k + v columns 413 k + u columns 40 k.
FLOPS_PER_LINEARIZED_KNOT = 1.15e6
FACTOR_BYTES_PER_KNOT = (25 * 11 + 25 + 25) * 8.0
# dram__bytes_read.sum + dram__bytes_write.sum of the same capture, per knot (14.2 GB / (4096 instances x 25 knots))
TRAFFIC_BYTES_PER_LINEARIZED_KNOT = 139.0e3

if __name__ == "__main__":
    main()
