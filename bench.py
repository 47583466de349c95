#!/usr/bin/env python3
"""bench.py — batched H1 MPC solves/s on N B200s (BASELINE.json metric), one process per GPU.

A "step" is one MPC step of every instance resident on a GPU: cold-start guess (gravity compensation) + the full
multi-iteration iLQR solve (<= 10 iterations: rollout, linearization, cost quadratics, Riccati backward pass,
8-alpha line search) + first control. Workload = BASELINE config 5 sharded by instance: `--batch` walking-MPC
instances per GPU (weak scaling; 8192/GPU x 8 GPUs = the 65,536-instance configuration), instance i tracking
the walking reference from window row t0_i = i mod 374 with a perturbed initial state.

  value        : solves/s with inputs resident in HBM (CUDA events on the solver's stream, max over ranks)
  e2e          : the same through the public C-ABI call h1ilqr_mpc_step with HOST buffers (page-locked once with
                 h1ilqr_host_register; H2D of x_measured + reference windows and D2H of u_apply + cost inside the timed region)
  roofline     : dominant kernel (k_backward, the Riccati pass on the fp64 tensor cores): algorithmic flops per knot x
                 knot passes / stage time measured live with CUDA events, against the fp64 tensor peak measured live;
                 its HBM view next to it. `roofline_linearize` does the same for the second stage (executed fp64
                 operations per knot taken from ncu, profiles/, not an estimate)
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
    """Instances [rank*batch, (rank+1)*batch) of the sharded walking workload (deterministic, seed 0): exactly the
    instances tests/test_gpu_workloads.py::test_bench_workload_parity checks against the oracle."""
    from mpc_ilqr_mujoco_b200 import workloads as wl
    win, x0, _ = wl.walking_instances(np.arange(rank * batch, (rank + 1) * batch), kinematics, N=N_HORIZON)
    return win, x0


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
        _, at_s = solver.solve_trace()
        first_passes = float((at_s[:, :, 0] != -2).sum())    # instance-iterations (linearization, first backward pass / line search)
        second_passes = float((at_s[:, :, 1] != -2).sum())   # second attempts after a failed line search
        knots = first_passes * N_HORIZON                     # linearized knots in that solve
        lin_s = tm["linearize_ms"] * 1e-3
        bwd_s = tm["backward_ms"] * 1e-3
        bwd_knots = (first_passes + second_passes) * N_HORIZON
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
        if not args.no_cpu_baseline and world == 1:   # (rank 0 at N = 1 only)
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
            # dominant kernel of the step: the Riccati backward pass (one CTA per instance, contractions on the fp64 tensor cores)
            "roofline": {"kernel": "k_backward (Riccati backward pass, five contractions per knot as mma.sync m8n8k4 f64 = SASS DMMA)",
                         "bound": "tensor", "achieved": bwd_knots * BWD_FLOPS_PER_KNOT / bwd_s / 1e12, "peak": dmma_peak,
                         "unit": "TFLOP/s", "frac": bwd_knots * BWD_FLOPS_PER_KNOT / bwd_s / 1e12 / max(dmma_peak, 1e-9),
                         "traffic": BWD_TRAFFIC_BYTES_PER_KNOT * B * N_HORIZON,
                         "traffic_note": "ncu --set full dram read+write of one full-batch launch (profiles/r01i_ncu_top_kernels.txt: 61.7 KB per knot, "
                                         "algorithmic 60.7 KB), scaled to this batch",
                         "peak_source": "fp64 tensor-core peak measured live in this run (mma.sync m8n8k4 f64 probe kernel); MEASURED_PEAKS.json "
                                        "carries HBM and bf16 figures only, and the bf16 tcgen05 peak does not apply to an fp64 path",
                         "flops_per_knot": BWD_FLOPS_PER_KNOT, "knot_passes": bwd_knots,
                         "share_of_step": tm["backward_ms"] / max(tm["total_ms"], 1e-9),
                         "hbm": {"achieved": bwd_knots * BWD_ALG_BYTES_PER_KNOT / bwd_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": bwd_knots * BWD_ALG_BYTES_PER_KNOT / bwd_s / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback"},
                         "note": "achieved = ALGORITHMIC flops (SURVEY 8(d), no symmetry credit) / time; the kernel executes 2028 DMMA = 1.04 MFLOP per knot (lower triangles of Qxx / Vxx only), DMMA pipe 57 % active in the r01i capture; bound by the sequential section of a knot (pivoted LDL^T of Quu) beside the A-block contractions and by two resident instances per SM, see DESIGN.md"},
            # second stage: analytic linearization = 3 tangent kernels (FMA pipe) + k_linearize_finish (DMMA)
            "roofline_linearize": {"kernel": "k_linearize_tangents<0|1|2> + k_linearize_finish",
                                   "bound": "fp64", "achieved_tflops": knots * LIN_FLOPS_PER_KNOT / lin_s / 1e12, "peak_tflops": fp64_peak,
                                   "frac": knots * LIN_FLOPS_PER_KNOT / lin_s / 1e12 / max(fp64_peak, 1e-9),
                                   "flops_per_knot": LIN_FLOPS_PER_KNOT, "peak_source": "measured live (DFMA kernel)",
                                   "hbm": {"achieved": knots * LIN_ALG_BYTES_PER_KNOT / lin_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                           "frac": knots * LIN_ALG_BYTES_PER_KNOT / lin_s / 1e9 / hbm_peak,
                                           "traffic": LIN_TRAFFIC_BYTES_PER_KNOT * B * N_HORIZON},
                                   "share_of_step": tm["linearize_ms"] / max(tm["total_ms"], 1e-9)},
            "stage_ms_per_solve": stage,
            "single_instance_ms_per_mpc_step": single_ms,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---- per-knot work figures (ncu numbers from profiles/r01i_ncu_top_kernels.txt, 4096 x 25 knots) ----
# Riccati backward pass: algorithmic flops of one knot (the five contractions with shared products, the
# factorisation and the solves) and its algorithmic bytes (A, B, lx, lu, lxx, luu read; K, kff written).
BWD_FLOPS_PER_KNOT = 1.153e6
BWD_ALG_BYTES_PER_KNOT = 52816.0 + 7904.0
BWD_TRAFFIC_BYTES_PER_KNOT = (5.500e9 + 0.817e9) / (4096 * 25)      # dram__bytes_read.sum + dram__bytes_write.sum of one launch
# Linearization: fp64 operations EXECUTED per knot (2 per DFMA, 1 per DADD / DMUL from the smsp__sass_thread_inst_executed_op_*
# counters: tangent kernels 10.22 + 10.03 + 3.71 GFLOP, finish 2.52 GFLOP, plus 255 DMMA m8n8k4 = 130.6 kflop per knot in finish)
LIN_FLOPS_PER_KNOT = (10.22e9 + 10.03e9 + 3.71e9 + 2.52e9) / (4096 * 25) + 255 * 512.0
# algorithmic bytes: A_k, B_k written; x_k, u_k and the factor (L, D, a) read; the parked tangents written and read once
LIN_ALG_BYTES_PER_KNOT = (51 * 51 + 51 * 19 + 51 + 19) * 8.0 + (25 * 11 + 25 + 25) * 8.0 + 2 * 48 * 25 * 8.0
LIN_TRAFFIC_BYTES_PER_KNOT = (0.171e9 + 1.201e9 + 0.181e9 + 0.809e9 + 0.117e9 + 0.299e9 + 1.954e9 + 2.870e9) / (4096 * 25)

if __name__ == "__main__":
    main()
