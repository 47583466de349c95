"""Experiment: does running the batch as G independent sub-batches on G streams (one host thread each) overlap
stages with different bottlenecks? Usage: explore_split.py B G [G ...]"""
import os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mpc_ilqr_mujoco_b200 import Config, gpu
import torch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
w = Config().build_weights()
probe = gpu.H1IlqrBatch(w, N=25, batch=1)
win, x0 = bench.workload(B, 0, probe.reference_kinematics)
ug = np.zeros(19); ug[:18] = probe.bias_forces(bench._standing()[None])[0][7:25]
for G in [int(a) for a in (sys.argv[2:] or ["1", "2"])]:
    sub = B // G
    hs = []
    for g in range(G):
        s = gpu.H1IlqrBatch(w, N=25, batch=sub)
        sl = slice(g * sub, (g + 1) * sub)
        s.set_reference_window(*(a[sl] for a in win), shared=False)
        s.upload_inputs(x0[sl], ug)
        hs.append(s)
    def run(s, n): s.run_resident_steps(n, True)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=run, args=(s, 2)) for s in hs]
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 2
        print(f"B {B} G {G} rep {rep}: {dt * 1e3:.1f} ms/step, {B / dt:.0f} solves/s", flush=True)
    for s in hs: s.close()
