#!/usr/bin/env python
"""bench.py -- forward+backward particle-substeps/s of the differentiable MLS-MPM hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload D|A|B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (N=1 and per rank for N>1, i.e. weak scaling over independent environments): BASELINE config D --
1 000 000 particles, 128^3 grid (dx=1/128, dt=2.5e-5), 19 hand-like primitives, one "step" = 80 substeps forward with
per-substep checkpoints + 80 substeps backward (loss = -mean(y) of the final state).  Synthetic, seeded.

One JSON line on stdout (rank 0).  `value` = particle-substeps/s with state resident in HBM (CUDA events, max over ranks);
`e2e` = the same metric through the public host API: state, poses and loss gradient uploaded from pinned host buffers every step
(plus the re-sort), loss taken on the device and read back with the pose gradients; `roofline` = the
dominant kernel's algorithmic bytes / its device time against the measured HBM peak; `cpu_baseline` = the reference's
kernels built for the host (oracle/_ref) or the C oracle port on a bounded sample.
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
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

METRIC = "particle-substeps/sec fwd+bwd"
UNIT = "particle-substeps/s"
ALG_BYTES_FWD, ALG_BYTES_BWD = 212, 308  # SURVEY.md 8(d), per particle-substep
# algorithmic bytes per particle of each kernel (DESIGN.md "kernels"): grid kernels work on the L2-resident grid
KERNEL_ALG_BYTES = {"p2g_tile": 152, "g2p_tile": 60, "g2p_grad_tile": 60, "p2g_grad_tile": 248, "g2p": 60, "p2g_grad": 248}


def workload_scene(name, seed=0):
    from dexdeform_b200.scenes import make_scene, scene_flip, scene_tutorial
    if name == "D":
        S = 80
        sc = make_scene(1000000, 128, box_center=(0.5, 0.3, 0.5), box_width=(0.4, 0.4, 0.4), steps=S, seed=seed, hand_scale=6.0)
        desc = "D: 1M particles, 128^3 grid, 19 primitives, 80 substeps fwd + 80 bwd with per-substep checkpoints"
    elif name == "A":
        S = 50
        sc = scene_tutorial(steps=S, seed=seed)
        desc = "A: tutorial scene, 10k particles, 64^3 grid, 19 primitives, 50 substeps fwd + bwd"
    else:
        S = 40
        sc = scene_flip(steps=S, seed=seed)
        desc = "B: flip scene, 50k particles, 64^3 grid, 19 primitives, 40 substeps fwd + bwd"
    return sc, S, desc


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def dist_setup(n_gpus):
    import torch
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier_sync(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    from dexdeform_b200.engine import FusedSim
    rank, world, local = dist_setup(args.gpus)
    sc, S, desc = workload_scene(args.workload, seed=rank)
    n, nb = sc["n"], sc["nb"]
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sim = FusedSim.from_scene(sc, n_envs=1, max_steps=S, stream=stream.cuda_stream)
    gx_host = np.zeros((1, n, 3), np.float32)
    gx_host[..., 1] = -1.0 / n  # d(-mean y)/dx
    gx_dev = torch.from_numpy(gx_host).cuda()
    packed = torch.zeros(1 + (S + 1) * nb * 7, dtype=torch.float32, device="cuda")  # [loss | pose grads] all-reduced over ranks

    def step(collective=True):
        sim.forward(0, S)
        sim.zero_grad(S)
        sim._check(sim.lib.dd_sim_add_state_grad(sim._h, S, gx_dev.data_ptr(), None, None, None, sim.stream))
        sim.backward(0, S)
        if world > 1 and collective:  # NCCL over NVLink: loss and pose (action) gradients only; environments never exchange state
            import torch.distributed as dist
            base = packed.data_ptr() + 4  # [0] is the loss; pose gradients are written device-to-device behind it
            sim._check(sim.lib.dd_sim_get_pose_grads(sim._h, 0, S + 1, base, base + 4 * 3 * (S + 1) * nb, sim.stream))
            dist.all_reduce(packed)

    for _ in range(args.warmup):
        step()
    barrier_sync(world)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier_sync(world)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier_sync(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    launches = sim.launch_count() - launches0  # kernels of the timed region only
    t_load = time.perf_counter()
    while len(sampler.samples) < 6 and time.perf_counter() - t_load < 4.0:  # short timed region: keep the same load up (untimed) for more clock samples
        step(collective=False)  # rank-local trip count: no collective in here, or the ranks' NCCL sequences diverge
    torch.cuda.synchronize()
    sampler.stop_flag = True
    units = float(world) * n * S * args.steps
    value = units / (ms * 1e-3)

    # ---- end to end through the public host API: pinned host buffers in, loss + pose gradients out, every step
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hx, hv, hF, hC = (pin(sc[k][None]) for k in ("x", "v", "F", "C"))
    hpos, hrot = pin(sc["pos"][:, None]), pin(sc["rot"][:, None])
    x_dev = torch.empty((1, n, 3), dtype=torch.float32, device="cuda")   # final positions stay on the device: the loss is taken there
    loss_out = torch.empty(1, dtype=torch.float32).pin_memory()
    gp_out = torch.empty((S + 1, 1, nb, 3), dtype=torch.float32).pin_memory()
    gr_out = torch.empty((S + 1, 1, nb, 4), dtype=torch.float32).pin_memory()
    h2d = sum(t.numel() * 4 for t in (hx, hv, hF, hC, hpos, hrot))
    d2h = sum(t.numel() * 4 for t in (loss_out, gp_out, gr_out))
    P = lambda t: t.data_ptr()

    def e2e_step():
        sim._check(sim.lib.dd_sim_set_state(sim._h, 0, P(hx), P(hv), P(hF), P(hC), sim.stream))   # H2D + re-sort
        sim._check(sim.lib.dd_sim_set_poses(sim._h, 0, S + 1, P(hpos), P(hrot), sim.stream))
        sim.forward(0, S)
        sim._check(sim.lib.dd_sim_get_state(sim._h, S, P(x_dev), None, None, None, sim.stream))     # caller's particle order, on the device
        loss_out.copy_(-x_dev[0, :, 1].mean(), non_blocking=True)                                     # loss on the device (as GradModel users do), D2H of the scalar
        torch.cuda.current_stream().synchronize()
        loss = float(loss_out[0])
        sim.zero_grad(S)
        sim._check(sim.lib.dd_sim_add_state_grad(sim._h, S, gx_dev.data_ptr(), None, None, None, sim.stream))  # d loss / d x, produced on the device like the loss
        sim.backward(0, S)
        sim._check(sim.lib.dd_sim_get_pose_grads(sim._h, 0, S + 1, P(gp_out), P(gr_out), sim.stream))
        return loss

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier_sync(world)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        loss = e2e_step()
    barrier_sync(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    e2e_value = float(world) * n * S * e2e_steps / e2e_s

    # ---- roofline of the dominant kernel: per-kernel device times of one fwd+bwd substep (CUDA events on our stream)
    sim.forward(0, S)
    sim.zero_grad(S)
    sim.add_state_grad(S, gx_host)
    prof = sim.profile_substep(S - 1, reps=10)
    sim.sync()
    peak, peak_src = measured_peak()
    total_ms = sum(ms_k for _, ms_k in prof)
    dom_name, dom_ms = max(prof, key=lambda kv: kv[1])
    key = next(k for k in KERNEL_ALG_BYTES if dom_name.startswith(k)) if any(dom_name.startswith(k) for k in KERNEL_ALG_BYTES) else None
    alg = (KERNEL_ALG_BYTES[key] if key else 0) * n
    achieved = alg / (dom_ms * 1e-3) / 1e9
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = None
    if os.path.isfile(traffic_path) and key:
        traffic = json.load(open(traffic_path)).get(key)
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_particle": KERNEL_ALG_BYTES.get(key),
                "share_of_substep": round(dom_ms / total_ms, 3),
                "path": {"algorithmic_bytes_per_particle_substep": ALG_BYTES_FWD + ALG_BYTES_BWD,
                         "achieved": round((ALG_BYTES_FWD + ALG_BYTES_BWD) * value / world / 1e9, 1),
                         "frac": round((ALG_BYTES_FWD + ALG_BYTES_BWD) * value / world / 1e9 / peak, 4)},
                "kernels_us": {k: round(v * 1e3, 1) for k, v in prof}}

    out = None
    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "impl": "ours",
               "config": {"workload": desc, "particles_per_gpu": n, "substeps_per_step": S, "grid": int(sc["grid_dim"][0]),
                          "l2_policy": "inputs larger than L2 (per-substep checkpoints: 180 MB/substep + 67 MB grids, 19.8 GB/step)",
                          "parallelism": f"env-batch x{world}, NCCL all-reduce of loss + pose gradients" if world > 1 else "single GPU"},
               "clocks": sampler.summary(),
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps, "loss": loss},
               "gpu_launches": int(launches), "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args.workload)
    sim.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------------------------------- CPU baseline
def cpu_baseline(workload, budget_substeps=24):
    """The reference's own kernels compiled for the host (oracle/_ref, kind "reference") -- else the C oracle (kind "port") --
    on a bounded sample of the same workload: the full scene, `budget_substeps` substeps forward + backward."""
    from abi1_driver import Abi1Sim
    from oracle import oracle_lib
    sc, S, desc = workload_scene(workload)
    n = sc["n"]
    if os.path.isfile(oracle_lib.REF_CPU):
        lib, kind = oracle_lib.load_ref_cpu(), "reference"
        cores = lib.ref_cpu_num_threads()
    else:
        lib, kind = oracle_lib.OracleLib(), "port"
        cores = lib.num_threads()
    K = budget_substeps
    sim = Abi1Sim(lib, sc, K)
    gx = np.zeros((n, 3), np.float32)
    gx[:, 1] = -1.0 / n
    t0 = time.perf_counter()
    for f in range(K):
        sim.substep(f)
    sim.states[K]["x_grad"].upload(gx)
    for f in range(K - 1, -1, -1):
        sim.substep_grad(f)
    sim.sync()
    dt = time.perf_counter() - t0
    return {"value": n * K / dt, "unit": UNIT, "cores": int(cores), "kind": kind,
            "sample": f"{K} of the workload's {S} substeps (forward + backward) on the full {n}-particle scene, OpenMP over {cores} host threads, {dt:.1f} s"}


# ------------------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The unmodified reference (oracle/_ref, built from /root/reference by oracle/build_ref.sh) through its own call
    sequence (mpm/simulator.py:553-585: set_pose upload + substep per substep, substep_grad + pose-gradient download per
    substep) on the same workload.  The reference has no CPU path (mpm/types.py:12-17 needs nvcc), so with a GPU present
    this runs its CUDA build; without one it runs the host build of the same sources on all cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return None
    from abi1_driver import Abi1Sim
    from oracle import oracle_lib
    sc, S, desc = workload_scene(args.workload)
    n, nb = sc["n"], sc["nb"]
    have_gpu = False
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    if have_gpu and os.path.isfile(oracle_lib.REF_GPU):
        lib, where, cores = oracle_lib.load_ref_gpu(), "reference CUDA build (oracle/_ref/libmaniskill_mpm.so) on the same GPU", 0
        K = S
    elif os.path.isfile(oracle_lib.REF_CPU):
        lib, where = oracle_lib.load_ref_cpu(), "reference host build (oracle/_ref/libmaniskill_mpm_cpu.so)"
        cores, K = lib.ref_cpu_num_threads(), 1
    else:
        return {"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}
    sim = Abi1Sim(lib, sc, K)
    gx = np.zeros((n, 3), np.float32)
    gx[:, 1] = -1.0 / n

    def step():
        for f in range(K):   # as shipped: pose upload then substep, every substep (simulator.py:626-634)
            sim.states[f + 1]["body_pos"].upload_async(sc["pos"][f + 1], sim.stream)
            sim.states[f + 1]["body_rot"].upload_async(sc["rot"][f + 1], sim.stream)
            sim.substep(f)
            for k in ("x_grad", "v_grad", "F_grad", "C_grad", "body_pos_grad", "body_rot_grad"):   # clear_grad=True (simulator.py:570-571)
                sim.states[f + 1][k].zero(sim.stream)
        sim.sync()
        sim.states[K]["x_grad"].upload(gx)
        for f in range(K - 1, -1, -1):   # torch_wrapper.py:128-134
            sim.substep_grad(f)
            sim.states[f + 1]["body_pos_grad"].download()
            sim.states[f + 1]["body_rot_grad"].download()
        sim.sync()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = n * K * args.steps / dt
    return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": desc, "particles_per_gpu": n, "substeps_per_step": K, "grid": int(sc["grid_dim"][0]), "ran": where},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(cores), "kind": "reference",
                             "sample": f"{K} substeps forward + backward per step, {where}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="D", choices=["D", "A", "B"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    out = run_ours(args) if args.impl == "ours" else run_reference(args)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
