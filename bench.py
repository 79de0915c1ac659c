#!/usr/bin/env python
"""bench.py -- forward+backward particle-substeps/s of the differentiable MLS-MPM hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|D|A|B|C|E]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (synthetic, seeded):
  D  1 000 000 particles, 128^3 grid (dx=1/128, dt=2.5e-5), 19 hand-like primitives; one step = 80 substeps forward with
     per-substep checkpoints + 80 substeps backward, loss = -mean(y).  The configuration the metric is quoted on: default at N=1.
  A  tutorial scene (10k particles, 64^3), 50 substeps fwd+bwd.      B  flip-sized scene (50k particles), 40 substeps fwd+bwd.
  C  64 flip scenes in one engine, forward replay only (demonstration scoring).
  E  512 tutorial scenes with the Shadow hand, 10 env steps x 40 substeps forward+backward through the batched GradModel
     (torch FK -> poses -> engine; device re-sort at every env step), environments split over the ranks (512/256/128/64 per GPU at
     1/2/4/8 GPUs, processed in sub-batches that fit HBM), one NCCL all-reduce of [loss | action gradients] per step: strong
     scaling.  Default for N>1 (a single scene is never decomposed, so D cannot use more than one GPU).

One JSON line on stdout (rank 0).  `value` = particle-substeps/s with state resident in HBM (CUDA events on the launch stream,
max over ranks); `e2e` = the same metric through the operator boundary the reference's users call -- MPMSimulator.set_state from
pinned host memory (H2D + cell sort), GradModel.get_obs / forward per env step, loss.backward(), action gradients and loss read
back to the host -- every step; `roofline` = the dominant kernel's algorithmic bytes / its device time against the measured HBM
peak; `cpu_baseline` = the reference's kernels built for the host (oracle/_ref) on a bounded sample.  Every run checks the loss
against a committed value for the workload, so a broken kernel cannot post a number.
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
# loss of one step of every workload, measured on a B200 with the parity-tested build (tests/test_parity_large_gpu.py); a run
# that does not reproduce it to 2e-4 is rejected.  None = not pinned (printed, not checked).
EXPECTED_LOSS = {"D": -0.2996545, "A": None, "B": None, "C": None, "E": None}
LOSS_FILE = os.path.join(ROOT, "profiles", "bench_expected_loss.json")
if os.path.isfile(LOSS_FILE):
    EXPECTED_LOSS.update(json.load(open(LOSS_FILE)))


def check_loss(workload, loss):
    exp = EXPECTED_LOSS.get(workload)
    if exp is None:
        return "unpinned"
    if not abs(loss - exp) <= 2e-4 * max(1.0, abs(exp)):
        raise SystemExit(f"bench.py: workload {workload} produced loss {loss!r}, expected {exp!r}: results are wrong, no number reported")
    return "ok"


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
    """SM clock and throttle reasons of THIS rank's GPU sampled while the timed region runs: NVML in-process (the library behind
    nvidia-smi; one cheap query every 100 ms), falling back to the nvidia-smi command line.  Spawning nvidia-smi from all eight
    ranks ten times a second serialises on the driver and showed up as lost throughput at N=8, hence NVML first."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.source = index, [], False, "nvidia-smi"
        self.nvml = self.handle = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(index).uuid)).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        n, h = self.nvml, self.handle
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(reasons_fn(h))
        return [str(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)), str(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)),
                str(n.nvmlDeviceGetPowerUsage(h) / 1000.0)] + ["Active" if mask & bit else "Not Active" for _, bit in self.REASONS]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self.sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1 if self.nvml is not None else 0.5)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def dist_setup():
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


def finish(world):
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def keep_load_for_clock_samples(sampler, fn, min_samples=6, budget_s=4.0):
    import torch
    t0 = time.perf_counter()
    while len(sampler.samples) < min_samples and time.perf_counter() - t0 < budget_s:  # short timed region: keep the same load up (untimed)
        fn()
    torch.cuda.synchronize()
    sampler.stop_flag = True


# --------------------------------------------------------------------------------------------- our arm: D, A, B
def free_tool_actions(sc, S):
    """Actions (nb, 6) that make MPMSimulator.compute_forward_kinematics (mpm/simulator.py:597-624) reproduce the scene's pose
    trajectory over one env step of S substeps: constant velocity and spin per primitive."""
    from dexdeform_b200.scenes import _qmul
    scale_t, scale_r = 0.05, 0.05
    dpos = sc["pos"][S] - sc["pos"][0]
    q0 = sc["rot"][0].astype(np.float64)
    dq = _qmul(q0 * np.array([1, -1, -1, -1.0]), sc["rot"][S].astype(np.float64))
    ang = 2 * np.arccos(np.clip(dq[:, :1], -1, 1))
    axis = dq[:, 1:] / np.maximum(np.sin(ang / 2), 1e-12)
    act = np.concatenate([dpos / scale_t, axis * ang / scale_r], 1).astype(np.float32)
    assert np.abs(act).max() < 1.0, "the scene's tools move too fast for the action scale"
    return act, [[scale_t] * 3 + [scale_r] * 3] * sc["nb"]


def run_single(args, workload, rank, world, local):
    import torch
    from dexdeform_b200.engine import FusedSim
    sc, S, desc = workload_scene(workload, seed=rank)
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
        sim.add_state_grad(S, gx=gx_dev)
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
    keep_load_for_clock_samples(sampler, lambda: step(collective=False))  # rank-local trip count: no collective in here
    units = float(world) * n * S * args.steps
    value = units / (ms * 1e-3)
    loss_dev = float(-sim.get_state(S, ("x",), device=True)["x"][0, :, 1].mean())
    loss_check = check_loss(workload, loss_dev) if rank == 0 else "n/a"

    # ---- roofline of the dominant kernel: per-kernel device times of one fwd+bwd substep (CUDA events on our stream)
    sim.forward(0, S)
    sim.zero_grad(S)
    sim.add_state_grad(S, gx=gx_dev)
    prof = sim.profile_substep(S - 1, reps=10)
    sim.sync()
    sim.close()
    del sim
    peak, peak_src = measured_peak()
    # per-kernel times come from events between individual launches and include the launch gaps the captured graph does not
    # have: only their SHARES are used, scaled to the substep time of the timed region (graph launches, CUDA events)
    total_ms = sum(ms_k for _, ms_k in prof)
    scale = (ms / args.steps / S) / total_ms
    prof = [(k, ms_k * scale) for k, ms_k in prof]
    total_ms = sum(ms_k for _, ms_k in prof)
    dom_name, dom_ms = max(prof, key=lambda kv: kv[1])
    key = next((k for k in KERNEL_ALG_BYTES if dom_name.startswith(k)), None)
    alg = (KERNEL_ALG_BYTES[key] if key else 0) * n
    achieved = alg / (dom_ms * 1e-3) / 1e9
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = None
    if os.path.isfile(traffic_path) and key and workload == "D":
        traffic = json.load(open(traffic_path)).get(key)
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_particle": KERNEL_ALG_BYTES.get(key),
                "share_of_substep": round(dom_ms / total_ms, 3),
                "path": {"algorithmic_bytes_per_particle_substep": ALG_BYTES_FWD + ALG_BYTES_BWD,
                         "achieved": round((ALG_BYTES_FWD + ALG_BYTES_BWD) * value / world / 1e9, 1),
                         "frac": round((ALG_BYTES_FWD + ALG_BYTES_BWD) * value / world / 1e9 / peak, 4)},
                "kernels_us": {k: round(v * 1e3, 1) for k, v in prof},
                "kernels_us_note": "shares from per-launch CUDA events, scaled to sum to the substep time of the timed region"}

    # ---- end to end through the operator boundary: MPMSimulator.set_state (pinned host -> device, cell sort) + GradModel
    e2e = e2e_gradmodel(args, sc, S, world, stream)
    if rank == 0 and abs(e2e["loss"] - loss_dev) > 2e-4 * max(1.0, abs(loss_dev)):
        raise SystemExit(f"bench.py: GradModel path loss {e2e['loss']} differs from the engine path loss {loss_dev}")

    out = None
    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "impl": "ours",
               "config": {"workload": desc, "particles_per_gpu": n, "substeps_per_step": S, "grid": int(sc["grid_dim"][0]),
                          "l2_policy": "inputs larger than L2 (per-substep checkpoints: 180 B/particle/substep + grids; 19.8 GB/step at D)" if n >= 500000
                          else "L2 flushed by the step itself only where the checkpoints exceed it; small scene, latency bound",
                          "parallelism": f"one scene per rank x{world}, NCCL all-reduce of loss + pose gradients" if world > 1 else "single GPU"},
               "clocks": sampler.summary(), "loss": loss_dev, "loss_check": loss_check,
               "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(workload)
    return out


def e2e_gradmodel(args, sc, S, world, stream):
    """One optimisation step exactly as a user of the reference writes it (tutorials/1_trajectory_optimization.ipynb:184-199):
    set_state from host memory, zero_grad, get_obs, forward per env step, loss.backward(), read loss and action gradients."""
    import torch
    from dexdeform_b200.simulator import MPMSimulator
    from dexdeform_b200.torch_wrapper import GradModel
    n, nb = sc["n"], sc["nb"]
    act0, scales = free_tool_actions(sc, S)
    sim = MPMSimulator(nb, ground_friction=sc["ground_friction"], gravity=tuple(sc["gravity"].reshape(3) / 30), n_particles=n, dx=sc["dx"],
                       dt=sc["dt"], max_steps=S, substeps=S, stream=stream.cuda_stream)
    sim.init_particles(sc["vol"], sc["mass"], sc["mu_lam_yield"])
    sim.init_bodies(sc["tfsr"][:, 0], sc["tfsr"][:, 2], sc["tfsr"][:, 1], sc["tfsr"][:, 3], sc["args"], action_scales=scales, pos=sc["pos"][0], rot=sc["rot"][0])
    model = GradModel(sim, return_grid=())
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hx, hv, hF, hC = (pin(sc[k][None]) for k in ("x", "v", "F", "C"))
    hact = pin(act0[None])                                        # (1 env step, nb, 6)
    hgrad, hloss = torch.empty_like(hact).pin_memory(), torch.empty(1).pin_memory()
    h2d = sum(t.numel() * 4 for t in (hx, hv, hF, hC, hact))
    d2h = sum(t.numel() * 4 for t in (hgrad, hloss))

    # Input pipeline: the state of step k+1 is copied host -> device on a copy stream while step k computes (two device buffer
    # sets); set_state then takes the device copy (device-to-device + cell sort).  Every step still moves its 96 MB over PCIe
    # inside the timed region -- the copy engine works beside the kernels instead of in front of them.
    copy_stream = torch.cuda.Stream()
    dev = [[torch.empty_like(t, device="cuda") for t in (hx, hv, hF, hC)] for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    count = [0]

    def prefetch(slot):
        copy_stream.wait_stream(torch.cuda.current_stream())   # the previous consumer of this buffer set (two steps ago) is done
        with torch.cuda.stream(copy_stream):
            for d, h in zip(dev[slot], (hx, hv, hF, hC)):
                d.copy_(h, non_blocking=True)
            ready[slot].record(copy_stream)

    prefetch(0)

    def step():
        slot = count[0] & 1
        count[0] += 1
        torch.cuda.current_stream().wait_event(ready[slot])
        sim.engine.set_state(0, *dev[slot])   # device copy of this step's state -> engine slot 0 (cell sort)
        prefetch(slot ^ 1)                   # next step's state: H2D beside this step's kernels
        model.zero_grad()
        action = hact.to("cuda", non_blocking=True).requires_grad_(True)
        obs = model.get_obs(0, "cuda")
        obs = model.forward(0, action[0], *obs)
        loss = -obs[0][:, 1].mean()
        loss.backward()
        hgrad.copy_(action.grad, non_blocking=True)
        hloss.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(hloss[0])

    k = max(2, min(args.steps, 5))
    step()
    barrier_sync(world)
    t0 = time.perf_counter()
    for _ in range(k):
        loss = step()
    barrier_sync(world)
    dt = max_over_ranks(time.perf_counter() - t0, world)
    sim.engine.close()
    return {"value": float(world) * n * S * k / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": k, "loss": loss,
            "path": "pinned host state -> device (copy stream, double buffered: the copy of step k+1 overlaps the kernels of step k) -> MPMSimulator.set_state (cell sort) -> GradModel.get_obs/forward -> loss.backward() -> action.grad, loss to host",
            "action_grad_norm": float(hgrad.norm())}


# --------------------------------------------------------------------------------------------- our arm: C (forward replay)
def run_C(args, rank, world, local):
    import torch
    from dexdeform_b200.engine import FusedSim
    from dexdeform_b200.scenes import scene_flip
    E_total, S = 64, 40
    from dexdeform_b200.batch import partition_envs
    first, E = partition_envs(E_total, world, rank)
    scs = [scene_flip(steps=S, seed=first + e) for e in range(E)]
    n, nb = scs[0]["n"], scs[0]["nb"]
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sc0 = scs[0]
    sim = FusedSim(E, n, nb, sc0["grid_dim"], sc0["dx"], sc0["dt"], S, sc0["ground_friction"], sc0["ground_height"], sc0["gravity"].reshape(3),
                   grid_ckpt=False, stream=stream.cuda_stream)
    st = lambda k: np.ascontiguousarray(np.stack([sc[k] for sc in scs]))
    sim.set_material(st("mass"), st("vol"), st("mu_lam_yield"))
    sim.set_bodies(sc0["tfsr"], sc0["args"])
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hpos, hrot = pin(np.stack([sc["pos"] for sc in scs], 1)), pin(np.stack([sc["rot"] for sc in scs], 1))
    hstate = [pin(st(k)) for k in ("x", "v", "F", "C")]
    sim.set_poses(0, hpos, hrot)
    sim.set_state(0, *hstate)
    for _ in range(args.warmup):
        sim.forward(0, S)
    barrier_sync(world)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        sim.forward(0, S)
    e1.record(stream)
    barrier_sync(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    launches = sim.launch_count() - l0
    keep_load_for_clock_samples(sampler, lambda: sim.forward(0, S))
    hscore = torch.empty(E).pin_memory()

    def e2e_step():   # demonstration scoring: states and poses from the host, one score per environment back
        sim.set_state(0, *hstate)
        sim.set_poses(0, hpos, hrot)
        sim.forward(0, S)
        x = sim.get_state(S, ("x",), device=True)["x"]
        hscore.copy_(-x[..., 1].mean(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(hscore.mean())

    k = max(2, min(args.steps, 5))
    e2e_step()
    barrier_sync(world)
    t0 = time.perf_counter()
    for _ in range(k):
        loss = e2e_step()
    barrier_sync(world)
    dt = max_over_ranks(time.perf_counter() - t0, world)
    out = None
    if rank == 0:
        check = check_loss("C", loss) if world == 1 else "n/a"
        peak, src = measured_peak()
        value = float(E_total) * n * S * args.steps / (ms * 1e-3)
        out = {"metric": "particle-substeps/sec fwd (forward replay)", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "ours",
               "config": {"workload": "C: 64 flip scenes (50k particles, 64^3) replayed forward, 40 substeps, one engine per rank", "envs_per_gpu": E,
                          "particles_per_gpu": E * n, "substeps_per_step": S, "l2_policy": "inputs larger than L2 (307 MB of state per substep)"},
               "clocks": sampler.summary(), "loss": loss, "loss_check": check, "gpu_launches": int(launches),
               "e2e": {"value": float(E_total) * n * S * k / dt, "unit": UNIT, "h2d_bytes_per_step": sum(t.numel() * 4 for t in hstate + [hpos, hrot]),
                       "d2h_bytes_per_step": E * 4, "steps": k},
               "roofline": {"bound": "hbm", "achieved": round(ALG_BYTES_FWD * value / world / 1e9, 1), "peak": peak, "unit": "GB/s",
                            "frac": round(ALG_BYTES_FWD * value / world / 1e9 / peak, 4), "traffic": None, "peak_source": src,
                            "note": "forward path figure (212 B per particle-substep) over the whole step"}}
    sim.close()
    return out


# --------------------------------------------------------------------------------------------- our arm: E (512 environments)
def shadow_tables(tag="rh15"):
    from dexdeform_b200.mujoco_parser import HandTables
    z = np.load(os.path.join(ROOT, "tests", "golden", "shadow_tables.npz"))
    return HandTables(**{k.split(".", 1)[1]: (int(z[k]) if k.endswith("n_hands") else z[k]) for k in z.files if k.startswith(tag + ".")})


def make_hand_batch(E, T, S_env, stream):
    """E copies of the lift_box scene (mpm/assets/env_cfgs/lift_box.yml) with the right Shadow hand (scale 1.5, 19 primitives)."""
    import torch
    from dexdeform_b200.hand import HandSimulator
    from dexdeform_b200.rotations import euler2mat
    from dexdeform_b200.scenes import scene_tutorial
    tables = shadow_tables("rh15")
    nb = len(tables.prim_type)
    cfg = dict(n_particles=10000, E=5e3, nu=0.2, yield_stress=50.0, ground_friction=0.3, quality=1, max_steps=T * S_env, gravity=(0.0, -2.0, 0.0),
               fixed_base=False)
    sim = HandSimulator(nb, {"tables": tables}, cfg=cfg, n_envs=E, stream=stream.cuda_stream)   # grids: brick checkpoints (a dense pair per substep would be 215 GB)
    assert sim.substeps == S_env
    sim.init_bodies(tables.prim_type.astype(np.float32), np.full(nb, 666.0, np.float32), np.full(nb, 0.9, np.float32), np.zeros(nb, np.float32),
                    tables.prim_size, action_scales=[()] * nb)
    sc = scene_tutorial(steps=1, seed=0)                       # the lift_box block (lift_box.yml:17-21)
    root = np.eye(4)
    root[:3, :3] = euler2mat(0.0, 0.0, np.pi)                   # lift_box.yml MANIPULATORS: init_pos (0.5, 0.2, 0.3), init_rot (0, 0, pi), qpos zero
    root[:3, 3] = (0.5, 0.2, 0.3)
    base = torch.tensor(root[None], dtype=torch.float32, device="cuda")       # (nh, 4, 4)
    q0 = torch.zeros((1, 24), dtype=torch.float32, device="cuda")
    pos, rot = sim.hand_forward_kinematics(base[None], q0[None])              # primitive poses of state 0 (hand.py:190-192)
    sim.set_poses(0, pos, rot)
    return sim, sc, base, q0


def run_E(args, rank, world, local):
    import torch
    from dexdeform_b200.batch import partition_envs
    from dexdeform_b200.torch_wrapper import GradModel
    E_total, T, S_env, n = 512, 10, 40, 10000
    first, E_local = partition_envs(E_total, world, rank)
    sub = min(E_local, int(os.environ.get("DD_BENCH_SUBBATCH", 128)))   # environments per engine pass: 411 slots x 180 B x sub x 10k particles of checkpoints
    assert E_local % sub == 0
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sim, sc, base0, q0 = make_hand_batch(sub, T, S_env, stream)
    model = GradModel(sim, return_grid=())
    rng = np.random.default_rng(7)
    act_shared = np.float32(rng.uniform(-0.4, 0.4, (T, 1, 26)))          # the action sequence being optimised ...
    act_shared[:, :, 21] = -0.5                                           # ... presses the hand onto the block
    noise = np.float32(rng.normal(size=(E_total, T, 1, 26)) * 0.2)       # per-environment exploration noise (same for any world size)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hstate = [pin(sc[k]) for k in ("x", "v", "F", "C")]                   # ONE environment's initial state; tiled on the device
    hact = pin(act_shared)
    hnoise = pin(noise)
    passes_local = range(first // sub, (first + E_local) // sub)         # engine passes of `sub` environments this rank owns
    packed = torch.zeros(1 + T * 26, dtype=torch.float32, device="cuda")  # [sum of losses | d sum / d shared action]
    hout = torch.empty_like(packed, device="cpu").pin_memory()
    state_dev = [None]
    nz_dev = hnoise.to("cuda")

    def one_pass(b, action):
        """sub environments, T env steps forward + backward through GradModel; returns their summed loss (graph attached)."""
        sim.engine.set_state(0, *state_dev[0])
        sim.base_pose[0], sim.joint_rot[0] = base0, q0
        model.zero_grad()
        a = (action[None] + nz_dev[b * sub:(b + 1) * sub]).clamp(-1, 1)   # (sub, T, 1, 26)
        obs = model.get_obs(0, "cuda")
        for j in range(T):
            obs = model.forward(j, a[:, j], *obs)
        return -obs[0][..., 1].mean(1).sum()

    def step(host_io, passes=passes_local, collective=True):
        if host_io or state_dev[0] is None:   # end to end: the initial state comes from the host every step
            state_dev[0] = [t.to("cuda", non_blocking=True)[None].expand(sub, -1, -1).contiguous() for t in hstate]
        action = (hact.to("cuda", non_blocking=True) if host_io else act_dev).clone().requires_grad_(True)
        total = 0.0
        for b in passes:
            loss = one_pass(b, action)
            loss.backward()
            total = total + loss.detach()
        packed[0] = total
        packed[1:] = action.grad.reshape(-1)
        if world > 1 and collective:   # NCCL over NVLink: loss and action gradients only
            import torch.distributed as dist
            dist.all_reduce(packed)
        if host_io:
            hout.copy_(packed, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(hout[0]) / E_total
        return None

    act_dev = hact.to("cuda")
    for _ in range(max(1, args.warmup - 1)):
        step(False)
    barrier_sync(world)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = sim.engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier_sync(world)
    e0.record(stream)
    for _ in range(args.steps):
        step(False)
    e1.record(stream)
    barrier_sync(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    launches = sim.engine.launch_count() - l0
    sampler.stop_flag = True
    step(True)
    barrier_sync(world)
    k = max(2, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(k):
        loss = step(True)
    barrier_sync(world)
    dt = max_over_ranks(time.perf_counter() - t0, world)
    units = float(E_total) * n * T * S_env
    # the same 512 environments on ONE GPU (rank 0 alone, after the timed region, the other ranks idle at the barrier): the
    # denominator of strong scaling for this workload, measured in the same run on the same box
    one_gpu = None
    if world > 1:
        if rank == 0:
            all_passes = range(E_total // sub)
            step(False, all_passes, False)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            f0.record(stream)
            for _ in range(2):
                step(False, all_passes, False)
            f1.record(stream)
            torch.cuda.synchronize()
            ms1 = f0.elapsed_time(f1) / 2
            one_gpu = {"value": units / (ms1 * 1e-3), "unit": UNIT, "ms_per_step": ms1, "steps": 2,
                       "note": f"all {E_total} environments on rank 0's GPU alone (same process and engine, {E_total // sub} passes of {sub}), other ranks idle"}
        barrier_sync(world)
    out = None
    if rank == 0:
        peak, src = measured_peak()
        value = units * args.steps / (ms * 1e-3)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "ours",
               "config": {"workload": "E: 512 tutorial scenes (10k particles, 64^3, Shadow hand 19 primitives), 10 env steps x 40 substeps fwd + bwd through the batched GradModel",
                          "envs_total": E_total, "envs_per_gpu": E_local, "envs_per_engine_pass": sub, "particles_per_gpu": E_local * n, "substeps_per_step": T * S_env,
                          "l2_policy": "inputs larger than L2 (per-substep checkpoints of the pass: 180 B x envs x 10k per substep)",
                          "parallelism": f"environments split over {world} ranks, one NCCL all-reduce of [loss | action gradients] per step" if world > 1 else "single GPU"},
               "clocks": sampler.summary(), "loss": loss, "loss_check": check_loss("E", loss), "gpu_launches": int(launches),
               "e2e": {"value": units * k / dt, "unit": UNIT, "h2d_bytes_per_step": sum(t.numel() * 4 for t in hstate + [hact]), "d2h_bytes_per_step": packed.numel() * 4, "steps": k,
                       "path": "HandSimulator.set_state(host state) -> GradModel.get_obs/forward x10 -> loss.backward() -> all-reduce -> loss + action gradients to host"},
               "roofline": {"bound": "hbm", "achieved": round((ALG_BYTES_FWD + ALG_BYTES_BWD) * value / world / 1e9, 1), "peak": peak, "unit": "GB/s",
                            "frac": round((ALG_BYTES_FWD + ALG_BYTES_BWD) * value / world / 1e9 / peak, 4), "traffic": None, "peak_source": src,
                            "note": "path figure (520 B per particle-substep) over the whole step incl. observations, FK and re-sorts; grids of the active bricks are checkpointed per substep (brick checkpoints), nothing is replayed in the adjoint"}}
        if one_gpu is not None:
            out["one_gpu_same_workload"] = one_gpu
    sim.engine.close()
    return out


def run_ours(args):
    rank, world, local = dist_setup()
    wl = args.workload
    if wl == "auto":
        wl = os.environ.get("DD_BENCH_WORKLOAD") or ("D" if world == 1 else "E")
    if wl == "E":
        out = run_E(args, rank, world, local)
    elif wl == "C":
        out = run_C(args, rank, world, local)
    else:
        out = run_single(args, wl, rank, world, local)
    finish(world)
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
    K = min(budget_substeps, S) if n >= 500000 else S
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
    sequence (mpm/simulator.py:553-585: set_pose = device->host->device pose round trip + substep per substep, substep_grad +
    two synchronous pose-gradient downloads per substep) on the same workload.  The reference has no CPU path
    (mpm/types.py:12-17 needs nvcc), so with a GPU present this runs its CUDA build; without one it runs the host build of the
    same sources on all cores.  Also reported: `kernel_only`, the same kernels without the per-substep host traffic."""
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return None
    from abi1_driver import Abi1Sim
    from oracle import oracle_lib
    wl = args.workload
    if wl == "auto":
        wl = os.environ.get("DD_BENCH_WORKLOAD") or ("D" if world == 1 else "E")
    if wl in ("E", "C"):   # the reference has no batch axis: one environment of the batch (its scenes run one after the other)
        from dexdeform_b200.scenes import scene_flip, scene_tutorial
        S = 400 if wl == "E" else 40
        sc = scene_tutorial(steps=S, seed=0) if wl == "E" else scene_flip(steps=S, seed=0)
        desc = ("E: 512 tutorial scenes ... the reference runs them one at a time; sample = 1 environment, 400 substeps fwd + bwd" if wl == "E"
                else "C: 64 flip scenes forward; sample = 1 environment, 40 substeps forward")
    else:
        sc, S, desc = workload_scene(wl)
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
        cores, K = lib.ref_cpu_num_threads(), (1 if n >= 500000 else min(S, 10))
    else:
        return {"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}
    sim = Abi1Sim(lib, sc, K)
    gx = np.zeros((n, 3), np.float32)
    gx[:, 1] = -1.0 / n
    fwd_only = wl == "C"
    pose_dev = None
    if have_gpu and nb:
        import torch
        pose_dev = (torch.tensor(sc["pos"][:K + 1], device="cuda"), torch.tensor(sc["rot"][:K + 1], device="cuda"))

    def step(as_shipped=True):
        for f in range(K):   # as shipped: pose round trip then substep, every substep (simulator.py:553-559, 626-634)
            if as_shipped and nb:
                if pose_dev is not None:   # pos.detach().cpu().numpy(): a device->host copy with a sync, then a pageable upload
                    p, r = pose_dev[0][f + 1].detach().cpu().numpy(), pose_dev[1][f + 1].detach().cpu().numpy()
                else:
                    p, r = sc["pos"][f + 1], sc["rot"][f + 1]
                sim.states[f + 1]["body_pos"].upload_async(p, sim.stream)
                sim.states[f + 1]["body_rot"].upload_async(r, sim.stream)
            sim.substep(f)
            if not fwd_only:
                for k in ("x_grad", "v_grad", "F_grad", "C_grad", "body_pos_grad", "body_rot_grad"):   # clear_grad=True (simulator.py:570-571)
                    sim.states[f + 1][k].zero(sim.stream)
        sim.sync()
        if fwd_only:
            return
        sim.states[K]["x_grad"].upload(gx)
        for f in range(K - 1, -1, -1):   # torch_wrapper.py:128-134
            sim.substep_grad(f)
            if as_shipped:
                sim.states[f + 1]["body_pos_grad"].download()
                sim.states[f + 1]["body_rot_grad"].download()
        sim.sync()

    def timed(as_shipped, reps):
        t0 = time.perf_counter()
        for _ in range(reps):
            step(as_shipped)
        return (time.perf_counter() - t0) / reps

    for _ in range(args.warmup):
        step()
    dt = timed(True, args.steps)
    dt_kernel = timed(False, max(1, min(args.steps, 3)))
    value = n * K / dt
    return {"metric": METRIC if not fwd_only else "particle-substeps/sec fwd (forward replay)", "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": desc, "particles_per_gpu": n, "substeps_per_step": K, "grid": int(sc["grid_dim"][0]), "ran": where},
            "kernel_only": {"value": n * K / dt_kernel, "unit": UNIT, "note": "same launches without the per-substep pose round trips and gradient downloads"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(cores), "kind": "reference",
                             "sample": f"{K} substeps forward{'' if fwd_only else ' + backward'} per step, {where}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "D", "A", "B", "C", "E"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1))
    wl = args.workload if args.workload != "auto" else (os.environ.get("DD_BENCH_WORKLOAD") or ("D" if world == 1 else "E"))
    if args.steps is None:
        args.steps = {"D": 20, "E": 3, "C": 10}.get(wl, 20)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    out = run_ours(args) if args.impl == "ours" else run_reference(args)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
