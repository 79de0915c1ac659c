"""Host-side wall time of every part of bench.py's end-to-end step (workload D by default), each bracketed by synchronisations,
next to the un-bracketed step time.  python tools/e2e_gradmodel_parts.py [workload]"""
import sys, time, numpy as np, torch
sys.path[:0] = ['/root/repo', '/root/repo/tests']
import bench
from dexdeform_b200.simulator import MPMSimulator
from dexdeform_b200.torch_wrapper import GradModel
wl = sys.argv[1] if len(sys.argv) > 1 else 'D'
sc, S, desc = bench.workload_scene(wl)
n, nb = sc['n'], sc['nb']
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
act0, scales = bench.free_tool_actions(sc, S)
sim = MPMSimulator(nb, ground_friction=sc["ground_friction"], gravity=tuple(sc["gravity"].reshape(3) / 30), n_particles=n, dx=sc["dx"],
                   dt=sc["dt"], max_steps=S, substeps=S, stream=stream.cuda_stream)
sim.init_particles(sc["vol"], sc["mass"], sc["mu_lam_yield"])
sim.init_bodies(sc["tfsr"][:, 0], sc["tfsr"][:, 2], sc["tfsr"][:, 1], sc["tfsr"][:, 3], sc["args"], action_scales=scales, pos=sc["pos"][0], rot=sc["rot"][0])
model = GradModel(sim, return_grid=())
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
hx, hv, hF, hC = (pin(sc[k][None]) for k in ("x", "v", "F", "C"))
hact = pin(act0[None])
hgrad, hloss = torch.empty_like(hact).pin_memory(), torch.empty(1).pin_memory()
acc = {}
def T(name, f, sync=True):
    if sync: torch.cuda.synchronize()
    t0 = time.perf_counter(); r = f()
    if sync: torch.cuda.synchronize()
    acc[name] = acc.get(name, 0) + time.perf_counter() - t0
    return r
def step(sync):
    T('set_state', lambda: sim.engine.set_state(0, hx, hv, hF, hC, non_blocking=True), sync)
    T('zero_grad', lambda: model.zero_grad(), sync)
    action = T('action_h2d', lambda: hact.to("cuda", non_blocking=True).requires_grad_(True), sync)
    obs = T('get_obs', lambda: model.get_obs(0, "cuda"), sync)
    obs = T('forward', lambda: model.forward(0, action[0], *obs), sync)
    loss = T('loss', lambda: -obs[0][:, 1].mean(), sync)
    T('backward', lambda: loss.backward(), sync)
    def out():
        hgrad.copy_(action.grad, non_blocking=True); hloss.copy_(loss.detach().reshape(1), non_blocking=True); torch.cuda.current_stream().synchronize()
    T('d2h', out, sync)
for mode in (True, False):
    step(mode); acc = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): step(mode)
    torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / 3
    print('bracketed by syncs' if mode else 'free running (host time of each call)', {k: round(v / 3 * 1e3, 2) for k, v in acc.items()}, 'step ms', round(tot * 1e3, 2))
# engine only
eng = sim.engine
gx = torch.zeros((1, n, 3), device='cuda'); gx[..., 1] = -1.0 / n
def estep():
    eng.forward(0, S); eng.zero_grad(S); eng.add_state_grad(S, gx=gx); eng.backward(0, S)
estep(); torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): estep()
torch.cuda.synchronize(); print('engine only step ms', round((time.perf_counter() - t0) / 3 * 1e3, 2))
