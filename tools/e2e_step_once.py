"""One warm end-to-end step of bench.py's GradModel path between cudaProfilerStart/Stop (for an ncu launch list):
ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file out.csv python tools/e2e_step_once.py [workload]"""
import sys, numpy as np, torch
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
def step():
    sim.engine.set_state(0, hx, hv, hF, hC, non_blocking=True)
    model.zero_grad()
    action = hact.to("cuda", non_blocking=True).requires_grad_(True)
    obs = model.get_obs(0, "cuda")
    obs = model.forward(0, action[0], *obs)
    loss = -obs[0][:, 1].mean()
    loss.backward()
    hgrad.copy_(action.grad, non_blocking=True); hloss.copy_(loss.detach().reshape(1), non_blocking=True); torch.cuda.current_stream().synchronize()
step(); step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(hloss[0]))
