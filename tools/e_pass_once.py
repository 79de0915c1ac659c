"""One warm engine pass of workload E (64 environments, 10 env steps x 40 substeps forward + backward through GradModel) between
cudaProfilerStart/Stop, for an ncu launch list; also prints its wall time.  python tools/e_pass_once.py [envs]"""
import sys, time, numpy as np, torch
sys.path[:0] = ['/root/repo', '/root/repo/tests']
import bench
from dexdeform_b200.torch_wrapper import GradModel
sub = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T, S_env = 10, 40
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
sim, sc, base0, q0 = bench.make_hand_batch(sub, T, S_env, stream)
model = GradModel(sim, return_grid=())
rng = np.random.default_rng(7)
act = np.float32(rng.uniform(-0.4, 0.4, (T, 1, 26))); act[:, :, 21] = -0.5
noise = torch.tensor(np.float32(rng.normal(size=(sub, T, 1, 26)) * 0.2), device="cuda")
state = [torch.tensor(sc[k], device="cuda")[None].expand(sub, -1, -1).contiguous() for k in ("x", "v", "F", "C")]
def one_pass():
    action = torch.tensor(act, device="cuda", requires_grad=True)
    sim.engine.set_state(0, *state)
    sim.base_pose[0], sim.joint_rot[0] = base0, q0
    model.zero_grad()
    a = (action[None] + noise).clamp(-1, 1)
    obs = model.get_obs(0, "cuda")
    for j in range(T):
        obs = model.forward(j, a[:, j], *obs)
    loss = -obs[0][..., 1].mean(1).sum()
    loss.backward()
    return loss
one_pass(); one_pass(); torch.cuda.synchronize()
t0 = time.perf_counter(); one_pass(); torch.cuda.synchronize(); print("pass ms", (time.perf_counter() - t0) * 1e3)
torch.cuda.cudart().cudaProfilerStart()
l = one_pass(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(l) / sub)
