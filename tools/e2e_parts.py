import sys, time, numpy as np, torch
sys.path[:0]=['/root/repo','/root/repo/tests']
import bench
from dexdeform_b200.engine import FusedSim
sc,S,desc=bench.workload_scene('D')
n,nb=sc['n'],sc['nb']
stream=torch.cuda.Stream(); torch.cuda.set_stream(stream)
sim=FusedSim.from_scene(sc,n_envs=1,max_steps=S,stream=stream.cuda_stream)
pin=lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
hx,hv,hF,hC=(pin(sc[k][None]) for k in ('x','v','F','C'))
hpos,hrot=pin(sc['pos'][:,None]),pin(sc['rot'][:,None])
gx=np.zeros((1,n,3),np.float32); gx[...,1]=-1.0/n; hgx=pin(gx)
x_out=torch.empty((1,n,3),dtype=torch.float32).pin_memory()
gp=torch.empty((S+1,1,nb,3),dtype=torch.float32).pin_memory(); gr=torch.empty((S+1,1,nb,4),dtype=torch.float32).pin_memory()
P=lambda t:t.data_ptr()
x_dev=torch.empty((1,n,3),dtype=torch.float32,device='cuda')
def T(f,name,acc):
    torch.cuda.synchronize(); t0=time.perf_counter(); r=f(); torch.cuda.synchronize(); acc[name]=acc.get(name,0)+time.perf_counter()-t0; return r
acc={}
for it in range(4):
    if it==1: acc={}
    T(lambda: sim._check(sim.lib.dd_sim_set_state(sim._h,0,P(hx),P(hv),P(hF),P(hC),sim.stream)),'set_state',acc)
    T(lambda: sim._check(sim.lib.dd_sim_set_poses(sim._h,0,S+1,P(hpos),P(hrot),sim.stream)),'set_poses',acc)
    T(lambda: sim.forward(0,S),'forward',acc)
    T(lambda: sim._check(sim.lib.dd_sim_get_state(sim._h,S,x_dev.data_ptr(),None,None,None,sim.stream)),'get_state_to_device_first',acc)
    T(lambda: sim._check(sim.lib.dd_sim_get_state(sim._h,S,P(x_out),None,None,None,sim.stream)),'get_state',acc)
    time.sleep(0.01)
    T(lambda: sim._check(sim.lib.dd_sim_get_state(sim._h,S,P(x_out),None,None,None,sim.stream)),'get_state_again',acc)
    T(lambda: sim._check(sim.lib.dd_sim_get_state(sim._h,S-1,P(x_out),None,None,None,sim.stream)),'get_state_prev_slot',acc)
    T(lambda: -float(x_out[0,:,1].mean()),'host_loss',acc)
    T(lambda: sim.zero_grad(S),'zero_grad',acc)
    T(lambda: sim._check(sim.lib.dd_sim_add_state_grad(sim._h,S,P(hgx),None,None,None,sim.stream)),'add_grad',acc)
    T(lambda: sim.backward(0,S),'backward',acc)
    T(lambda: sim._check(sim.lib.dd_sim_get_pose_grads(sim._h,0,S+1,P(gp),P(gr),sim.stream)),'get_pose_grads',acc)
print({k:round(v/3*1e3,2) for k,v in acc.items()}, 'total ms', round(sum(acc.values())/3*1e3,2))
t=torch.empty(108*1024*1024//4,dtype=torch.float32).pin_memory(); d=torch.empty_like(t,device='cuda')
torch.cuda.synchronize(); t0=time.perf_counter(); d.copy_(t,non_blocking=True); torch.cuda.synchronize(); print('H2D GB/s', 108*1.048576e-3/(time.perf_counter()-t0))
for name,src,dst in (('H2D',t,d),('D2H',d,t)):
    for _ in range(3):
        torch.cuda.synchronize(); t0=time.perf_counter(); dst.copy_(src,non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print(name,'GB/s', round(108*1.048576e-3/dt,1))
small=torch.empty(3*1000000,dtype=torch.float32).pin_memory(); ds=torch.empty_like(small,device='cuda')
for _ in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter(); small.copy_(ds,non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t0
print('D2H 12MB ms', round(dt*1e3,3))
