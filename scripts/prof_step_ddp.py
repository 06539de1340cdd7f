"""torch.profiler timeline of one data-parallel benchmark step (rank 0): where the gradient exchange sits (debugging aid).
   torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/prof_step_ddp.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
import sdface_gan_b200 as sg
from torch.profiler import profile, ProfilerActivity

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
B = 32
torch.manual_seed(1234 + rank)
mo, ro = sg.default_options("ngp", renderer_res=64, n_samples=24, perturb=1.0, no_features_output=True, return_sdf=True)
g = sg.Generator(mo, ro, full_pipeline=False).to(dev)
for p_ in g.parameters():
    dist.broadcast(p_.data, 0)
model = sg.distributed.data_parallel(g, device_ids=[local])
opt = torch.optim.Adam(g.parameters(), lr=2e-5, betas=(0.0, 0.9), fused=True)
cam, focal, near, far, _ = sg.generate_camera_params(64, dev, batch=B)
z = torch.randn(B, 256, device=dev)

def step():
    opt.zero_grad(set_to_none=True)
    _, thumb, sdf, eik = model([z], cam, focal, near, far, return_sdf=True, return_eikonal=True)
    loss = bench.g_losses(thumb, sdf, eik)
    loss.backward()
    opt.step()

for _ in range(5):
    step()
torch.cuda.synchronize()
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    last_fwd = [e for e in evs if "tc_fchain_fwd" in e.name][-1].time_range.start
    for e in evs:
        d = e.time_range.end - e.time_range.start
        if e.time_range.start >= last_fwd and (d > 80 or "nccl" in e.name.lower()):
            print("TL %9.1f %8.1f  %s" % (e.time_range.start - last_fwd, d, e.name[:70]))
dist.destroy_process_group()
