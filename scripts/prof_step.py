"""torch.profiler view of the benchmark step: per-kernel device time and idle gaps of the stream (debugging aid)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import sdface_gan_b200 as sg
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mo, ro = sg.default_options("ngp", renderer_res=64, n_samples=24, perturb=1.0, no_features_output=True, return_sdf=True)
g = sg.Generator(mo, ro, full_pipeline=False).to(dev)
opt = torch.optim.Adam(g.parameters(), lr=2e-5, betas=(0.0, 0.9))
cam, focal, near, far, _ = sg.generate_camera_params(64, dev, batch=B)
z = torch.randn(B, 256, device=dev)

def step():
    opt.zero_grad(set_to_none=True)
    _, thumb, sdf, eik = g([z], cam, focal, near, far, return_sdf=True, return_eikonal=True)
    loss = bench.g_losses(thumb, sdf, eik)
    loss.backward()
    opt.step()

for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print("span_us", t1 - t0, "busy_us", busy, "n", len(evs))
gaps = []
last = evs[0].time_range.end
for a, b in zip(evs[:-1], evs[1:]):
    gap = b.time_range.start - max(last, a.time_range.end)
    last = max(last, a.time_range.end)
    if gap > 30:
        gaps.append((gap, a.name[:50], b.name[:50]))
gaps.sort(reverse=True)
print("total gap >30us:", sum(x[0] for x in gaps))
for x in gaps[:25]:
    print(x)
agg = {}
for e in evs:
    a = agg.setdefault(e.name[:70], [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%-72s %4d %10.1f" % (k, n, t / 3))
if os.environ.get("PROF_TIMELINE"):
    # timeline of the last step's long kernels: start (us, relative), duration, stream -- shows whether two streams really overlap
    last_fwd = [e for e in evs if "chain_fwd" in e.name][-1].time_range.start
    for e in evs:
        d = e.time_range.end - e.time_range.start
        if e.time_range.start >= last_fwd and d > 60:
            print("TL %9.1f %8.1f  stream %s  %s" % (e.time_range.start - last_fwd, d, getattr(e, "stream", "?"), e.name[:60]))
