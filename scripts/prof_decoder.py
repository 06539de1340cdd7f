"""torch.profiler kernel table of one configs[2] pass (B = 64): renderer + decoder."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdface_gan_b200 as sg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda"
mo, ro = sg.default_options("ngp", size=256, renderer_res=64, n_samples=24, perturb=0.)
g = sg.Generator(mo, ro, full_pipeline=True, ema=True).to(dev).eval()
cam, focal, near, far, _ = sg.generate_camera_params(64, dev, batch=B)
z = torch.randn(B, 256, device=dev)
with torch.no_grad():
    for _ in range(3):
        g([z], cam, focal, near, far)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            g([z], cam, focal, near, far)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
