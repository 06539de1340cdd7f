"""The bench workload without the measuring: W warm-up + K training steps (configs[1], B = 32), K inference passes with features, or (infer256,
PROF_B=64) K passes of the full generator.
The program ncu is wrapped around (scripts/gpu_profile3.sh); prints the number of kernel launches per step for -s / -c."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdface_gan_b200 as sg
from bench import g_losses, R, S, STYLE

B = int(os.environ.get("PROF_B", "32"))
mode = sys.argv[1] if len(sys.argv) > 1 else "train"
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
if mode == "train":
    mo, ro = sg.default_options("ngp", renderer_res=R, n_samples=S, perturb=1.0, no_features_output=True, return_sdf=True)
    g = sg.Generator(mo, ro, full_pipeline=False).to(dev)
    g.renderer.network.precision = "tc16"
    opt = torch.optim.Adam(g.parameters(), lr=2e-5, betas=(0.0, 0.9), fused=True)
    cam, focal, near, far, _ = sg.generate_camera_params(R, dev, batch=B)
    z = torch.randn(B, STYLE, device=dev)
    for i in range(5):
        opt.zero_grad(set_to_none=True)
        _, thumb, sdf, eik = g([z], cam, focal, near, far, return_sdf=True, return_eikonal=True)
        g_losses(thumb, sdf, eik).backward()
        opt.step()
        torch.cuda.synchronize()
elif mode == "infer256":
    # configs[2]: full generator (renderer + decoder), 256^2 images
    mo, ro = sg.default_options("ngp", size=256, renderer_res=R, n_samples=S, perturb=0.)
    g = sg.Generator(mo, ro, full_pipeline=True, ema=True).to(dev).eval()
    cam, focal, near, far, _ = sg.generate_camera_params(R, dev, batch=B)
    z = torch.randn(B, STYLE, device=dev)
    with torch.no_grad():
        for i in range(3):
            g([z], cam, focal, near, far)
            torch.cuda.synchronize()
else:
    mo, ro = sg.default_options("ngp", renderer_res=R, n_samples=S, perturb=0.)
    g = sg.Generator(mo, ro, full_pipeline=False, ema=True).to(dev).eval()
    g.renderer.network.precision = "tc16"
    cam, focal, near, far, _ = sg.generate_camera_params(R, dev, batch=B)
    z = torch.randn(B, STYLE, device=dev)
    with torch.no_grad():
        for i in range(5):
            g([z], cam, focal, near, far)
            torch.cuda.synchronize()
print("ok")
