"""Launch every non-GEMM kernel of the path (and the reference's kernels from oracle/_ref) a few times at the BASELINE size on real
ray samples -- the program `ncu --set full` is wrapped around by scripts/gpu_kernels_ncu.sh.  No timing here."""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import sdface_gan_b200 as sg
from sdface_gan_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev, R, S = "cuda", 64, 24
N = B * R * R * S
torch.manual_seed(0)
mo, ro = sg.default_options("ngp", renderer_res=R, n_samples=S, perturb=0.)
g = sg.Generator(mo, ro, full_pipeline=False, ema=True).to(dev).eval()
enc = g.renderer.network.encoder
tab = enc.embeddings.detach()
cam, focal, near, far, _ = sg.generate_camera_params(R, dev, batch=B)
with torch.no_grad():
    smp, _, _ = g.renderer._sample(cam, focal, near, far, t_rand=None)
pts = smp["npts"].reshape(-1, 3).contiguous()
Sg, H = ops.log2_scale(enc.per_level_scale), enc.base_resolution
feats = torch.empty(N, 32, device=dev)
dy = torch.empty(96, N, device=dev)
grad = torch.randn(N, 32, device=dev)
gt = torch.zeros_like(tab)
sdf = torch.randn(N, device=dev) * 0.05
rgb = torch.randn(N, 3, device=dev)
f16 = torch.randn(N, 256, device=dev).half()
f32 = f16.float()
zv = smp["z_vals"].reshape(-1).contiguous()
rd = smp["rays_d"].reshape(-1, 3).contiguous()
ptsw = smp["pts"].reshape(-1, 3).contiguous()
sb = g.renderer.sigmoid_beta.detach()
NR = N // S
d_rgb_map, d_feat_map = torch.randn(NR, 3, device=dev), torch.randn(NR, 256, device=dev)
vd_ray = smp["viewdirs"].reshape(-1, 3).contiguous()


def ref(name):
    path = os.path.join(ROOT, "oracle", "_ref", name + ".so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


ref_g, ref_s = ref("_gridencoder_ref"), ref("_shencoder_ref")
u = ((pts + 2.0) / 4.0).contiguous()
o_ref = torch.empty(16, N, 2, device=dev)
g_ref = grad.view(N, 16, 2).permute(1, 0, 2).contiguous()
S_ref = float(np.log2(enc.per_level_scale))
for _ in range(reps):
    g.renderer._sample(cam, focal, near, far, t_rand=None)                                                         # sample_rays_kernel
    ops.grid_encode_forward(pts, tab, enc.offsets, Sg, H, bound=2.0, outputs=feats)                                  # grid_forward_kernel
    ops.grid_encode_forward(pts, tab, enc.offsets, Sg, H, bound=2.0, calc_dy_dx=True, outputs=feats, dy_dx=dy)       # ... with dy_dx
    ops.grid_encode_backward(grad, pts, tab, enc.offsets, Sg, H, bound=2.0, grad_embeddings=gt)                      # grid_backward_kernel
    ops.grid_encode_backward(grad, pts, tab, enc.offsets, Sg, H, bound=2.0, dy_dx=dy, grad_embeddings=None, want_grad_inputs=True)   # grid_input_backward_kernel
    ops.sh_encode_forward(vd_ray, 4)                                                                                 # sh_forward_kernel
    ops.composite_forward(sdf, rgb, f16, zv, rd, ptsw, None, sb, S, True, False, False)                              # composite_forward (fp16 features)
    ops.composite_forward(sdf, rgb, f32, zv, rd, ptsw, None, sb, S, True, False, False)                              # composite_forward (fp32 features)
    ops.composite_forward(sdf, rgb, None, zv, rd, ptsw, None, sb, S, True, False, False)                             # stage-1 variant
    ops.composite_backward(sdf, rgb, f32, zv, rd, None, None, sb, S, True, False, d_rgb_map, d_feat_map, None, None, True)
    ops.composite_backward(sdf, rgb, None, zv, rd, None, None, sb, S, True, False, d_rgb_map, None, None, None, False)
    if ref_g is not None:
        ref_g.grid_encode_forward(u, tab, enc.offsets, o_ref, N, 3, 2, 16, S_ref, H, None, 0, False, 0)              # kernel_grid (reference)
        ref_g.grid_encode_backward(g_ref, u, tab, enc.offsets, gt, N, 3, 2, 16, S_ref, H, None, None, 0, False, 0)   # kernel_grid_backward (reference)
    if ref_s is not None:
        ref_s.sh_encode_forward(vd_ray, torch.empty(NR, 16, device=dev), NR, 3, 4, None)                             # kernel_sh (reference), per ray here
torch.cuda.synchronize()
print("ok")
