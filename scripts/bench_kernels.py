"""Per-kernel roofline measurements of the non-GEMM kernels of the path (encoder, compositing) plus the inference forward,
at the BASELINE config sizes.  Device-timed with CUDA events after warm-up; prints one JSON object.

    python scripts/bench_kernels.py [B]      # B = images per pass (default 32; configs[2] uses 64 / n_gpus)
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdface_gan_b200 as sg
from sdface_gan_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda"
R, S = 64, 24
N = B * R * R * S
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm_peak = next((v for k, v in peaks.items() if "hbm" in k.lower() and isinstance(v, (int, float))), 6547.0)


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


torch.manual_seed(0)
mo, ro = sg.default_options("ngp", renderer_res=R, n_samples=S, perturb=0.)
g = sg.Generator(mo, ro, full_pipeline=False, ema=True).to(dev).eval()
net = g.renderer.network
enc = net.encoder
cam, focal, near, far, _ = sg.generate_camera_params(R, dev, batch=B)
z = torch.randn(B, 256, device=dev)
out = {"B": B, "samples": N, "hbm_peak_GBps": hbm_peak}

# --- L2 random-gather roofline: table-sized buffer (50.6 MB), 8 B per gather
tab = enc.embeddings.detach()
threads, rounds = 148 * 2048 * 4, 64
ms = timed(lambda: ops.l2_gather_probe(tab, threads, rounds))
gathers = threads * 8 * rounds
out["l2_gather_probe"] = {"table_MB": tab.numel() * 4 / 1e6, "Ggathers_per_s": gathers / ms / 1e6, "GBps_8B": gathers * 8 / ms / 1e6,
                          "note": "rate of INDEPENDENT random 8-byte gathers over a table-sized buffer -- a yardstick, not a ceiling: the encoder's "
                                  "coarse levels beat it because neighbouring samples share sectors (x_random_gather_probe > 1).  The bound "
                                  "fractions are the hardware's own counters in ncu_bound."}
# the kernels' bound as the hardware counts it (ncu --set full, profiles/r02k_summary.txt): busiest unit, percent of its peak
out["ncu_bound"] = {}
try:
    import re
    blocks = open(os.path.join(ROOT, "profiles", "r02k_summary.txt")).read().split("---\n")
    for blk in blocks:
        m = re.search(r"Kernel Name\s+(?:void )?(?:sdfg::)?(\w+)", blk)
        if not m or not m.group(1).startswith(("grid_", "composite_")):
            continue
        g1 = lambda key: float(re.search(key + r"\s+([\d.]+)", blk).group(1)) if re.search(key + r"\s+([\d.]+)", blk) else None
        out["ncu_bound"].setdefault(m.group(1), {"l1tex_pct": g1("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                                                 "lts_pct": g1("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                                                 "dram_pct": g1("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")})
except Exception as e:                                  # the summary is a committed file; absent only in a stripped checkout
    out["ncu_bound"] = {"unavailable": str(e)}

# --- encoder on the real ray samples (the points are NOT uniform: SURVEY 8d)
with torch.no_grad():
    smp, _, _ = g.renderer._sample(cam, focal, near, far, t_rand=None)
pts = smp["npts"].reshape(-1, 3).contiguous()
Sg, H = ops.log2_scale(enc.per_level_scale), enc.base_resolution
feats = torch.empty(N, 32, device=dev)
dy = torch.empty(96, N, device=dev)
ms_f = timed(lambda: ops.grid_encode_forward(pts, tab, enc.offsets, Sg, H, bound=2.0, outputs=feats))
ms_fd = timed(lambda: ops.grid_encode_forward(pts, tab, enc.offsets, Sg, H, bound=2.0, calc_dy_dx=True, outputs=feats, dy_dx=dy))
grad = torch.randn(N, 32, device=dev)
gt = torch.zeros_like(tab)
ms_b = timed(lambda: ops.grid_encode_backward(grad, pts, tab, enc.offsets, Sg, H, bound=2.0, grad_embeddings=gt))
probe = out["l2_gather_probe"]["Ggathers_per_s"]
out["grid_forward"] = {"ms": ms_f, "hbm_GBps": N * 140 / ms_f / 1e6, "Ggathers_per_s": N * 128 / ms_f / 1e6,
                       "x_random_gather_probe": N * 128 / ms_f / 1e6 / probe}
out["grid_forward_dydx"] = {"ms": ms_fd, "hbm_GBps": N * (140 + 384) / ms_fd / 1e6, "frac_of_hbm": N * (140 + 384) / ms_fd / 1e6 / hbm_peak,
                            "x_random_gather_probe": N * 128 / ms_fd / 1e6 / probe}
out["grid_backward"] = {"ms": ms_b, "Greductions_8B_per_s": N * 128 / ms_b / 1e6, "x_random_gather_probe": N * 128 / ms_b / 1e6 / probe}

# --- the reference's own kernels (gridencoder.cu / shencoder.cu compiled UNCHANGED for sm_100a into oracle/_ref/, same box, same
#     inputs): SURVEY 2.3 sets "beat the reference kernel recompiled for sm_100a" as the bar
import importlib.util
import numpy as np


def _load_ref(name):
    path = os.path.join(ROOT, "oracle", "_ref", name + ".so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


ref_g, ref_s = _load_ref("_gridencoder_ref"), _load_ref("_shencoder_ref")
if ref_g is not None:
    u = ((pts + 2.0) / 4.0).contiguous()                  # the reference maps to [0,1] in Python (grid.py:149): not timed
    S_ref = float(np.log2(enc.per_level_scale))
    o_ref = torch.empty(16, N, 2, device=dev)
    dd_ref = torch.empty(N, 96, device=dev)
    g_ref = grad.view(N, 16, 2).permute(1, 0, 2).contiguous()
    ge_ref = torch.zeros_like(tab)
    gi_ref = torch.zeros(N, 3, device=dev)
    r_f = timed(lambda: ref_g.grid_encode_forward(u, tab, enc.offsets, o_ref, N, 3, 2, 16, S_ref, H, None, 0, False, 0))
    r_fd = timed(lambda: ref_g.grid_encode_forward(u, tab, enc.offsets, o_ref, N, 3, 2, 16, S_ref, H, dd_ref, 0, False, 0))
    r_perm = timed(lambda: o_ref.permute(1, 0, 2).reshape(N, 32))        # grid.py:57, the transposing copy the reference adds
    r_b = timed(lambda: ref_g.grid_encode_backward(g_ref, u, tab, enc.offsets, ge_ref, N, 3, 2, 16, S_ref, H, None, None, 0, False, 0))
    r_bi = timed(lambda: ref_g.grid_encode_backward(g_ref, u, tab, enc.offsets, ge_ref, N, 3, 2, 16, S_ref, H, dd_ref, gi_ref, 0, False, 0))
    dsdf = torch.randn(N, 32, device=dev)
    ms_ib = timed(lambda: ops.grid_encode_backward(dsdf, pts, tab, enc.offsets, Sg, H, bound=2.0, dy_dx=dy, grad_embeddings=None, want_grad_inputs=True))
    out["reference_kernels_same_box"] = {
        "kernel_grid_ms": r_f, "kernel_grid_dydx_ms": r_fd, "permute_copy_ms": r_perm, "kernel_grid_backward_ms": r_b,
        "kernel_grid_backward_plus_input_ms": r_bi,
        "ours_grid_forward_ms": ms_f, "ours_grid_forward_dydx_ms": ms_fd, "ours_grid_backward_ms": ms_b, "ours_grid_input_backward_ms": ms_ib,
        "speedup_forward": (r_f + r_perm) / ms_f, "speedup_forward_dydx": (r_fd + r_perm) / ms_fd, "speedup_backward": r_b / ms_b,
        "speedup_input_backward": (r_bi - r_b) / ms_ib}
    del o_ref, dd_ref, g_ref, ge_ref, gi_ref, u
if ref_s is not None:
    vd_ray = smp["viewdirs"].reshape(-1, 3).contiguous()
    vd_smp = vd_ray[:, None, :].expand(-1, S, 3).reshape(-1, 3).contiguous()      # the reference encodes every SAMPLE (sdf_model.py:304)
    so_ref = torch.empty(N, 16, device=dev)
    r_sh = timed(lambda: ref_s.sh_encode_forward(vd_smp, so_ref, N, 3, 4, None))
    ms_sh = timed(lambda: ops.sh_encode_forward(vd_ray, 4))
    out.setdefault("reference_kernels_same_box", {}).update({"kernel_sh_per_sample_ms": r_sh, "ours_sh_per_ray_ms": ms_sh, "speedup_sh": r_sh / ms_sh})
    del so_ref, vd_smp
torch.cuda.empty_cache()

# --- compositing: full-feature inference variant (fp16 features from the field chain) and the stage-1 variant (no features)
sdf = torch.randn(N, device=dev) * 0.05
rgb = torch.randn(N, 3, device=dev)
f16 = torch.randn(N, 256, device=dev).half()
zv = smp["z_vals"].reshape(-1).contiguous()
rd = smp["rays_d"].reshape(-1, 3).contiguous()
ptsw = smp["pts"].reshape(-1, 3).contiguous()
sb = g.renderer.sigmoid_beta.detach()
ms_c = timed(lambda: ops.composite_forward(sdf, rgb, f16, zv, rd, ptsw, None, sb, S, True, False, False))
bytes_c = N * (4 + 12 + 512 + 4) + (N // S) * (12 + 1024 + 12)
out["composite_forward_feat16"] = {"ms": ms_c, "hbm_GBps": bytes_c / ms_c / 1e6, "frac_of_hbm": bytes_c / ms_c / 1e6 / hbm_peak}
f32 = f16.float()
ms_c32 = timed(lambda: ops.composite_forward(sdf, rgb, f32, zv, rd, ptsw, None, sb, S, True, False, False))
bytes_c32 = N * (4 + 12 + 1024 + 4) + (N // S) * (12 + 1024 + 12)
out["composite_forward_feat32"] = {"ms": ms_c32, "hbm_GBps": bytes_c32 / ms_c32 / 1e6, "frac_of_hbm": bytes_c32 / ms_c32 / 1e6 / hbm_peak}
ms_c0 = timed(lambda: ops.composite_forward(sdf, rgb, None, zv, rd, ptsw, None, sb, S, True, False, False))
bytes_c0 = N * (4 + 12 + 4) + (N // S) * 24
out["composite_forward_nofeat"] = {"ms": ms_c0, "hbm_GBps": bytes_c0 / ms_c0 / 1e6, "frac_of_hbm": bytes_c0 / ms_c0 / 1e6 / hbm_peak}

# --- inference forward of the renderer (configs[2] field part): rays -> encoder -> field chain -> compositing
sg._lib.prof_enable(True, "gemm")
with torch.no_grad():
    ms_fw = timed(lambda: g([z], cam, focal, near, far), reps=5)
sg._lib.prof_enable(False, "")
kms, kn = sg._lib.prof_collect()
flop = N * 2 * (32 * 256 + 3 * 256 * 256 + 256 + 272 * 256 + 768)
tc_peak = peaks.get("bf16_tflops_sustained", 1393.1)
out["inference_forward"] = {"ms": ms_fw, "images_per_s": B / ms_fw * 1e3, "Msamples_per_s": N / ms_fw / 1e3,
                            "field_chain_ms": kms / max(kn, 1), "field_chain_TFLOPs": flop / (kms / max(kn, 1) * 1e-3) / 1e12,
                            "field_chain_frac_of_tensor_peak": flop / (kms / max(kn, 1) * 1e-3) / 1e12 / tc_peak}
# --- configs[4]: the sdf_mesh.py frustum query (sdf_mesh.py:243-253): R' x R' rays x R' samples, return_sdf + return_xyz,
#     static view directions, forced background; R' = 128 is the reference call, R' = 256 the BASELINE "256^3" size
out["sdf_mesh_query"] = {}
for Rm in (128, 256):
    mo_m, ro_m = sg.default_options("ngp", renderer_res=Rm, n_samples=Rm, perturb=0., return_sdf=True, return_xyz=True,
                                    static_viewdirs=True, force_background=True)
    gm = sg.Generator(mo_m, ro_m, full_pipeline=False, ema=True).to(dev).eval()
    cam_m, focal_m, near_m, far_m, _ = sg.generate_camera_params(Rm, dev, batch=1)
    with torch.no_grad():
        ms_m = timed(lambda: gm([z[:1]], cam_m, focal_m, near_m, far_m, return_sdf=True, return_xyz=True), reps=3, warm=2)
    out["sdf_mesh_query"]["R%d" % Rm] = {"points": Rm ** 3, "ms": ms_m, "Mpoints_per_s": Rm ** 3 / ms_m / 1e3}
    # the SDF-only fast path: trunk + sigma head only -- no view layer, rgb or feature map -- then the frustum -> box resample
    gm.renderer.sdf_only = True
    with torch.no_grad():
        ms_s = timed(lambda: gm([z[:1]], cam_m, focal_m, near_m, far_m, return_sdf=True, return_xyz=True), reps=3, warm=2)
        sdf_vol = gm([z[:1]], cam_m, focal_m, near_m, far_m, return_sdf=True, return_xyz=True)[3]
        ms_a = timed(lambda: sg.align_volume(sdf_vol), reps=3, warm=1)
    out["sdf_mesh_query"]["R%d" % Rm].update({"sdf_only_ms": ms_s, "sdf_only_Mpoints_per_s": Rm ** 3 / ms_s / 1e3, "align_volume_ms": ms_a})
    del gm, sdf_vol
    torch.cuda.empty_cache()
print(json.dumps(out))
