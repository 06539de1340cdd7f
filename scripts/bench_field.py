"""Quick device-timed look at the field forward (fp32 SIMT vs tcgen05) at the config-3 per-GPU shape."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdface_gan_b200 as sg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda"
torch.manual_seed(0)
mo, ro = sg.default_options("ngp", renderer_res=64, n_samples=24, perturb=0.)
g = sg.Generator(mo, ro, full_pipeline=False, ema=True).to(dev).eval()
cam, focal, near, far, _ = sg.generate_camera_params(64, dev, batch=B)
z = torch.randn(B, 256, device=dev)
N = B * 64 * 64 * 24
for prec in ([os.environ["SDFG_ONLY"]] if os.environ.get("SDFG_ONLY") else ("tc16", "fp32")):
    g.renderer.network.precision = prec
    with torch.no_grad():
        for _ in range(3):
            g([z], cam, focal, near, far)
        torch.cuda.synchronize()
        sg._lib.prof_enable(True, "gemm")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g([z], cam, focal, near, far)
        e1.record()
        torch.cuda.synchronize()
        sg._lib.prof_enable(False, "")
        kms, kn = sg._lib.prof_collect()
    ms = e0.elapsed_time(e1) / 5
    flop = N * 2 * (32 * 256 + 3 * 256 * 256 + 272 * 256)
    print(json.dumps({"precision": prec, "B": B, "ms_forward": ms, "img_s": B / ms * 1e3, "msamples_s": N / ms / 1e3,
                      "gemm_ms": kms / 5, "gemm_launches": kn / 5, "gemm_tflops": flop / (kms / 5 * 1e-3) / 1e12}))
