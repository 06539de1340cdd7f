"""CUDA-event timing of the decoder's kernels per level (B = 64, size 256)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdface_gan_b200 as sg
from sdface_gan_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda"
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (H, C) in ((64, 512), (128, 256), (256, 128)):
    x = torch.randn(B, H, H, C, device=dev).half()
    w = torch.randn(1, 3, C, 1, 1, device=dev); st = torch.randn(B, C, device=dev); bias = torch.zeros(1, 3, 1, 1, device=dev)
    skip = torch.randn(B, H // 2, H // 2, 3, device=dev)
    t0 = timed(lambda: ops.to_rgb(x, w, st, 0.1, bias, None, want_nhwc=True))
    t1 = timed(lambda: ops.to_rgb(x, w, st, 0.1, bias, skip, want_nhwc=True))
    t2 = timed(lambda: ops.to_rgb(x, w, st, 0.1, bias, skip, want_nhwc=False, want_nchw=True))
    gb = x.numel() * 2 / 1e9
    print("to_rgb H=%d C=%d: no skip %.3f ms, skip %.3f ms, skip+nchw %.3f ms (%.2f GB in -> %.0f GB/s)" % (H, C, t0, t1, t2, gb, gb / t0 * 1e3))
for (H, Cin, Cout) in ((64, 256, 512), (128, 256, 256), (256, 128, 128)):
    x = torch.randn(B, H, H, Cin, device=dev).half()
    wf = torch.randn(B, 9, Cout, Cin, device=dev).half() * 0.02
    t = timed(lambda: ops.conv_forward(x, wf))
    fl = 2 * 9 * B * H * H * Cin * Cout
    print("conv3x3 H=%d %d->%d: %.3f ms = %.0f TFLOP/s" % (H, Cin, Cout, t, fl / t / 1e9))
for (H, Cin, Cout) in ((64, 512, 256), (128, 256, 128)):
    x = torch.randn(B, H, H, Cin, device=dev).half()
    wf = torch.randn(B, 9, Cout, Cin, device=dev).half() * 0.02
    t = timed(lambda: ops.upconv_forward(x, wf))
    fl = 2 * 9 * B * H * H * Cin * Cout
    print("upconv H=%d %d->%d: %.3f ms (transposed convolution + blur; %.2f TFLOP, T %.2f GB, out %.2f GB)" % (H, Cin, Cout, t, fl / 1e12, B * (2 * H + 1) ** 2 * Cout * 2 / 1e9, B * 4 * H * H * Cout * 2 / 1e9))
