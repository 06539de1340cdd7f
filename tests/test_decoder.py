"""the StyleGAN2 decoder forward.

CPU: the oracle (oracle/decoder_oracle.py) against the golden produced by the reference's own Decoder (tests/golden/make_golden.py).
GPU: the sm_100a decoder (decoder.py on csrc/tc_conv.cuh + conv.cu) against that golden and, at the BASELINE size (64^2 features ->
256^2 image), against the oracle.  Tolerance: fp16 operands / activations with fp32 accumulation through 5 convolutions --
max-abs 2e-2 of the image's value range (the image is O(1): mean |value| 0.86 in the golden)."""
import numpy as np
import pytest
import torch

import helpers as H
import param_fill as pf
from oracle import decoder_oracle as do


def _golden():
    z = H.load_fixture("decoder")
    tab = pf.table_from_npz(z)
    params = {name: torch.from_numpy(pf.values_for(name, shape, std, 0, mean)) for name, shape, std, mean in tab}
    noise = [torch.from_numpy(z[f"noise_{i}"]) for i in range(5)]
    return z, tab, params, noise


def test_decoder_oracle_matches_reference_golden():
    z, tab, params, noise = _golden()
    with torch.no_grad():
        img = do.decoder_forward(params, torch.from_numpy(z["features"]), torch.from_numpy(z["z"]), noise)
        assert H.max_abs(do.mapping(params, torch.from_numpy(z["z"])), z["latent"][:, 0]) < 1e-5
        assert H.max_abs(img, z["image"]) < 2e-4 * float(np.abs(z["image"]).max())
        bufs = [torch.from_numpy(z[f"buf_noise_{i}"]) for i in range(5)]
        img_b = do.decoder_forward(params, torch.from_numpy(z["features"]), torch.from_numpy(z["z"]), bufs)
        assert H.max_abs(img_b, z["image_buffers"]) < 2e-4 * float(np.abs(z["image_buffers"]).max())


def _product_decoder(size, res, tab=None, seed=0, channel_multiplier=2):
    import sdface_gan_b200 as sg
    mo, _ = sg.default_options("ngp", size=size, renderer_res=res)
    mo.feature_encoder_in_channels = 256
    mo.channel_multiplier = channel_multiplier
    dec = sg.Decoder(mo)
    if tab is not None:
        pf.fill_state(dec, tab, seed)
    return dec.cuda().eval()


@pytest.mark.gpu
def test_decoder_matches_reference_golden():
    z, tab, params, noise = _golden()
    dec = _product_decoder(32, 8, tab)
    assert set(k for k in dec.state_dict()) >= set(params)                      # the reference's parameter names, all of them
    with torch.no_grad():
        img, latent = dec(torch.from_numpy(z["features"]).cuda(), [torch.from_numpy(z["z"]).cuda()], noise=[n.cuda() for n in noise],
                          return_latents=True)
        assert H.max_abs(latent, z["latent"]) < 1e-4
        scale = float(np.abs(z["image"]).max())
        assert img.shape == (2, 3, 32, 32) and H.max_abs(img, z["image"]) < 2e-2 * scale
        assert H.rel_err(img, z["image"]) < 1e-2
        # registered noise buffers, broadcast over the batch (randomize_noise = False)
        for i in range(5):
            getattr(dec.noises, f"noise_{i}").copy_(torch.from_numpy(z[f"buf_noise_{i}"]))
        img_b, none = dec(torch.from_numpy(z["features"]).cuda(), [torch.from_numpy(z["z"]).cuda()], randomize_noise=False)
        assert none is None and H.rel_err(img_b, z["image_buffers"]) < 1e-2
    with pytest.raises(NotImplementedError):                                     # forward-only kernels: loud, not silent
        dec(torch.from_numpy(z["features"]).cuda().requires_grad_(True), [torch.from_numpy(z["z"]).cuda()])


@pytest.mark.gpu
@pytest.mark.parametrize("channel_multiplier", [2, 1])
def test_decoder_full_size_matches_oracle(channel_multiplier):
    """BASELINE configs[2] decoder shape: [B, 256, 64, 64] features -> [B, 3, 256, 256], seeded weights with live noise / biases.
    channel_multiplier 1 halves the widths (256 / 128 / 64 channels: the N = 64 tile of the conv kernel, 64-channel ToRGB and blur)."""
    import sdface_gan_b200 as sg
    torch.manual_seed(1)
    dec = _product_decoder(256, 64, channel_multiplier=channel_multiplier)
    with torch.no_grad():
        for n, p in dec.named_parameters():
            if n.endswith("noise.weight"):
                p.fill_(0.2)
            elif p.abs().max() == 0:
                p.normal_(0, 0.1)
    Bn = 2
    feats = torch.randn(Bn, 256, 64, 64)
    zl = torch.randn(Bn, 256)
    noise = [torch.randn(Bn, 1, r, r) for r in (64, 128, 128, 256, 256)]
    params = {k: v.detach().cpu() for k, v in dec.state_dict().items()}
    with torch.no_grad():
        ref = do.decoder_forward(params, feats, zl, noise)
        img, _ = dec(feats.cuda(), [zl.cuda()], noise=[n.cuda() for n in noise])
    assert img.shape == (Bn, 3, 256, 256)
    assert H.rel_err(img, ref) < 1e-2 and H.max_abs(img, ref) < 2e-2 * float(ref.abs().max())


@pytest.mark.gpu
def test_generator_full_pipeline_runs_and_matches_parts():
    """Generator(full_pipeline=True): renderer -> (channels-last view) -> decoder; the image equals the decoder applied to the renderer's
    features, and the thumbnail equals the renderer-only generator's."""
    import sdface_gan_b200 as sg
    torch.manual_seed(2)
    mo, ro = sg.default_options("ngp", size=32, renderer_res=8, n_samples=16, perturb=0.)
    g = sg.Generator(mo, ro, full_pipeline=True, ema=True).cuda().eval()
    g.renderer.network.encoder.embeddings.data.uniform_(-0.5, 0.5)
    cam, focal, near, far, _ = sg.generate_camera_params(8, "cuda", batch=2)
    zl = torch.randn(2, 256, device="cuda")
    with torch.no_grad():
        img, thumb = g([zl], cam, focal, near, far, randomize_noise=False)
        style = g.style(zl)
        t2, feats, _, _, _, _ = g.renderer(cam, focal, near, far, styles=style)
        img2, _ = g.decoder(feats, [style], randomize_noise=False)
    assert img.shape == (2, 3, 32, 32) and thumb.shape == (2, 3, 8, 8)
    assert feats.shape == (2, 256, 8, 8) and feats.permute(0, 2, 3, 1).is_contiguous()       # logically NCHW, channels-last memory
    assert torch.equal(thumb, t2) and torch.equal(img, img2) and torch.isfinite(img).all()


@pytest.mark.gpu
def test_graphed_generator_replays_the_eager_forward():
    """GraphedGenerator: the CUDA-graph replay of the full pipeline returns exactly what the eager call returns (fixed noise), follows its
    inputs (a second latent batch gives the second eager result), and draws fresh noise per replay when asked to."""
    import sdface_gan_b200 as sg
    torch.manual_seed(3)
    mo, ro = sg.default_options("ngp", size=32, renderer_res=8, n_samples=16, perturb=0.)
    g = sg.Generator(mo, ro, full_pipeline=True, ema=True).cuda().eval()
    g.renderer.network.encoder.embeddings.data.uniform_(-0.5, 0.5)
    cam, focal, near, far, _ = sg.generate_camera_params(8, "cuda", batch=2)
    cam2, focal2, near2, far2, _ = sg.generate_camera_params(8, "cuda", batch=2)
    z1, z2 = torch.randn(2, 256, device="cuda"), torch.randn(2, 256, device="cuda")
    with torch.no_grad():
        e1 = [t.clone() for t in g([z1], cam, focal, near, far, randomize_noise=False)]
        e2 = [t.clone() for t in g([z2], cam2, focal2, near2, far2, randomize_noise=False)]
    gg = sg.GraphedGenerator(g, [z1], cam, focal, near, far, randomize_noise=False)
    r1 = [t.clone() for t in gg([z1], cam, focal, near, far)]
    r2 = [t.clone() for t in gg([z2], cam2, focal2, near2, far2)]
    assert all(torch.equal(a, b) for a, b in zip(e1, r1)) and all(torch.equal(a, b) for a, b in zip(e2, r2))
    assert not torch.equal(r1[0], r2[0])
    for m in g.modules():
        if isinstance(m, sg.decoder.NoiseInjection):
            m.weight.data.fill_(0.5)                                        # the reference initialises the noise strength to 0
    gn = sg.GraphedGenerator(g, [z1], cam, focal, near, far)                # randomize_noise=True
    n1 = gn([z1], cam, focal, near, far)[0].clone()
    n2 = gn([z1], cam, focal, near, far)[0].clone()
    assert torch.isfinite(n1).all() and not torch.equal(n1, n2)            # fresh noise per replay
    assert H.max_abs(gn([z1], cam, focal, near, far)[1], e1[1]) == 0.0      # the thumbnail has no noise input


def _ref_weights(wf):
    """wf [B, 9 | 1, Cout, Cin] fp16 (what sdfg_modconv_fold produces) -> fp32 [B, Cout, Cin, k, k]"""
    B, taps, Cout, Cin = wf.shape
    k = 3 if taps == 9 else 1
    return wf.float().permute(0, 2, 3, 1).reshape(B, Cout, Cin, k, k)


@pytest.mark.gpu
@pytest.mark.parametrize("B,Hh,Ww,Cin,Cout", [(2, 8, 8, 64, 128), (1, 16, 32, 128, 128), (3, 64, 64, 64, 256), (1, 6, 16, 64, 128), (2, 16, 16, 128, 64)])
def test_conv_and_upconv_ops_match_torch(B, Hh, Ww, Cin, Cout):
    """Operator level, against plain torch fp32 on the same fp16-rounded operands: the 3 x 3 convolution and the up-sampling layer
    (conv_transpose2d stride 2 as four parity classes + two edge strips -> blur -> noise + bias -> leaky ReLU * sqrt 2), at shapes with one
    pixel tile per sample, odd tile counts, a height that is not a multiple of the tile height, and all N tiles (64 / 128 / 256).
    Tolerance: fp16 output (+ one fp16 intermediate for the up-sampling layer): 3e-3 of the output's max."""
    import torch.nn.functional as F
    from sdface_gan_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(Hh * 131 + Cout)
    x = (torch.randn(B, Hh, Ww, Cin, generator=g) * 0.5).half().cuda()
    wf = (torch.randn(B, 9, Cout, Cin, generator=g) * (1.0 / (3 * Cin ** 0.5))).half().cuda()
    bias = torch.randn(Cout, generator=g).cuda() * 0.1
    nw = torch.tensor([0.3]).cuda()
    w5 = _ref_weights(wf)
    xn = x.float().permute(0, 3, 1, 2)                                       # NCHW fp32
    lrelu = lambda t: F.leaky_relu(t, 0.2) * 2 ** 0.5
    # 3 x 3
    noise = torch.randn(B, Hh, Ww, generator=g).cuda()
    got = ops.conv_forward(x, wf, bias=bias, noise=noise, noise_w=nw).float().permute(0, 3, 1, 2)
    ref = torch.stack([F.conv2d(xn[b:b + 1], w5[b], padding=1)[0] for b in range(B)])
    ref = lrelu(ref + nw * noise[:, None] + bias.view(1, -1, 1, 1))
    assert H.max_abs(got, ref) < 3e-3 * float(ref.abs().max())
    # up-sampling layer
    noise2 = torch.randn(B, 2 * Hh, 2 * Ww, generator=g).cuda()
    got = ops.upconv_forward(x, wf, bias=bias, noise=noise2, noise_w=nw).float().permute(0, 3, 1, 2)
    t = torch.stack([F.conv_transpose2d(xn[b:b + 1], w5[b].transpose(0, 1), stride=2)[0] for b in range(B)])      # [B, Cout, 2H+1, 2W+1]
    ref = do.upfirdn2d(t.cpu(), do._blur_kernel(4.0), pad=(1, 1)).cuda()
    ref = lrelu(ref + nw * noise2[:, None] + bias.view(1, -1, 1, 1))
    assert got.shape == ref.shape == (B, Cout, 2 * Hh, 2 * Ww)
    assert H.max_abs(got, ref) < 3e-3 * float(ref.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("B,Hh,C,with_skip", [(2, 8, 64, False), (3, 16, 128, True), (1, 64, 512, True)])
def test_to_rgb_op_matches_torch(B, Hh, C, with_skip):
    """ToRGB as the N = 16 GEMM of the conv kernel: modulated 1 x 1 convolution (no demodulation) + bias + up-sampled skip, both output
    layouts, against torch fp32 on the fp16-rounded activation (weights are rounded to fp16 inside: 2e-3 of the output's max)."""
    from sdface_gan_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(C + Hh)
    x = (torch.randn(B, Hh, Hh, C, generator=g) * 0.5).half().cuda()
    w = torch.randn(1, 3, C, 1, 1, generator=g).cuda()
    style = (1 + 0.3 * torch.randn(B, C, generator=g)).cuda()
    bias = torch.randn(1, 3, 1, 1, generator=g).cuda() * 0.1
    skip = torch.randn(B, Hh // 2, Hh // 2, 3, generator=g).cuda() if with_skip else None
    scale = 1 / C ** 0.5
    nhwc, nchw = ops.to_rgb(x, w, style, scale, bias, skip, want_nhwc=True, want_nchw=True)
    ref = torch.einsum("bhwc,oc,bc->bohw", x.float(), w.view(3, C), style) * scale + bias
    if with_skip:
        ref = ref + do.upfirdn2d(skip.permute(0, 3, 1, 2).cpu(), do._blur_kernel(4.0), up=2, pad=(2, 1)).cuda()
    assert torch.equal(nhwc.permute(0, 3, 1, 2), nchw)
    assert H.max_abs(nchw, ref) < 2e-3 * float(ref.abs().max())
