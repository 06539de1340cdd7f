#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (build container only; needs /root/reference).

    python tests/golden/make_golden.py

Each fixture holds: the option overrides, the (name, shape, std) parameter table (see param_fill.py), the inputs
(latents, camera poses, focals, near/far, injected jitter) and the reference's outputs.  The reference classes are imported
through ref_harness.py; for `--ngp 1` cases the two CUDA-only extensions are served by the C restatement in oracle/
(the reference has no CPU implementation), so those fixtures pin the reference's Python wiring + torch arithmetic and the
restatement together; the restatement alone is pinned on the GPU box against oracle/_ref (tests/test_gpu_parity_ref.py).

Also writes sh_deg8.npz by parsing the 64 polynomial expressions (and their 192 derivatives) out of the reference's
shencoder.cu and evaluating them in float32 on seeded unit vectors -- i.e. the reference's own arithmetic, interpreted.
"""
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import param_fill as pf  # noqa: E402
import ref_harness as rh  # noqa: E402

SEED = 0


def cameras(B, res, seed):
    from im2scene.sdf.models.sdf_utils import generate_camera_params
    g = torch.Generator().manual_seed(seed)
    loc = torch.stack([0.3 * torch.randn(B, generator=g), 0.15 * torch.randn(B, generator=g)], 1)
    cam, focal, near, far, vp = generate_camera_params(res, "cpu", locations=loc)
    return loc, cam, focal, near, far


def run_case(name, net_type, B, res, S, *, fc=0, table_std=None, perturb=0.0, grads=False, gen_kwargs=None,
             no_features_output=False, init_pass=False, **rend):
    sm = rh.install()
    torch.manual_seed(SEED)
    mo, ro = rh.default_opts(net_type, res=res, S=S, perturb=perturb, fc=fc, **rend)
    if no_features_output:
        ro["no_features_output"] = True
    g = sm.Generator(mo, ro, full_pipeline=False)
    tab = pf.param_table(g)
    pf.fill_state(g, tab, SEED, table_std_override=table_std)
    loc, cam, focal, near, far = cameras(B, res, 100 + B)
    zg = torch.Generator().manual_seed(7)
    z = torch.randn(B, mo.style_dim, generator=zg)
    out = dict(pf.table_to_npz(tab))
    out.update(cfg_net_type=net_type, cfg_B=B, cfg_res=res, cfg_S=S, cfg_fc=fc, cfg_perturb=perturb,
               cfg_no_features_output=int(no_features_output),
               cfg_table_std=-1.0 if table_std is None else table_std,
               cfg_rend_keys=np.array(list(rend.keys())), cfg_rend_vals=np.array([float(v) for v in rend.values()]),
               loc=loc.numpy(), cam=cam.numpy(), focal=focal.numpy(), near=near.numpy(), far=far.numpy(), z=z.numpy())
    gen_kwargs = dict(gen_kwargs or {})
    if init_pass:
        torch.manual_seed(31337)
        t_rand = torch.rand(B, res, res, S)
        torch.manual_seed(31337)
        sdf, target = g.init_forward([z], cam, focal, near, far)
        out.update(t_rand=t_rand.numpy(), init_sdf=sdf.detach().numpy(), init_target=target.detach().numpy())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "init_pass", sdf.shape)
        return
    if perturb > 0:
        torch.manual_seed(31337)
        if ro.no_offset_sampling:
            t_rand = torch.rand(B, res, res, S)
        else:
            t_rand = torch.rand(B, res, res)
        out["t_rand"] = t_rand.numpy()
        torch.manual_seed(31337)        # the jitter draw is the first RNG use inside forward (sdf_model.py:331/338)
    res_t = g([z], cam, focal, near, far, **gen_kwargs)
    # tuple protocol, sdf_model.py:1203-1216
    names = ["rgb", "thumb_rgb"]
    if gen_kwargs.get("return_xyz"):
        names.append("xyz")
    if gen_kwargs.get("return_sdf"):
        names.append("sdf")
    if gen_kwargs.get("return_eikonal"):
        names.append("eikonal")
    if gen_kwargs.get("return_xyz"):
        names.append("mask")
    named = dict(zip(names, res_t))
    # features are not part of Generator's tuple; fetch them from the renderer directly with the same inputs
    if perturb > 0:
        torch.manual_seed(31337)
    with torch.no_grad():
        style = g.style(z)
        r_rgb, r_feat, r_sdf, r_mask, r_xyz, _ = g.renderer(cam, focal, near, far, styles=style)
    out["style"] = style.numpy()
    if r_feat is not None:
        out["features"] = r_feat.numpy()
    for k, v in named.items():
        if v is not None:
            out["out_" + k] = v.detach().numpy()
    if grads:
        # a fixed, seeded linear functional of every differentiable output
        lg = torch.Generator().manual_seed(99)
        loss = 0
        for k in ("thumb_rgb", "sdf", "xyz", "mask"):
            v = named.get(k)
            if v is not None and v.requires_grad:
                w = torch.randn(v.shape, generator=lg)
                out["lossw_" + k] = w.numpy()
                loss = loss + (w * v).sum() / v.numel() ** 0.5
        g.zero_grad()
        loss.backward()
        out["loss"] = float(loss)
        for pname, p in g.named_parameters():
            if p.grad is None:
                continue
            d = pf.grad_digest(pname, p.grad.numpy())
            out["g_norm_" + pname] = d["norm"]
            out["g_proj_" + pname] = d["proj"]
            out["g_idx_" + pname] = d["idx"]
            out["g_val_" + pname] = d["val"]
        ge = dict(g.named_parameters()).get("renderer.network.encoder.embeddings")
        if ge is not None and ge.grad is not None:
            flat = ge.grad.reshape(-1).numpy()
            top = np.argsort(-np.abs(flat))[:256]
            out["g_top_idx_embeddings"] = top.astype(np.int64)
            out["g_top_val_embeddings"] = flat[top]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if k.startswith("out_") or k == "features"})


def sh_from_reference_source():
    """Evaluate the reference's SH expressions (shencoder.cu:49-123 values, :130-354 derivatives) in float32."""
    src = open(os.path.join(rh.REF, "im2scene/sdf/models/shencoder/src/shencoder.cu")).read()
    rs = np.random.RandomState(5)
    v = rs.standard_normal((257, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v = v.astype(np.float32)
    v[0] = [0, 0, 1]
    v[1] = [1, 0, 0]
    v[2] = [0, -1, 0]
    f32 = np.float32
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    env = dict(x=x, y=y, z=z)
    env.update(xy=x * y, xz=x * z, yz=y * z, x2=x * x, y2=y * y, z2=z * z)
    env["xyz"] = env["xy"] * z
    env.update(x4=env["x2"] * env["x2"], y4=env["y2"] * env["y2"], z4=env["z2"] * env["z2"])
    env.update(x6=env["x4"] * env["x2"], y6=env["y4"] * env["y2"], z6=env["z4"] * env["z2"])

    def ev(expr):
        # float literals "1.5f" -> np.float32(1.5); keeps float32 arithmetic throughout
        e = re.sub(r"(\d+\.\d+(?:[eE][-+]?\d+)?)f", r"f32(\1)", expr)
        r = eval(e, dict(f32=f32), env)
        return np.broadcast_to(np.asarray(r, np.float32), x.shape).copy()

    outs = np.zeros((v.shape[0], 64), np.float32)
    dd = np.zeros((v.shape[0], 3, 64), np.float32)
    for arr, tag in ((outs, "outputs"), (dd[:, 0], "dx"), (dd[:, 1], "dy"), (dd[:, 2], "dz")):
        found = re.findall(r"^\s*%s\[(\d+)\]\s*=\s*(.*?);" % tag, src, flags=re.M)
        assert len(found) == 64, (tag, len(found))
        for k, expr in found:
            arr[:, int(k)] = ev(expr.strip())
    np.savez_compressed(os.path.join(HERE, "sh_deg8.npz"), dirs=v, outputs=outs, dy_dx=dd)
    print("sh_deg8", outs.shape, dd.shape)


def camera_goldens():
    rh.install()
    from im2scene.sdf.models.sdf_utils import generate_camera_params
    loc = torch.tensor([[0.0, 0.0], [0.3, -0.15], [-0.45, 0.2], [1.2, 0.6], [0.0, 1.5607], [0.0004, -1.5703]])
    cam, focal, near, far, vp = generate_camera_params(64, "cpu", locations=loc, fov_ang=6, dist_radius=0.12)
    np.savez_compressed(os.path.join(HERE, "camera.npz"), loc=loc.numpy(), cam=cam.numpy(), focal=focal.numpy(),
                        near=near.numpy(), far=far.numpy(), vp=vp.numpy())
    print("camera", cam.shape)


def align_volume_golden():
    """The reference's own align_volume (sdf_utils.py:164-184) on a seeded, non-cubic volume (CPU torch; the reference function only accepts batch 1: its grid has batch 1)."""
    rh.install()
    from im2scene.sdf.models.sdf_utils import align_volume
    rs = np.random.RandomState(5)
    vol = torch.from_numpy(rs.standard_normal((1, 12, 10, 14, 1)).astype(np.float32))
    out = align_volume(vol.clone())
    out2 = align_volume(vol.clone(), near=0.7, far=1.3)
    np.savez_compressed(os.path.join(HERE, "align_volume.npz"), volume=vol.numpy(), out=out.numpy(), out_near07_far13=out2.numpy())
    print("align_volume", out.shape)


def decoder_table(dec):
    """(name, shape, std, mean) like pf.param_table, with the zero-initialised tensors (noise weights, activation / rgb biases)
    made non-zero so that every term of the decoder is exercised."""
    tab = []
    for name, shape, std, mean in pf.param_table(dec):
        if name.endswith("noise.weight"):
            mean = 0.3
        elif std == 0.0 and int(np.prod(shape)) > 1:
            std = 0.1
        tab.append((name, shape, std, mean))
    return tab


def decoder_golden():
    """The reference's own Decoder (sdf_model.py:883-1056; CPU branches of sdf_op.py) at 8^2 -> 32^2: 5 styled convolutions (2 of them
    up-sampling), 3 ToRGB with skip up-sampling, fixed noise buffers."""
    sm = rh.install()
    torch.manual_seed(SEED)
    mo, _ = rh.default_opts("ngp", res=8, size=32)
    mo["feature_encoder_in_channels"] = 256
    dec = sm.Decoder(mo)
    tab = decoder_table(dec)
    pf.fill_state(dec, tab, SEED)
    g = torch.Generator().manual_seed(11)
    feats = torch.randn(2, 256, 8, 8, generator=g)
    z = torch.randn(2, 256, generator=g)
    noise = [torch.randn(1, 1, 2 ** ((i + 2 * 3 + 1) // 2), 2 ** ((i + 2 * 3 + 1) // 2), generator=g) for i in range(dec.num_layers)]
    with torch.no_grad():
        img, latent = dec(feats, [z], noise=noise, return_latents=True)
        img_b, _ = dec(feats, [z], randomize_noise=False)               # the registered noise buffers, broadcast over the batch
    out = dict(pf.table_to_npz(tab))
    out.update(features=feats.numpy(), z=z.numpy(), image=img.numpy(), latent=latent.numpy(), image_buffers=img_b.numpy(),
               **{f"noise_{i}": n.numpy() for i, n in enumerate(noise)},
               **{f"buf_noise_{i}": getattr(dec.noises, f"noise_{i}").numpy() for i in range(dec.num_layers)})
    np.savez_compressed(os.path.join(HERE, "decoder.npz"), **out)
    print("decoder", img.shape, float(img.abs().mean()))


def main():
    global run_case
    only = sys.argv[1:]
    if only:
        _rc = run_case
        run_case = lambda name, *a, **k: _rc(name, *a, **k) if name in only else None
    if not only or "align_volume" in only:
        align_volume_golden()
    if not only or "decoder" in only:
        decoder_golden()
    if only and all(o in ("align_volume", "decoder") for o in only):
        return
    sh_from_reference_source()
    camera_goldens()
    # config-1 family: --sdf 1 --ngp 0 --fc 0 (SIREN 8x256), forward
    run_case("siren_fwd", "sdf", 2, 8, 24)
    # --fc 1 ablation
    run_case("fc_fwd", "sdf", 2, 6, 24, fc=1)
    # config-2/3 family: --ngp 1, forward, reference table init and a non-vacuous table
    run_case("ngp_fwd_init", "ngp", 2, 8, 24)
    run_case("ngp_fwd_tab1", "ngp", 2, 8, 24, table_std=1.0 / np.sqrt(3.0))
    # stage-1 training step shape (training_utils.py:396-445): jitter, no_features_output, sdf + eikonal, gradients
    run_case("ngp_train", "ngp", 2, 8, 24, table_std=1.0 / np.sqrt(3.0), perturb=1.0, grads=True, no_features_output=True,
             return_sdf=True, gen_kwargs=dict(return_sdf=True, return_eikonal=True))
    # full-feature backward (features consumed) -- gradients of thumb only via Generator, features via renderer fwd
    run_case("ngp_train_feat", "ngp", 2, 6, 24, table_std=1.0 / np.sqrt(3.0), perturb=1.0, grads=True)
    # the same at a tensor-core-eligible size (8 x 8 x 24 = 1536 = 12 x 128 samples per image)
    run_case("ngp_train_feat8", "ngp", 2, 8, 24, table_std=1.0 / np.sqrt(3.0), perturb=1.0, grads=True)
    run_case("siren_train", "sdf", 2, 6, 24, perturb=1.0, grads=True, no_features_output=True, return_sdf=True,
             gen_kwargs=dict(return_sdf=True, return_eikonal=True))
    # sdf_mesh.py surface generator (sdf_mesh.py:211-214,243-253): static dirs, forced background, xyz/sdf out, S = R
    run_case("ngp_mesh", "ngp", 1, 8, 32, table_std=1.0 / np.sqrt(3.0), static_viewdirs=True, force_background=True,
             return_xyz=True, return_sdf=True, gen_kwargs=dict(return_sdf=True, return_xyz=True))
    # stratified sampling + NeRF density branch
    run_case("ngp_nosdf_strat", "ngp", 2, 6, 24, table_std=1.0 / np.sqrt(3.0), perturb=1.0, no_sdf=True,
             no_offset_sampling=True, no_z_normalize=True)
    # sphere-init pass (sdf_model.py:380-409)
    # (the reference's own split([3,1]) at :403 only works with no_features_output, as stage 1 sets it)
    run_case("ngp_init_pass", "ngp", 3, 6, 24, init_pass=True, no_features_output=True)


if __name__ == "__main__":
    main()
