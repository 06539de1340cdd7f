"""Shared test plumbing: golden fixtures -> (oracle parameter dict, inputs) and -> a product Generator on the GPU."""
import os

import numpy as np
import torch

import param_fill as pf

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def fixture_cfg(z):
    rend = {str(k): float(v) for k, v in zip(z["cfg_rend_keys"], z["cfg_rend_vals"])}
    return dict(net_type=str(z["cfg_net_type"]), B=int(z["cfg_B"]), res=int(z["cfg_res"]), S=int(z["cfg_S"]), fc=int(z["cfg_fc"]),
                perturb=float(z["cfg_perturb"]), no_features_output=bool(int(z["cfg_no_features_output"])),
                table_std=None if float(z["cfg_table_std"]) < 0 else float(z["cfg_table_std"]), rend=rend)


def fixture_params(z, seed=0, requires_grad=False):
    """All parameters of the fixture's Generator as {full_name: tensor} (CPU, float32)."""
    cfg = fixture_cfg(z)
    out = {}
    for name, shape, std, mean in pf.table_from_npz(z):
        if cfg["table_std"] is not None and name.endswith("encoder.embeddings"):
            std = cfg["table_std"]
        t = torch.from_numpy(pf.values_for(name, shape, std, seed, mean))
        out[name] = t.requires_grad_(requires_grad)
    return out


def oracle_param_dicts(params):
    """Split full names into (renderer-dict for oracle.field_oracle, style-dict for field_oracle.mapping)."""
    import oracle
    rp = {k[len("renderer."):]: v for k, v in params.items() if k.startswith("renderer.")}
    sp = {k[len("style."):]: v for k, v in params.items() if k.startswith("style.")}
    if "network.encoder.embeddings" in rp:
        from oracle import field_oracle as fo
        offsets, _ = oracle.grid_offsets(**fo.NGP_GRID)
        rp["network.encoder.offsets"] = torch.from_numpy(offsets)
    return rp, sp


def fixture_inputs(z, device="cpu"):
    t = lambda k: torch.from_numpy(z[k]).to(device)
    d = dict(cam=t("cam"), focal=t("focal"), near=t("near"), far=t("far"), z=t("z"))
    d["t_rand"] = t("t_rand") if "t_rand" in z.files else None
    return d


def render_kwargs(cfg):
    r = cfg["rend"]
    return dict(res=cfg["res"], S=cfg["S"], net_type="ngp" if cfg["net_type"] == "ngp" else "sdf", fc=bool(cfg["fc"]),
                offset_sampling=not bool(r.get("no_offset_sampling", 0)), z_normalize=not bool(r.get("no_z_normalize", 0)),
                static_viewdirs=bool(r.get("static_viewdirs", 0)), with_sdf=not bool(r.get("no_sdf", 0)),
                output_features=not cfg["no_features_output"], force_background=bool(r.get("force_background", 0)),
                return_sdf=bool(r.get("return_sdf", 0)), return_xyz=bool(r.get("return_xyz", 0)))


def product_generator(z, device, precision="fp32", seed=0):
    """The product Generator (sdface-gan_b200) built with the fixture's options and filled with the fixture's weights."""
    import sdface_gan_b200 as sg
    cfg = fixture_cfg(z)
    r = cfg["rend"]
    over = dict(perturb=cfg["perturb"])
    for k in ("no_offset_sampling", "no_z_normalize", "static_viewdirs", "no_sdf", "force_background", "return_sdf", "return_xyz"):
        if k in r:
            over[k] = bool(r[k])
    if cfg["no_features_output"]:
        over["no_features_output"] = True
    if cfg.get("fc"):
        over["fc"] = 1
    mo, ro = sg.default_options(cfg["net_type"], renderer_res=cfg["res"], n_samples=cfg["S"], **over)
    g = sg.Generator(mo, ro, full_pipeline=False)
    pf.fill_state(g, pf.table_from_npz(z), seed, table_std_override=cfg["table_std"])
    g = g.to(device)
    g.renderer.network.precision = precision
    return g


def max_abs(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0


def rel_err(a, b):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
