"""GPU debug: where does the eikonal chain diverge from the oracle?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import numpy as np, torch
import helpers as H
import oracle
from oracle import field_oracle as fo
import sdface_gan_b200 as sg

z = H.load_fixture("ngp_train")
dev = "cuda"
g = H.product_generator(z, dev)
net = g.renderer.network
params = H.fixture_params(z)
rp, sp = H.oracle_param_dicts(params)
torch.manual_seed(0)
B, R, S = 2, 4, 6
npts = (torch.rand(B, R, R, S, 3) * 2 - 1)
vd = torch.nn.functional.normalize(torch.randn(B, R, R, 3), dim=-1)
style = torch.randn(B, 256) * 0.5
# oracle: features as a leaf
cfg = dict(fo.NGP_GRID); offsets = rp["network.encoder.offsets"]; _, pls = oracle.grid_offsets(**cfg)
pts_leaf = npts.clone().requires_grad_(True)
feat_o = fo.hash_encode(pts_leaf, rp["network.encoder.embeddings"], offsets, pls, 16, 2.0)
feat_o.retain_grad()
h = fo.linear_layer(rp, "network.input_linear", feat_o)
for i in range(3):
    h = fo.film_siren(rp, f"network.pts_linears.{i}", h, style)
sdf_o = fo.linear_layer(rp, "network.sigma_linear", h)
sdf_o.sum().backward()
dfeat_o = feat_o.grad.reshape(-1, 32)
dpts_o = pts_leaf.grad.reshape(-1, 3)
# product
sdf, rgb, feat, dsdf = net.forward_rays(npts.to(dev), vd.to(dev), style.to(dev), want_dsdf=True)
print("sdf max abs", H.max_abs(sdf.view(-1), sdf_o.view(-1)))
print("dpts rel", H.rel_err(dsdf, dpts_o), "max", dsdf.abs().max().item(), dpts_o.abs().max().item())
# stage by stage
flat = npts.reshape(-1, 3).to(dev)
x_in, view_feat, dy_dx = net._encode_rays(flat, vd.reshape(-1, 3).to(dev), True)
print("feat max abs", H.max_abs(x_in, feat_o.reshape(-1, 32)))
gamma, beta = net._modulation(style.to(dev))
from importlib import import_module
sm = import_module("sdface-gan_b200.sdf_model")
spec = net._spec
wts = sm._unpack_weights(spec, [w.detach().contiguous() for w in sm._pack_weights(spec, net)])
sdf2, _, _, ws = sg.ops.field_forward(spec, x_in.detach(), view_feat, gamma.detach().contiguous(), beta.detach().contiguous(), wts, R * R * S, S,
                                      want_rgb=True, want_feat=True, save_for_backward=True)
dx = sg.ops.field_backward(spec, x_in.detach(), view_feat, gamma.detach().contiguous(), beta.detach().contiguous(), wts, R * R * S, S, ws, _,
                           torch.ones_like(sdf2), None, None, grads=None, want_dx=True)
print("dfeat rel", H.rel_err(dx, dfeat_o), dx.abs().max().item(), dfeat_o.abs().max().item())
# grid input backward with the ORACLE's dfeat
enc = net.encoder
_, gi = sg.ops.grid_encode_backward(dfeat_o.to(dev).contiguous(), flat, enc.embeddings.detach(), enc.offsets, sg.ops.log2_scale(enc.per_level_scale), 16,
                                    bound=2.0, dy_dx=dy_dx, want_grad_inputs=True)
print("grid input bwd rel", H.rel_err(gi, dpts_o))
