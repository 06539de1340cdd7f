"""Debug aid: per-parameter gradient errors of the tc16 path vs the CPU oracle at full size, table amplitude from argv."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch
import test_gpu_fullsize as T
import helpers as H
amp = float(sys.argv[1]); feats = bool(int(sys.argv[2]))
g, (cam, focal, near, far, z, t_rand), lw, ref = T._setup(amp, feats)
g = g.cuda(); g.renderer.network.precision = "tc16"
d = lambda t: t.cuda()
_, thumb, sdf, eik = g([d(z)], d(cam), d(focal), d(near), d(far), return_sdf=True, return_eikonal=True, t_rand=d(t_rand))
feat = None
if feats:
    _, feat, _, _, _, _ = g.renderer(d(cam), d(focal), d(near), d(far), styles=g.style(d(z)), t_rand=d(t_rand))
T._loss(thumb, sdf, feat, lw).backward()
errs = {n: H.rel_err(p.grad, ref["grads"][n]) for n, p in g.named_parameters() if n in ref["grads"] and float(ref["grads"][n].abs().max()) > 0 and n != "renderer.sigmoid_beta"}
top = sorted(errs.items(), key=lambda kv: -kv[1])
import numpy as np
print(os.environ.get("TAG", ""), "amp", amp, "sdf err %.2e" % H.rel_err(sdf, ref["sdf"]), "median %.2e" % float(np.median(list(errs.values()))), " ".join("%s=%.1e" % (n.replace("renderer.network.", ""), e) for n, e in top[:6]))
