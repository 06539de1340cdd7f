import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import sdface_gan_b200 as sg
import oracle
from oracle import field_oracle as fo
offsets, pls = oracle.grid_offsets(**fo.NGP_GRID)
S = float(np.float32(np.log2(pls)))
dev = sg.ops.grid_level_scales(16, S, 16, "cuda").cpu().numpy()
lib = oracle.grid_level_scales(16, S, 16)
out = {"S": S, "S_hex": np.float32(S).view(np.uint32).item(), "H": 16, "L": 16,
       "device_hex": [int(v) for v in dev.view(np.uint32)], "libm_hex": [int(v) for v in lib.view(np.uint32)],
       "device": [float(v) for v in dev], "libm": [float(v) for v in lib]}
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "level_scales.json"), "w"), indent=1)
print(out)
