"""Golden vectors for gaussian_renderer.sh_to_rgb from the reference's own eval_sh (gaussian_splatting/utils/sh_utils.py:55-118)
and the colour expression of its render() (gaussian_splatting/gaussian_renderer/__init__.py:108-117).  Run in the build
container (the reference tree does not travel to the GPU box):  python tests/golden/make_sh_golden.py"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_sh_utils", "/root/reference/gaussian_splatting/utils/sh_utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

g = torch.Generator().manual_seed(11)
P, max_deg = 257, 3
M = (max_deg + 1) ** 2
features = torch.randn(P, M, 3, generator=g, dtype=torch.float64) * 1.5          # pc.get_features: [P, M, 3]
xyz = torch.randn(P, 3, generator=g, dtype=torch.float64) * 2.0
center = torch.tensor([0.3, -0.2, 1.5], dtype=torch.float64)
out = {"features": features.numpy(), "xyz": xyz.numpy(), "camera_center": center.numpy()}
for deg in range(max_deg + 1):
    # the reference's lines, verbatim in meaning: shs_view, dir_pp, dir_pp_normalized, eval_sh, clamp_min(. + 0.5, 0)
    shs_view = features.transpose(1, 2).view(-1, 3, M)
    dir_pp = xyz - center.repeat(features.shape[0], 1)
    dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
    sh2rgb = ref.eval_sh(deg, shs_view, dir_pp_normalized)
    out["rgb_deg%d" % deg] = torch.clamp_min(sh2rgb + 0.5, 0.0).numpy()
np.savez_compressed(os.path.join(HERE, "sh_eval_golden.npz"), **out)
print("wrote", os.path.join(HERE, "sh_eval_golden.npz"), {k: v.shape for k, v in out.items()})
