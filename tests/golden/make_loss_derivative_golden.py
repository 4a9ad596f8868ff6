"""Generates tests/golden/loss_derivative_2d_kat.json by IMPORTING the reference's own CPU script
/root/reference/Loss_Derivative_wrt_mu_and_cov.py (BASELINE.json configs[0]; pure numpy) and running its
compute_gradients_2D on a small seeded case.  Run in the build container only:
    python tests/golden/make_loss_derivative_golden.py
"""
import importlib.util
import json
import os

import numpy as np

REF = "/root/reference/Loss_Derivative_wrt_mu_and_cov.py"
spec = importlib.util.spec_from_file_location("ref_ld", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(4)
H, W, N = 12, 16, 5
gaussians = []
for i in range(N):
    A = rng.normal(size=(2, 2))
    gaussians.append(dict(mu_I=np.array([rng.uniform(2, W - 2), rng.uniform(2, H - 2)]), Sigma_I=A @ A.T * 3.0 + 2.0 * np.eye(2),
                          opacity=float(rng.uniform(0.2, 0.9)), color=rng.uniform(0, 1, 3), depth=float(rng.uniform(0.5, 4.0))))
rendered_color = rng.uniform(0, 1, (H, W, 3))
rendered_depth = rng.uniform(0, 4, (H, W))
gt_color = rng.uniform(0, 1, (H, W, 3))
gt_depth = rng.uniform(0, 4, (H, W))
grad_mu, grad_Sigma = ref.compute_gradients_2D(gaussians, rendered_color, rendered_depth, gt_color, gt_depth, (H, W))
alphas = [[ref.compute_alpha_at_pixel(g, np.array([u, v])) for u in (0, 5, W - 1)] for g in gaussians for v in (0, H - 1)]
out = dict(source=REF + ":3-145", H=H, W=W,
           gaussians=[{k: (np.asarray(v).tolist() if not isinstance(v, float) else v) for k, v in g.items()} for g in gaussians],
           rendered_color=rendered_color.tolist(), rendered_depth=rendered_depth.tolist(), gt_color=gt_color.tolist(),
           gt_depth=gt_depth.tolist(), grad_mu_I=np.asarray(grad_mu).tolist(), grad_Sigma_I=np.asarray(grad_Sigma).tolist(),
           alpha_samples=np.asarray(alphas).tolist())
p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "loss_derivative_2d_kat.json")
json.dump(out, open(p, "w"))
print("wrote", p, np.asarray(grad_mu)[0], np.asarray(grad_Sigma)[0])
