"""Generates tests/golden/analytic_jacobian_kat.json from the REFERENCE's own numpy functions.

Run in the build container only (needs /root/reference):
    python tests/golden/make_kat_golden.py
The reference's closed-form single-Gaussian pose Jacobians live in Loss_Derivative_script.py
(pi_func .. GetAnalyticalJcobian, :454-567; the same code as 3DGS_Analytical_Jacobian.ipynb
cells 3-7).  The module cannot be imported (module-level import of the CUDA extension), so the
function block is sliced out of the source text and exec'd unchanged.  Cases: the notebook's two
poses (cells 1 and 8, whose printed outputs this reproduces) plus seeded random poses.
"""
import json
import os
import re

import numpy as np

REF = "/root/reference/Loss_Derivative_script.py"
src = open(REF).read()
start = src.index("def pi_func(v):")
end = src.index("def compute_analytical_jacobians_all_gaussians")
ns = {"np": np}
exec(compile(src[start:end], REF, "exec"), ns)
GetAnalyticalJcobian = ns["GetAnalyticalJcobian"]


def rot(theta):
    a = np.linalg.norm(theta)
    K = np.array([[0, -theta[2], theta[1]], [theta[2], 0, -theta[0]], [-theta[1], theta[0], 0]])
    return np.eye(3) + np.sin(a) / a * K + (1 - np.cos(a)) / a**2 * K @ K


cases = []
T1 = np.array([[0.8047, -0.3106, 0.5059, 1.0], [0.5059, 0.8047, -0.3106, 1.0], [-0.3106, 0.5059, 0.8047, 1.0], [0, 0, 0, 1.0]])
T2 = np.array([[1, 0, 0, 1.0], [0, 1, 0, 1.0], [0, 0, 1, 1.0], [0, 0, 0, 1.0]])
mu = np.array([2, 3, 4, 1.0])
cov = np.array([[1, 2, 3], [2, 4, 5], [3, 5, 9.0]])
inputs = [("notebook_cell1_rotated", T1, mu, cov), ("notebook_cell8_translation", T2, mu, cov)]
rng = np.random.default_rng(7)
for i in range(6):
    T = np.eye(4)
    T[:3, :3] = rot(rng.normal(size=3) * 0.7)
    T[:3, 3] = rng.normal(size=3)
    p_c = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(1.5, 5.0)])
    p_w = np.linalg.inv(T) @ np.append(p_c, 1.0)
    A = rng.normal(size=(3, 3))
    inputs.append(("random_%d" % i, T, p_w, A @ A.T + 0.1 * np.eye(3)))
for name, T, m, c in inputs:
    dmu, dcov = GetAnalyticalJcobian(T, m, c)
    cases.append(dict(name=name, T_cw=T.tolist(), mu_w=m.tolist(), cov_3D=c.tolist(),
                      dmuI_dTcw=np.asarray(dmu).tolist(), dcovI_dTcw=np.asarray(dcov).tolist()))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "analytic_jacobian_kat.json")
json.dump(dict(source=REF + ":454-567", cases=cases), open(out, "w"), indent=1)
print("wrote", out, len(cases), "cases")
print(np.asarray(cases[0]["dcovI_dTcw"]))
