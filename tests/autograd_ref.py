"""Independent differentiable re-statement of the forward pass in torch (float64, dense over
pixels x Gaussians, CPU) used to cross-check the ANALYTIC gradients of the oracle / CUDA path
with torch.autograd, the way the reference's VerifyJacobian.ipynb (cells 7-31) does for one
Gaussian.  The pose enters as the left perturbation T_cw <- Exp(tau) T_cw evaluated at tau=0
(utils/pose_utils.py:61-93).  Hard decisions of the kernels (tile membership, power>0,
alpha<1/255, T<1e-4 termination, 0.99 clamp, 1.3*tanfov clamp) are reproduced as constant
masks, so autograd yields exactly the gradient the analytic kernels claim to compute.
Test infrastructure only."""
import math

import numpy as np
import torch

SH_C0 = 0.28209479177387814


def _hat(v):
    z = torch.zeros((), dtype=v.dtype)
    return torch.stack([torch.stack([z, -v[2], v[1]]), torch.stack([v[2], z, -v[0]]), torch.stack([-v[1], v[0], z])])


def se3_exp_first_order(tau):
    """Exp(tau) to first order at tau=0 is enough for the gradient AT tau=0; use the full series
    (pose_utils.py:12-73 small-angle branch) to stay faithful."""
    rho, theta = tau[:3], tau[3:]
    W = _hat(theta)
    W2 = W @ W
    I = torch.eye(3, dtype=tau.dtype)
    R = I + W + 0.5 * W2
    V = I + 0.5 * W + W2 / 6.0
    T = torch.eye(4, dtype=tau.dtype)
    T = T.clone()
    T[:3, :3] = R
    T[:3, 3] = V @ rho
    return T


def render_autograd(sc, w2c, proj_raw_math, st, params=None, tau=None):
    """sc: scene dict (numpy), w2c: 4x4 float64 world->camera, proj_raw_math: 4x4 projection P
    (NOT transposed), st: oracle forward state at the same pose (for the constant masks/order).
    params: dict of torch leaf tensors overriding means3D/scales/rotations/opacities/shs (deg 0).
    Returns color[3,H,W], depth[H,W]."""
    dt = torch.float64
    W, H = sc["image_width"], sc["image_height"]
    g = lambda k: (params[k] if params and k in params else torch.tensor(np.asarray(sc[k]), dtype=dt))
    means, scales, rots, opac, shs = g("means3D"), g("scales"), g("rotations"), g("opacities"), g("shs")
    T = torch.tensor(w2c, dtype=dt)
    if tau is not None:
        T = se3_exp_first_order(tau) @ T
    Pm = torch.tensor(proj_raw_math, dtype=dt)
    full = Pm @ T
    P = means.shape[0]
    hom = torch.cat([means, torch.ones(P, 1, dtype=dt)], 1)
    p_view = (T @ hom.T).T[:, :3]
    p_h = (full @ hom.T).T
    p_w = 1.0 / (p_h[:, 3] + 1e-7)
    ndc = p_h[:, :2] * p_w[:, None]
    pix = torch.stack([((ndc[:, 0] + 1.0) * W - 1.0) * 0.5, ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5], 1)
    # cov3D (forward.cu:120-154), quaternion not normalised
    r, x, y, z = rots[:, 0], rots[:, 1], rots[:, 2], rots[:, 3]
    Rq = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], 1),
        torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], 1),
        torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1)], 1)
    S = torch.diag_embed(scales * float(sc["scale_modifier"]))
    L = Rq @ S
    Sigma = L @ L.transpose(1, 2)
    fx = W / (2.0 * sc["tanfovx"])
    fy = H / (2.0 * sc["tanfovy"])
    tz = p_view[:, 2]
    limx, limy = 1.3 * sc["tanfovx"], 1.3 * sc["tanfovy"]
    tx = torch.clamp(p_view[:, 0] / tz, -limx, limx) * tz
    ty = torch.clamp(p_view[:, 1] / tz, -limy, limy) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack([torch.stack([fx / tz, zero, -fx * tx / (tz * tz)], 1),
                     torch.stack([zero, fy / tz, -fy * ty / (tz * tz)], 1)], 1)
    A = J @ T[:3, :3]
    cov = A @ Sigma @ A.transpose(1, 2)
    a = cov[:, 0, 0] + 0.3
    b = cov[:, 0, 1]
    c = cov[:, 1, 1] + 0.3
    det = a * c - b * b
    conic = torch.stack([c / det, -b / det, a / det], 1)
    color = torch.clamp(SH_C0 * shs[:, 0, :] + 0.5, min=0.0)
    depth = p_view[:, 2]
    # dense compositing in the oracle's per-tile order with the oracle's constant decisions
    out_c = torch.zeros(3, H, W, dtype=dt)
    out_d = torch.zeros(H, W, dtype=dt)
    ranges, plist = st["ranges"], st["point_list"].astype(np.int64)
    gx = (W + 15) // 16
    bg = torch.tensor(np.asarray(sc["bg"]), dtype=dt)
    for tile in range(ranges.shape[0]):
        r0, r1 = int(ranges[tile, 0]), int(ranges[tile, 1])
        tx0, ty0 = (tile % gx) * 16, (tile // gx) * 16
        xs = torch.arange(tx0, min(tx0 + 16, W), dtype=dt)
        ys = torch.arange(ty0, min(ty0 + 16, H), dtype=dt)
        if len(xs) == 0 or len(ys) == 0:
            continue
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        npix = yy.numel()
        if r1 <= r0:
            for ch in range(3):
                out_c[ch, ty0:ty0 + len(ys), tx0:tx0 + len(xs)] = bg[ch]
            continue
        ids = torch.tensor(plist[r0:r1])
        dx = pix[ids, 0][None, :] - xx.reshape(-1, 1)
        dy = pix[ids, 1][None, :] - yy.reshape(-1, 1)
        co = conic[ids]
        power = -0.5 * (co[:, 0] * dx * dx + co[:, 2] * dy * dy) - co[:, 1] * dx * dy
        alpha_raw = opac[ids, 0][None, :] * torch.exp(power)
        alpha = torch.clamp(alpha_raw, max=0.99)
        with torch.no_grad():
            ok = (power <= 0) & (alpha >= 1.0 / 255.0)
            # sequential termination exactly like forward.cu:500-505
            Tn = torch.ones(npix, dtype=dt)
            live = torch.ones(npix, dtype=torch.bool)
            use = torch.zeros_like(ok)
            for j in range(len(ids)):
                cand = ok[:, j] & live
                test = Tn * (1 - alpha[:, j])
                stop = cand & (test < 1e-4)
                live = live & ~stop
                take = cand & ~stop
                use[:, j] = take
                Tn = torch.where(take, test, Tn)
        am = torch.where(use, alpha, torch.zeros_like(alpha))
        one_m = 1 - am
        Tcum = torch.cumprod(torch.cat([torch.ones(npix, 1, dtype=dt), one_m[:, :-1]], 1), 1)
        w = am * Tcum
        Tfin = Tcum[:, -1] * one_m[:, -1]
        C = w @ color[ids] + Tfin[:, None] * bg[None, :]
        D = w @ depth[ids]
        for ch in range(3):
            out_c[ch, ty0:ty0 + len(ys), tx0:tx0 + len(xs)] = C[:, ch].reshape(len(ys), len(xs))
        out_d[ty0:ty0 + len(ys), tx0:tx0 + len(xs)] = D.reshape(len(ys), len(xs))
    return out_c, out_d
