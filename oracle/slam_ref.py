"""TEST INFRASTRUCTURE (oracle) -- NOT product code: plain-torch restatement of the reference code around the rasterizer, used to check the fused
slam_ops kernels.  Each function follows the reference lines it cites (repo root = /root/reference).
Parity pinning: tests/golden/slam_golden.npz is produced by IMPORTING the reference's own utils/slam_utils.py, utils/pose_utils.py and
utils/camera_utils.py (tests/golden/make_slam_golden.py); tests/test_slam_golden.py checks this restatement against it on the CPU."""
import torch


# ---- utils/slam_utils.py -------------------------------------------------------------------------------------------
def loss_tracking_rgb(image_ab, opacity, gt_image, grad_mask, rgb_boundary_threshold):
    """get_loss_tracking_rgb, utils/slam_utils.py:63-73."""
    _, h, w = gt_image.shape
    rgb_pixel_mask = (gt_image.sum(dim=0) > rgb_boundary_threshold).view(1, h, w)
    rgb_pixel_mask = rgb_pixel_mask * grad_mask
    l1 = opacity * torch.abs(image_ab * rgb_pixel_mask - gt_image * rgb_pixel_mask)
    return l1.mean()


def loss_tracking(image, depth, opacity, gt_image, gt_depth, grad_mask, exposure_a, exposure_b, rgb_boundary_threshold, alpha,
                  monocular):
    """get_loss_tracking / get_loss_tracking_rgbd, utils/slam_utils.py:56-61, 76-89."""
    image_ab = torch.exp(exposure_a) * image + exposure_b
    l1_rgb = loss_tracking_rgb(image_ab, opacity, gt_image, grad_mask, rgb_boundary_threshold)
    if monocular:
        return l1_rgb
    depth_pixel_mask = (gt_depth > 0.01).view(*depth.shape)
    opacity_mask = (opacity > 0.95).view(*depth.shape)
    depth_mask = depth_pixel_mask * opacity_mask
    l1_depth = torch.abs(depth * depth_mask - gt_depth * depth_mask)
    return alpha * l1_rgb + (1 - alpha) * l1_depth.mean()


def loss_mapping(image, depth, gt_image, gt_depth, exposure_a, exposure_b, rgb_boundary_threshold, alpha, monocular,
                 initialization=False):
    """get_loss_mapping / _rgb / _rgbd, utils/slam_utils.py:92-128."""
    image_ab = image if initialization else torch.exp(exposure_a) * image + exposure_b
    _, h, w = gt_image.shape
    rgb_pixel_mask = (gt_image.sum(dim=0) > rgb_boundary_threshold).view(1, h, w)
    l1_rgb = torch.abs(image_ab * rgb_pixel_mask - gt_image * rgb_pixel_mask)
    if monocular:
        return l1_rgb.mean()
    depth_pixel_mask = (gt_depth > 0.01).view(*depth.shape)
    l1_depth = torch.abs(depth * depth_pixel_mask - gt_depth * depth_pixel_mask)
    return alpha * l1_rgb.mean() + (1 - alpha) * l1_depth.mean()


# ---- utils/pose_utils.py ---------------------------------------------------------------------------------------------
def skew_sym_mat(x):
    """utils/pose_utils.py:12-23."""
    ssm = torch.zeros(3, 3, device=x.device, dtype=x.dtype)
    ssm[0, 1] = -x[2]; ssm[0, 2] = x[1]; ssm[1, 0] = x[2]; ssm[1, 2] = -x[0]; ssm[2, 0] = -x[1]; ssm[2, 1] = x[0]
    return ssm


def SO3_exp(theta):
    """utils/pose_utils.py:26-41."""
    W = skew_sym_mat(theta)
    W2 = W @ W
    angle = torch.norm(theta)
    I = torch.eye(3, device=theta.device, dtype=theta.dtype)
    if angle < 1e-5:
        return I + W + 0.5 * W2
    return I + (torch.sin(angle) / angle) * W + ((1 - torch.cos(angle)) / (angle**2)) * W2


def V(theta):
    """utils/pose_utils.py:44-58."""
    I = torch.eye(3, device=theta.device, dtype=theta.dtype)
    W = skew_sym_mat(theta)
    W2 = W @ W
    angle = torch.norm(theta)
    if angle < 1e-5:
        return I + 0.5 * W + (1.0 / 6.0) * W2
    return I + W * ((1.0 - torch.cos(angle)) / (angle**2)) + W2 * ((angle - torch.sin(angle)) / (angle**3))


def SE3_exp(tau):
    """utils/pose_utils.py:61-73."""
    T = torch.eye(4, device=tau.device, dtype=tau.dtype)
    T[:3, :3] = SO3_exp(tau[3:])
    T[:3, 3] = V(tau[3:]) @ tau[:3]
    return T


def update_pose(R, T, cam_trans_delta, cam_rot_delta, converged_threshold=1e-4):
    """utils/pose_utils.py:76-93 -> (new_R, new_T, converged)."""
    tau = torch.cat([cam_trans_delta, cam_rot_delta], axis=0)
    T_w2c = torch.eye(4, device=tau.device, dtype=tau.dtype)
    T_w2c[0:3, 0:3] = R
    T_w2c[0:3, 3] = T
    new_w2c = SE3_exp(tau) @ T_w2c
    return new_w2c[0:3, 0:3], new_w2c[0:3, 3], bool(tau.norm() < converged_threshold)      # bool(): the reference's `if converged` sync


def camera_tensors(R, T, projection_matrix):
    """world_view_transform, full_proj_transform, camera_center: utils/camera_utils.py:96-109 with getWorld2View2
    (gaussian_splatting/utils/graphics_utils.py:33-46, translate 0, scale 1)."""
    Rt = torch.zeros((4, 4), device=R.device, dtype=R.dtype)
    Rt[:3, :3] = R
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    C2W = torch.linalg.inv(Rt)
    Rt = torch.linalg.inv(C2W)
    wvt = Rt.transpose(0, 1)
    full = wvt.unsqueeze(0).bmm(projection_matrix.unsqueeze(0)).squeeze(0)
    center = wvt.inverse()[3, :3]
    return wvt, full, center
