"""Host-side layout of RasterEngine's flat gradient buffer (the tensor a keyframe window all-reduces), checked without a GPU:
flat_size() is what the constructor lays out, every per-Gaussian segment starts on a 16-byte boundary (vector REDs / float4
row stores of the backward), segments do not overlap, the [tau_slots, 8] pose-gradient block is the tail."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gs-slam-analytica_jacobian_b200"))


@pytest.mark.parametrize("P", [1, 7, 1001, 4096])
@pytest.mark.parametrize("variant", ["sh16_scales", "sh1_scales", "colors_cov", "sh1_cov"])
def test_flat_gradient_buffer_layout(P, variant):
    from diff_gaussian_rasterization.engine import RasterEngine

    M = 16 if variant.startswith("sh16") else 1
    g = dict(means3D=torch.zeros(P, 3), opacities=torch.zeros(P, 1))
    if variant.startswith("sh"):
        g["shs"] = torch.zeros(P, M, 3)
    else:
        g["colors_precomp"] = torch.zeros(P, 3)
    if variant.endswith("scales"):
        g["scales"], g["rotations"] = torch.zeros(P, 3), torch.zeros(P, 4)
    else:
        g["cov3D_precomp"] = torch.zeros(P, 6)
    e = RasterEngine(g, 200, 136, 0.5, 0.4, [0, 0, 0], sh_degree=0, device="cpu", tau_slots=40)
    n = e.grad_flat.numel()
    assert n == RasterEngine.flat_size(P, M, colors_precomp="colors_precomp" in g, cov3D_precomp="cov3D_precomp" in g, tau_slots=40)
    base = e.grad_flat.data_ptr()
    spans = []
    for name in ("g_means3D", "g_sh", "g_colors", "g_opacity", "g_rot", "g_scales", "g_cov"):
        t = getattr(e, name)
        if t is None:
            continue
        off = (t.data_ptr() - base) // 4
        assert (t.data_ptr() - base) % 16 == 0, name
        spans.append((off, off + t.numel(), name))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    tau0 = (e.tau_block.data_ptr() - base) // 4
    assert spans[-1][1] <= tau0 and tau0 + 8 * 40 == n and e.tau_block.shape == (40, 8)
    # a second engine over the same Gaussians may share the buffer (several streams adding into one window gradient)
    e2 = RasterEngine(g, 200, 136, 0.5, 0.4, [0, 0, 0], sh_degree=0, device="cpu", tau_slots=40, grad_flat=e.grad_flat)
    assert e2.g_means3D.data_ptr() == e.g_means3D.data_ptr() and e2.tau_block.data_ptr() == e.tau_block.data_ptr()


def test_pack_camera_block_layout():
    from diff_gaussian_rasterization.engine import RasterEngine

    vm, pm, pr = (torch.arange(16, dtype=torch.float32).reshape(4, 4) + k for k in (0, 100, 200))
    blk = RasterEngine.pack_camera(vm, pm, pr, torch.tensor([1.0, 2.0, 3.0]))
    assert blk.shape == (52,) and blk.dtype == torch.float32
    assert torch.equal(blk[0:16], vm.reshape(-1)) and torch.equal(blk[16:32], pm.reshape(-1)) and torch.equal(blk[32:48], pr.reshape(-1))
    assert blk[48:52].tolist() == [1.0, 2.0, 3.0, 0.0]
