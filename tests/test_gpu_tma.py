"""GPU: the TMA-staged render / backward (dense layout, VL3D_VIEW_RECT_PLANES) against the per-thread-load kernels
and the CPU oracle.

The TMA path must never change results: a tap outside the staged box is read from global memory, border tiles
take the per-thread path.  Views: near 1:1 (everything from the box), zoom-out (footprint larger than the 40x12 box:
most taps fall back), a 0.6 rad roll (the axis-aligned box of a rotated tile), an oblique view (planes leave the image:
mixed tiles); frame counts that are not multiples of the frame chunk; image sizes that are not multiples of the tiles."""
import numpy as np
import pytest
import torch

from oracle import mpv_oracle as MO
from test_gpu_parity import state_tensors

pytestmark = pytest.mark.gpu

VIEWS = {
    "near_identity": dict(fmul=1.0, rot=("y", 0.03), trans=(0.05, -0.02, 0.01)),
    "zoom_out": dict(fmul=0.45, rot=("x", -0.02), trans=(0.0, 0.02, -0.05)),
    "zoom_in": dict(fmul=2.3, rot=("y", 0.01), trans=(0.01, 0.0, 0.0)),
    "roll": dict(fmul=1.05, rot=("z", 0.6), trans=(0.03, -0.02, 0.01)),
    "oblique": dict(fmul=0.9, rot=("y", 0.35), trans=(0.15, 0.0, 0.05)),
}


def _rot(ax, ang):
    c, s = np.cos(ang), np.sin(ang)
    R = {"x": [[1, 0, 0], [0, c, -s], [0, s, c]], "y": [[c, 0, s], [0, 1, 0], [-s, 0, c]],
         "z": [[c, -s, 0], [s, c, 0], [0, 0, 1]]}[ax]
    return torch.tensor(R, dtype=torch.float32)


@pytest.mark.parametrize("vname", sorted(VIEWS))
@pytest.mark.parametrize("T", [3, 7])
def test_tma_paths_equal_per_thread_loads(vname, T):
    from videoloop3d_b200 import ops
    from videoloop3d_b200.testing import model_from_tensors
    dev = torch.device("cuda:0")
    H, W, D = 75, 133, 6
    st = MO.dense_state(H, W, D, 5, 8, 2, T, 1.0, 10.0, 1.15, 1.15, seed=11)
    st.atlas = st.atlas[:, :, :1, :1].clone()
    m = model_from_tensors(state_tensors(st), H, W, dev)
    v = VIEWS[vname]
    ext = torch.eye(4)
    ext[:3, :3] = _rot(*v["rot"])
    ext[:3, 3] = torch.tensor(v["trans"])
    f = 0.8 * W * v["fmul"]
    intr = torch.tensor([[f, 0, W / 2 + 0.37], [0, f, H / 2 - 0.21], [0, 0, 1.]])
    atlas_dyn, atlas = m._texels()
    view = m.make_view(H, W, (ext @ torch.inverse(st.ref_extrin)).double().numpy(), intr[None])
    pack = m._pack
    assert pack.rect_planes and view.flags == 1
    g = torch.Generator(device=dev).manual_seed(5)
    grad_rgb = torch.randn((T, 3, H, W), device=dev, generator=g)
    w_smooth = torch.tensor([0.011, 0.013, 0.017, 0.019], device=dev)

    def run(view):
        rgb = torch.empty((T, 3, H, W), device=dev)
        ops.composite_fwd(view, pack, atlas_dyn.data, atlas.data, None, T, 0, rgb_out=rgb)
        g_dyn, g_sta = torch.zeros_like(atlas_dyn.data), torch.zeros_like(atlas.data)
        sums = torch.zeros(4, dtype=torch.float64, device=dev)
        ops.composite_bwd(view, pack, atlas_dyn.data, atlas.data, None, T, 0, grad_rgb, rgb, w_smooth, g_dyn, g_sta,
                          smooth_sums=sums)
        torch.cuda.synchronize()
        return rgb, g_dyn, sums

    # VL3D_VIEW_RECT_PLANES is a hint ("results never depend on it", include/vl3d.h): without it the kernels use
    # per-thread loads for every tile
    plain = type(view).from_buffer_copy(view)
    plain.flags = 0
    rgb0, g0, s0 = run(plain)
    rgb1, g1, s1 = run(view)                                         # TMA staging
    assert torch.equal(rgb0, rgb1)                                   # bit-identical render
    assert float((g1 - g0).abs().max()) <= 2e-6 * float(g0.abs().max())       # same terms, RED order differs
    assert float(((s1 - s0).abs() / s0.abs().clamp_min(1e-30)).max()) < 1e-9
    # and the render against the oracle
    rgb_o, _ = MO.render(st, H, W, ext[None], intr[None], list(range(T)))
    assert float((rgb1.permute(0, 2, 3, 1).cpu().double() - rgb_o).abs().max()) < 1e-4
