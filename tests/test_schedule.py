"""CPU: work-item tables of the fused backward + Adam kernel (videoloop3d_b200/schedule.py).

The queue invariant (an item only waits for counters that earlier items complete — otherwise the persistent kernel
could deadlock), coverage (every texel of the atlas gets exactly one Adam item, every tile exactly one BWD item) and
that every rectangle a tile row can touch is covered by the dependencies are checked on the host for several views."""
import numpy as np
import pytest
import torch

from oracle import mpv_oracle as MO
from videoloop3d_b200 import ops, schedule
from videoloop3d_b200.schedule import C_A, C_B, C_C, C_TYPE, ITEM_ADAM, ITEM_BWD, ITEM_ZERO

VIEWS = {
    "near_identity": dict(fmul=1.0, rot=("y", 0.03), trans=(0.05, -0.02, 0.01)),
    "zoom_out": dict(fmul=0.45, rot=("x", -0.02), trans=(0.0, 0.02, -0.05)),
    "roll": dict(fmul=1.05, rot=("z", 0.6), trans=(0.03, -0.02, 0.01)),
    "oblique": dict(fmul=0.9, rot=("y", 0.35), trans=(0.15, 0.0, 0.05)),
}


def _rot(ax, ang):
    c, s = np.cos(ang), np.sin(ang)
    return np.array({"x": [[1, 0, 0], [0, c, -s], [0, s, c]], "y": [[c, 0, s], [0, 1, 0], [-s, 0, c]],
                     "z": [[c, -s, 0], [s, c, 0], [0, 0, 1]]}[ax])


def _geometry(H, W, D, vname):
    st = MO.dense_state(H, W, D, 5, 8, 2, 2, 1.0, 10.0, 1.15, 1.15, seed=11)
    hd, wd = st.atlas_dyn.shape[-2:]
    pack = ops.make_mesh_pack(dict(verts=st.verts, faces=st.faces, uvs=st.uvs, uvfaces=st.uvfaces, atlas_hw=(1, 1),
                                   faces_dyn=st.faces_dyn, uvs_dyn=st.uvs_dyn, uvfaces_dyn=st.uvfaces_dyn,
                                   atlas_dyn_hw=(hd, wd)), D, 5, 8, "cpu")
    v = VIEWS[vname]
    ext = np.eye(4)
    ext[:3, :3] = _rot(*v["rot"])
    ext[:3, 3] = v["trans"]
    f = 0.8 * W * v["fmul"]
    intr = np.array([[f, 0, W / 2 + 0.37], [0, f, H / 2 - 0.21], [0, 0, 1.]])
    view = ops.make_view(pack, H, W, ext @ np.linalg.inv(st.ref_extrin.double().numpy()), intr, np.eye(4), (hd, wd), (1, 1))
    homs = np.ctypeslib.as_array(view.hom).reshape(-1)[:D * 9].astype(np.float64)
    return pack, view, homs, hd, wd


def _coverage(s, gx, gy, hd, wd):
    it = s.items
    kind = it[:, C_TYPE] & 15
    tiles = it[kind == ITEM_BWD]
    assert len(tiles) == gx * gy and len({(a, b) for a, b in tiles[:, [C_A, C_B]]}) == gx * gy
    cover = np.zeros((hd, wd), dtype=np.int32)
    for base, width, rows in it[kind == ITEM_ADAM][:, [C_A, C_B, C_C]]:
        r0, c0 = divmod(int(base), wd)
        cover[r0:r0 + rows, c0:c0 + width] += 1
    assert cover.min() == 1 and cover.max() == 1                    # every texel optimised exactly once per round


@pytest.mark.parametrize("vname", sorted(VIEWS))
@pytest.mark.parametrize("use_zero", [False, True])
def test_band_schedule_invariants(vname, use_zero):
    H, W, D = 75, 133, 6
    pack, view, homs, hd, wd = _geometry(H, W, D, vname)
    assert pack.rect_planes
    gx, gy, _, _ = schedule.tile_grid(H, W, True)
    for rb, za, lag in ((8, 2, 2), (3, 1, 0), (5, 4, 6)):
        s = schedule.band_schedule(homs, view.cx, view.cy, H, W, pack.table, D, pack.qh, pack.qw, hd, wd, True,
                                   row_block=rb, zero_ahead=za, adam_lag=lag, use_zero=use_zero)
        assert schedule.validate(s)
        _coverage(s, gx, gy, hd, wd)
        if use_zero:                                                # a zeroed rectangle is never re-zeroed by its Adam item
            it = s.items
            z = {(a, b, c) for a, b, c in it[(it[:, C_TYPE] & 15) == ITEM_ZERO][:, [C_A, C_B, C_C]]}
            ad = it[(it[:, C_TYPE] & 15) == ITEM_ADAM]
            for row in ad:
                fl = row[C_TYPE] >> 4
                key = (row[C_A], row[C_B], row[C_C])
                if fl & schedule.FLAG_HAS_GRAD:
                    assert (key in z) != bool(fl & schedule.FLAG_REZERO)


def test_generic_schedule_invariants():
    for H, W, hd, wd, smooth in ((75, 133, 172, 456, True), (720, 1280, 2880, 10240, True), (45, 80, 64, 96, False)):
        s = schedule.generic_schedule(H, W, hd, wd, smooth)
        assert schedule.validate(s) and s.extra_round
        gx, gy, _, _ = schedule.tile_grid(H, W, smooth)
        _coverage(s, gx, gy, hd, wd)


def test_band_dependencies_cover_exact_footprints():
    """Brute force on a small view: for every pixel of every tile row, the atlas rows its bilinear taps touch on every
    plane must lie in rectangles whose Adam item waits for that tile row."""
    H, W, D = 40, 64, 4
    pack, view, homs, hd, wd = _geometry(H, W, D, "oblique")
    s = schedule.band_schedule(homs, view.cx, view.cy, H, W, pack.table, D, pack.qh, pack.qw, hd, wd, True, row_block=4)
    gx, gy, sx, sy = schedule.tile_grid(H, W, True)
    it = s.items
    ad = it[(it[:, C_TYPE] & 15) == ITEM_ADAM]
    first = np.full((hd, wd), 1 << 30)
    last = np.full((hd, wd), -1)
    for row in ad:
        r0, c0 = divmod(int(row[C_A]), wd)
        if (row[C_TYPE] >> 4) & schedule.FLAG_HAS_GRAD:
            first[r0:r0 + row[C_C], c0:c0 + row[C_B]] = row[schedule.C_W0]
            last[r0:r0 + row[C_C], c0:c0 + row[C_B]] = row[schedule.C_W0] + row[schedule.C_WN] - 1
    X0, Y0, qsx, qsy = schedule._plane_rects(pack.table, D, pack.qh, pack.qw)
    h = homs.reshape(D, 3, 3)
    for R in range(gy):
        ys = np.arange(R * sy, min(R * sy + schedule.BY, H))
        xs = np.arange(W)
        u, v = np.meshgrid(xs + 0.5 - view.cx, ys + 0.5 - view.cy)
        for d in range(D):
            w = h[d, 2, 0] * u + h[d, 2, 1] * v + h[d, 2, 2]
            gxq = (h[d, 0, 0] * u + h[d, 0, 1] * v + h[d, 0, 2]) / w
            gyq = (h[d, 1, 0] * u + h[d, 1, 1] * v + h[d, 1, 2]) / w
            hit = (w > 0) & (gxq > 0) & (gxq < pack.qw) & (gyq > 0) & (gyq < pack.qh)
            lx, ly = X0[d] + gxq[hit] * qsx[d], Y0[d] + gyq[hit] * qsy[d]
            for dx in (0, 1):
                for dy in (0, 1):
                    cx = np.clip(np.floor(lx).astype(int) + dx, 0, wd - 1)
                    cy = np.clip(np.floor(ly).astype(int) + dy, 0, hd - 1)
                    assert np.all(first[cy, cx] <= R) and np.all(last[cy, cx] >= R), (R, d)
