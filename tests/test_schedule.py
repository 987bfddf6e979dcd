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


# ------------------------------------------------------------------------------------------------
# owner mode (tiles.own_descriptor): the host-side facts the kernel's ownership predicate rests on
# ------------------------------------------------------------------------------------------------
def _taps_of_pixels(pack, homs_d, cx, cy, H, W, d, D, qh, qw):
    """float64 restatement of the kernels' tap geometry: per pixel the top-left tap (ix, iy) on plane d, or -1."""
    t = pack.table.reshape(D, qh, qw)
    px, py = np.meshgrid(np.arange(W), np.arange(H))
    u, v = px + 0.5 - cx, py + 0.5 - cy
    Hm = homs_d.reshape(3, 3)
    w = Hm[2, 0] * u + Hm[2, 1] * v + Hm[2, 2]
    ok = w > 1e-12
    ws = np.where(ok, w, 1.0)
    gx = (Hm[0, 0] * u + Hm[0, 1] * v + Hm[0, 2]) / ws
    gy = (Hm[1, 0] * u + Hm[1, 1] * v + Hm[1, 2]) / ws
    ok &= (gx > 0) & (gx < qw) & (gy > 0) & (gy < qh)
    qx = np.clip(gx.astype(np.int64), 0, qw - 1)
    qy = np.clip(gy.astype(np.int64), 0, qh - 1)
    q = t[d][qy, qx]
    lx = q["x0i"] + (q["x0f"].astype(np.float64) + (gx - qx) * q["sx"])
    ly = q["y0i"] + (q["y0f"].astype(np.float64) + (gy - qy) * q["sy"])
    return np.where(ok, np.floor(lx), -1).astype(np.int64), np.where(ok, np.floor(ly), -1).astype(np.int64), ok


@pytest.mark.parametrize("vname", sorted(VIEWS))
def test_own_descriptor_reach_bound_and_partition(vname):
    """(1) private rectangles are disjoint, lie inside their cell and the cells partition the atlas; (2) for every pixel
    and each of its four taps, the screen position of the tapped texel (inverse map of the descriptor) is closer than
    `reach` to the pixel — the property that makes 'texel centre >= reach inside the tile' imply 'every contributing
    pixel is inside the tile'; (3) the owner-mode schedule covers every texel exactly once."""
    from videoloop3d_b200 import tiles
    H, W, D = 75, 133, 6
    pack, view, homs, hd, wd = _geometry(H, W, D, vname)
    qh, qw = pack.qh, pack.qw
    hinv, reach, rect, cells = tiles.own_descriptor(pack.table, D, qh, qw, homs.reshape(D, 9).astype(np.float32), view.cx, view.cy,
                                                    H, W, (hd, wd))
    cover = np.zeros((hd, wd), dtype=np.int32)
    for (x0, y0, x1, y1, plane) in cells:
        cover[y0:y1 + 1, x0:x1 + 1] += 1
        if plane >= 0 and rect[plane, 2] >= rect[plane, 0]:
            assert x0 <= rect[plane, 0] and rect[plane, 2] <= x1 and y0 <= rect[plane, 1] and rect[plane, 3] <= y1
    assert cover.min() == 1 and cover.max() == 1
    planes_in_cells = [c[4] for c in cells if c[4] >= 0]
    assert len(planes_in_cells) == len(set(planes_in_cells))
    touched_by = np.zeros((hd, wd), dtype=np.int32)
    for d in range(D):
        ix, iy, ok = _taps_of_pixels(pack, homs[d * 9:(d + 1) * 9], view.cx, view.cy, H, W, d, D, qh, qw)
        if not ok.any():
            continue
        mark = np.zeros((hd, wd), dtype=bool)
        for dx in (0, 1):
            for dy in (0, 1):
                tx, ty = np.minimum(ix[ok] + dx, wd - 1), np.minimum(iy[ok] + dy, hd - 1)
                mark[ty, tx] = True
                if reach[d] < 16:
                    # screen position of the tapped texel vs the pixel that taps it
                    Mi = hinv[d].astype(np.float64).reshape(3, 3)
                    lxx, lyy = tx - rect[d, 0], ty - rect[d, 1]
                    wq = Mi[2, 0] * lxx + Mi[2, 1] * lyy + Mi[2, 2]
                    inside = (tx >= rect[d, 0]) & (tx <= rect[d, 2]) & (ty >= rect[d, 1]) & (ty <= rect[d, 3])
                    assert np.all(wq[inside] > 0)
                    qx = (Mi[0, 0] * lxx + Mi[0, 1] * lyy + Mi[0, 2]) / wq
                    qy = (Mi[1, 0] * lxx + Mi[1, 1] * lyy + Mi[1, 2]) / wq
                    pxs, pys = np.meshgrid(np.arange(W), np.arange(H))
                    dist = np.maximum(np.abs(qx - pxs[ok]), np.abs(qy - pys[ok]))
                    assert float(dist[inside].max(initial=0.0)) < reach[d] - 0.03, (d, float(dist[inside].max()), reach[d])
        touched_by += mark
        # texels inside the private rectangle are tapped by this plane only (checked below via touched_by)
    for d in range(D):
        if rect[d, 2] >= rect[d, 0]:
            assert touched_by[rect[d, 1]:rect[d, 3] + 1, rect[d, 0]:rect[d, 2] + 1].max() <= 1
    s = schedule.generic_schedule(H, W, hd, wd, True, seg_texels=2048, cells=cells)
    assert s.kind == "own" and schedule.validate(s)
    gx, gy, _, _ = schedule.tile_grid(H, W, True)
    _coverage(s, gx, gy, hd, wd)
