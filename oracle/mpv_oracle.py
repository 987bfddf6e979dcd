"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the reference's MPV
render / forward path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module.  The product (`videoloop3d_b200/`) never does and fails loudly without its
CUDA library.

Pinned against: golden vectors in `tests/golden/*.npz`, produced by running the UNMODIFIED
reference (`/root/reference/MPV.py`, `utils_mpi.py`, `utils_vid.py`) behind `oracle/ref_shims`
(script: `oracle/make_golden.py`).  Caveat ("parity unpinned" boundary): the reference rasterises
with pytorch3d, which is neither vendored in `/root/reference` nor installable here; the golden
vectors therefore used `oracle/ref_shims/pytorch3d` (a naive edge-function rasteriser restated
from pytorch3d's published algorithm).  Everything downstream of the rasteriser is verbatim
reference code.

What is restated (reference file:line):
  * geometry      MPV.py:351-405  (NDC transform + rasterize_meshes + get_uvs)  -> analytic
                  ray/plane intersection, quad lookup through faces[::2,0], barycentric uv
  * sampling      MPV.py:413-436  (grid_sample align_corners=True zeros + sigmoid, MPI.py:21-31)
  * slots         utils.py:64-69, MPV.py:441-449 (hits compacted front-to-back into K slots)
  * compositing   utils_mpi.py:92-107 (overcompose), MPV.py:454 (alpha)
  * forward       MPV.py:477-556 (loop pad, scale-invariant gain, loss call, smoothness terms)
  * step          train_3dvid.py:214-244 (loss sum, backward, Adam eps=6e-8 MPV.py:213)
  * optional      MPV.py:455-466 (background blend, disparity), 511-515 / 533-551 (sparsity, density, d_smooth) — pinned
                  to `step_dense_terms` / `step_sparse_terms` / `render_bg_random`
  * stage 1       MPI.py:452-652 (`MPMesh.render` / `forward`: loop-mask label, normalised disparity, six extra terms) —
                  pinned to `stage1_loopmask` / `stage1_bg_normdepth`, and on a culled state to `stage1_sparsify`
All maths is done with torch on the CPU in the dtype of `dtype` (float64 default for geometry).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from . import looploss_oracle as LL


# --------------------------------------------------------------------------------------
# model state (plain tensors; mirrors the reference's parameter / buffer names)
# --------------------------------------------------------------------------------------
@dataclass
class MPVState:
    """The tensors of a reference `MPMeshVid` (MPV.py:47-104), held as plain CPU tensors."""
    verts: torch.Tensor        # (D*hv*wv, 3)  `_verts`
    planedepth: torch.Tensor   # (D,)
    faces: torch.Tensor        # (2*Ns, 3) static
    faces_dyn: torch.Tensor    # (2*Nd, 3) dynamic
    uvs: torch.Tensor          # (., 2) in [-1, 1]
    uvs_dyn: torch.Tensor
    uvfaces: torch.Tensor      # (2*Ns, 3)
    uvfaces_dyn: torch.Tensor
    atlas: torch.Tensor        # (1, 4, Hs, Ws)
    atlas_dyn: torch.Tensor    # (T, 4, Hd, Wd)
    ref_extrin: torch.Tensor   # (4, 4)
    ref_intrin: torch.Tensor   # (3, 3)
    mpi_d: int
    hv: int
    wv: int

    @staticmethod
    def from_state_dict(sd, mpi_d, hv, wv):
        g = lambda k: torch.as_tensor(np.asarray(sd[k])) if not torch.is_tensor(sd[k]) else sd[k].detach().cpu()
        return MPVState(verts=g("_verts"), planedepth=g("planedepth"), faces=g("faces").long(),
                        faces_dyn=g("faces_dyn").long(), uvs=g("uvs"), uvs_dyn=g("uvs_dyn"),
                        uvfaces=g("uvfaces").long(), uvfaces_dyn=g("uvfaces_dyn").long(),
                        atlas=g("atlas"), atlas_dyn=g("atlas_dyn"), ref_extrin=g("ref_extrin"),
                        ref_intrin=g("ref_intrin"), mpi_d=int(mpi_d), hv=int(hv), wv=int(wv))


def make_depths(num_plane, min_depth, max_depth):
    """utils_mpi.py:210-211."""
    return torch.reciprocal(torch.linspace(1. / max_depth, 1. / min_depth, num_plane, dtype=torch.float32))


def dense_state(H, W, D, hv, wv, grid_h, T, near, far, h_scale=1.0, w_scale=1.0, ref_intrin=None,
                ref_extrin=None, seed=0, alpha_mean=-1.0):
    """Dense (un-culled) model exactly as MPMeshVid.__init__ lays it out (MPV.py:27-104)."""
    g = torch.Generator().manual_seed(seed)
    mpi_h, mpi_w = int(h_scale * H), int(w_scale * W)
    grid_w = D // grid_h
    assert grid_h * grid_w == D
    if ref_intrin is None:
        f = 0.8 * W
        ref_intrin = torch.tensor([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.]])
    if ref_extrin is None:
        ref_extrin = torch.eye(4)
    planedepth = make_depths(D, near, far).float().flip(0)                      # MPV.py:51
    H_start, W_start = (mpi_h - H) // 2, (mpi_w - W) // 2                       # MPV.py:55
    intr = ref_intrin.clone().float()
    intr[0, 2] += W_start                                                      # utils.py:196-200 with -start
    intr[1, 2] += H_start
    # utils_mpi.py:80-89
    ys, xs = torch.meshgrid(torch.linspace(0, mpi_h - 1, hv), torch.linspace(0, mpi_w - 1, wv), indexing="ij")
    v2 = torch.stack([xs, ys], -1).reshape(1, -1, 2)
    v2 = (v2 - intr[None, None, :2, 2]) * planedepth[:, None, None]
    v2 = v2 / intr[None, None, [0, 1], [0, 1]]
    zs = planedepth[:, None, None].expand_as(v2[..., :1])
    verts = torch.cat([v2.reshape(-1, 2), zs.reshape(-1, 1)], -1)
    vi = torch.arange(len(verts)).reshape(D, hv, wv)                            # MPV.py:66-71
    f013 = torch.stack([vi[:, :-1, :-1], vi[:, :-1, 1:], vi[:, 1:, 1:]], -1)
    f320 = torch.stack([vi[:, 1:, 1:], vi[:, 1:, :-1], vi[:, :-1, :-1]], -1)
    faces = torch.cat([f013.reshape(-1, 1, 3), f320.reshape(-1, 1, 3)], 1).reshape(-1, 3)
    gy, gx = torch.meshgrid(torch.arange(grid_h) / grid_h, torch.arange(grid_w) / grid_w, indexing="ij")
    uvs_plane = torch.stack([gx, gy], -1) * 2 - 1                               # MPV.py:75-81
    vsz = (-uvs_plane[-1, -1] + 1).reshape(1, 1, 2)
    vy, vx = torch.meshgrid(torch.linspace(0, 1, hv), torch.linspace(0, 1, wv), indexing="ij")
    uvs_voxel = torch.stack([vx, vy], -1).reshape(1, -1, 2) * vsz
    uvs = (uvs_plane.reshape(-1, 1, 2) + uvs_voxel.reshape(1, -1, 2)).reshape(-1, 2)
    Hd, Wd = int(grid_h * mpi_h), int(grid_w * mpi_w)
    atlas_dyn = torch.randn((T, 4, Hd, Wd), generator=g)
    atlas_dyn[:, 3] += alpha_mean
    atlas = torch.rand((1, 4, Hd, Wd), generator=g)
    return MPVState(verts=verts, planedepth=planedepth, faces=faces[:0].clone(), faces_dyn=faces,
                    uvs=uvs[:0].clone(), uvs_dyn=uvs, uvfaces=faces[:0].clone(), uvfaces_dyn=faces.clone(),
                    atlas=atlas, atlas_dyn=atlas_dyn, ref_extrin=ref_extrin.float(), ref_intrin=ref_intrin.float(),
                    mpi_d=D, hv=hv, wv=wv)


def sparse_state(H, W, D, hv, wv, T, near, far, tile, occupancy=0.5, dyn_frac=0.5, h_scale=1.0, w_scale=1.0,
                 seed=0, alpha_mean=-1.0):
    """Tile-culled model with the layout `MPI.sparsify_faces` produces (MPI.py:367-439):
    kept quads own 4 private uv vertices, tiles of `tile`x`tile` texels packed row-major on an
    n_h x n_w grid (last tile repeated as filler), faces still index the shared vertex grid."""
    st = dense_state(H, W, D, hv, wv, 1, T, near, far, h_scale, w_scale, seed=seed)
    # dense_state needs grid_h | D; grid_h = 1 always divides.
    g = torch.Generator().manual_seed(seed + 1)
    nq = D * (hv - 1) * (wv - 1)
    keep = torch.rand(nq, generator=g) < occupancy
    dyn = torch.rand(nq, generator=g) < dyn_frac
    quads = st.faces_dyn.reshape(-1, 2, 3)

    def pack(mask, frames):
        n = int(mask.sum())
        if n == 0:
            return (torch.zeros(frames, 4, 1, 1), torch.zeros(0, 2), torch.zeros(0, 3, dtype=torch.long),
                    torch.zeros(0, 3, dtype=torch.long))
        n_min, n_max = int(np.sqrt(n / 4)), int(np.sqrt(n))                  # MPI.py:369-381
        n_try = np.arange(max(n_min, 1), max(n_max, 2))
        sel = np.argmin(n_try - n % n_try)
        nh = int(n_try[sel]); nw = n // nh + 1
        ah, aw = nh * tile, nw * tile
        atlas = torch.randn((frames, 4, ah, aw), generator=g)
        atlas[:, 3] += alpha_mean
        qh, qw = 2 / (ah - 1) * (tile - 1), 2 / (aw - 1) * (tile - 1)         # MPI.py:403-418
        off = torch.tensor([[0, 0], [qw, 0], [0, qh], [qw, qh]])
        uy, ux = torch.meshgrid(torch.arange(0, ah, tile) / (ah - 1) * 2 - 1,
                                torch.arange(0, aw, tile) / (aw - 1) * 2 - 1, indexing="ij")
        uv0 = torch.stack([ux, uy], -1).reshape(-1, 1, 2)
        quv = (uv0 + off[None]).reshape(-1, 4, 2)[:n]
        uvid = torch.arange(n)[:, None, None] * 4 + torch.tensor([[0, 1, 3], [3, 2, 0]])[None]
        return atlas, quv.reshape(-1, 2).float(), uvid.reshape(-1, 3).long(), quads[mask].reshape(-1, 3).long()

    atlas, uvs, uvfaces, faces = pack(keep & ~dyn, 1)
    atlas_dyn, uvs_dyn, uvfaces_dyn, faces_dyn = pack(keep & dyn, T)
    st.atlas, st.uvs, st.uvfaces, st.faces = atlas, uvs, uvfaces, faces
    st.atlas_dyn, st.uvs_dyn, st.uvfaces_dyn, st.faces_dyn = atlas_dyn, uvs_dyn, uvfaces_dyn, faces_dyn
    return st


# --------------------------------------------------------------------------------------
# geometry: analytic restatement of MPV.py:353-405 (+ pytorch3d raster contract)
# --------------------------------------------------------------------------------------
def _quad_lookup(st: MPVState):
    """quad (d,qy,qx) -> (kind, local quad id).  kind 0 none, 1 static, 2 dynamic.
    `faces[2i, 0]` is the v00 vertex of kept quad i (MPV.py:68-71, MPI.py:385-399)."""
    D, hv, wv = st.mpi_d, st.hv, st.wv
    kind = torch.zeros(D, hv - 1, wv - 1, dtype=torch.long)
    local = torch.full((D, hv - 1, wv - 1), -1, dtype=torch.long)
    for k, fs in ((1, st.faces), (2, st.faces_dyn)):
        if len(fs) == 0:
            continue
        v00 = fs[0::2, 0]
        d = v00 // (hv * wv)
        r = (v00 % (hv * wv)) // wv
        c = v00 % wv
        kind[d, r, c] = k
        local[d, r, c] = torch.arange(len(v00))
    return kind, local


def geometry(st: MPVState, H, W, tar_extrin, tar_intrin, dtype=torch.float64):
    """Per (pixel, plane): hit?, kind, atlas pixel coordinates.

    Ray through image point (c+0.5, r+0.5) (the NDC "strange trick", MPV.py:357-371, maps pixel
    centres there), intersected with plane z_ref = planedepth[d]; a hit needs view depth
    > 0 and a point strictly inside a present quad (pytorch3d: all barycentrics > 0, pz >= 0).
    Returns dict of (H*W, D) tensors: hit(bool), kind, ax, ay (atlas pixel coords, align_corners),
    depth (view-space z) and the plane order flag.
    """
    D, hv, wv = st.mpi_d, st.hv, st.wv
    ext = (tar_extrin.to(dtype).reshape(4, 4) @ torch.inverse(st.ref_extrin.to(dtype)))    # MPV.py:478
    R, Tt = ext[:3, :3], ext[:3, 3]
    Kinv = torch.inverse(tar_intrin.to(dtype).reshape(3, 3))
    r, c = torch.meshgrid(torch.arange(H, dtype=dtype), torch.arange(W, dtype=dtype), indexing="ij")
    pix = torch.stack([c + 0.5, r + 0.5, torch.ones_like(c)], -1).reshape(-1, 3)            # (P,3)
    dirs = (R.T @ (Kinv @ pix.T)).T                                                          # ref-frame dir, view z = 1
    org = -(R.T @ Tt)
    verts = st.verts.to(dtype).reshape(D, hv, wv, 3)
    z = verts[:, 0, 0, 2]                                                                    # plane depths in ref frame
    s = (z[None, :] - org[2]) / dirs[:, 2:3]                                                 # (P,D) == view depth
    Px = org[0] + s * dirs[:, 0:1]
    Py = org[1] + s * dirs[:, 1:2]
    X0, X1 = verts[:, 0, 0, 0], verts[:, 0, -1, 0]
    Y0, Y1 = verts[:, 0, 0, 1], verts[:, -1, 0, 1]
    gx = (Px - X0[None]) / (X1 - X0)[None] * (wv - 1)
    gy = (Py - Y0[None]) / (Y1 - Y0)[None] * (hv - 1)
    inside = (s > 0) & (gx > 0) & (gx < wv - 1) & (gy > 0) & (gy < hv - 1)
    qx = gx.floor().clamp(0, wv - 2).long()
    qy = gy.floor().clamp(0, hv - 2).long()
    a = gx - qx
    b = gy - qy
    kind_t, local_t = _quad_lookup(st)
    dd = torch.arange(D)[None, :].expand_as(qx)
    kind = kind_t[dd, qy, qx]
    local = local_t[dd, qy, qx]
    hit = inside & (kind > 0)
    kind = torch.where(hit, kind, torch.zeros_like(kind))

    ax = torch.zeros_like(gx)
    ay = torch.zeros_like(gy)
    for k, uvs, uvfaces, atl in ((1, st.uvs, st.uvfaces, st.atlas), (2, st.uvs_dyn, st.uvfaces_dyn, st.atlas_dyn)):
        m = kind == k
        if not m.any():
            continue
        li = local[m]
        am, bm = a[m], b[m]
        tri1 = am > bm                      # face 2i = (v0, v1, v3), face 2i+1 = (v3, v2, v0)  (MPV.py:68-71)
        fidx = torch.where(tri1, 2 * li, 2 * li + 1)
        uvt = uvs.to(dtype)[uvfaces[fidx]]  # (N,3,2)                                        (MPV.py:394-400)
        w = torch.where(tri1[:, None], torch.stack([1 - am, am - bm, bm], -1),
                        torch.stack([am, bm - am, 1 - bm], -1))
        uv = (w[..., None] * uvt).sum(1)
        hA, wA = atl.shape[-2:]
        ax[m] = (uv[:, 0] + 1) / 2 * (wA - 1)   # grid_sample align_corners=True                (MPV.py:425-427)
        ay[m] = (uv[:, 1] + 1) / 2 * (hA - 1)
    return dict(hit=hit, kind=kind, ax=ax, ay=ay, depth=s, forward_order=bool((dirs[:, 2] > 0).all()))


def _bilinear_zeros(atlas, ax, ay):
    """atlas (B,4,h,w); ax,ay (N,) pixel coords -> (B,N,4); grid_sample bilinear / zeros padding."""
    Bn, C, h, w = atlas.shape
    x0 = ax.floor(); y0 = ay.floor()
    fx = (ax - x0).to(atlas.dtype); fy = (ay - y0).to(atlas.dtype)
    x0 = x0.long(); y0 = y0.long()
    out = torch.zeros(Bn, ax.numel(), C, dtype=atlas.dtype)
    flat = atlas.permute(0, 2, 3, 1).reshape(Bn, h * w, C)
    for dx, dy, wgt in ((0, 0, (1 - fx) * (1 - fy)), (1, 0, fx * (1 - fy)), (0, 1, (1 - fx) * fy), (1, 1, fx * fy)):
        xx, yy = x0 + dx, y0 + dy
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        idx = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1))
        val = flat[:, idx, :]
        out = out + val * (wgt * ok.to(atlas.dtype))[None, :, None]
    return out


def parse_bg_color(bg_color):
    """'r#g#b' -> tensor (MPV.py:455-460); 'random' is drawn by the caller (it consumes torch's CPU generator)."""
    return torch.tensor([float(v) for v in bg_color.split('#')], dtype=torch.float64)


def render(st: MPVState, H, W, tar_extrin, tar_intrin, ts, dtype=torch.float64, atlas=None, atlas_dyn=None,
           geo=None, bg_color=None):
    """Restates MPMeshVid.render (MPV.py:351-475).  Returns rgb (T,H,W,3) and variables with
    slot-indexed `mpi` (T,H,W,K,4), `blend_weight`, `alpha`, `disp_norm` (MPV.py:385,463-464: sum of
    blend weight / view depth), `K`.  `bg_color`: optional (3,) tensor (MPV.py:455-461).
    `atlas`/`atlas_dyn` may be autograd leaves (dtype `dtype`)."""
    if geo is None:
        geo = geometry(st, H, W, tar_extrin, tar_intrin)
    atlas = st.atlas.to(dtype) if atlas is None else atlas
    atlas_dyn = st.atlas_dyn.to(dtype) if atlas_dyn is None else atlas_dyn
    ts = torch.as_tensor(np.asarray(ts)).long()
    T = len(ts)
    D = st.mpi_d
    P = H * W
    hit, kind = geo["hit"], geo["kind"]
    order = torch.arange(D) if geo["forward_order"] else torch.arange(D - 1, -1, -1)
    hit_o, kind_o = hit[:, order], kind[:, order]
    slot = torch.cumsum(hit_o.long(), 1) - 1                                   # k-th hit along the ray (utils.py:64-69)
    K = int(hit_o.sum(1).max().item()) if P > 0 else 0
    mpi = torch.zeros(T, P, max(K, 1), 4, dtype=dtype)
    pidx = torch.arange(P)[:, None].expand(P, D)
    for k, atl in ((1, atlas), (2, atlas_dyn[ts])):
        m = kind_o == k
        if not m.any():
            continue
        ax = geo["ax"][:, order][m]
        ay = geo["ay"][:, order][m]
        samp = torch.sigmoid(_bilinear_zeros(atl, ax, ay))                      # MPV.py:434-435, MPI.py:22
        if samp.shape[0] == 1:
            samp = samp.expand(T, -1, -1)                                       # MPV.py:445
        mpi = mpi.index_put((torch.arange(T)[:, None], pidx[m][None, :], slot[m][None, :]), samp)
    mpi = mpi[:, :, :K].reshape(T, H, W, K, 4)
    alpha, content = mpi[..., -1], mpi[..., :-1]
    if K > 0:
        bw = torch.cumprod((1 - alpha)[..., :-1], -1)                           # utils_mpi.py:100-104
        bw = torch.cat([alpha[..., :1], alpha[..., 1:] * bw], -1)
    else:
        bw = alpha
    rgb = (content * bw.unsqueeze(-1)).sum(-2)                                  # utils_mpi.py:106
    alpha_sum = bw.sum(-1)                                                      # MPV.py:454
    if bg_color is not None:                                                    # MPV.py:455-461
        bg = bg_color.to(dtype)
        rgb = rgb * alpha_sum[..., None] + bg[None, None, None] * (-alpha_sum[..., None] + 1)
    # depths = 1 / zbuf (MPV.py:384-385), compacted into the same slots; empty slots carry weight 0
    inv_d = torch.zeros(P, max(K, 1), dtype=dtype)
    depth_o = geo["depth"][:, order].to(dtype)
    inv_d = inv_d.index_put((pidx[hit_o], slot[hit_o]), 1.0 / depth_o[hit_o])
    disp = (inv_d[:, :K].reshape(1, H, W, K) * bw).sum(-1)                      # MPV.py:463-464
    return rgb, dict(mpi=mpi, blend_weight=bw, alpha=alpha_sum, disp_norm=disp, K=K, hitmask=hit)


def forward_train(st: MPVState, h, w, tar_extrin, tar_intrin, res, losscfg, *, isloop=True, scale_invariant=True,
                  swd_patcht_size=3, rgb_smooth=True, a_smooth=True, dtype=torch.float64, atlas=None,
                  atlas_dyn=None, nn_mode="exact64", ts=None, sparsity=False, density=False, d_smooth=False,
                  bg_color=None):
    """Restates MPMeshVid.forward, training branch (MPV.py:477-553).  `losscfg` is the un-batched
    dict (loss_name, loss_gain, patch_size, ..).  Returns dict of 0-dim tensors + aux."""
    T_all = (st.atlas_dyn if atlas_dyn is None else atlas_dyn).shape[0]
    ts = torch.arange(T_all) if ts is None else ts
    rgb, var = render(st, h, w, tar_extrin, tar_intrin, ts, dtype, atlas, atlas_dyn, bg_color=bg_color)
    rgb = rgb.permute(0, 3, 1, 2)                                               # (T,3,h,w)  MPV.py:484
    cfg = dict(losscfg)
    loss_name = cfg.pop("loss_name")
    gain = float(cfg.pop("loss_gain", 1.0))
    rgb_pad = rgb
    if isloop:                                                                  # MPV.py:490-492
        rgb_pad = torch.cat([rgb, rgb[:swd_patcht_size - 1]], 0)
    res = res.to(dtype)
    scale = None
    if scale_invariant:                                                         # MPV.py:499-504
        res_avg = res[0].mean(0)
        rgb_avg = rgb.detach().mean(0)
        scale = torch.exp(torch.log((res_avg + 0.01) / (rgb_avg + 0.01)).mean())
        scale = (scale + 3) / 4
        rgb_pad = rgb_pad * scale
    x = rgb_pad.permute(1, 0, 2, 3)[None]
    y = res.permute(0, 2, 1, 3, 4)
    loss, aux = LL.LOSSES[loss_name](x, y, nn_mode=nn_mode, **cfg)
    out = {"swd": loss * gain}
    mpi, K, D = var["mpi"], var["K"], st.mpi_d
    if rgb_smooth:                                                              # MPV.py:517-523
        sm = mpi[..., :-1]
        out["rgb_smooth"] = ((sm[:, :, :-1] - sm[:, :, 1:]).abs().mean() +
                             (sm[:, :-1] - sm[:, 1:]).abs().mean()) * (gain * K / D)
    if a_smooth:                                                                # MPV.py:525-531
        sm = mpi[..., -1]
        out["a_smooth"] = ((sm[:, :, :-1] - sm[:, :, 1:]).abs().mean() +
                           (sm[:, :-1] - sm[:, 1:]).abs().mean()) * (gain * K / D)
    if sparsity:                                                                # MPV.py:511-515
        al = mpi[..., -1]
        sp = al.norm(dim=-1, p=1) / al.norm(dim=-1, p=2).clamp_min(1e-4)
        out["sparsity"] = sp.mean() / math.sqrt(D) * gain
    if density:                                                                 # MPV.py:533-536
        out["density"] = (var["alpha"] - 1).abs().mean()
    if d_smooth:                                                                # MPV.py:538-551
        disp = var["disp_norm"]
        gx = (disp[:, 1:, :-1] - disp[:, 1:, 1:]).abs()
        gy = (disp[:, :-1, 1:] - disp[:, 1:, 1:]).abs()
        out["d_smooth"] = (gx + gy).mean()
    return out, dict(rgb=rgb, scale=scale, K=K, **aux)


# --------------------------------------------------------------------------------------
# stage 1: MPMesh.render / forward (MPI.py:452-652) — SURVEY §8(f) N4, second half
# --------------------------------------------------------------------------------------
def stage1_state(H, W, D, hv, wv, grid_h, near, far, h_scale=1.0, w_scale=1.0, seed=0, alpha_mean=-1.0):
    """The dense stage-1 model as MPMesh.__init__ lays it out (MPI.py:38-124): the same vertex grid / faces / uvs as the
    stage-2 model, but every quad STATIC (one atlas, no dynamic part) plus a one-channel loop-mask atlas."""
    st = dense_state(H, W, D, hv, wv, grid_h, 1, near, far, h_scale, w_scale, seed=seed, alpha_mean=alpha_mean)
    st.faces, st.uvs, st.uvfaces, st.atlas = st.faces_dyn, st.uvs_dyn, st.uvfaces_dyn, st.atlas_dyn[:1].clone()
    st.faces_dyn, st.uvs_dyn, st.uvfaces_dyn = st.faces[:0].clone(), st.uvs[:0].clone(), st.uvfaces[:0].clone()
    st.atlas_dyn = torch.zeros(1, 4, 1, 1)
    g = torch.Generator().manual_seed(seed + 11)
    atlas_mask = torch.randn((1, 1) + tuple(st.atlas.shape[-2:]), generator=g)
    return st, atlas_mask


def render_stage1(st: MPVState, H, W, tar_extrin, tar_intrin, near, far, dtype=torch.float64, atlas=None, atlas_dyn=None,
                  atlas_mask=None, bg_color=None, normalize_blendweight_fordepth=False):
    """Restates MPMesh.render (MPI.py:452-594) for rgb_mlp_type='direct', one view (the reference's own batching of views
    does not run: MPI.py:472 broadcasts (B,3,3) against (1,N,3,1)).  Returns rgbl (1,H,W,3|4) and variables with `mpi`,
    `blend_weight`, `alpha`, `disp_norm` (normalised disparity, MPI.py:552-553,566), `loopmask3d` (MPI.py:568-580)."""
    atlas = st.atlas.to(dtype) if atlas is None else atlas
    geo = geometry(st, H, W, tar_extrin, tar_intrin)
    rgb, var = render(st, H, W, tar_extrin, tar_intrin, [0], dtype, atlas, atlas_dyn, geo=geo, bg_color=bg_color)
    bw, alpha = var["blend_weight"], var["alpha"]
    span = 1.0 / near - 1.0 / far
    if normalize_blendweight_fordepth:                                          # MPI.py:564-566
        bwn = bw / alpha.clamp_min(1e-10)[..., None]
        disp = (_slot_inv_depth(st, geo, H, W, var["K"], dtype) * bwn).sum(-1)
        disp = (disp - bwn.sum(-1) / far) / span
        var["blend_weight"] = bwn
    else:
        disp = (var["disp_norm"] - alpha / far) / span                          # sum_k bw_k (1/z_k - 1/far) / span
    var["disp_norm"] = disp
    var["loopmask3d"] = None
    rgbl = rgb
    if atlas_mask is not None:                                                  # MPI.py:568-580
        assert len(st.faces_dyn) == 0
        fake = torch.cat([atlas_mask.to(dtype).expand(-1, 3, -1, -1), atlas[:, 3:4].detach()], 1)
        lab, var_l = render(st, H, W, tar_extrin, tar_intrin, [0], dtype, fake, None, geo=geo)
        var["loopmask3d"] = var_l["mpi"][..., :1]
        rgbl = torch.cat([rgb, lab[..., :1]], -1)
    return rgbl, var


def _slot_inv_depth(st, geo, H, W, K, dtype):
    """1 / view depth of each ray's hits, compacted into its K slots (0 for empty slots)."""
    D, P = st.mpi_d, H * W
    order = torch.arange(D) if geo["forward_order"] else torch.arange(D - 1, -1, -1)
    hit_o = geo["hit"][:, order]
    slot = torch.cumsum(hit_o.long(), 1) - 1
    pidx = torch.arange(P)[:, None].expand(P, D)
    inv_d = torch.zeros(P, max(K, 1), dtype=dtype)
    inv_d = inv_d.index_put((pidx[hit_o], slot[hit_o]), 1.0 / geo["depth"][:, order].to(dtype)[hit_o])
    return inv_d[:, :K].reshape(1, H, W, K)


def forward_stage1(st: MPVState, h, w, tar_extrin, tar_intrin, near, far, *, sparsity=True, rgb_smooth=True, a_smooth=True,
                   d_smooth=True, l_smooth=True, density=True, edge_scale=4.0, **render_kw):
    """Restates MPMesh.forward, training branch (MPI.py:596-652).  Returns rgbl (1,C,h,w) and the dict of extra terms."""
    rgbl, var = render_stage1(st, h, w, tar_extrin, tar_intrin, near, far, **render_kw)
    rgbl = rgbl.permute(0, 3, 1, 2)
    mpi, K, D = var["mpi"], var["K"], st.mpi_d
    out = {}
    if sparsity:                                                                # MPI.py:603-607 (1e-6, no gain)
        al = mpi[..., -1]
        out["sparsity"] = (al.norm(dim=-1, p=1) / al.norm(dim=-1, p=2).clamp_min(1e-6)).mean() / math.sqrt(D)
    sm_mean = lambda t: (t[:, :, :-1] - t[:, :, 1:]).abs().mean() + (t[:, :-1] - t[:, 1:]).abs().mean()
    if rgb_smooth:                                                              # MPI.py:609-615
        out["rgb_smooth"] = sm_mean(mpi[..., :-1]) * (K / D)
    if a_smooth:                                                                # MPI.py:617-623
        out["a_smooth"] = sm_mean(mpi[..., -1]) * (K / D)
    if d_smooth:                                                                # MPI.py:625-638 (edge-aware)
        disp = var["disp_norm"]
        dg = (disp[:, 1:, :-1] - disp[:, 1:, 1:]).abs() + (disp[:, :-1, 1:] - disp[:, 1:, 1:]).abs()
        rgb = rgbl[:, :3]
        edge = ((rgb[..., 1:, :-1] - rgb[..., 1:, 1:]).abs().sum(dim=1) + (rgb[..., :-1, 1:] - rgb[..., 1:, 1:]).abs().sum(dim=1))
        out["d_smooth"] = (dg * (-edge * edge_scale + 1).clamp_min(0)).mean()
    if l_smooth and var["loopmask3d"] is not None:                              # MPI.py:640-646
        out["l_smooth"] = sm_mean(var["loopmask3d"][..., 0]) * (K / D)
    if density:                                                                 # MPI.py:648-651
        out["density"] = (var["alpha"] - 1).abs().mean()
    return rgbl, out, var


def total_loss(extra, rgb_smooth_w=0.2, a_smooth_w=0.2, **weights):
    """train_3dvid.py:230-240.  `weights`: sparsity= / density= / d_smooth= loss weights of the optional terms."""
    for k, wgt in weights.items():
        if k in extra and wgt > 0:
            extra = dict(extra)
            extra["swd"] = extra["swd"] + extra[k] * wgt
    loss = extra["swd"]
    if "rgb_smooth" in extra and rgb_smooth_w > 0:
        loss = loss + extra["rgb_smooth"] * rgb_smooth_w
    if "a_smooth" in extra and a_smooth_w > 0:
        loss = loss + extra["a_smooth"] * a_smooth_w
    return loss


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=6e-8):
    """torch.optim.Adam single-tensor maths (MPV.py:213: betas=(0.9,0.999), eps=6e-8)."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


def get_lrate(lrate, lrate_decay, step, n_dataset=None):
    """MPV.py:220-229 + train_3dvid.py:281-284."""
    lr = lrate * (0.1 ** (step / (lrate_decay * 1000)))
    return lr / n_dataset if n_dataset else lr
