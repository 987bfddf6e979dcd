"""Restatement of pytorch3d.renderer.rasterize_meshes (naive CPU algorithm, from the published
pytorch3d source as remembered; pytorch3d is NOT available here => parity unpinned).

Conventions reproduced: NDC +X left / +Y up, the short image side spans [-1, 1], pixel (yi, xi)
samples NDC x = W/H - (2 xi + 1)/H, y = 1 - (2 yi + 1)/H when H < W (PixToNonSquareNdc with the
reversed index), edge-function barycentrics, perspective correction with the vertex z, a hit needs
all barycentrics > 0 and pz >= 0 (blur_radius == 0), the K nearest hits sorted by ascending z,
-1 / -1 / -1 padding.
"""
import torch
from .mesh.rasterizer import Fragments  # noqa: F401


class RasterizationSettings:
    def __init__(self, image_size=256, blur_radius=0.0, faces_per_pixel=1, bin_size=None,
                 max_faces_per_bin=None, perspective_correct=None, clip_barycentric_coords=None,
                 cull_backfaces=False, z_clip_value=None, cull_to_frustum=False):
        self.image_size = image_size
        self.blur_radius = blur_radius
        self.faces_per_pixel = faces_per_pixel
        self.bin_size = bin_size
        self.max_faces_per_bin = max_faces_per_bin
        self.perspective_correct = perspective_correct
        self.clip_barycentric_coords = clip_barycentric_coords
        self.cull_backfaces = cull_backfaces
        self.z_clip_value = z_clip_value
        self.cull_to_frustum = cull_to_frustum


def _pix_to_ndc(i, s1, s2):
    rng = 2.0 * s1 / s2 if s1 > s2 else 2.0
    off = rng / 2.0
    return -off + (rng * i + off) / s1


def rasterize_meshes(meshes, image_size, blur_radius=0.0, faces_per_pixel=8, bin_size=None,
                     max_faces_per_bin=None, perspective_correct=False,
                     clip_barycentric_coords=False, cull_backfaces=False, z_clip_value=None,
                     cull_to_frustum=False, face_chunk=4096):
    assert blur_radius == 0.0 and not clip_barycentric_coords and not cull_backfaces
    verts, faces = meshes.verts, meshes.faces
    assert verts.shape[0] == 1 and faces.shape[0] == 1
    v = verts[0].float()
    f = faces[0].long()
    H, W = image_size
    K = faces_per_pixel
    dev = v.device
    yi = torch.arange(H, device=dev, dtype=torch.float32)
    xi = torch.arange(W, device=dev, dtype=torch.float32)
    yf = _pix_to_ndc((H - 1) - yi, H, W)
    xf = _pix_to_ndc((W - 1) - xi, W, H)
    px = xf[None, :].expand(H, W).reshape(-1, 1)  # P,1
    py = yf[:, None].expand(H, W).reshape(-1, 1)

    best_z = torch.full((H * W, K), float("inf"), device=dev)
    best_f = torch.full((H * W, K), -1, dtype=torch.long, device=dev)
    best_b = torch.full((H * W, K, 3), -1.0, device=dev)

    for s in range(0, f.shape[0], face_chunk):
        fc = f[s:s + face_chunk]
        v0, v1, v2 = v[fc[:, 0]], v[fc[:, 1]], v[fc[:, 2]]

        def edge(px_, py_, a, b):  # EdgeFunctionForward(p, a, b)
            return (px_ - a[:, 0]) * (b[:, 1] - a[:, 1]) - (py_ - a[:, 1]) * (b[:, 0] - a[:, 0])

        area = (v2[:, 0] - v0[:, 0]) * (v1[:, 1] - v0[:, 1]) - (v2[:, 1] - v0[:, 1]) * (v1[:, 0] - v0[:, 0])
        area = area + 1e-8  # pytorch3d adds kEpsilon to the denominator
        w0 = edge(px, py, v1, v2) / area
        w1 = edge(px, py, v2, v0) / area
        w2 = edge(px, py, v0, v1) / area
        if perspective_correct:
            z0, z1, z2 = v0[:, 2], v1[:, 2], v2[:, 2]
            t0, t1, t2 = w0 * z1 * z2, z0 * w1 * z2, z0 * z1 * w2
            den = (t0 + t1 + t2).clamp_min(1e-8)
            b0, b1, b2 = t0 / den, t1 / den, t2 / den
        else:
            b0, b1, b2 = w0, w1, w2
        pz = b0 * v0[:, 2] + b1 * v1[:, 2] + b2 * v2[:, 2]
        inside = (b0 > 0) & (b1 > 0) & (b2 > 0) & (pz >= 0) & (area.abs() > 1e-8)
        zc = torch.where(inside, pz, torch.full_like(pz, float("inf")))
        k = min(K, zc.shape[1])
        zt, it = torch.topk(zc, k, dim=1, largest=False)
        ft = torch.where(torch.isinf(zt), torch.full_like(it, -1), it + s)
        bt = torch.stack([torch.gather(b0, 1, it), torch.gather(b1, 1, it), torch.gather(b2, 1, it)], -1)
        bt = torch.where(torch.isinf(zt)[..., None], torch.full_like(bt, -1.0), bt)
        # merge with running best
        allz = torch.cat([best_z, zt], 1)
        allf = torch.cat([best_f, ft], 1)
        allb = torch.cat([best_b, bt], 1)
        zs, order = torch.sort(allz, dim=1, stable=True)
        order = order[:, :K]
        best_z = zs[:, :K]
        best_f = torch.gather(allf, 1, order)
        best_b = torch.gather(allb, 1, order[..., None].expand(-1, -1, 3))

    zbuf = torch.where(torch.isinf(best_z), torch.full_like(best_z, -1.0), best_z)
    dists = torch.where(best_f >= 0, torch.zeros_like(zbuf), torch.full_like(zbuf, -1.0))
    shp = (1, H, W, K)
    return best_f.reshape(shp), zbuf.reshape(shp), best_b.reshape(1, H, W, K, 3), dists.reshape(shp)


# names imported (unused) at MPV.py:14-22 / MPI.py:10-19
def look_at_view_transform(*a, **k):
    raise NotImplementedError


class FoVPerspectiveCameras:  # noqa: D401
    pass


class PerspectiveCameras:
    pass


class TexturesUV:
    pass


class Textures:
    pass
