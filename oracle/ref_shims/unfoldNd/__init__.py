"""Restatement of the two unfoldNd classes the reference uses (utils_vid.py:3,66,218).

UnfoldNd(x:(N,C,t,h,w)) -> (N, C*kt*kh*kw, L): row = c*(kt*kh*kw) + (it*kh + ih)*kw + iw,
L row-major over output positions (t',h',w').  FoldNd is the adjoint (sum of overlaps).
No padding / dilation (the reference never sets them).
"""
import torch


def _triple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


class UnfoldNd(torch.nn.Module):
    def __init__(self, kernel_size, dilation=1, padding=0, stride=1):
        super().__init__()
        assert dilation == 1 and padding == 0
        self.k, self.s = _triple(kernel_size), _triple(stride)

    def forward(self, x):
        n, c = x.shape[:2]
        (kt, kh, kw), (st, sh, sw) = self.k, self.s
        p = x.unfold(2, kt, st).unfold(3, kh, sh).unfold(4, kw, sw)  # n,c,t',h',w',kt,kh,kw
        p = p.permute(0, 1, 5, 6, 7, 2, 3, 4)
        return p.reshape(n, c * kt * kh * kw, -1)


class FoldNd(torch.nn.Module):
    def __init__(self, output_size, kernel_size, dilation=1, padding=0, stride=1):
        super().__init__()
        assert dilation == 1 and padding == 0
        self.o, self.k, self.s = tuple(output_size), _triple(kernel_size), _triple(stride)

    def forward(self, z):
        n = z.shape[0]
        (kt, kh, kw), (st, sh, sw), (t, h, w) = self.k, self.s, self.o
        c = z.shape[1] // (kt * kh * kw)
        idx = torch.arange(t * h * w, device=z.device).reshape(1, 1, t, h, w).float()
        idx = UnfoldNd(self.k, stride=self.s)(idx).long()[0]  # (kt*kh*kw, L)
        out = torch.zeros(n, c, t * h * w, dtype=z.dtype, device=z.device)
        z = z.reshape(n, c, kt * kh * kw * idx.shape[1])
        out.index_add_(2, idx.reshape(-1), z)
        return out.reshape(n, c, t, h, w)
