"""Work-item tables for the fused backward + Adam kernel (csrc/fused_bwd_adam.cu, `vl3d_fused_bwd_adam`).

The kernel is a persistent grid that pulls items from ONE ordered queue; an item is a screen tile of the backward
(BWD), Adam on a rectangle of texels (ADAM) or the zeroing of a rectangle of the gradient (ZERO), each with an
optional wait (a range of counters, each >= target) and an optional counter to bump when done.  An item may only
wait for items that precede it in the queue.  One table describes one round (= one chunk of 2 frames) and is replayed
for every chunk with per-chunk counters.

Two schedules:

* `generic_schedule` (any layout): the tiles of chunk c are interleaved with the Adam rectangles of chunk c-1, which
  wait for "all tiles of chunk c-1 done".  The gradient buffer is all-zero between steps (Adam writes the zeros back),
  so there is no separate fill; issue-bound tiles and DRAM-bound Adam overlap on every SM.
* `band_schedule` (dense layout, VL3D_VIEW_RECT_PLANES): tiles are walked in screen-row order; from the plane
  homographies the host derives, for every atlas row block, the first and the last tile row that touches it, and
  places its ZERO item shortly before the first and its ADAM item shortly after the last.  A band's gradient rows
  are then meant to live in L2 from zeroing to consumption (Adam reads them with ld.global.cg).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

ITEM_BWD, ITEM_ADAM, ITEM_ZERO = 0, 1, 2
FLAG_HAS_GRAD, FLAG_REZERO, FLAG_PREV_ROUND = 1, 2, 8
BX, BY, TF = 32, 8, 2                  # tile shape / frames per chunk of the kernels (composite_common.cuh)
# item columns (12 int32 = three int4 of the kernel)
C_TYPE, C_A, C_B, C_C, C_W0, C_WN, C_WT, C_SIG, C_V0, C_VN, C_VT, C_PLANE = range(12)


@dataclass
class Schedule:
    items: np.ndarray        # (n_items, 12) int32, see include/vl3d.h
    counter_init: np.ndarray  # (n_counters,) int32: initial counter values of every round
    extra_round: bool        # items of the last chunk run in one more round (generic schedule)
    kind: str
    stats: dict

    @property
    def n_items(self):
        return int(self.items.shape[0])

    @property
    def n_counters(self):
        return int(self.counter_init.shape[0])


def tile_grid(H, W, smooth=True):
    sx, sy = (BX - 1, BY - 1) if smooth else (BX, BY)
    return (W + sx - 1) // sx, (H + sy - 1) // sy, sx, sy


def _rows(n, kind, flags=0):
    it = np.zeros((n, 12), dtype=np.int64)
    it[:, C_TYPE] = kind | (flags << 4)
    it[:, C_W0] = -1
    it[:, C_SIG] = -1
    it[:, C_V0] = -1
    it[:, C_PLANE] = -1
    return it


def generic_schedule(H, W, dyn_h, dyn_w, smooth=True, seg_texels=32768, lead_tiles=888, cells=None):
    """Any layout.  Counter 0 = finished tiles of the chunk.  `cells` (owner mode, tiles.atlas_cells): the Adam
    rectangles are cut along the atlas cells and carry their cell's plane, so that the kernel can skip the texels a
    screen tile owns (and has already updated)."""
    gx, gy, _, _ = tile_grid(H, W, smooth)
    n_tiles = gx * gy
    kind = "generic" if cells is None else "own"
    if cells is None:
        cells = [(0, 0, dyn_w - 1, dyn_h - 1, -1)]
    recs = []
    for (x0, y0, x1, y1, plane) in cells:
        cw = x1 - x0 + 1
        rows_per = max(1, seg_texels // max(cw, 1))
        for r in range(y0, y1 + 1, rows_per):
            recs.append((r * dyn_w + x0, cw, min(rows_per, y1 + 1 - r), plane))
    recs = np.asarray(recs, dtype=np.int64).reshape(-1, 4)
    adam = _rows(len(recs), ITEM_ADAM, FLAG_HAS_GRAD | FLAG_REZERO | FLAG_PREV_ROUND)
    adam[:, C_A], adam[:, C_B], adam[:, C_C], adam[:, C_PLANE] = recs[:, 0], recs[:, 1], recs[:, 2], recs[:, 3]
    adam[:, C_W0], adam[:, C_WN], adam[:, C_WT] = 0, 1, n_tiles
    tiles = _rows(n_tiles, ITEM_BWD)
    tiles[:, C_A], tiles[:, C_B] = np.tile(np.arange(gx), gy), np.repeat(np.arange(gy), gx)
    tiles[:, C_SIG] = 0
    lead = min(lead_tiles, n_tiles // 4)
    # merge: no Adam among the first `lead` tiles (the previous chunk's last tiles are still running), then evenly
    pos = lead + (np.arange(len(adam)) + 0.5) * (n_tiles - lead) / max(len(adam), 1)
    order = np.argsort(np.concatenate([np.arange(n_tiles, dtype=np.float64), pos]), kind="stable")
    items = np.concatenate([tiles, adam])[order].astype(np.int32)
    return Schedule(items=items, counter_init=np.zeros(1, np.int32), extra_round=True, kind=kind,
                    stats=dict(tiles=n_tiles, adam=len(adam), zero=0))


def _plane_rects(table, D, qh, qw):
    """Per plane: atlas position of quad-grid coordinate (gx, gy) is (X0 + gx*sx, Y0 + gy*sy) (dense layout)."""
    t = table.reshape(D, qh, qw)
    X0 = t["x0i"][:, 0, 0].astype(np.float64) + t["x0f"][:, 0, 0]
    Y0 = t["y0i"][:, 0, 0].astype(np.float64) + t["y0f"][:, 0, 0]
    sx = t["sx"][:, 0, 0].astype(np.float64)
    sy = t["sy"][:, 0, 0].astype(np.float64)
    return X0, Y0, sx, sy


def band_schedule(view_homs, cx, cy, H, W, table, D, qh, qw, dyn_h, dyn_w, smooth=True, row_block=8, col_blocks=None,
                  zero_ahead=2, adam_lag=2, use_zero=True, margin=2, long_span=12):
    """Dense layout.  Counters of a round: [0, gy) finished tiles per tile row; (zero-ahead schedules) [gy, 2 gy)
    finished ZERO items whose first tile row is R, pre-biased so that every one of them is complete at `zmax`, and
    2 gy = finished ZERO items of the few long-lived rectangles (atlas rows shared by two planes: touched by the first
    and the last tile rows), which are zeroed at the start of the round.
    Queue order: tile row R at key R; ZERO of a rectangle `zero_ahead` tile rows before the first tile row touching it;
    Adam `adam_lag` tile rows after the last."""
    gx, gy, sx_t, sy_t = tile_grid(H, W, smooth)
    homs = np.asarray(view_homs, dtype=np.float64).reshape(D, 3, 3)
    X0, Y0, qsx, qsy = _plane_rects(table, D, qh, qw)
    # pixel rectangle of tile row R (threads outside the image replicate the border pixel)
    R = np.arange(gy)
    ytop = (R * sy_t).astype(np.float64)
    ybot = np.minimum(R * sy_t + BY - 1, H - 1).astype(np.float64)
    us = np.array([0.0, W - 1.0]) + 0.5 - cx
    lo = np.full((D, gy), np.inf)
    hi = np.full((D, gy), -np.inf)
    xlo = np.full((D, gy), np.inf)
    xhi = np.full((D, gy), -np.inf)
    behind = np.zeros((D, gy), dtype=bool)
    front = np.zeros((D, gy), dtype=bool)
    for yv in (ytop, ybot):
        v = yv + 0.5 - cy
        for u in us:
            w = homs[:, 2, 0, None] * u + homs[:, 2, 1, None] * v[None] + homs[:, 2, 2, None]
            ok = w > 1e-9
            ws = np.where(ok, w, 1.0)
            g_x = (homs[:, 0, 0, None] * u + homs[:, 0, 1, None] * v[None] + homs[:, 0, 2, None]) / ws
            g_y = (homs[:, 1, 0, None] * u + homs[:, 1, 1, None] * v[None] + homs[:, 1, 2, None]) / ws
            behind |= ~ok
            front |= ok
            lo = np.where(ok, np.minimum(lo, g_y), lo)
            hi = np.where(ok, np.maximum(hi, g_y), hi)
            xlo = np.where(ok, np.minimum(xlo, g_x), xlo)
            xhi = np.where(ok, np.maximum(xhi, g_x), xhi)
    mixed = behind & front                     # the plane's horizon crosses the tile row: assume it touches everything
    lo = np.where(mixed, 0.0, lo)
    hi = np.where(mixed, float(qh), hi)
    xlo = np.where(mixed, 0.0, xlo)
    xhi = np.where(mixed, float(qw), xhi)
    eps = 1e-2
    touches = front & (hi > -eps) & (lo < qh + eps) & (xhi > -eps) & (xlo < qw + eps)
    ly0 = Y0[:, None] + np.clip(lo, 0.0, qh) * qsy[:, None]
    ly1 = Y0[:, None] + np.clip(hi, 0.0, qh) * qsy[:, None]
    lx0 = X0[:, None] + np.clip(xlo, 0.0, qw) * qsx[:, None]
    lx1 = X0[:, None] + np.clip(xhi, 0.0, qw) * qsx[:, None]
    row_lo = np.clip(np.floor(ly0) - margin, 0, dyn_h - 1).astype(np.int64)      # taps: floor(y), floor(y) + 1
    row_hi = np.clip(np.floor(ly1) + 1 + margin, 0, dyn_h - 1).astype(np.int64)
    col_lo = np.clip(np.floor(lx0) - margin, 0, dyn_w - 1).astype(np.int64)
    col_hi = np.clip(np.floor(lx1) + 1 + margin, 0, dyn_w - 1).astype(np.int64)

    if col_blocks is None:
        pw = max(float(np.median(qsx * qw)), 8.0)
        col_blocks = max(1, int(round(dyn_w / pw)))
    cb_edges = [(dyn_w * i // col_blocks) // 8 * 8 for i in range(col_blocks)] + [dyn_w]
    n_rb = (dyn_h + row_block - 1) // row_block
    BIG = 1 << 30
    first = np.full((dyn_h, col_blocks), BIG, dtype=np.int64)       # per atlas row and column block
    last = np.full((dyn_h, col_blocks), -1, dtype=np.int64)
    ys = np.arange(dyn_h)
    for d in range(D):
        idx = np.nonzero(touches[d])[0]
        if len(idx) == 0:
            continue
        # widen to monotone interval ends so that "tile rows touching atlas row y" is one contiguous range
        lo_d = np.minimum.accumulate(row_lo[d, idx][::-1])[::-1]
        hi_d = np.maximum.accumulate(row_hi[d, idx])
        f_i = np.searchsorted(hi_d, ys, side="left")                # first tile row (index into idx) with hi >= y
        l_i = np.searchsorted(lo_d, ys, side="right") - 1           # last tile row with lo <= y
        hit = (f_i <= l_i) & (f_i < len(idx)) & (l_i >= 0)
        if not hit.any():
            continue
        fR = idx[np.clip(f_i, 0, len(idx) - 1)]
        lR = idx[np.clip(l_i, 0, len(idx) - 1)]
        c0, c1 = int(col_lo[d, idx].min()), int(col_hi[d, idx].max())
        for cb in range(col_blocks):
            if c1 < cb_edges[cb] or c0 >= cb_edges[cb + 1]:
                continue
            first[hit, cb] = np.minimum(first[hit, cb], fR[hit])
            last[hit, cb] = np.maximum(last[hit, cb], lR[hit])
    # row blocks
    pad = n_rb * row_block - dyn_h
    fb = np.pad(first, ((0, pad), (0, 0)), constant_values=BIG).reshape(n_rb, row_block, col_blocks).min(1)
    lb = np.pad(last, ((0, pad), (0, 0)), constant_values=-1).reshape(n_rb, row_block, col_blocks).max(1)

    # ---- items (vectorised: this runs once per view)
    rb_i, cb_i = np.meshgrid(np.arange(n_rb), np.arange(col_blocks), indexing="ij")
    rb_i, cb_i = rb_i.reshape(-1), cb_i.reshape(-1)
    edges = np.asarray(cb_edges)
    r0 = rb_i * row_block
    base = r0 * dyn_w + edges[cb_i]
    width = edges[cb_i + 1] - edges[cb_i]
    nr = np.minimum(row_block, dyn_h - r0)
    f, l = fb.reshape(-1), lb.reshape(-1)
    keep = width > 0
    rb_i, cb_i, base, width, nr, f, l = (v[keep] for v in (rb_i, cb_i, base, width, nr, f, l))
    touched = l >= 0
    n_untouched = int((~touched).sum())
    aligned = (base % 8 == 0) & (width % 8 == 0) & (dyn_w % 8 == 0)
    long_lived = touched & ((l - f) > long_span)

    tiles = _rows(gx * gy, ITEM_BWD)
    tiles[:, C_A], tiles[:, C_B] = np.tile(np.arange(gx), gy), np.repeat(np.arange(gy), gx)
    tiles[:, C_SIG] = tiles[:, C_B]
    tile_keys = tiles[:, C_B].astype(np.float64)

    adam = _rows(len(base), ITEM_ADAM)
    adam[:, C_A], adam[:, C_B], adam[:, C_C] = base, width, nr
    fl = np.where(touched, FLAG_HAS_GRAD | np.where(use_zero & aligned, 0, FLAG_REZERO), 0)
    adam[:, C_TYPE] = ITEM_ADAM | (fl << 4)
    adam[touched, C_W0] = f[touched]
    adam[touched, C_WN] = (l - f + 1)[touched]
    adam[touched, C_WT] = gx
    adam_keys = np.where(touched, l + adam_lag + 0.25, 0.0)
    adam_keys[~touched] = (np.arange(n_untouched) + 0.5) * gy / max(n_untouched, 1)     # g = 0: any time

    zb = gy                                                         # first ZERO counter
    init = np.zeros(zb, dtype=np.int64)
    parts, keys = [tiles, adam], [tile_keys, adam_keys]
    n_zero = 0
    if use_zero:
        zsel = touched & aligned
        zshort, zlong = zsel & ~long_lived, zsel & long_lived
        zero = _rows(int(zsel.sum()), ITEM_ZERO)
        zero[:, C_A], zero[:, C_B], zero[:, C_C] = base[zsel], width[zsel], nr[zsel]
        is_long = zlong[zsel]
        zero[:, C_SIG] = np.where(is_long, zb + gy, zb + f[zsel])
        zero_keys = np.where(is_long, -1e9, f[zsel] - zero_ahead - 0.5)
        per_first = np.bincount(f[zshort], minlength=gy)[:gy]
        zmax = int(per_first.max()) if per_first.size else 0
        n_long = int(zlong.sum())
        init = np.concatenate([init, zmax - per_first, [0]])        # pre-bias: every Z counter is complete at zmax
        # tile row R waits for the ZERO items of every rectangle it touches: first in [fmin(R), R]
        fmin = np.arange(gy)
        fs, ls = f[zshort], l[zshort]
        if fs.size:
            reach = np.full(gy + 1, gy, dtype=np.int64)             # reach[R] = min first among rectangles with last >= R
            np.minimum.at(reach, ls, fs)
            fmin = np.minimum(np.minimum.accumulate(reach[::-1])[::-1][:gy], np.arange(gy))
        tR = tiles[:, C_B]
        if zmax > 0:
            tiles[:, C_W0], tiles[:, C_WN], tiles[:, C_WT] = zb + fmin[tR], tR - fmin[tR] + 1, zmax
        if n_long > 0:
            tiles[:, C_V0], tiles[:, C_VN], tiles[:, C_VT] = zb + gy, 1, n_long
        parts.append(zero)
        keys.append(zero_keys)
        n_zero = len(zero)
    allitems = np.concatenate(parts)
    order = np.argsort(np.concatenate(keys), kind="stable")
    out = allitems[order].astype(np.int32)
    return Schedule(items=out, counter_init=init.astype(np.int32), extra_round=False, kind="band-zero" if use_zero else "band",
                    stats=dict(tiles=gx * gy, adam=len(adam), zero=n_zero, untouched=n_untouched,
                               long_lived=int(long_lived.sum()), max_wait=int(out[:, C_WN].max())))


def validate(s: Schedule):
    """Host-side check of the queue invariant: an item only waits for counters that earlier items of the same round
    complete (items flagged "previous round" wait for the whole previous round)."""
    it = s.items
    cnt = s.counter_init.astype(np.int64).copy()
    for k in range(len(it)):
        fl = it[k, C_TYPE] >> 4
        if not (fl & FLAG_PREV_ROUND):
            for w0, wn, wt in ((it[k, C_W0], it[k, C_WN], it[k, C_WT]), (it[k, C_V0], it[k, C_VN], it[k, C_VT])):
                if wn > 0:
                    assert np.all(cnt[w0:w0 + wn] >= wt), f"item {k} waits for counters {w0}..{w0 + wn - 1} >= {wt}: {cnt[w0:w0 + wn]}"
        if it[k, C_SIG] >= 0:
            cnt[it[k, C_SIG]] += 1
    final = cnt
    for k in range(len(it)):                                        # previous-round waits: against the final counts
        if (it[k, C_TYPE] >> 4) & FLAG_PREV_ROUND and it[k, C_WN] > 0:
            assert np.all(final[it[k, C_W0]:it[k, C_W0] + it[k, C_WN]] >= it[k, C_WT]), f"item {k}"
    return True
