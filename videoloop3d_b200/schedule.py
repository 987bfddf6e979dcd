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
  then live in L2 from zeroing to consumption (Adam reads them with ld.global.cg and drops the lines with
  discard.global.L2), so the texel gradient never crosses HBM.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

ITEM_BWD, ITEM_ADAM, ITEM_ZERO = 0, 1, 2
FLAG_HAS_GRAD, FLAG_REZERO, FLAG_DISCARD, FLAG_PREV_ROUND, FLAG_ORDERED = 1, 2, 4, 8, 16
BX, BY, TF = 32, 8, 2                  # tile shape / frames per chunk of the kernels (composite_common.cuh)


@dataclass
class Schedule:
    items: np.ndarray        # (n_items, 8) int32: [type | flags << 4, a, b, c, wait_first, wait_count, wait_target, signal]
    n_counters: int
    extra_round: bool        # items of the last chunk run in one more round (generic schedule)
    kind: str
    stats: dict

    @property
    def n_items(self):
        return int(self.items.shape[0])


def tile_grid(H, W, smooth=True):
    sx, sy = (BX - 1, BY - 1) if smooth else (BX, BY)
    return (W + sx - 1) // sx, (H + sy - 1) // sy, sx, sy


def _item(kind, flags, a=0, b=0, c=0, wait=(-1, 0, 0), signal=-1):
    return (kind | (flags << 4), a, b, c, wait[0], wait[1], wait[2], signal)


def generic_schedule(H, W, dyn_h, dyn_w, smooth=True, seg_texels=32768, lead_tiles=888):
    """Any layout.  Counter 0 = finished tiles of the chunk."""
    gx, gy, _, _ = tile_grid(H, W, smooth)
    n_tiles = gx * gy
    rows_per = max(1, seg_texels // max(dyn_w, 1))
    segs = [(r0 * dyn_w, dyn_w, min(rows_per, dyn_h - r0)) for r0 in range(0, dyn_h, rows_per)]
    adam = [_item(ITEM_ADAM, FLAG_HAS_GRAD | FLAG_REZERO | FLAG_PREV_ROUND, b, w, r, wait=(0, 1, n_tiles)) for b, w, r in segs]
    tiles = [_item(ITEM_BWD, 0, bx, by, 0, signal=0) for by in range(gy) for bx in range(gx)]
    lead = min(lead_tiles, n_tiles // 4)
    # merge: no Adam among the first `lead` tiles (the previous chunk's last tiles are still running), then evenly
    pos = lead + (np.arange(len(adam)) + 0.5) * (n_tiles - lead) / max(len(adam), 1)
    order = np.argsort(np.concatenate([np.arange(n_tiles, dtype=np.float64), pos]), kind="stable")
    allitems = tiles + adam
    items = np.asarray([allitems[i] for i in order], dtype=np.int32)
    return Schedule(items=items, n_counters=1, extra_round=True, kind="generic",
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
                  zero_ahead=2, adam_lag=2, use_zero=True, margin=2):
    """Dense layout.  Counters: [0, gy) finished tiles per tile row; gy = in-order count of finished ZERO items."""
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

    keys, items = [], []
    zero_first = []
    for by in range(gy):
        for bx in range(gx):
            keys.append(float(by))
            items.append([ITEM_BWD, 0, bx, by, 0, -1, 0, 0, by])   # wait filled below
    n_untouched = int((lb < 0).sum())
    ui = 0
    for rb in range(n_rb):
        r0 = rb * row_block
        nr = min(row_block, dyn_h - r0)
        for cb in range(col_blocks):
            base, width = r0 * dyn_w + cb_edges[cb], cb_edges[cb + 1] - cb_edges[cb]
            if width <= 0:
                continue
            f, l = int(fb[rb, cb]), int(lb[rb, cb])
            if l < 0:                                               # never touched: Adam with g = 0, any time
                keys.append((ui + 0.5) * gy / max(n_untouched, 1))
                ui += 1
                items.append([ITEM_ADAM, 0, base, width, nr, -1, 0, 0, -1])
                continue
            aligned = (base % 8 == 0) and (width % 8 == 0) and (dyn_w % 8 == 0)
            if use_zero:
                keys.append(f - zero_ahead - 0.5)
                items.append([ITEM_ZERO, FLAG_ORDERED, base, width, nr, -1, 0, 0, gy])
                zero_first.append(f)
                fl = FLAG_HAS_GRAD | (FLAG_DISCARD if aligned else 0)
            else:
                fl = FLAG_HAS_GRAD | FLAG_REZERO
            keys.append(l + adam_lag + 0.25)
            items.append([ITEM_ADAM, fl, base, width, nr, f, l - f + 1, gx, -1])
    order = np.argsort(np.asarray(keys), kind="stable")
    out = np.zeros((len(items), 8), dtype=np.int32)
    zseq = 0
    nz_before_row = np.zeros(gy, dtype=np.int64)
    if use_zero:
        zf = np.sort(np.asarray(zero_first, dtype=np.int64))
        nz_before_row = np.searchsorted(zf, np.arange(gy), side="right")       # ZERO items with first <= R
    for k, i in enumerate(order):
        kind, fl, a, b, c, w0, wn, wt, sig = items[i]
        if kind == ITEM_BWD and use_zero:
            w0, wn, wt = gy, 1, int(nz_before_row[b])
            if wt == 0:
                w0, wn = -1, 0
        if kind == ITEM_ZERO:
            wt = zseq                                               # in-order commit: bump counter gy when it equals zseq
            zseq += 1
        out[k] = _item(kind, fl, a, b, c, wait=(w0, wn, wt), signal=sig)
    # the in-order ZERO sequence must match the order in which `nz_before_row` counts them: ZERO items are queued by
    # ascending `first` (their keys), so the first nz_before_row[R] of them are exactly those with first <= R
    return Schedule(items=out, n_counters=gy + 1, extra_round=False, kind="band-zero" if use_zero else "band",
                    stats=dict(tiles=gx * gy, adam=int((out[:, 0] & 15 == ITEM_ADAM).sum()), zero=zseq,
                               untouched=n_untouched, max_wait=int(out[:, 5].max())))


def validate(s: Schedule):
    """Host-side check of the queue invariant: an item only waits for counters that earlier items complete."""
    it = s.items
    done = np.zeros(s.n_counters, dtype=np.int64)
    for k in range(len(it)):
        kind, fl = it[k, 0] & 15, it[k, 0] >> 4
        if fl & FLAG_PREV_ROUND:
            continue                                                # waits for the previous round: checked separately
        w0, wn, wt, sig = (int(x) for x in it[k, 4:8])
        if kind == ITEM_ZERO and (fl & FLAG_ORDERED):
            assert done[sig] == wt, f"item {k}: ZERO sequence {wt} but {done[sig]} committed"
        elif wn > 0:
            assert np.all(done[w0:w0 + wn] >= wt), f"item {k} waits for counters {w0}..{w0 + wn - 1} >= {wt}: {done[w0:w0 + wn]}"
        if sig >= 0:
            done[sig] += 1
    return True
