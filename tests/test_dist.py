"""N>1 host logic on CPU (gloo, world_size 2) + the multi-GPU equivalence check (needs >= 2 GPUs)."""
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from videoloop3d_b200.train_step import (band_layout, bands_to_frames, exchange_row_bands, frames_to_bands, gather_frames,
                                         owned_frame_ranges, partition, rows_equal)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partitions_cover_everything_once():
    for n, world in [(48, 1), (48, 2), (48, 8), (50, 4), (7, 3), (43, 8)]:
        b = partition(n, world)
        assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(world))
        assert sum(b[i + 1] - b[i] for i in range(world)) == n
    # every frame of the padded video is owned by exactly one rank
    for T, pad, world in [(48, 2, 8), (48, 2, 1), (6, 2, 4), (7, 4, 3)]:
        b = partition(T, world)
        seen = []
        for r in range(world):
            for lo, hi in owned_frame_ranges(b, r, T, pad):
                seen += list(range(lo, hi))
        assert sorted(seen) == list(range(T + pad))


def test_band_layout_covers_rows_and_votes():
    """Row bands of the loss: the owned pixel rows partition [0, h); a band buffer holds every row its patch rows and
    its owned rows touch; every patch row that covers an owned pixel row is inside the band's NN slice."""
    for h, p, s, world in [(720, 11, 4, 8), (720, 3, 2, 8), (180, 11, 4, 4), (46, 5, 2, 3), (43, 7, 3, 2), (64, 3, 4, 4), (37, 15, 4, 2)]:
        hf = (h - p) // s * s + p
        ho = (hf - p) // s + 1
        if ho < world:
            continue
        bands = band_layout(h, p, s, ho, world)
        assert bands[0]["own0"] == 0 and bands[-1]["own1"] == h and bands[0]["pr0"] == 0 and bands[-1]["pr1"] == ho
        for a, b in zip(bands[:-1], bands[1:]):
            assert a["own1"] == b["own0"] and a["pr1"] == b["pr0"]
        for b in bands:
            assert b["ya"] <= b["own0"] <= b["own1"] <= b["yb"] <= h and b["ya"] == (b["pr0"] - b["halo"]) * s
            assert b["yb"] >= min((b["pr1"] - 1) * s + p, h)
            for y in range(b["own0"], b["own1"]):
                if y >= hf:
                    continue
                i0, i1 = max(0, -(-(y - p + 1) // s)), min(y // s, ho - 1)     # patch rows covering pixel row y
                assert b["pr0"] - b["halo"] <= i0 and i1 < b["pr1"], (h, p, s, world, y)


def _band_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for T, h, w, p, s in ((6, 46, 5, 5, 2), (7, 43, 4, 7, 3)):
            hf = (h - p) // s * s + p
            ho = (hf - p) // s + 1
            bounds = partition(T, world)
            bands = band_layout(h, p, s, ho, world)
            me = bands[rank]
            truth = torch.arange(T * 3 * h * w, dtype=torch.float32).reshape(T, 3, h, w)
            mine = truth[bounds[rank]:bounds[rank + 1]].clone()
            hb = me["yb"] - me["ya"]
            band = torch.full((T + 2, 3, hb, w), -1.0)
            frames_to_bands(mine, band, bounds, bands, rank, None)
            ok &= torch.equal(band[:T], truth[:, :, me["ya"]:me["yb"]])
            # adjoint: every rank contributes its owned rows of every frame
            gband = torch.zeros(T + 2, 3, hb, w)
            gband[:T] = 2 * truth[:, :, me["ya"]:me["yb"]] + 1
            out = torch.full((bounds[rank + 1] - bounds[rank], 3, h, w), -7.0)
            bands_to_frames(gband, out, bounds, bands, rank, None)
            ok &= torch.equal(out, 2 * mine + 1)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_band_exchanges_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_band_worker, args=(r, world, 29535 + world, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=10) for _ in range(world))
    assert got == {r: True for r in range(world)}


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the collective pattern of FusedLoopStep on CPU tensors: all-gather rendered frames into their slots,
        # sum the disjoint NN row bands, all-reduce the 5 partial sums
        T, pad, h, w, ho = 6, 2, 3, 4, 5
        b = partition(T, world)
        t0, t1 = b[rank], b[rank + 1]
        video = torch.zeros(T + pad, 3, h, w)
        truth = torch.arange(T * 3 * h * w, dtype=torch.float32).reshape(T, 3, h, w)
        video[t0:t1] = truth[t0:t1]
        dist.all_gather_into_tensor(video[:T], video[t0:t1].clone())
        video[T:T + pad] = video[:pad]
        rows = partition(ho, world)
        nn = torch.zeros(ho, 2, 3, dtype=torch.int32)
        nn[rows[rank]:rows[rank + 1]] = rank + 7
        dist.all_reduce(nn)
        sums = torch.tensor([1.0, 2.0, 3.0, 4.0, float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(sums)
        # scale-invariant gain: block-wise sums of the target frames, all-reduced
        F_ = 5
        fb = partition(F_, world)
        target = torch.arange(F_ * 3 * h * w, dtype=torch.float32).reshape(F_, 3, h, w) / 7
        rsum = target[fb[rank]:fb[rank + 1]].sum(0)
        dist.all_reduce(rsum)
        # the step's own exchange helpers (in-place all_gather_into_tensor for equal blocks, list all-gather /
        # band sum for ragged ones), synchronous and asynchronous
        ok = True
        for T2, ho2 in ((6, 4), (7, 5)):
            b2 = partition(T2, world)
            vid = torch.full((T2 + pad, 3, h, w), -1.0)
            tr = torch.arange(T2 * 3 * h * w, dtype=torch.float32).reshape(T2, 3, h, w)
            vid[b2[rank]:b2[rank + 1]] = tr[b2[rank]:b2[rank + 1]]
            if T2 % world == 0:                    # (gloo has no ragged list all-gather; NCCL does: check_sharded.py)
                for use_async in (False, True):
                    v = vid.clone()
                    work = gather_frames(v, b2, rank, T2, None, async_op=use_async)
                    if use_async:
                        work.wait()
                    ok &= torch.equal(v[:T2], tr)
            rows2 = partition(ho2, world)
            nn2 = torch.zeros(ho2, 2, 3, dtype=torch.int32)
            if rows_equal(rows2):
                nn2.fill_(-5)                      # stale contents of the other bands must be overwritten
            nn2[rows2[rank]:rows2[rank + 1]] = 11 + rank
            exchange_row_bands(nn2, rows2, rank, None)
            ok &= all(int(nn2[r, 1, 2]) == 11 + (0 if r < rows2[1] else 1) for r in range(ho2))
        ok &= torch.equal(video[:T], truth) and torch.equal(video[T:], truth[:pad])
        ok &= all(int(nn[r, 0, 0]) == 7 + (0 if r < rows[1] else 1) for r in range(ho))
        ok &= sums.tolist() == [2.0, 4.0, 6.0, 8.0, 3.0]
        ok &= torch.allclose(rsum, target.sum(0), rtol=1e-6)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,T,pad,h,p,s", [(2, 7, 2, 46, 5, 2), (3, 8, 2, 43, 11, 4), (4, 6, 0, 30, 3, 2)])
def test_peer_exchange_box_lists_single_process(world, T, pad, h, p, s, monkeypatch):
    """PeerExchange (the NVLink peer-memory exchanges) with every rank simulated in ONE process on the CPU: the symmetric
    buffers are plain tensors, `vl3d_copy_boxes` is replaced by the torch copy it stands for.  After frames_to_bands every
    rank's band buffer holds its rows of cat(video, video[:pad]); after bands_to_frames every rank's gradient buffer holds,
    for its frames, the owned rows of every band with the loop-pad adjoint folded in; share_nn delivers whole maps or just
    the halo rows."""
    from videoloop3d_b200 import ops, train_step
    from videoloop3d_b200.train_step import PeerExchange, partition
    w = 12
    ho = (h - p) // s + 1
    bands = band_layout(h, p, s, ho, world)
    bounds = partition(T, world)
    nn_shape = (ho, 5, 3)
    store = {}

    class FakePeer(PeerExchange):
        def __init__(self, rank):
            self.rank, self.world = rank, world
            self.key = None

        def layout(self, *a):
            self.bands, self.bounds, self.dims, self.nn_shape = bands, bounds, (T, pad, h, w), nn_shape
            return self

        def _region(self, rank, which, shape):
            key = (rank, which)
            if key not in store:
                store[key] = torch.zeros(int(torch.tensor(shape).prod()), dtype=torch.float32)
            return store[key].view(shape)

        def barrier(self):
            pass

    def copy_boxes(boxes):
        for src, dst, src2 in boxes:
            dst.copy_(src if src2 is None else src + src2)
    monkeypatch.setattr(ops, "copy_boxes", copy_boxes)
    g = torch.Generator().manual_seed(5)
    video = torch.rand(T, 3, h, w, generator=g)
    peers = [FakePeer(r).layout() for r in range(world)]
    for r, pe in enumerate(peers):
        pe.frames_to_bands(video[bounds[r]:bounds[r + 1]])
    padded = torch.cat([video, video[:pad]])
    for r, b in enumerate(bands):
        assert torch.equal(peers[r].x_band(r), padded[:, :, b["ya"]:b["yb"]])
    # gradient: every rank produces dL/dx for its band (only owned rows matter)
    gfull = torch.rand(T + pad, 3, h, w, generator=g)
    for r, (pe, b) in enumerate(zip(peers, bands)):
        pe.bands_to_frames(gfull[:, :, b["ya"]:b["yb"]].contiguous())
    want = gfull[:T].clone()
    want[:pad] += gfull[T:T + pad]
    for r in range(world):
        assert torch.allclose(peers[r].grad_frames(r), want[bounds[r]:bounds[r + 1]])
    # NN maps
    rows = [b["pr0"] for b in bands] + [ho]
    truth = torch.randint(0, 1000, nn_shape, generator=g, dtype=torch.int32)
    for everything in (True, False):
        for key in [k for k in store if k[1] == 1]:
            store[key].zero_()
        for r, pe in enumerate(peers):
            pe.nn(r)[rows[r]:rows[r + 1]] = truth[rows[r]:rows[r + 1]]
        for pe in peers:
            pe.share_nn(rows, everything=everything)
        for r, b in enumerate(bands):
            got = peers[r].nn(r)
            if everything:
                assert torch.equal(got, truth)
            else:
                assert torch.equal(got[b["pr0"] - b["halo"]:b["pr1"]], truth[b["pr0"] - b["halo"]:b["pr1"]])


def test_collective_pattern_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29533, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=10) for _ in range(2))
    assert got == {0: True, 1: True}


@pytest.mark.gpu
def test_sharded_step_equals_single_gpu_step():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run tests/check_sharded.py under gpurun --gpus 2)")
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29544", os.path.join(ROOT, "tests", "check_sharded.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
