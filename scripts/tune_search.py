"""GPU tuning / A-B aid for the patch-NN search and vote kernels (development tool).

    python scripts/tune_search.py [--H 720 --W 1280 --T 48 --F 258 --p 11 --pt 3 --s 4]

Times vl3d_patchnn_search and vl3d_vote_loss on random videos.  `--search a,b` / `--vote a,b` set
VL3D_NN_VARIANT / VL3D_VOTE_VARIANT for experimental kernel variants (none are compiled in at the moment: the
packed-fp32x2, three-CTA and batched-gather variants were measured and removed, see profiles/README.md) and
report how many NN indices / gradient values differ from the first one."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from videoloop3d_b200 import ops  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    for k, v in dict(H=720, W=1280, T=48, F=258, p=11, pt=3, s=4, st=1, reps=3).items():
        ap.add_argument(f"--{k}", type=int, default=v)
    ap.add_argument("--alpha", type=float, default=0.0)
    ap.add_argument("--search", default="0")
    ap.add_argument("--vote", default="0")
    ap.add_argument("--dump", default="", help="save the NN map here (A/B across processes: VL3D_LIB=... builds)")
    ap.add_argument("--check", default="", help="compare the NN map with one saved by --dump")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    pad = a.pt - 1
    x = torch.rand((a.T + pad, 3, a.H, a.W), device=dev, generator=g)
    raw = torch.rand((a.F + 8, 3, a.H, a.W), device=dev, generator=g)
    cs = torch.cumsum(raw, 0)
    y = ((cs[9:] - cs[:-9]) / 9).contiguous()[: a.F]
    del raw, cs
    xscale = torch.tensor([1.03], device=dev)
    desc = ops.make_loss_desc(x.shape, (x.stride(0), x.stride(1), x.stride(2)), y.shape, (y.stride(0), y.stride(1), y.stride(2)),
                              a.p, a.pt, a.s, a.st, a.alpha)
    print(f"x {tuple(x.shape)} y {tuple(y.shape)}: ho={desc.ho} wo={desc.wo} n1={desc.n1} n2={desc.n2}")
    ws = torch.empty_like(x)
    ref = None
    for v in a.search.split(","):
        os.environ["VL3D_NN_VARIANT"] = v
        nn = torch.empty((desc.ho, desc.wo, desc.n1), dtype=torch.int32, device=dev)
        ms = timed(lambda: ops.patchnn_search(desc, x, xscale, y, nn_out=nn, scaled_ws=ws), a.reps)
        if ref is None:
            ref = nn.clone()
        print(f"search variant {v}: {ms:8.3f} ms   indices differing from variant {a.search.split(',')[0]}: "
              f"{int((nn != ref).sum())} / {nn.numel()}", flush=True)
    os.environ.pop("VL3D_NN_VARIANT", None)
    if a.dump:
        torch.save(ref.cpu(), a.dump)
    if a.check:
        other = torch.load(a.check)
        print(f"indices differing from {a.check}: {int((ref.cpu() != other).sum())} / {ref.numel()}", flush=True)
    gref = None
    for v in a.vote.split(","):
        os.environ["VL3D_VOTE_VARIANT"] = v
        grad = torch.zeros_like(x)

        def vote():
            return ops.vote_loss(desc, x, xscale, y, ref, "-2", 0.1, 3.5, (a.T + pad, a.H, a.W), grad_out=grad)

        ms = timed(vote, a.reps)
        loss = float(vote()[0])
        if gref is None:
            gref = grad.clone()
        err = float((grad - gref).abs().max() / gref.abs().max())
        print(f"vote variant {v}: {ms:8.3f} ms   loss {loss:.7f}  max grad diff vs first {err:.1e}", flush=True)


if __name__ == "__main__":
    main()
