"""Multi-GPU check (run under torchrun): the T-sharded FusedLoopStep must reproduce the single-GPU step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/check_sharded.py

Every rank builds the same small model; rank r optimises only its frame block (both the "full model on
every rank" and the memory-sharded variants are exercised); after 3 steps the parameters, the losses and
the NN maps are compared with a single-GPU run of the same steps on rank 0.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

from oracle import mpv_oracle as MO
from videoloop3d_b200 import FusedLoopStep
from videoloop3d_b200.testing import model_from_tensors


def tensors(st):
    return dict(verts=st.verts, planedepth=st.planedepth, faces=st.faces, faces_dyn=st.faces_dyn, uvs=st.uvs,
                uvs_dyn=st.uvs_dyn, uvfaces=st.uvfaces, uvfaces_dyn=st.uvfaces_dyn, atlas=st.atlas, atlas_dyn=st.atlas_dyn,
                ref_extrin=st.ref_extrin, ref_intrin=st.ref_intrin, mpi_d=st.mpi_d, hv=st.hv, wv=st.wv)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    # ragged blocks (T=7, odd number of patch rows: list all-gather / all-reduce of the NN map) and equal blocks
    # (T=8, patch rows divisible by the world size up to 4: the in-place all_gather_into_tensor exchanges)
    for H, W, D, T, F in ((46, 83, 8, 7, 13), (43, 83, 8, 8, 13)):
        ok &= run_case(H, W, D, T, F, rank, world, dev)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


def run_case(H, W, D, T, F, rank, world, dev):
    st = MO.sparse_state(H, W, D, 6, 9, T, 1.0, 10.0, tile=6, occupancy=0.7, dyn_frac=0.5, h_scale=1.2, w_scale=1.2, seed=4)
    ext = torch.eye(4)[None]
    ext[0, 0, 3] = 0.05
    f = 0.8 * W
    intr = torch.tensor([[f, 0, W / 2 + 0.3], [0, f, H / 2 - 0.2], [0, 0, 1.]])[None]
    res = torch.rand(1, F, 3, H, W, generator=torch.Generator().manual_seed(0)).to(dev)
    cfg = dict(loss_name="gpnn_lm", loss_gain=3.5, patch_size=5, patcht_size=3, stride=2, stridet=1, alpha=0.0,
               rou="-2", scaling=0.1, dist_fn="mse", macro_block=65, factor=1)
    ok = True
    # exchanges through NCCL all-to-alls and through stores into the peers' symmetric-memory buffers (PeerExchange)
    for variant, exchange in (("replicated", "nccl"), ("memory-sharded", "nccl"), ("replicated", "p2p"), ("memory-sharded", "p2p")):
        m = model_from_tensors(tensors(st), H, W, dev)
        if variant == "replicated":
            step = FusedLoopStep(m, group=dist.group.WORLD, exchange=exchange)
        else:
            b = [(T * r) // world for r in range(world + 1)]
            m.atlas_dyn.data = m.atlas_dyn.data[b[rank]:b[rank + 1]].clone(memory_format=torch.preserve_format)
            step = FusedLoopStep(m, group=dist.group.WORLD, global_frames=T, exchange=exchange)
        losses = [step.step(H, W, ext.to(dev), intr.to(dev), res, cfg, lr=0.01)["loss"] for _ in range(3)]
        t0, t1 = step.t0, step.t1
        mine = m.atlas_dyn.data if variant == "memory-sharded" else m.atlas_dyn.data[t0:t1]
        full = [torch.empty((b1 - b0,) + tuple(mine.shape[1:]), device=dev) for b0, b1 in zip(step.bounds[:-1], step.bounds[1:])]
        dist.all_gather(full, mine.contiguous())
        nn_sh = step._buf["nn"].clone()
        if rank == 0:
            m1 = model_from_tensors(tensors(st), H, W, dev)
            s1 = FusedLoopStep(m1)
            l1 = [s1.step(H, W, ext.to(dev), intr.to(dev), res, cfg, lr=0.01)["loss"] for _ in range(3)]
            dp = float((torch.cat(full) - m1.atlas_dyn.data.contiguous()).abs().max())
            ds = float((m.atlas.data - m1.atlas.data).abs().max())
            dl = max(abs(float(a) - float(b)) / abs(float(b)) for a, b in zip(losses, l1))
            nn_eq = bool(torch.equal(nn_sh, s1._buf["nn"]))
            good = dp < 2e-5 and ds < 2e-5 and dl < 1e-5 and nn_eq
            ok &= good
            print(f"[{variant}/{exchange}] T={T} H={H} world={world}: max|d atlas_dyn|={dp:.2e} max|d atlas|={ds:.2e} rel d loss={dl:.2e} "
                  f"nn identical={nn_eq} -> {'OK' if good else 'MISMATCH'}", flush=True)
    return ok


if __name__ == "__main__":
    main()
