"""videoloop3d_b200 — B200-native (sm_100a) hot path of VideoLoop3D's stage-2 optimisation step.

Public surface mirrors the reference's Python operators:
    MPMeshVid                    (reference MPV.py)       render / forward / lod / get_optimizer / ...
    MPMesh                       (reference MPI.py:38-124,452-652) stage-1 model: render / forward
    Patch3DGPNNLowMemLoss, ...   (reference utils_vid.py) loop-loss callables
    make_run_iter, FusedLoopStep (reference train_3dvid.py:214-255) the optimisation step
    MVVidPatchDataset, generate_patchinfo (reference train_3dvid.py:22-66, utils.py:115-134) the step's items
The numerical work is done by hand-written CUDA kernels in libvl3d.so (C ABI: include/vl3d.h).
"""
from ._lib import Vl3dError, load as load_library  # noqa: F401
from .loop_loss import (Patch3DAvg, Patch3DGPNNDirectLoss,   # noqa: F401
                        Patch3DGPNNLowMemLoss, Patch3DMSE)
from .mpi import MPMesh  # noqa: F401
from .mpv import MPMeshVid, get_new_intrin, make_depths, gen_mpi_vertices, pose2extrin_torch  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .evaluations import compute_nnerr, to8b  # noqa: F401
from .dataset import MVVidPatchDataset, generate_patchinfo  # noqa: F401
from .train_step import (FusedLoopStep, make_run_iter, make_run_iter_stage1, default_args,  # noqa: F401
                         default_args_stage1)

__version__ = "0.1.0"
