"""TEST INFRASTRUCTURE — makes the unmodified reference importable in THIS container.

Only `oracle/make_golden.py` (and ad-hoc validation run here) uses this; it needs
`/root/reference`, which does not exist on the GPU box.  Nothing in the product imports it.
"""
import os
import sys

REFERENCE_DIR = os.environ.get("VL3D_REFERENCE_DIR", "/root/reference")
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "MPV.py"))


def enable():
    """Put the shims ahead of the reference on sys.path."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_DIR}")
    sys.dont_write_bytecode = True  # the reference dir is read-only
    for p in (REFERENCE_DIR, SHIM_DIR):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_DIR)
    sys.path.insert(0, SHIM_DIR)


def make_args(config="configs/mpv_base.txt", **overrides):
    """Namespace from the reference's own parser + one of its config files + overrides."""
    enable()
    import config_parser  # reference module
    argv = []
    if config:
        argv += ["--config", os.path.join(REFERENCE_DIR, config)]
    args = config_parser.config_parser().parse_args(argv)
    for k, v in overrides.items():
        assert hasattr(args, k), k
        setattr(args, k, v)
    return args
