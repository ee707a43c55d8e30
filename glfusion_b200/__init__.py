"""glfusion_b200 — B200-native (sm_100a) implementation of GL-Fusion's global-local cross-view fusion hot path.

Public surface (mirrors the reference, R/models/ours.py):
    TPAVIModule          drop-in for models.ours.TPAVIModule / models.TPAVI.TPAVIModule
    GlobalLocalFusion    gate + view concat + MGFM + MLFM + sum as one fused autograd node
    cycle                spatial_sum / seg_cycle / dense_seg_cycle: the trainer's cycle-consistency step (R/main.py:229-235)
    install()            monkey-patch the reference's module namespace so Global_and_Local builds on this block
"""
from .tpavi import TPAVIModule, set_default_precision  # noqa: F401
from .fusion import GlobalLocalFusion  # noqa: F401
from . import dp  # noqa: F401
from . import cycle  # noqa: F401
from ._lib import GlfError, load as load_library  # noqa: F401

__version__ = "0.1.0"


def install(reference_models_module) -> None:
    """``install(models.ours)`` before constructing ``Global_and_Local``: the class name is resolved at call time
    (ours.py:1746-1747), so the reference network is then built on the B200 block."""
    reference_models_module.TPAVIModule = TPAVIModule
