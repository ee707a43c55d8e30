"""glfusion_b200 — B200-native (sm_100a) implementation of GL-Fusion's global-local cross-view fusion hot path.

Public surface (mirrors the reference, R/models/ours.py):
    TPAVIModule          drop-in for models.ours.TPAVIModule / models.TPAVI.TPAVIModule
    GlobalLocalFusion    gate + view concat + MGFM + MLFM + sum as one fused autograd node
    cycle                spatial_sum / seg_cycle / dense_seg_cycle: the trainer's cycle-consistency step (R/main.py:229-235)
    install_trainer()    the same for main.Trainer.seg_cycle / dense_seg_cycle
    install()            monkey-patch the reference's module namespace so Global_and_Local builds on this block
"""
from .tpavi import TPAVIModule, set_default_precision  # noqa: F401
from .fusion import GlobalLocalFusion  # noqa: F401
from . import dp  # noqa: F401
from . import cycle  # noqa: F401
from ._lib import GlfError, load as load_library  # noqa: F401

__version__ = "0.1.0"


def install_trainer(trainer_cls) -> None:
    """``install_trainer(main.Trainer)``: the trainer's two cycle-consistency methods (R/main.py:650-717, :719-798) are
    replaced by the fused loss + gradient kernels of ``glfusion_b200.cycle`` — same names, arguments, defaults and the
    same single ``np.random.choice`` draw in ``seg_cycle``; ``self.device`` is not needed any more."""
    def seg_cycle(self, feat_out, target_region, cyc_off, chunk_size, temperature):
        return cycle.seg_cycle(feat_out, target_region, cyc_off, chunk_size, temperature)

    def dense_seg_cycle(self, feat_out, target_region, cyc_off, chunk_size, temperature, soft_label=False,
                        is_overlap=True):
        return cycle.dense_seg_cycle(feat_out, target_region, cyc_off, chunk_size, temperature, soft_label=soft_label,
                                     is_overlap=is_overlap)
    trainer_cls.seg_cycle = seg_cycle
    trainer_cls.dense_seg_cycle = dense_seg_cycle


def install(reference_models_module) -> None:
    """``install(models.ours)`` before constructing ``Global_and_Local``: the class name is resolved at call time
    (ours.py:1746-1747), so the reference network is then built on the B200 block."""
    reference_models_module.TPAVIModule = TPAVIModule
