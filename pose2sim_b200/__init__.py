"""pose2sim_b200 — B200-native (sm_100a) triangulation / person-association hot path of Pose2Sim.

Drop-in entry points (same signature and side effects as the reference's stage functions):

    from pose2sim_b200 import triangulate_all, associate_all
    triangulate_all(config_dict)      # Pose2Sim/triangulation.py:656
    associate_all(config_dict)        # Pose2Sim/personAssociation.py:642

The arithmetic runs in hand-written CUDA kernels behind the C ABI of include/pose2sim_b200.h
(`libp2s_b200.so`, built in-tree by `make -C pose2sim_b200/csrc`); there is no CPU fallback.
"""


def triangulate_all(config_dict):
    from .triangulation import triangulate_all as _f
    return _f(config_dict)


def associate_all(config_dict):
    from .personAssociation import associate_all as _f
    return _f(config_dict)


def install_into_reference():
    """Rebind the stage functions of an importable `Pose2Sim` package so that the unchanged
    orchestrator (`Pose2Sim.triangulation()`, `Pose2Sim.personAssociation()`; lazy imports at
    Pose2Sim/Pose2Sim.py:233,242) runs this implementation.  See INTEGRATION.md."""
    import Pose2Sim.personAssociation as _pa
    import Pose2Sim.triangulation as _tri
    _tri.triangulate_all = triangulate_all
    _pa.associate_all = associate_all


__all__ = ["triangulate_all", "associate_all", "install_into_reference"]
