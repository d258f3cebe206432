"""B200-native exhaustive rigid-body docking search: a drop-in for the hot path of
lupoglaz/DeepLocalProteinDocking (`src/Docker/Docker.py:184-238`).  See DESIGN.md."""
from .rotations import Rotations, euler_to_matrix, read_eul, write_eul, soi_euler_angles  # noqa: F401
from .search import DockingSearch  # noqa: F401
from .volume import VolumeRotation, VolumeConvolution, TypedCoords2Volume  # noqa: F401
from .models import SimpleFilter, GlobalDockingModel, E3MultiResRepr4x4, select_model  # noqa: F401
from .se3_scalar import SE3MultiResReprScalar  # noqa: F401
from .docker import Docker, Atoms  # noqa: F401
from . import rotations, synthetic, distributed, tpl, poses  # noqa: F401
from ._lib import set_option  # noqa: F401

__all__ = ["Rotations", "DockingSearch", "VolumeRotation", "VolumeConvolution", "TypedCoords2Volume", "SimpleFilter",
           "GlobalDockingModel", "E3MultiResRepr4x4", "SE3MultiResReprScalar", "select_model", "Docker", "Atoms"]
