"""toycrystals_b200 — B200 (sm_100a) native implementation of the ToyCrystals VP-SDE sampling hot
path.  Mirrors ``toycrystals.models.sde_score_model`` of the reference for that path only."""
from . import _cabi  # noqa: F401

__all__ = ["_cabi"]
__version__ = "0.1.0"
