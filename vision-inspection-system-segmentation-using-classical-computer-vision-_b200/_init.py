from . import grid, synth  # noqa: F401
from ._lib import (RECORD_DTYPE, STATUS_NG, STATUS_OK, STATUS_ROI_EMPTY, ViError, ViParams,  # noqa: F401
                   default_params)
from .api import Inspector, default_inspector  # noqa: F401
from .grid import Grid, generate_grid, load_grid, parse_grid, save_grid  # noqa: F401

__all__ = ["grid", "synth", "RECORD_DTYPE", "STATUS_NG", "STATUS_OK", "STATUS_ROI_EMPTY", "ViError", "ViParams",
           "default_params", "Inspector", "default_inspector", "Grid", "generate_grid", "load_grid", "parse_grid",
           "save_grid"]
