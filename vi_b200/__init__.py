"""Import alias for the product package.

The package directory carries the repository's mandated (hyphenated, hence
not importable) name; this shim makes it importable as ``vi_b200`` by pointing
``__path__`` at it.  All code lives in that directory.
"""
import os as _os

_PKG_DIR = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "vision-inspection-system-segmentation-using-classical-computer-vision-_b200",
)
__path__.append(_PKG_DIR)

from ._init import *  # noqa: E402,F401,F403
from ._init import __all__  # noqa: E402,F401
