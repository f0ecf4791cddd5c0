"""Import shim: the package directory is named `eigen-zeth_b200` (not a valid Python identifier),
so `import eigen_zeth_b200` resolves its submodules from there."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "eigen-zeth_b200")]
from .lib import *  # noqa: F401,F403,E402
from . import lib  # noqa: F401,E402
