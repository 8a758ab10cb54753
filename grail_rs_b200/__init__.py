"""Import shim: the product package lives in the directory `grail-rs_b200/` (the name the build
contract asks for, which is not a valid Python identifier).  `import grail_rs_b200` resolves here and
extends the package path to that directory."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "grail-rs_b200"))

from ._pkg import *  # noqa: F401,F403,E402
from ._pkg import __all__  # noqa: F401,E402
