"""Import shim: the product package lives in ``rigidbody-simulation_b200/`` (a hyphen is not a valid
module name), so this package re-points its ``__path__`` there and executes that ``__init__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "rigidbody-simulation_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
