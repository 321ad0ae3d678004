"""Mirror of the reference's ``src`` package for the accelerated path.

Put ``<repo>/rigidbody-simulation_b200`` on ``sys.path`` and the reference's import lines keep working:
``from src.physics.collision import compute_collision_impulse_friction`` etc. (the same modules are also
importable as ``rigidbody_simulation_b200.src...``).
"""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:          # the import shim ``rigidbody_simulation_b200`` lives at the repo root
    _sys.path.append(_root)
