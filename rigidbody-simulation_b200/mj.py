"""The handful of ``mujoco`` symbols the reference's physics path touches, for batched device state.

``import rigidbody_simulation_b200.mj as mj`` lets a scenario written against the reference
(``mj.MjModel.from_xml_path``, ``mj.MjData``, ``mj.mj_name2id``, ``mj.mj_forward``, ``mj.mj_resetData``,
``mj.mju_mulQuat``) run on the B200 path.  MuJoCo's own dynamics (``mj_step``) and rendering are out of
scope.
"""
import enum

import numpy as np

from .batched import BatchedData, BatchedModel

MjModel = BatchedModel


def MjData(model, layout=None):
    """Batched MjData.  Scenes with more than two free spheres use the thread-per-body layout."""
    if layout is None:
        layout = "body" if model.nfree > 2 else "env"
    return BatchedData(model, layout=layout)


class mjtObj(enum.IntEnum):
    mjOBJ_UNKNOWN = 0
    mjOBJ_BODY = 1
    mjOBJ_GEOM = 5


def mj_name2id(model, objtype, name):
    """-1 for an unknown name, like MuJoCo (the reference relies on it: body_mass[-1] is the last body,
    src/physics/collision.py:58-60 with obj="sphere" in src/simulation/single_sphere_bounce.py:67)."""
    if int(objtype) == int(mjtObj.mjOBJ_BODY):
        return model.scene.body_id(name)
    if int(objtype) == int(mjtObj.mjOBJ_GEOM):
        for i, g in enumerate(model.scene.geoms):
            if g.name == name:
                return i
    return -1


def mj_forward(model, data):
    """The reference calls mj_forward only to obtain the contact list (src/physics/collision.py:57).  On this
    path the narrow phase runs inside the step kernels on the start-of-step pose, so this is a no-op kept for
    source compatibility."""
    return None


def mj_resetData(model, data):
    data.reset()


def mj_step(model, data):
    raise NotImplementedError("MuJoCo's soft-contact solver (compare_builtin) is outside the accelerated path")


def mju_mulQuat(res, a, b):
    """Hamilton product res = a (x) b, scalar first (host helper)."""
    a0, a1, a2, a3 = (float(x) for x in a)
    b0, b1, b2, b3 = (float(x) for x in b)
    out = np.array([a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3, a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
                    a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1, a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0])
    res[:] = out
