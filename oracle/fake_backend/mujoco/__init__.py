"""Fake ``mujoco`` backend -- TEST INFRASTRUCTURE ONLY (oracle side, never shipped).

The reference (pratyay2510/RigidBody-Simulation) depends on the third-party PyPI
package ``mujoco`` (Google DeepMind; UNPINNED in the reference's requirements.txt:1)
for exactly three things on its custom-physics path:

  1. XML -> model   (``MjModel.from_xml_path``: mass, diagonal inertia, gravity, timestep, qpos0)
  2. narrow-phase contact generation inside ``mj_forward``
     (call sites: src/physics/collision.py:57, src/physics/time_integeration.py:29,95,
      src/simulation/multi_sphere_bounce.py:43, src/simulation/ball_collision.py:74)
  3. ``mju_mulQuat`` (src/physics/collision.py:93)

``mujoco`` is not installed in the build container and cannot be installed (no
network), so this module restates the published behaviour of those three pieces
(MuJoCo's engine_collision_primitive.c plane-sphere / plane-box / sphere-sphere
routines, the XML compiler's inertia-from-geom rule, and the Hamilton product) in
pure Python/NumPy.  SURVEY.md Appendix A is the normative spec.  With this module
first on ``sys.path`` the reference's *unmodified* source files run headless; that
is how the golden vectors under tests/golden/ are produced (oracle/make_golden.py).

Parity statement: results are "vs the fake-MuJoCo oracle"; parity against the real
MuJoCo C library is unverified (the only external pins are the two published plots,
see tests/test_oracle_golden.py).
"""
import enum
import math
import xml.etree.ElementTree as ET

import numpy as np

__version__ = "0.0-fake"


# --------------------------------------------------------------------------- enums
class mjtObj(enum.IntEnum):
    mjOBJ_UNKNOWN = 0
    mjOBJ_BODY = 1
    mjOBJ_GEOM = 5


class mjtGeom(enum.IntEnum):
    mjGEOM_PLANE = 0
    mjGEOM_SPHERE = 2
    mjGEOM_BOX = 6


class _ValueEnum(enum.Enum):
    pass


class mjtMouse(enum.IntEnum):
    mjMOUSE_NONE = 0
    mjMOUSE_ROTATE_V = 1
    mjMOUSE_ROTATE_H = 2
    mjMOUSE_MOVE_V = 3
    mjMOUSE_MOVE_H = 4
    mjMOUSE_ZOOM = 5


class mjtFontScale(enum.Enum):
    mjFONTSCALE_50 = 50
    mjFONTSCALE_100 = 100
    mjFONTSCALE_150 = 150


class mjtCatBit(enum.Enum):
    mjCAT_STATIC = 1
    mjCAT_DYNAMIC = 2
    mjCAT_DECOR = 4
    mjCAT_ALL = 7


# ----------------------------------------------------------------- small math helpers
def _euler_xyz_to_quat(e):
    """Intrinsic x-y-z euler (radians) -> unit quaternion wxyz (MuJoCo default eulerseq 'xyz')."""
    q = np.array([1.0, 0.0, 0.0, 0.0])
    for axis, ang in enumerate(e):
        h = 0.5 * float(ang)
        r = np.array([math.cos(h), 0.0, 0.0, 0.0])
        r[1 + axis] = math.sin(h)
        out = np.zeros(4)
        mju_mulQuat(out, q, r)
        q = out
    return q


def _quat_to_mat(q):
    w, x, y, z = (float(v) for v in q)
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mju_mulQuat(res, a, b):
    """Hamilton product res = a (x) b, scalar-first."""
    a0, a1, a2, a3 = (float(v) for v in a)
    b0, b1, b2, b3 = (float(v) for v in b)
    res[0] = a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3
    res[1] = a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2
    res[2] = a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1
    res[3] = a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0


def _floats(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=float)
    v = np.array([float(t) for t in s.split()], dtype=float)
    if n is not None and v.size < n:
        v = np.concatenate([v, np.zeros(n - v.size)])
    return v


# --------------------------------------------------------------------------- model
class _Opt:
    def __init__(self):
        self.gravity = np.array([0.0, 0.0, -9.81])
        self.timestep = 0.002


class _Geom:
    __slots__ = ("name", "type", "size", "body", "pos", "quat", "density")


class MjModel:
    """XML subset compiler: worldbody with geoms and one level of
    ``<body>`` children, each optionally carrying a free joint.  angle=radian, local
    coordinates, inertiafromgeom=true, density from the geom, euler = intrinsic xyz."""

    def __init__(self):
        self.opt = _Opt()
        self.body_names = ["world"]
        self.body_mass = np.zeros(1)
        self.body_inertia = np.zeros((1, 3))
        self.body_pos = np.zeros((1, 3))
        self.body_quat = np.array([[1.0, 0, 0, 0]])
        self.body_jntadr = [-1]
        self.geoms = []
        self.qpos0 = np.zeros(0)
        self.nq = 0
        self.nv = 0

    @property
    def nbody(self):
        return len(self.body_names)

    @property
    def ngeom(self):
        return len(self.geoms)

    @classmethod
    def from_xml_path(cls, path):
        with open(path, "r") as f:
            return cls.from_xml_string(f.read())

    @classmethod
    def from_xml_string(cls, text):
        root = ET.fromstring(text)
        m = cls()
        opt = root.find("option")
        if opt is not None:
            if opt.get("gravity") is not None:
                m.opt.gravity = _floats(opt.get("gravity"))
            if opt.get("timestep") is not None:
                m.opt.timestep = float(opt.get("timestep"))
        wb = root.find("worldbody")
        masses, inertias, bpos, bquat, jadr = [0.0], [np.zeros(3)], [np.zeros(3)], [np.array([1.0, 0, 0, 0])], [-1]
        qpos0 = []
        nv = 0

        def add_geom(el, body_id):
            g = _Geom()
            g.name = el.get("name")
            g.type = el.get("type", "sphere")
            g.size = _floats(el.get("size"), 3, [0, 0, 0])
            g.body = body_id
            g.pos = _floats(el.get("pos"), 3, [0, 0, 0])
            if el.get("quat") is not None:
                g.quat = _floats(el.get("quat"))
            else:
                g.quat = _euler_xyz_to_quat(_floats(el.get("euler"), 3, [0, 0, 0]))
            g.density = float(el.get("density", 1000.0))
            m.geoms.append(g)
            return g

        # world geoms come first in id order, then bodies in document order
        for el in wb.findall("geom"):
            add_geom(el, 0)
        for b in wb.findall("body"):
            bid = len(m.body_names)
            m.body_names.append(b.get("name"))
            p = _floats(b.get("pos"), 3, [0, 0, 0])
            if b.get("quat") is not None:
                q = _floats(b.get("quat"))
            else:
                q = _euler_xyz_to_quat(_floats(b.get("euler"), 3, [0, 0, 0]))
            bpos.append(p)
            bquat.append(q)
            mass, inertia = 0.0, np.zeros(3)
            for el in b.findall("geom"):
                g = add_geom(el, bid)
                if g.type == "sphere":
                    r = g.size[0]
                    gm = g.density * (4.0 / 3.0) * math.pi * r ** 3
                    gi = np.full(3, 0.4 * gm * r * r)
                elif g.type == "box":
                    a, bb, c = g.size
                    gm = g.density * 8.0 * a * bb * c
                    gi = gm / 3.0 * np.array([bb * bb + c * c, a * a + c * c, a * a + bb * bb])
                else:  # planes carry no mass
                    gm, gi = 0.0, np.zeros(3)
                mass += gm
                inertia = inertia + gi
            masses.append(mass)
            inertias.append(inertia)
            free = any(j.get("type") == "free" for j in b.findall("joint")) or b.find("freejoint") is not None
            if free:
                jadr.append(len(qpos0) // 7)
                qpos0.extend(list(p) + list(q))
                nv += 6
            else:
                jadr.append(-1)
        m.body_mass = np.array(masses)
        m.body_inertia = np.array(inertias)
        m.body_pos = np.array(bpos)
        m.body_quat = np.array(bquat)
        m.body_jntadr = jadr
        m.qpos0 = np.array(qpos0, dtype=float)
        m.nq = m.qpos0.size
        m.nv = nv
        return m


class _Contact:
    __slots__ = ("dist", "pos", "frame", "geom1", "geom2", "geom")

    def __init__(self, dist, pos, normal, g1, g2):
        self.dist = float(dist)
        self.pos = np.array(pos, dtype=float)
        fr = np.zeros(9)
        fr[:3] = normal
        self.frame = fr
        self.geom1 = g1
        self.geom2 = g2
        self.geom = np.array([g1, g2])


class MjData:
    def __init__(self, model):
        self.qpos = model.qpos0.copy()
        self.qvel = np.zeros(model.nv)
        self.xfrc_applied = np.zeros((model.nbody, 6))
        self.contact = []
        self.ncon = 0
        self.time = 0.0


def mj_resetData(model, data):
    data.qpos[:] = model.qpos0
    data.qvel[:] = 0.0
    data.xfrc_applied[:] = 0.0
    data.time = 0.0
    data.contact = []
    data.ncon = 0


def mj_name2id(model, objtype, name):
    if int(objtype) == int(mjtObj.mjOBJ_BODY):
        names = model.body_names
    elif int(objtype) == int(mjtObj.mjOBJ_GEOM):
        names = [g.name for g in model.geoms]
    else:
        return -1
    try:
        return names.index(name)
    except ValueError:
        return -1


# ------------------------------------------------------------------- narrow phase
def _geom_world(model, data, g):
    """World position and rotation matrix of geom g from the current qpos."""
    b = g.body
    adr = model.body_jntadr[b]
    if adr >= 0:
        bp = np.array(data.qpos[7 * adr: 7 * adr + 3], dtype=float)
        bq = np.array(data.qpos[7 * adr + 3: 7 * adr + 7], dtype=float)
        bq = bq / np.linalg.norm(bq)          # mj_kinematics normalises joint quaternions
    else:
        bp, bq = model.body_pos[b], model.body_quat[b]
    R = _quat_to_mat(bq)
    if not (g.pos == 0).all() or not (g.quat == np.array([1.0, 0, 0, 0])).all():
        return bp + R @ g.pos, R @ _quat_to_mat(g.quat)
    return bp.copy(), R


def _plane_sphere(pp, pR, c, r):
    n = pR[:, 2]
    dist = float(np.dot(c - pp, n)) - r
    if dist > 0:
        return []
    return [(dist, c - n * (r + 0.5 * dist), n.copy())]


def _plane_box(pp, pR, c, R, half):
    n = pR[:, 2]
    d0 = float(np.dot(c - pp, n))
    out = []
    for i in range(8):
        vert = np.array([half[0] if i & 1 else -half[0],
                         half[1] if i & 2 else -half[1],
                         half[2] if i & 4 else -half[2]])
        corner = R @ vert
        ld = float(np.dot(n, corner))
        if d0 + ld > 0 or ld > 0:
            continue
        dist = d0 + ld
        out.append((dist, c + corner - n * (0.5 * dist), n.copy()))
        if len(out) >= 4:
            break
    return out


def _sphere_sphere(c1, r1, c2, r2):
    d = c2 - c1
    L = float(np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
    dist = L - r1 - r2
    if dist > 0:
        return []
    n = d / L if L >= 1e-15 else np.array([1.0, 0.0, 0.0])
    return [(dist, c1 + n * (r1 + 0.5 * dist), n)]


def mj_forward(model, data):
    """Only the collision stage matters to the reference: fill data.contact / data.ncon.
    Order: ascending (body1, body2), then geom pair."""
    gw = [_geom_world(model, data, g) for g in model.geoms]
    pairs = []
    ng = len(model.geoms)
    for i in range(ng):
        for j in range(i + 1, ng):
            gi, gj = model.geoms[i], model.geoms[j]
            if gi.body == gj.body:
                continue
            static_i = model.body_jntadr[gi.body] < 0
            static_j = model.body_jntadr[gj.body] < 0
            if static_i and static_j:
                continue
            pairs.append((min(gi.body, gj.body), max(gi.body, gj.body), i, j))
    pairs.sort()
    contacts = []
    for _, _, i, j in pairs:
        gi, gj = model.geoms[i], model.geoms[j]
        # canonical type order: plane < sphere < box; geom1 is the lower type
        order = {"plane": 0, "sphere": 2, "box": 6}
        a, b = (i, j) if order[gi.type] <= order[gj.type] else (j, i)
        ga, gb = model.geoms[a], model.geoms[b]
        (pa, Ra), (pb, Rb) = gw[a], gw[b]
        if ga.type == "plane" and gb.type == "sphere":
            res = _plane_sphere(pa, Ra, pb, gb.size[0])
        elif ga.type == "plane" and gb.type == "box":
            res = _plane_box(pa, Ra, pb, Rb, gb.size)
        elif ga.type == "sphere" and gb.type == "sphere":
            res = _sphere_sphere(pa, ga.size[0], pb, gb.size[0])
        else:
            res = []          # plane-plane, box-box ...: not on the reference's path
        for dist, pos, n in res:
            contacts.append(_Contact(dist, pos, n, a, b))
    data.contact = contacts
    data.ncon = len(contacts)


def mj_step(model, data):  # only compare_builtin uses MuJoCo's own solver: out of scope
    raise NotImplementedError("fake mujoco: mj_step (soft-contact solver) is not modelled")


# ------------------------------------------------------------------- viewer stubs
class MjvCamera:
    def __init__(self):
        self.azimuth = 90.0
        self.elevation = -45.0
        self.distance = 5.0
        self.lookat = np.zeros(3)


class MjvOption:
    pass


class MjvScene:
    def __init__(self, model=None, maxgeom=1000):
        self.maxgeom = maxgeom


class MjrContext:
    def __init__(self, model=None, fontscale=150):
        pass


class MjrRect:
    def __init__(self, left, bottom, width, height):
        self.left, self.bottom, self.width, self.height = left, bottom, width, height


def mjv_defaultCamera(cam):
    pass


def mjv_defaultOption(opt):
    pass


def mjv_moveCamera(*a, **k):
    pass


def mjv_updateScene(*a, **k):
    pass


def mjr_render(*a, **k):
    pass


def mjr_readPixels(*a, **k):
    pass
