"""Host-side parser for the MuJoCo-XML subset the reference's scenes use (models/*.xml).

Parsed ONCE on the host into plain numbers; nothing here runs per step.  The only fields that reach
the reference's custom physics are:

  option/@gravity, option/@timestep; for each geom: type (plane | sphere | box), size, density, pos,
  euler / quat; for each body: name, pos, euler / quat and whether it has a free joint.

Rules reproduced from the MuJoCo XML compiler: ``angle="radian"``; euler is intrinsic x-y-z; body ids
are world = 0 then ``<body>`` elements in document order; ``inertiafromgeom`` with the geom's density:
sphere m = rho*4/3*pi*r^3, I = 2/5 m r^2; box (half sizes a,b,c) m = 8*rho*a*b*c,
I = m/3 * (b^2+c^2, a^2+c^2, a^2+b^2).  solref / solimp / damping / friction attributes belong to
MuJoCo's own solver, which the reference's path never runs, and are ignored.

The reference templates ``{INCLINE_ANGLE}`` / ``{TIMESTEP}`` into the XML text before compiling
(src/simulation/single_sphere_bounce.py:29-30, cube_incline.py:33-34); ``render_template`` does that.
"""
import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import List, Optional

GEOM_TYPES = ("plane", "sphere", "box")


@dataclass
class Geom:
    name: Optional[str]
    type: str
    size: List[float]
    density: float
    pos: List[float]
    quat: List[float]          # wxyz, relative to the owning body
    body: int


@dataclass
class Body:
    name: Optional[str]
    pos: List[float]
    quat: List[float]          # wxyz
    free: bool
    geoms: List[int] = field(default_factory=list)
    mass: float = 0.0
    inertia: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])


@dataclass
class Scene:
    gravity: List[float]
    timestep: float
    bodies: List[Body]         # index = MuJoCo body id; bodies[0] is the world
    geoms: List[Geom]          # index = MuJoCo geom id (world geoms first)

    @property
    def free_bodies(self):
        return [i for i, b in enumerate(self.bodies) if b.free]

    def body_id(self, name):
        """mj_name2id(model, mjOBJ_BODY, name): -1 when absent (relied upon by
        src/simulation/single_sphere_bounce.py:67, where "sphere" is not a body of sphere.xml)."""
        for i, b in enumerate(self.bodies):
            if b.name == name:
                return i
        return -1

    def planes(self):
        return [g for g in self.geoms if g.type == "plane"]


def render_template(text, incline_angle=None, timestep=None):
    if incline_angle is not None:
        text = text.replace("{INCLINE_ANGLE}", str(incline_angle))
    if timestep is not None:
        text = text.replace("{TIMESTEP}", str(timestep))
    return text


def _vec(text, n, default):
    if text is None:
        return list(default)
    vals = [float(t) for t in text.split()]
    if len(vals) > n:
        raise ValueError(f"expected at most {n} numbers, got {text!r}")
    return vals + list(default[len(vals):])


def quat_mul(a, b):
    """Hamilton product, wxyz (same convention as mju_mulQuat)."""
    return [a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
            a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
            a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]]


def euler_to_quat(euler):
    q = [1.0, 0.0, 0.0, 0.0]
    for axis, angle in enumerate(euler):
        elem = [math.cos(0.5 * angle), 0.0, 0.0, 0.0]
        elem[1 + axis] = math.sin(0.5 * angle)
        q = quat_mul(q, elem)
    return q


def quat_to_mat(q):
    w, x, y, z = q
    return [[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]]


def _orientation(el):
    if el.get("quat") is not None:
        q = _vec(el.get("quat"), 4, [1.0, 0.0, 0.0, 0.0])
        nrm = math.sqrt(sum(c * c for c in q))
        if nrm == 0.0:
            raise ValueError("zero quaternion in XML")
        return [c / nrm for c in q]
    return euler_to_quat(_vec(el.get("euler"), 3, [0.0, 0.0, 0.0]))


def geom_mass_inertia(g):
    if g.type == "sphere":
        r = g.size[0]
        m = g.density * (4.0 / 3.0) * math.pi * r ** 3
        i = 0.4 * m * r * r
        return m, [i, i, i]
    if g.type == "box":
        a, b, c = g.size
        m = g.density * 8.0 * a * b * c
        return m, [m / 3.0 * (b * b + c * c), m / 3.0 * (a * a + c * c), m / 3.0 * (a * a + b * b)]
    return 0.0, [0.0, 0.0, 0.0]


def parse_string(text):
    root = ET.fromstring(text)
    if root.tag != "mujoco":
        raise ValueError("not a MuJoCo XML document")
    comp = root.find("compiler")
    if comp is not None and comp.get("angle", "radian") != "radian":
        raise ValueError("only angle=\"radian\" scenes are supported (all reference scenes use it)")
    gravity, timestep = [0.0, 0.0, -9.81], 0.002
    opt = root.find("option")
    if opt is not None:
        gravity = _vec(opt.get("gravity"), 3, gravity)
        timestep = float(opt.get("timestep", timestep))
    world = root.find("worldbody")
    if world is None:
        raise ValueError("scene has no <worldbody>")
    bodies = [Body("world", [0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0], False)]
    geoms: List[Geom] = []

    def add_geom(el, owner):
        gtype = el.get("type", "sphere")
        if gtype not in GEOM_TYPES:
            raise ValueError(f"geom type {gtype!r} is outside the supported subset {GEOM_TYPES}")
        g = Geom(el.get("name"), gtype, _vec(el.get("size"), 3, [0.0, 0.0, 0.0]), float(el.get("density", 1000.0)),
                 _vec(el.get("pos"), 3, [0.0, 0.0, 0.0]), _orientation(el), owner)
        bodies[owner].geoms.append(len(geoms))
        geoms.append(g)
        return g

    for el in world.findall("geom"):
        add_geom(el, 0)
    for el in world.findall("body"):
        if el.find("body") is not None:
            raise ValueError("nested bodies are outside the supported subset")
        free = el.find("freejoint") is not None or any(j.get("type") == "free" for j in el.findall("joint"))
        body = Body(el.get("name"), _vec(el.get("pos"), 3, [0.0, 0.0, 0.0]), _orientation(el), free)
        bodies.append(body)
        owner = len(bodies) - 1
        for gel in el.findall("geom"):
            g = add_geom(gel, owner)
            m, inertia = geom_mass_inertia(g)
            if free and m > 0.0 and (any(g.pos) or g.quat != [1.0, 0.0, 0.0, 0.0]):
                raise ValueError("offset geoms on a free body are outside the supported subset")
            body.mass += m
            body.inertia = [a + b for a, b in zip(body.inertia, inertia)]
        if free and len(body.geoms) != 1:
            raise ValueError("a free body must carry exactly one geom (as in every reference scene)")
    return Scene(gravity, timestep, bodies, geoms)


def parse_file(path, incline_angle=None, timestep=None):
    with open(path, "r") as f:
        return parse_string(render_template(f.read(), incline_angle, timestep))


def plane_frame(scene, geom):
    """World point and unit normal (z axis of the geom frame) of a plane geom."""
    body = scene.bodies[geom.body]
    rb = quat_to_mat(body.quat)
    point = [body.pos[i] + sum(rb[i][k] * geom.pos[k] for k in range(3)) for i in range(3)]
    rg = quat_to_mat(quat_mul(body.quat, geom.quat)) if body.quat != [1.0, 0.0, 0.0, 0.0] else quat_to_mat(geom.quat)
    return point, [rg[0][2], rg[1][2], rg[2][2]]
