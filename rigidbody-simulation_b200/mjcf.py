"""Host-side parser for the MuJoCo-XML subset the reference's scenes use (models/*.xml).

Parsed ONCE on the host into plain numbers; nothing here runs per step.  The only fields that reach
the reference's custom physics are:

  option/@gravity, option/@timestep; for each geom: type (plane | sphere | box), size, density, pos,
  euler / quat; for each body: name, pos, euler / quat and whether it has a free joint.

Rules reproduced from the MuJoCo XML compiler: ``compiler/@angle`` (MuJoCo's default is degree; every reference
scene says radian) and ``@eulerseq`` (default "xyz"; lower case = rotating frame, upper case = fixed frame); the
orientation specifiers quat / axisangle / xyaxes / zaxis / euler; body ids are world = 0 then ``<body>`` elements in
document order; ``inertiafromgeom`` (true: always from the geoms, as in every reference scene; auto: an
``<inertial>`` element wins when present; false: ``<inertial>`` required) with the geom's density or explicit mass:
sphere m = rho*4/3*pi*r^3, I = 2/5 m r^2; box (half sizes a,b,c) m = 8*rho*a*b*c,
I = m/3 * (b^2+c^2, a^2+c^2, a^2+b^2); top-level ``<default><geom .../></default>`` attributes.  solref / solimp /
damping / friction attributes belong to MuJoCo's own solver, which the reference's path never runs, and are ignored.
A geom may sit at an offset (pos / orientation) in its free body's frame: mass and principal moments are the geom's own
(about its centre, as ``inertiafromgeom`` computes them), the reference's step keeps taking the impulse arm from the
body origin qpos[:3] (collision.py:75), and only ``stepper.step_multi_body`` generates contacts for such a geom -- the
single-body, two-ball and multi-sphere steppers raise ValueError (``BatchedModel.has_offset_geoms``).
Outside the subset (rejected with ValueError, never approximated): other geom types, nested bodies, default classes,
inertial frames offset from their free body's origin, several geoms on a free body.

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
    mass: Optional[float] = None      # explicit geom mass (overrides density)


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


def euler_to_quat(euler, seq="xyz"):
    """MuJoCo's eulerseq: one letter per rotation; lower case rotates about the axes of the rotating frame
    (q <- q * r), upper case about the fixed frame (q <- r * q)."""
    q = [1.0, 0.0, 0.0, 0.0]
    if len(seq) != 3 or any(ch not in "xyzXYZ" for ch in seq):
        raise ValueError(f"bad eulerseq {seq!r}")
    for ch, angle in zip(seq, euler):
        elem = [math.cos(0.5 * angle), 0.0, 0.0, 0.0]
        elem[1 + "xyz".index(ch.lower())] = math.sin(0.5 * angle)
        q = quat_mul(q, elem) if ch.islower() else quat_mul(elem, q)
    return q


def mat_to_quat(cols):
    """Unit quaternion (wxyz, w >= 0 branch of Shepperd's method) of the rotation whose COLUMNS are ``cols``."""
    (r00, r10, r20), (r01, r11, r21), (r02, r12, r22) = cols
    tr = r00 + r11 + r22
    if tr > 0:
        s4 = 2.0 * math.sqrt(tr + 1.0)
        q = [0.25 * s4, (r21 - r12) / s4, (r02 - r20) / s4, (r10 - r01) / s4]
    elif r00 > r11 and r00 > r22:
        s4 = 2.0 * math.sqrt(1.0 + r00 - r11 - r22)
        q = [(r21 - r12) / s4, 0.25 * s4, (r01 + r10) / s4, (r02 + r20) / s4]
    elif r11 > r22:
        s4 = 2.0 * math.sqrt(1.0 + r11 - r00 - r22)
        q = [(r02 - r20) / s4, (r01 + r10) / s4, 0.25 * s4, (r12 + r21) / s4]
    else:
        s4 = 2.0 * math.sqrt(1.0 + r22 - r00 - r11)
        q = [(r10 - r01) / s4, (r02 + r20) / s4, (r12 + r21) / s4, 0.25 * s4]
    nrm = math.sqrt(sum(c * c for c in q))
    return [c / nrm for c in q]


def _unit(v, what):
    nrm = math.sqrt(sum(c * c for c in v))
    if nrm < 1e-14:
        raise ValueError(f"zero-length {what} in XML")
    return [c / nrm for c in v]


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def quat_to_mat(q):
    w, x, y, z = q
    return [[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]]


class _Compiler:
    """compiler/@angle, @eulerseq, @inertiafromgeom"""

    def __init__(self, el):
        get = (lambda k, d: d) if el is None else (lambda k, d: el.get(k, d))
        angle = get("angle", "degree")                 # MuJoCo's default unit is the degree
        if angle not in ("degree", "radian"):
            raise ValueError(f"compiler angle={angle!r}")
        self.to_rad = math.pi / 180.0 if angle == "degree" else 1.0
        self.eulerseq = get("eulerseq", "xyz")
        self.inertiafromgeom = get("inertiafromgeom", "auto")
        if self.inertiafromgeom not in ("true", "false", "auto"):
            raise ValueError(f"compiler inertiafromgeom={self.inertiafromgeom!r}")
        if get("coordinate", "local") != "local":
            raise ValueError("only coordinate=\"local\" scenes are supported")


def _orientation(el, comp, attrs=None):
    """quat | axisangle | xyaxes | zaxis | euler -> unit quaternion wxyz (at most one may be given, as in MuJoCo)."""
    get = el.get if attrs is None else (lambda k: el.get(k, attrs.get(k)))
    given = [k for k in ("quat", "axisangle", "xyaxes", "zaxis", "euler") if get(k) is not None]
    if len(given) > 1:
        raise ValueError(f"several orientation specifiers on <{el.tag}>: {given}")
    if not given:
        return [1.0, 0.0, 0.0, 0.0]
    kind = given[0]
    if kind == "quat":
        return _unit(_vec(get("quat"), 4, [1.0, 0.0, 0.0, 0.0]), "quaternion")
    if kind == "axisangle":
        x, y, z, ang = _vec(get("axisangle"), 4, [0.0, 0.0, 1.0, 0.0])
        ax = _unit([x, y, z], "axisangle axis")
        h = 0.5 * ang * comp.to_rad
        return [math.cos(h)] + [math.sin(h) * c for c in ax]
    if kind == "xyaxes":
        v = _vec(get("xyaxes"), 6, [1.0, 0.0, 0.0, 0.0, 1.0, 0.0])
        x = _unit(v[:3], "xyaxes x")
        d = sum(a * b for a, b in zip(x, v[3:]))
        y = _unit([b - d * a for a, b in zip(x, v[3:])], "xyaxes y")       # made orthogonal to x, as MuJoCo does
        return mat_to_quat((x, y, _cross(x, y)))
    if kind == "zaxis":
        z = _unit(_vec(get("zaxis"), 3, [0.0, 0.0, 1.0]), "zaxis")
        # minimal rotation taking (0,0,1) to z
        ax = _cross([0.0, 0.0, 1.0], z)
        s_, c_ = math.sqrt(sum(a * a for a in ax)), z[2]
        if s_ < 1e-14:
            return [1.0, 0.0, 0.0, 0.0] if c_ > 0 else [0.0, 1.0, 0.0, 0.0]
        h = 0.5 * math.atan2(s_, c_)
        return [math.cos(h)] + [math.sin(h) * a / s_ for a in ax]
    return euler_to_quat([a * comp.to_rad for a in _vec(get("euler"), 3, [0.0, 0.0, 0.0])], comp.eulerseq)


def geom_mass_inertia(g):
    """inertiafromgeom: uniform density, or the geom's explicit ``mass`` (which overrides the density)."""
    if g.type == "sphere":
        r = g.size[0]
        m = g.density * (4.0 / 3.0) * math.pi * r ** 3 if g.mass is None else g.mass
        i = 0.4 * m * r * r
        return m, [i, i, i]
    if g.type == "box":
        a, b, c = g.size
        m = g.density * 8.0 * a * b * c if g.mass is None else g.mass
        return m, [m / 3.0 * (b * b + c * c), m / 3.0 * (a * a + c * c), m / 3.0 * (a * a + b * b)]
    return 0.0, [0.0, 0.0, 0.0]


def parse_string(text):
    root = ET.fromstring(text)
    if root.tag != "mujoco":
        raise ValueError("not a MuJoCo XML document")
    comp = _Compiler(root.find("compiler"))
    gdef = {}
    dflt = root.find("default")
    if dflt is not None:
        if dflt.find("default") is not None:
            raise ValueError("default classes are outside the supported subset")
        if dflt.find("geom") is not None:
            gdef = dict(dflt.find("geom").attrib)
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
        if el.get("class") is not None:
            raise ValueError("default classes are outside the supported subset")
        gtype = el.get("type", gdef.get("type", "sphere"))
        if gtype not in GEOM_TYPES:
            raise ValueError(f"geom type {gtype!r} is outside the supported subset {GEOM_TYPES}")
        if el.get("fromto", gdef.get("fromto")) is not None:
            raise ValueError("geom fromto is outside the supported subset")
        mass = el.get("mass", gdef.get("mass"))
        g = Geom(el.get("name"), gtype, _vec(el.get("size", gdef.get("size")), 3, [0.0, 0.0, 0.0]),
                 float(el.get("density", gdef.get("density", 1000.0))), _vec(el.get("pos", gdef.get("pos")), 3, [0.0, 0.0, 0.0]),
                 _orientation(el, comp, gdef), owner, None if mass is None else float(mass))
        if gtype != "plane" and not g.size[0] > 0.0:
            raise ValueError(f"geom {g.name!r}: size must be positive")
        if gtype == "box" and not (g.size[1] > 0.0 and g.size[2] > 0.0):
            raise ValueError(f"box geom {g.name!r} needs three positive half sizes")
        bodies[owner].geoms.append(len(geoms))
        geoms.append(g)
        return g

    for el in world.findall("geom"):
        add_geom(el, 0)
    for el in world.findall("body"):
        if el.find("body") is not None:
            raise ValueError("nested bodies are outside the supported subset")
        free = el.find("freejoint") is not None or any(j.get("type") == "free" for j in el.findall("joint"))
        if any(j.get("type", "hinge") != "free" for j in el.findall("joint")):
            raise ValueError("only free joints are inside the supported subset")
        body = Body(el.get("name"), _vec(el.get("pos"), 3, [0.0, 0.0, 0.0]), _orientation(el, comp), free)
        bodies.append(body)
        owner = len(bodies) - 1
        for gel in el.findall("geom"):
            g = add_geom(gel, owner)
            m, inertia = geom_mass_inertia(g)
            # a geom offset from its free body's origin is kept (Geom.pos / Geom.quat in the body frame); only the
            # multi-body stepper (stepper.step_multi_body) places geoms that way, the others refuse such a scene
            body.mass += m
            body.inertia = [a + b for a, b in zip(body.inertia, inertia)]
        if free and len(body.geoms) != 1:
            raise ValueError("a free body must carry exactly one geom (as in every reference scene)")
        inertial = el.find("inertial")
        if comp.inertiafromgeom == "false" and free and inertial is None:
            raise ValueError("inertiafromgeom=\"false\" needs an <inertial> element on every free body")
        if inertial is not None and comp.inertiafromgeom != "true":       # "true" overrides <inertial> with the geoms
            if any(_vec(inertial.get("pos"), 3, [0.0, 0.0, 0.0])) or _orientation(inertial, comp) != [1.0, 0.0, 0.0, 0.0]:
                raise ValueError("an inertial frame offset from the body origin is outside the supported subset")
            if inertial.get("fullinertia") is not None:
                raise ValueError("fullinertia is outside the supported subset (give diaginertia in the body frame)")
            body.mass = float(inertial.get("mass"))
            body.inertia = _vec(inertial.get("diaginertia"), 3, [0.0, 0.0, 0.0])
            if not (body.mass > 0.0 and all(i > 0.0 for i in body.inertia)):
                raise ValueError("<inertial> needs a positive mass and diaginertia")
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
