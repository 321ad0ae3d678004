"""Scene construction helpers: the four shipped scenes (models/*.xml) and their scaled variants."""
import os

import torch

from .batched import BatchedModel

MODELS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")


def model_path(name):
    return os.path.join(MODELS_DIR, name + ".xml")


def single_body_xml(geom, size, plane_euler=(0.0, 0.0, 0.0), body_pos=(0.0, 0.0, 1.0), body_euler=(0.0, 0.0, 0.0),
                    timestep=0.009, gravity=(0.0, 0.0, -9.8), density=50.0, name="obj"):
    """Same shape as models/sphere.xml / models/cube.xml: a plane body plus one free body."""
    fmt = lambda v: " ".join(repr(float(x)) for x in v)
    return (f'<mujoco><compiler angle="radian" inertiafromgeom="true"/>'
            f'<option gravity="{fmt(gravity)}" timestep="{float(timestep)!r}"/><worldbody>'
            f'<body name="inclined_plane" pos="0 0 0"><geom name="ground" type="plane" size="5 5 0.1" '
            f'euler="{fmt(plane_euler)}"/></body>'
            f'<body name="{name}" pos="{fmt(body_pos)}" euler="{fmt(body_euler)}"><joint name="j" type="free"/>'
            f'<geom name="g" type="{geom}" size="{fmt(size)}" density="{float(density)!r}"/></body>'
            f'</worldbody></mujoco>')


def multi_sphere_xml(n_body, radius=0.1, timestep=0.01, gravity=(0.0, 0.0, -9.8), density=50.0, plane_euler=(0.0, 0.0, 0.0)):
    """Same shape as models/multi_sphere.xml, scaled to ``n_body`` spheres ball1..ballN (``plane_euler`` tilts the ground)."""
    fmt = lambda v: " ".join(repr(float(x)) for x in v)
    balls = "".join(f'<body name="ball{i + 1}" pos="0 0 {1 + i}"><joint name="ball_joint{i + 1}" type="free"/>'
                    f'<geom name="ball_geom{i + 1}" type="sphere" size="{float(radius)!r}" density="{float(density)!r}"/>'
                    f'</body>' for i in range(n_body))
    return (f'<mujoco><compiler angle="radian" inertiafromgeom="true"/>'
            f'<option gravity="{fmt(gravity)}" timestep="{float(timestep)!r}"/><worldbody>'
            f'<geom name="ground" type="plane" size="5 5 0.1" euler="{fmt(plane_euler)}"/>{balls}</worldbody></mujoco>')


def multi_body_xml(bodies, timestep=0.005, gravity=(0.0, 0.0, -9.8), density=50.0, plane_euler=(0.0, 0.0, 0.0)):
    """A ground plane plus free bodies body1..bodyN for stepper.step_multi_body (N4).  ``bodies`` is a list of dicts:
    {"type": "sphere" | "box", "size": [...], optional "pos" (body), "geom_pos", "geom_euler" (geom in the body frame)}."""
    fmt = lambda v: " ".join(repr(float(x)) for x in v)
    out = []
    for i, b in enumerate(bodies):
        off = ""
        if "geom_pos" in b:
            off += f' pos="{fmt(b["geom_pos"])}"'
        if "geom_euler" in b:
            off += f' euler="{fmt(b["geom_euler"])}"'
        out.append(f'<body name="body{i + 1}" pos="{fmt(b.get("pos", (0, 0, 1 + i)))}"><joint name="joint{i + 1}" type="free"/>'
                   f'<geom name="geom{i + 1}" type="{b["type"]}" size="{fmt(b["size"])}" density="{float(density)!r}"{off}/></body>')
    return (f'<mujoco><compiler angle="radian" inertiafromgeom="true"/>'
            f'<option gravity="{fmt(gravity)}" timestep="{float(timestep)!r}"/><worldbody>'
            f'<geom name="ground" type="plane" size="5 5 0.1" euler="{fmt(plane_euler)}"/>{"".join(out)}</worldbody></mujoco>')


def sphere_on_incline(nenv, theta=0.7, device=None, dtype=torch.float64):
    """config 2 scene: the sphere of models/sphere.xml over a plane tilted ``theta`` rad about x."""
    return BatchedModel.from_xml_string(single_body_xml("sphere", [0.2], plane_euler=(theta, 0, 0), name="ball"),
                                        nenv=nenv, device=device, dtype=dtype)


def cube_on_plane(nenv, theta=0.7, device=None, dtype=torch.float64):
    """config 4 scene: the cube of models/cube.xml; theta = 0 for 'bounce', 0.7 for 'incline'."""
    return BatchedModel.from_xml_string(
        single_body_xml("box", [0.4, 0.4, 0.4], plane_euler=(theta, 0, 0), body_pos=(0, 0, 0.4),
                        body_euler=(theta, 0, 0), name="cube"), nenv=nenv, device=device, dtype=dtype)
