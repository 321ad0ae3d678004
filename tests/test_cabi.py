"""The C-ABI shared library loads without a GPU and exports exactly what include/rbsim_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rbsim_b200.h")).read()
    return sorted(set(re.findall(r"RBS_API\s+[\w\s\*]+?\b(rbs_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    from rigidbody_simulation_b200 import _lib
    assert declared_symbols() == sorted(_lib.PROTOTYPES)
    assert len(declared_symbols()) == 25


def test_library_exports_every_declared_symbol():
    from rigidbody_simulation_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    loaded = _lib.load()
    assert loaded.rbs_version() == 1
    assert loaded.rbs_last_error() is not None
    assert loaded.rbs_launch_count() >= 0


def test_struct_layout_matches_header():
    """sizeof of the argument structs as the C compiler sees them == ctypes' view."""
    import subprocess
    import tempfile
    from rigidbody_simulation_b200 import _lib
    src = ('#include <stdio.h>\n#include "rbsim_b200.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(rbs_body_plane_args),'
           ' sizeof(rbs_two_ball_args), sizeof(rbs_multi_sphere_args)); return 0;}\n')
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(tmp, "t")
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_lib.BodyPlaneArgs), ctypes.sizeof(_lib.TwoBallArgs), ctypes.sizeof(_lib.MultiSphereArgs)]


def test_integration_doc_stub_matches_the_abi():
    """The ctypes stub INTEGRATION.md tells a maintainer to paste has the field order and size of the real struct."""
    from rigidbody_simulation_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    start = text.index("class BodyPlaneArgs(ctypes.Structure)")
    block = text[start:text.index("\n\n", start)]
    ns = {"ctypes": ctypes}
    exec(block, ns)
    doc = ns["BodyPlaneArgs"]
    assert [f[0] for f in doc._fields_] == [f[0] for f in _lib.BodyPlaneArgs._fields_]
    assert ctypes.sizeof(doc) == ctypes.sizeof(_lib.BodyPlaneArgs)
    for name, *_ in doc._fields_:
        assert getattr(doc, name).offset == getattr(_lib.BodyPlaneArgs, name).offset, name


def test_argument_validation_needs_no_gpu():
    from rigidbody_simulation_b200 import _lib
    lib = _lib.load()
    a = _lib.BodyPlaneArgs()
    a.dtype = 9
    assert lib.rbs_step_body_plane(ctypes.byref(a)) == _lib.RBS_EINVAL
    assert lib.rbs_step_body_plane(None) == _lib.RBS_EINVAL
    assert lib.rbs_step_two_ball(None) == _lib.RBS_EINVAL
    m = _lib.MultiSphereArgs()
    m.dtype, m.n_body, m.substeps = 1, 4096, 1
    assert lib.rbs_step_multi_sphere(ctypes.byref(m)) == _lib.RBS_EINVAL and b"n_body" in lib.rbs_last_error()
    assert lib.rbs_impulse_friction(5, 1, None, 1.0, None, None, None, None, None, 1.0, None, 0.5, None, None, None, None) == _lib.RBS_EINVAL
    with pytest.raises(ValueError):
        _lib.check(_lib.RBS_EINVAL)


def _build_c_client(tmp_path):
    import subprocess
    from rigidbody_simulation_b200 import _lib
    exe = str(tmp_path / "cabi_client")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["/usr/bin/gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cabi_client.c"),
                    "-o", exe, "-L", libdir, "-lrbsim_b200", "-Wl,-rpath," + libdir], check=True)
    return exe


def test_plain_c_client_links_and_validates(tmp_path):
    """A C program that includes only include/rbsim_b200.h links against the library and gets the documented
    status codes (no GPU involved)."""
    import subprocess
    _lib_ok = test_library_exports_every_declared_symbol  # noqa: F841  (ensures the library is built)
    exe = _build_c_client(tmp_path)
    r = subprocess.run([exe, "validate"], capture_output=True, text=True)
    assert r.returncode == 0 and "validate ok" in r.stdout, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_plain_c_client_runs_config1(tmp_path, golden):
    """configs[0] (single sphere from rest, 2000 steps) driven from C through rbs_run_body_plane_host."""
    import subprocess
    exe = _build_c_client(tmp_path)
    r = subprocess.run([exe, "sphere"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    vals = [float(x) for x in r.stdout.split()]
    assert vals[0] == 0.0 and vals[1] == 0.0 and vals[3:7] == [1.0, 0.0, 0.0, 0.0]
    assert vals[2] == pytest.approx(0.20844281676054977, rel=1e-12)       # SURVEY Appendix B, from-rest variant
    assert vals[7] == pytest.approx(0.1051418877747299, rel=1e-11)


@pytest.mark.gpu
def test_integration_doc_stub_runs(monkeypatch):
    """Section 2 of INTEGRATION.md, pasted as is: its step_many advances host qpos / qvel like the oracle does."""
    import sys
    import types
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle as co
    from rigidbody_simulation_b200 import synth
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    start = text.index("```python", text.index("## 2. Call the C ABI directly")) + len("```python")
    block = text[start:text.index("```", start)]
    monkeypatch.chdir(ROOT)
    ns = {}
    exec(block, ns)
    E = 3000
    s = synth.sphere_incline(E)
    m, inertia = 50 * 4 / 3 * np.pi * 0.2 ** 3, 0.4 * (50 * 4 / 3 * np.pi * 0.2 ** 3) * 0.04
    model = types.SimpleNamespace(body_mass=[0.0, m], body_inertia=[[0.0] * 3, [inertia] * 3],
                                  opt=types.SimpleNamespace(gravity=[0.0, 0.0, -9.8], timestep=0.009))
    qp, qv = s["qpos"].copy(), s["qvel"].copy()
    ns["step_many"](model, qp, qv, 100, 0.8, 0.4)
    rq, rv = s["qpos"].copy(), s["qvel"].copy()
    co.step_body_plane(rq, rv, 100, geom="sphere", mass=m, inertia=[inertia] * 3, size=0.2, plane_pos=[0, 0, 0],
                       plane_normal=[0, 0, 1], gravity=[0, 0, -9.8], dt=0.009, restitution=0.8, friction=0.4, threshold=0.0)
    assert np.max(np.abs(qp - rq) / np.maximum(np.abs(rq), 1e-3)) <= 1e-10
    assert np.max(np.abs(qv - rv) / np.maximum(np.abs(rv), 1e-3)) <= 1e-10
