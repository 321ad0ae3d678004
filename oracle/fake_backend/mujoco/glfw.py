"""Fake ``mujoco.glfw`` -- test infrastructure. The reference does ``from mujoco.glfw import glfw``
(src/viewer/mujoco_viewer.py:4); the object it gets is the fake top-level ``glfw`` module."""
import glfw  # noqa: F401  (the fake one next to this package)
