"""Fake ``mpl_toolkits.mplot3d`` -- test infrastructure."""


class Axes3D:
    pass
