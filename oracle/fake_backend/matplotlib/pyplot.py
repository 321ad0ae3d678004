"""Fake ``matplotlib.pyplot`` -- every call is accepted and ignored."""


class _Anything:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __iter__(self):
        return iter(())


_a = _Anything()


def __getattr__(name):
    return _a
