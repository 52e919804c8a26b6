"""import-only stub (test infrastructure): the real package is not installed and is not on the step path.
robogym's rearrange mesh utilities annotate with trimesh.Trimesh and call trimesh.load / sample / remesh / util;
mesh objects for rearrange are outside this round's scope, so everything refuses to run."""
import types


class Trimesh:
    def __init__(self, *a, **k):
        raise NotImplementedError("trimesh is not installed (rearrange mesh objects are out of scope)")


def load(*a, **k):
    raise NotImplementedError("trimesh is not installed (rearrange mesh objects are out of scope)")


sample = types.SimpleNamespace()
remesh = types.SimpleNamespace()
util = types.SimpleNamespace()
