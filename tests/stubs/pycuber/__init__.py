"""import-only stub (test infrastructure): pycuber is only needed by the Rubik's-cube goal logic of the
full-cube envs; the locked / reach envs import the module but never call it."""


class Cube:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError("pycuber is not installed")


class Cubie:  # pragma: no cover
    pass
