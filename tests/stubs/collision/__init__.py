"""import-only stub (test infrastructure): the real package is not installed and is not on the step path.
robogym's rearrange placement code imports Poly / Vector / collide from it; object placement is outside this round's
scope, so they refuse to run."""


class _Unsupported:
    def __init__(self, *a, **k):
        raise NotImplementedError("the `collision` package is not installed (rearrange placement is out of scope)")


class Poly(_Unsupported):
    pass


class Vector(_Unsupported):
    pass


def collide(*a, **k):
    raise NotImplementedError("the `collision` package is not installed (rearrange placement is out of scope)")
