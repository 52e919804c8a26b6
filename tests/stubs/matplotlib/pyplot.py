"""matplotlib.pyplot stand-in: every attribute is a function that does nothing (reach_helper.py's optional plotting)."""


def __getattr__(name):
    def _noop(*args, **kwargs):
        return None

    return _noop
