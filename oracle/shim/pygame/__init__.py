"""pygame stand-in so the reference env modules import; render=True is unsupported (oracle only)."""


def init():
    raise RuntimeError("pygame stand-in: rendering is not available in the oracle")


class _NoRender(object):
    def __getattr__(self, name):
        raise RuntimeError("pygame stand-in: rendering is not available in the oracle")


display = _NoRender()
time = _NoRender()
draw = _NoRender()
surfarray = _NoRender()
image = _NoRender()
