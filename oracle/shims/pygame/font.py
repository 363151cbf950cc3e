from .surface import Surface


class _Font:
    def render(self, *a, **k):
        return Surface((1, 1))


def init():
    return None


def SysFont(*a, **k):
    return _Font()


Font = SysFont
