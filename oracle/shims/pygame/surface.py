"""Size-only Surface (no pixels): the hot path only ever asks a surface for its rectangle."""
from .rect import Rect


class Surface:
    def __init__(self, size=(0, 0), *a, **k):
        self._w, self._h = int(size[0]), int(size[1])

    def get_width(self):
        return self._w

    def get_height(self):
        return self._h

    def get_size(self):
        return (self._w, self._h)

    def get_rect(self, **kwargs):
        # surface.c surf_get_rect: Rect(0, 0, w, h), then setattr for each kwarg in call order.
        r = Rect(0, 0, self._w, self._h)
        for k, v in kwargs.items():
            setattr(r, k, v)
        return r

    def convert(self, *a):
        return self

    convert_alpha = convert

    def fill(self, *a, **k):
        return None

    def blit(self, *a, **k):
        return None

    def copy(self):
        return Surface((self._w, self._h))
