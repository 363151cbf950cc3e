"""Integer rectangle with pygame 2.1.2 semantics (src_c/rect.c) -- see package docstring."""


def _c_int(v):
    """pg_IntFromObj: floats (and numpy scalars via __int__) truncate toward zero."""
    return int(v)


class Rect:
    __slots__ = ("x", "y", "w", "h")

    def __init__(self, *args):
        if len(args) == 1:
            a = args[0]
            if isinstance(a, Rect):
                args = (a.x, a.y, a.w, a.h)
            else:
                args = tuple(a)
        if len(args) == 2:
            (x, y), (w, h) = args
        else:
            x, y, w, h = args
        self.x, self.y, self.w, self.h = _c_int(x), _c_int(y), _c_int(w), _c_int(h)

    # --- scalar attributes -------------------------------------------------------------------
    width = property(lambda s: s.w, lambda s, v: setattr(s, "w", _c_int(v)))
    height = property(lambda s: s.h, lambda s, v: setattr(s, "h", _c_int(v)))
    left = property(lambda s: s.x, lambda s, v: setattr(s, "x", _c_int(v)))
    top = property(lambda s: s.y, lambda s, v: setattr(s, "y", _c_int(v)))
    right = property(lambda s: s.x + s.w, lambda s, v: setattr(s, "x", _c_int(v) - s.w))
    bottom = property(lambda s: s.y + s.h, lambda s, v: setattr(s, "y", _c_int(v) - s.h))
    centerx = property(lambda s: s.x + (s.w >> 1), lambda s, v: setattr(s, "x", _c_int(v) - (s.w >> 1)))
    centery = property(lambda s: s.y + (s.h >> 1), lambda s, v: setattr(s, "y", _c_int(v) - (s.h >> 1)))

    # --- point attributes --------------------------------------------------------------------
    @property
    def center(self):
        return (self.x + (self.w >> 1), self.y + (self.h >> 1))

    @center.setter
    def center(self, v):
        cx, cy = v
        self.x = _c_int(cx) - (self.w >> 1)
        self.y = _c_int(cy) - (self.h >> 1)

    @property
    def size(self):
        return (self.w, self.h)

    @size.setter
    def size(self, v):
        self.w, self.h = _c_int(v[0]), _c_int(v[1])

    topleft = property(lambda s: (s.x, s.y))
    topright = property(lambda s: (s.x + s.w, s.y))
    bottomleft = property(lambda s: (s.x, s.y + s.h))
    bottomright = property(lambda s: (s.x + s.w, s.y + s.h))
    midtop = property(lambda s: (s.x + (s.w >> 1), s.y))
    midbottom = property(lambda s: (s.x + (s.w >> 1), s.y + s.h))
    midleft = property(lambda s: (s.x, s.y + (s.h >> 1)))
    midright = property(lambda s: (s.x + s.w, s.y + (s.h >> 1)))

    # --- methods -----------------------------------------------------------------------------
    def copy(self):
        return Rect(self.x, self.y, self.w, self.h)

    def move_ip(self, *args):
        if len(args) == 1:
            dx, dy = args[0]
        else:
            dx, dy = args
        self.x += _c_int(dx)
        self.y += _c_int(dy)

    def move(self, *args):
        r = self.copy()
        r.move_ip(*args)
        return r

    def collidepoint(self, *args):
        if len(args) == 1:
            px, py = args[0]
        else:
            px, py = args
        px, py = _c_int(px), _c_int(py)
        return self.x <= px < self.x + self.w and self.y <= py < self.y + self.h

    def colliderect(self, other):
        o = other if isinstance(other, Rect) else Rect(other)
        if self.w == 0 or self.h == 0 or o.w == 0 or o.h == 0:
            return False
        return (self.x < o.x + o.w and self.y < o.y + o.h and
                self.x + self.w > o.x and self.y + self.h > o.y)

    def collidelist(self, rects):
        for i, r in enumerate(rects):
            if self.colliderect(r):
                return i
        return -1

    def __iter__(self):
        return iter((self.x, self.y, self.w, self.h))

    def __len__(self):
        return 4

    def __getitem__(self, i):
        return (self.x, self.y, self.w, self.h)[i]

    def __eq__(self, other):
        try:
            return tuple(self) == tuple(other)
        except TypeError:
            return False

    def __repr__(self):
        return "<rect(%d, %d, %d, %d)>" % (self.x, self.y, self.w, self.h)
