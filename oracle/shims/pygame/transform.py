"""pygame.transform.scale / rotate restated for sizes only (src_c/transform.c, pygame 2.1.2)."""
import math
import struct

from .surface import Surface


def scale(surface, size, dest=None):
    # "(ii)" argument parsing: floats truncate (python 3.7 behaviour, conda.yml:7 pins 3.7.12).
    return Surface((int(size[0]), int(size[1])))


def _c_float(v):
    return struct.unpack("f", struct.pack("f", float(v)))[0]


def rotated_size(w, h, angle):
    """Size of the surface surf_rotate() would allocate for a w x h source."""
    angle = _c_float(angle)                      # "f" format: C float
    if math.fmod(angle, 90.0) == 0.0:
        # C: numturns = ((int)angle / 90) % 4, shifted by +4 when negative == python's % on the
        # (exact) quotient.
        numturns = int(int(angle) / 90) % 4
        if numturns % 2:
            return h, w
        return w, h
    rad = angle * .01745329251994329
    s, c = math.sin(rad), math.cos(rad)
    cx, cy, sx, sy = c * w, c * h, s * w, s * h
    nx = int(max(abs(cx + sy), abs(cx - sy), abs(-cx + sy), abs(-cx - sy)))
    ny = int(max(abs(sx + cy), abs(sx - cy), abs(-sx + cy), abs(-sx - cy)))
    return nx, ny


def rotate(surface, angle):
    w, h = rotated_size(surface.get_width(), surface.get_height(), angle)
    return Surface((w, h))
