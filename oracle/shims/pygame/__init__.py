"""Minimal stand-in for pygame==2.1.2 -- TEST INFRASTRUCTURE ONLY.

Purpose: let the UNMODIFIED reference modules under /root/reference import and run head-less in
this container so that golden traces can be exported (oracle/gen_golden.py).  Nothing in the
product package imports this.

Only the integer-rectangle arithmetic the hot path depends on is restated (from the published
behaviour of pygame 2.1.2: src_c/rect.c, src_c/transform.c, src_c/base.c, src_c/time.c).  The
reference pins pygame==2.1.2 in requirements.txt:5; its sources are not vendored under
/root/reference, so these semantics are "parity unpinned" (SURVEY.md section 8(c)):

* ``Rect`` holds four C ints; float arguments truncate toward zero (pg_IntFromObj).
* ``center`` setter: x = int(cx) - (w >> 1), y = int(cy) - (h >> 1); getter adds (w >> 1, h >> 1).
* ``colliderect``: false when either rect has a zero dimension, else strict-inequality AABB overlap.
* ``collidepoint``: x <= px < x + w and y <= py < y + h.
* ``transform.rotate``: angle is parsed as a C float; multiples of 90 swap w/h on odd quarter turns;
  otherwise the new size is the truncated bounding box of the rotated w x h rectangle.
* time: a virtual clock advancing 1 ms per ``Clock.tick`` and reset by ``pygame.init`` (the
  reference keys the leader-trail append on ``get_ticks() % 5``,
  follow_the_leader_continuous_env.py:1074 -- wall-clock in the original, virtual here so that runs
  are reproducible; the CUDA path uses the same rule: append when frame_index % 5 == 0).
"""
import math
import struct

from . import transform, image, font, display, time, draw, event, joystick, surfarray  # noqa: F401
from .rect import Rect  # noqa: F401
from .surface import Surface  # noqa: F401

K_LEFT, K_RIGHT, K_UP, K_DOWN = 1, 2, 3, 4
KEYDOWN, KEYUP, QUIT = 10, 11, 12
JOYAXISMOTION = 13


def init():
    """pygame.init(): the reference calls it at every reset (ENV:517); restart the virtual clock."""
    time._reset()
    return (6, 0)


def quit():
    return None
