import numpy as np


def array3d(surface):
    return np.zeros((surface.get_width(), surface.get_height(), 3), dtype=np.uint8)
