from .surface import Surface


def set_mode(size=(0, 0), *a, **k):
    return Surface(size)


def set_caption(*a, **k):
    return None


def update(*a, **k):
    return None


flip = update
