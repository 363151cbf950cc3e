def _noop(*a, **k):
    return None


circle = line = lines = rect = polygon = aaline = aalines = ellipse = arc = _noop
