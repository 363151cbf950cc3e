"""Virtual clock: one millisecond per Clock.tick(), restarted by pygame.init() (package docstring)."""
_ticks = 0


def _reset():
    global _ticks
    _ticks = 0


def get_ticks():
    return _ticks


class Clock:
    def tick(self, framerate=0):
        global _ticks
        _ticks += 1
        return 1

    def get_fps(self):
        return 0.0
