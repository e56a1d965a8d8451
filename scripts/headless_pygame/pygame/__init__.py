"""Headless stand-in for the part of pygame the reference demos use, so that
visualization.py and robot-visualization.py can run UNMODIFIED on a machine without a display or pygame:
drawing calls are no-ops, the event queue is scripted (QUIT after PYGAME_STUB_FRAMES frames, optional held
keys), Clock.tick really sleeps (the robot demo's worker needs wall-clock time to answer).
Tooling for scripts/run_demo.py and the tests - not part of the engine."""
import os
import time as _time

QUIT, KEYDOWN = 256, 768
K_LEFT, K_RIGHT, K_UP, K_DOWN, K_e, K_c, K_ESCAPE = 1073741904, 1073741903, 1073741906, 1073741905, 101, 99, 27
SRCALPHA, RESIZABLE = 65536, 16

_state = {"frames": 0, "limit": int(os.environ.get("PYGAME_STUB_FRAMES", "30")),
          "tick_s": float(os.environ.get("PYGAME_STUB_TICK", "0.05")),
          "held": {int(k) for k in os.environ.get("PYGAME_STUB_KEYS", str(K_UP)).split(",") if k}, "calls": {}}


def _count(name):
    _state["calls"][name] = _state["calls"].get(name, 0) + 1


def init():
    _count("init")


def quit():  # noqa: A001
    _count("quit")


class Rect:
    def __init__(self, x, y, w, h):
        self.x, self.y, self.w, self.h = x, y, w, h
        self.topleft, self.topright = (x, y), (x + w, y)
        self.bottomleft, self.bottomright = (x, y + h), (x + w, y + h)
        self.center = (x + w // 2, y + h // 2)


class Surface:
    def __init__(self, size=(1, 1), flags=0):
        self.size = tuple(size)

    def fill(self, *a, **k):
        _count("fill")

    def blit(self, *a, **k):
        _count("blit")

    def set_alpha(self, *a):
        pass

    def get_rect(self, **k):
        r = Rect(0, 0, *self.size)
        if "center" in k:
            r.center = k["center"]
        return r

    def get_size(self):
        return self.size

    def get_width(self):
        return self.size[0]

    def get_height(self):
        return self.size[1]


class _Event:
    def __init__(self, type_, key=None):
        self.type, self.key = type_, key


class _Keys(dict):
    def __getitem__(self, k):
        return k in _state["held"]


class _Clock:
    def tick(self, fps=0):
        _time.sleep(_state["tick_s"])
        return int(1000 * _state["tick_s"])


class display:
    @staticmethod
    def set_mode(size=(1, 1), flags=0):
        return Surface(size)

    @staticmethod
    def set_caption(*a):
        pass

    @staticmethod
    def flip():
        _count("flip")

    @staticmethod
    def update(*a):
        _count("flip")


class time:
    Clock = _Clock


class event:
    @staticmethod
    def get():
        _state["frames"] += 1
        if _state["frames"] > _state["limit"]:
            return [_Event(QUIT)]
        if _state["frames"] % 7 == 0:
            return [_Event(KEYDOWN, K_RIGHT)]     # the static viewer steps through its iterations on arrow keys
        return []


class key:
    @staticmethod
    def get_pressed():
        return _Keys()


class draw:
    @staticmethod
    def circle(*a, **k):
        _count("draw")

    line = rect = ellipse = circle


class transform:
    @staticmethod
    def rotate(surf, angle):
        return surf


class font:
    class Font:
        def __init__(self, *a):
            pass

        def render(self, text, aa, color, *a):
            _count("text")
            return Surface((8 * len(str(text)), 16))

    @staticmethod
    def init():
        pass

    SysFont = Font
