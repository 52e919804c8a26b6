"""Stand-in for the `collision` package (not installed in this image; test infrastructure).  robogym's rearrange placement code
uses exactly Vector(x, y), Poly.from_box(center, width, height) and collide(a, b) on such axis-aligned rectangles
(robogym/envs/rearrange/common/utils.py:600-620), for which the separating-axis test is an interval-overlap test."""


class Vector:
    def __init__(self, x, y):
        self.x, self.y = float(x), float(y)


class Poly:
    def __init__(self, center, width, height):
        self.pos, self.w, self.h = center, float(width), float(height)

    @classmethod
    def from_box(cls, center, width, height):
        return cls(center, width, height)


def collide(a, b):
    return abs(a.pos.x - b.pos.x) * 2.0 < a.w + b.w and abs(a.pos.y - b.pos.y) * 2.0 < a.h + b.h
