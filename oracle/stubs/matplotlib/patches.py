"""No-op patches stand-in (see matplotlib/__init__.py)."""
class Ellipse:
    def __init__(self, *a, **k):
        pass
class Patch:
    def __init__(self, *a, **k):
        pass
