"""No-op seaborn stand-in (see matplotlib/__init__.py)."""
def color_palette(*a, n_colors=1, **k):
    return [(0.0, 0.0, 0.0)] * int(n_colors)
