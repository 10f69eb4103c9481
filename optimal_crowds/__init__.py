"""Import alias so that the reference's own lines work unedited on the B200 implementation:

    from optimal_crowds import simulations            # README.md:32
    from optimal_crowds import pedestrians, optimals  # simulations.py:10-11

The reference is a namespace directory called ``optimal_crowds`` whose parent is the working directory, and it
reads ``optimal_crowds/config.json`` and ``rooms/<room>.json`` relative to that working directory
(simulations.py:42,47).  Running from this repository's root reproduces that layout: this package maps the three
module names onto ``optimal_crowds_b200`` (the CUDA-backed implementation; there is no second implementation and no
CPU fallback behind this alias) and ``config.json`` next to this file is the reference's file, verbatim.
"""
import sys as _sys

from optimal_crowds_b200 import optimals, pedestrians, simulations  # noqa: F401

for _name, _mod in (("simulations", simulations), ("optimals", optimals), ("pedestrians", pedestrians)):
    _sys.modules[__name__ + "." + _name] = _mod
