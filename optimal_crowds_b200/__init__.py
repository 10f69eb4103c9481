"""optimal_crowds_b200 -- B200-native (sm_100a) drop-in for the two hot paths of matteobutano/optimal_crowds.

    from optimal_crowds_b200 import simulations
    simu = simulations.simulation('room_test', 30.0, recompute=False)
    simu.run(verbose=False, draw=False)

mirrors `from optimal_crowds import simulations` (reference README.md:31-35).  All heavy arithmetic runs in
liboc_b200.so (CUDA, C ABI in include/optimal_crowds.h); there is no CPU fallback.
"""
import os as _os

# Ensembles run up to 128 members on their own CUDA streams; the driver maps streams onto 8 hardware queues by default
# (measured: member sweeps overlap 8-fold and no further).  Must be set before the CUDA context exists; an explicit
# setting of the user wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

__version__ = "0.1.0"
