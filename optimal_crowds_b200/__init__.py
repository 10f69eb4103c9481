"""optimal_crowds_b200 -- B200-native (sm_100a) drop-in for the two hot paths of matteobutano/optimal_crowds.

    from optimal_crowds_b200 import simulations
    simu = simulations.simulation('room_test', 30.0, recompute=False)
    simu.run(verbose=False, draw=False)

mirrors `from optimal_crowds import simulations` (reference README.md:31-35).  All heavy arithmetic runs in
liboc_b200.so (CUDA, C ABI in include/optimal_crowds.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
