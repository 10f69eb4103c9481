"""No-op stand-in for matplotlib (absent from this image) so that the unmodified reference can be
imported by oracle/ref_harness.py. Test infrastructure only."""
def __getattr__(name):
    def _noop(*a, **k):
        return None
    return _noop
