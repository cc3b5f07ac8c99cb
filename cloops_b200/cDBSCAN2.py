"""Drop-in for ``cLoops.cDBSCAN2.cDBSCAN`` (the default clusterer, cLoops/pipe.py:42)."""
from ._facade import _GpuDBSCAN


class cDBSCAN(_GpuDBSCAN):
    """cLoops/cDBSCAN2.py:7-35 -- same constructor, same ``labels``; computed on the GPU."""
    _variant = 2
