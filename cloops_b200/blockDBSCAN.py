"""Drop-in for ``cLoops.blockDBSCAN.blockDBSCAN`` (cell-level DBSCAN, alternative import at cLoops/pipe.py:43)."""
from ._facade import _GpuDBSCAN


class blockDBSCAN(_GpuDBSCAN):
    """cLoops/blockDBSCAN.py:6-41 -- same constructor, same ``labels``; computed on the GPU."""
    _variant = 3
