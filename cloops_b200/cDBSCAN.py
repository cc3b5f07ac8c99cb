"""Drop-in for ``cLoops.cDBSCAN.cDBSCAN`` (v1; used by scripts/jd2saturation:23, scripts/callStripes:29)."""
from ._facade import _GpuDBSCAN


class cDBSCAN(_GpuDBSCAN):
    """cLoops/cDBSCAN.py:6-40 -- same constructor, same ``labels``; computed on the GPU."""
    _variant = 1
