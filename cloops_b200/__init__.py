"""cloops_b200 -- B200-native (sm_100a) implementation of the cLoops clustering + scoring hot path.

Module and class names mirror the reference package so that ``from cLoops.cDBSCAN2 import cDBSCAN``
becomes ``from cloops_b200.cDBSCAN2 import cDBSCAN``.
"""
__version__ = "0.1.0"
