"""quake_b200 -- B200-native (sm_100a) implementation of the Quake partitioned-IVF search hot path.

Drop-in for the reference's Python surface (`from quake import QuakeIndex, IndexBuildParams, SearchParams`,
/root/reference/src/python/__init__.py:1-8). Importing the package does not need a GPU; every compute call
does, and raises if the CUDA library or a compute-capability-10.x device is missing.
"""
from .params import (BuildTimingInfo, IndexBuildParams, MaintenancePolicyParams, MaintenanceTimingInfo,  # noqa: F401
                     ModifyTimingInfo, SearchParams, SearchResult, SearchTimingInfo)
from .index import QuakeIndex  # noqa: F401
from . import _lib  # noqa: F401

__version__ = "0.1.0"
