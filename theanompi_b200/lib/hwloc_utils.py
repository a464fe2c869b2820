from ..parallel.hwloc_utils import *  # noqa: F401,F403
