"""Import-path parity with the reference's loader process ``models/data/proc_load_mpi.py``;
the B200 loader (pinned ring + copy stream + fused crop kernel) is :class:`ParaLoader`."""
from .loader import LoadedBatch, ParaLoader  # noqa: F401
