"""The reference's ``lib/proc_comm_mpi.py`` is a docstring-only TODO for a *parallel
communication process* that would overlap exchange with compute (``proc_comm_mpi.py:1-17``).
That idea is implemented here without an extra process: the BSP exchanger launches the
fused allreduce+SGD kernels per bucket on a side CUDA stream from the backward's
grad-ready callbacks (:class:`theanompi_b200.parallel.exchanger.BSP_Exchanger`)."""
from ..parallel.exchanger import BSP_Exchanger  # noqa: F401
