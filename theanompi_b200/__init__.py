"""theanompi_b200 — a B200-native data-parallel training framework with the
capabilities and API of Theano-MPI (``theanompi/__init__.py:1``)::

    from theanompi_b200 import BSP, EASGD, GOSGD
    rule = BSP(); rule.init(devices=['cuda0', 'cuda1'], modelfile=..., modelclass=...); rule.wait()
"""
__version__ = "0.1.0"

from .rules import ASGD, BSP, EASGD, GOSGD, Rule  # noqa: E402,F401

__all__ = ["BSP", "EASGD", "GOSGD", "ASGD", "Rule"]
