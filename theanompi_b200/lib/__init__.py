"""Import-path parity with ``theanompi.lib`` — the implementations live in
:mod:`theanompi_b200.parallel` and :mod:`theanompi_b200.utils`."""
