"""Overlay of ``ppnp.data`` (see ../__init__.py): sparsegraph.py here, everything else from the reference."""
import os
import sys

__path__ = [os.path.dirname(os.path.abspath(__file__))]
for _d in sys.path:
    _cand = os.path.join(os.path.abspath(_d or "."), "ppnp", "data")
    if os.path.isdir(_cand) and _cand not in __path__:
        __path__.append(_cand)
