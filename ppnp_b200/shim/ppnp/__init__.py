"""Overlay of the reference's ``ppnp`` package: this directory precedes the reference on sys.path
(INTEGRATION.md), so ``import ppnp`` lands here; every other ``ppnp`` directory on sys.path is
appended to the package path, which keeps ``ppnp.preprocessing`` (main.py:28) and the data files
resolving to the reference while ``ppnp.data.sparsegraph`` (main.py:27) is the GPU-backed drop-in."""
import os
import sys

__path__ = [os.path.dirname(os.path.abspath(__file__))]
for _d in sys.path:
    _cand = os.path.join(os.path.abspath(_d or "."), "ppnp")
    if os.path.isdir(_cand) and _cand not in __path__:
        __path__.append(_cand)
