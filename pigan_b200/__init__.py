"""Importable alias of the ``pi-gan-thz_b200/`` package directory (a hyphen cannot be imported).

``import pigan_b200`` resolves submodules (``pigan_b200.native``, ``pigan_b200.engine`` …) from
``pi-gan-thz_b200/``; the drop-in ``core`` / ``config`` packages live in that same directory and are used
by putting it on ``sys.path``.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pi-gan-thz_b200")
__path__.insert(0, _PKG_DIR)
PACKAGE_DIR = _PKG_DIR

__all__ = ["PACKAGE_DIR"]
