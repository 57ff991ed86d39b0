"""Importable alias for the package directory `differentiable-point-clouds_b200/`
(a hyphen cannot appear in a Python module name).  `import dpc_b200.util.point_cloud`
resolves to `differentiable-point-clouds_b200/util/point_cloud.py`.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "differentiable-point-clouds_b200")
__path__.insert(0, _PKG_DIR)
PACKAGE_DIR = _PKG_DIR
CSRC_DIR = _os.path.join(_PKG_DIR, "csrc")
__version__ = "0.1.0"
