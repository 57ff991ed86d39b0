"""TEST INFRASTRUCTURE ONLY.  Imports the reference's hot-path modules, unmodified, from
/root/reference/dpc with the TF1 shim (oracle/tf1_shim) standing in for TensorFlow, so the
reference's own source text can be executed in this container.  Used by
tests/golden/make_golden.py (fixture generation) and by the CPU tests that compare the
oracle restatement with the reference live (skipped where /root/reference is absent,
e.g. on the GPU box).  Never imported by the product path.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("DPC_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf1_shim")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dpc", "util", "point_cloud.py"))


def load():
    """Return a namespace with the reference modules (point_cloud, drc, gauss_kernel,
    quaternion, camera) and the shim `tf`."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    dpc = os.path.join(REFERENCE_ROOT, "dpc")
    for p in (dpc, _SHIM):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    tf = importlib.import_module("tensorflow")
    assert "tf1_shim" in tf.__file__, "a real tensorflow shadows the shim: %s" % tf.__file__

    class NS:
        pass

    ns = NS()
    ns.tf = tf
    ns.point_cloud = importlib.import_module("util.point_cloud")
    ns.drc = importlib.import_module("util.drc")
    ns.gauss_kernel = importlib.import_module("util.gauss_kernel")
    ns.quaternion = importlib.import_module("util.quaternion")
    ns.camera = importlib.import_module("util.camera")
    ns.point_cloud_distance = importlib.import_module("util.point_cloud_distance")
    assert ns.point_cloud.__file__.startswith(REFERENCE_ROOT)
    return ns
