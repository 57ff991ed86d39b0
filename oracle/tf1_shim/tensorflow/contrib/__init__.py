"""TEST INFRASTRUCTURE ONLY -- tf.contrib stubs: summaries are no-ops; slim / layers only need to be importable (the
reference's loss code never calls them; its networks are rebuilt in torch, not run through the shim)."""
from . import slim  # noqa: F401


class _Summary:
    @staticmethod
    def scalar(*a, **k):
        return None

    @staticmethod
    def histogram(*a, **k):
        return None

    @staticmethod
    def image(*a, **k):
        return None


summary = _Summary()


class _Layers:
    @staticmethod
    def variance_scaling_initializer(*a, **k):
        raise NotImplementedError("networks are not run through the shim")


layers = _Layers()
