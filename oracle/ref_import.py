"""TEST INFRASTRUCTURE ONLY -- imports the *reference* Python package from /root/reference without
running its broken top-level __init__ (data_sets.loaders imports a missing module,
data_sets/loaders/__init__.py:2) and with the torch_scatter / torch_cluster shims on sys.path.
Only usable in the build container (the GPU box has no /root/reference)."""
import os
import sys
import types

REF_ROOT = os.environ.get("SE3_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "point_cloud_lib", "point_cloud_lib"))


def import_reference():
    if not available():
        raise ImportError("reference tree not present at " + REF_ROOT)
    for p in (os.path.join(HERE, "shims"), os.path.join(HERE, "_ref")):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "point_cloud_lib" not in sys.modules:
        pkg = types.ModuleType("point_cloud_lib")
        pkg.__path__ = [os.path.join(REF_ROOT, "point_cloud_lib", "point_cloud_lib")]
        sys.modules["point_cloud_lib"] = pkg
    import point_cloud_lib.custom_ops  # noqa: F401  (needs oracle/_ref/point_cloud_lib_ops*.so)
    import point_cloud_lib.pc  # noqa: F401
    import point_cloud_lib.layers  # noqa: F401
    return sys.modules["point_cloud_lib"]
