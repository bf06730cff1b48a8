"""Stages the UNMODIFIED reference model definitions (models/*.py, tasks/SemSeg/seg_models.py,
tasks/Classification/class_models.py) from /root/reference into the git-ignored baseline/_ref/ so that they travel
to the GPU box with the working tree (the box has no /root/reference).  Nothing is copied into the tracked tree or
into the package; tests and `bench.py --workload fpn` import the models from there with `point_cloud_lib` aliased
to se3conv3d_b200.  Run by __graft_entry__.build() whenever the reference tree is present."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SE3_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["tasks/SemSeg/seg_models.py", "tasks/Classification/class_models.py"]


def stage():
    if not os.path.isdir(os.path.join(REF, "models")):
        return None
    os.makedirs(DST, exist_ok=True)
    dst_models = os.path.join(DST, "models")
    if os.path.isdir(dst_models):
        shutil.rmtree(dst_models)
    shutil.copytree(os.path.join(REF, "models"), dst_models, ignore=shutil.ignore_patterns("__pycache__"))
    for f in FILES:
        d = os.path.join(DST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(REF, f), d)
    return DST


def import_models(which="seg_models"):
    """Imports the staged (or, in the build container, the original) reference model module with `point_cloud_lib`
    aliased to se3conv3d_b200.  Returns the module (seg_models or class_models)."""
    import importlib
    import warnings
    import se3conv3d_b200
    import se3conv3d_b200.layers
    import se3conv3d_b200.pc
    base = DST if os.path.isdir(os.path.join(DST, "models")) else (REF if os.path.isdir(os.path.join(REF, "models")) else None)
    if base is None:
        raise ImportError("reference models are not staged (run tools/stage_reference_models.py in the build container)")
    sys.modules["point_cloud_lib"] = se3conv3d_b200
    sys.modules["point_cloud_lib.layers"] = se3conv3d_b200.layers
    sys.modules["point_cloud_lib.pc"] = se3conv3d_b200.pc
    sub = "SemSeg" if which == "seg_models" else "Classification"
    for p in (base, os.path.join(base, "tasks", sub)):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module(which)


if __name__ == "__main__":
    print(stage())
