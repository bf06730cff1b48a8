"""Generates tests/golden/fpn_faust.npz by running the REFERENCE end to end on the CPU of the build container:
the reference's own Python for the hierarchy (Pointcloud, GridSubSample, PointcloudRotEquiv, PointHierarchyRotEquiv:
tasks/SemSeg/train_dfaust_rot.py:108-158, 249-259), its unmodified model code (models/FPNSegUNet.py through
tasks/SemSeg/seg_models.py: FPNSegUNetMLPGeluRotEqFAUST), forward, cross-entropy loss and backward
(train_dfaust_rot.py:262-264), in float64.

    python tests/golden/gen_fpn_golden.py

Substitutions (the reference has no CPU implementation of its CUDA ops): `point_cloud_lib_ops` is
oracle/shims_cpu/point_cloud_lib_ops.py (C oracle for keys / ball query / kNN), torch_scatter / torch_cluster are
oracle/shims, and FeatBasisProj is its scatter formulation so that the float64 graph is differentiable (the
wrapper casts to float32, custom_ops/FeatBasisProj.py:36-40).  Drop-path probability is 0 (the test needs a
deterministic forward); parameters come from tests/fpn_fixture.py::reinit_by_name.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle.shims_cpu.point_cloud_lib_ops as cpu_ops  # noqa: E402
sys.modules["point_cloud_lib_ops"] = cpu_ops
from oracle.ref_import import import_reference, REF_ROOT  # noqa: E402
from fpn_fixture import FPN_CFG, FULL_GRADS, reinit_by_name  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    pclib = import_reference()
    refmod = sys.modules["point_cloud_lib.layers.PNEConvLayerRotEquiv"]
    from einops import repeat
    sys.path[:0] = [REF_ROOT, os.path.join(REF_ROOT, "tasks", "SemSeg")]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import seg_models

    class ScatterFeatBasisProj:
        @staticmethod
        def apply(basis, feats, neighbors, ends):
            return cpu_ops.feat_basis_proj(basis, feats, neighbors, ends)
    refmod.FeatBasisProj = ScatterFeatBasisProj

    from se3conv3d_b200 import workloads as wl          # synthetic bodies only (no kernels involved)
    pts, batch_ids = wl.synthetic_bodies(2, 1400, seed=5)
    gen = torch.Generator().manual_seed(11)
    feats = torch.ones(pts.shape[0], 1)
    labels = torch.randint(0, 20, (pts.shape[0],), generator=gen)
    cfg = FPN_CFG
    torch.manual_seed(3)                                 # frame shuffles (torch.multinomial) and the rnd sub-sample
    with torch.no_grad():
        pc = pclib.pc.Pointcloud(pts, batch_ids)
        samp = pclib.pc.GridSubSample(pc, cfg["init_subsample"])
        new_pts = samp.__subsample_tensor__(pc.pts_, "avg")
        new_b = samp.__subsample_tensor__(pc.batch_ids_, "max")
        new_feats = samp.__subsample_tensor__(feats, "avg")
        new_pc = pclib.pc.PointcloudRotEquiv(new_pts, new_b, cfg["RefFrames"])
        hierarchy = pclib.pc.PointHierarchyRotEquiv(new_pc, len(cfg["grid_subsamples"]), "grid_avg",
                                                    grid_radii=cfg["grid_subsamples"])
        radii = [cfg["init_subsample"]] + cfg["grid_subsamples"]
        osamp = pclib.pc.GridSubSample(pc, cfg["init_subsample"], p_rnd_sample=True)
        out_pts = osamp.__subsample_tensor__(pc.pts_, "avg")
        out_b = osamp.__subsample_tensor__(pc.batch_ids_, "max")
        out_labels = osamp.__subsample_tensor__(labels, "max")
        out_pc = pclib.pc.PointcloudRotEquiv(out_pts, out_b, cfg["RefFrames"])
    features = repeat(new_feats, "n d -> (n times) d", times=cfg["RefFrames"]["n_frames"])
    sizes = [int(p.pts_.shape[0]) for p in hierarchy.pcs_]
    print("levels", sizes, "out", int(out_pc.pts_.shape[0]))

    model = seg_models.FPNSegUNetMLPGeluRotEqFAUST(1, 20, 0.0)
    reinit_by_name(model)
    # what the pre-process epoch does (train_dfaust_rot.py:172-218): one no-grad forward with the EMA switch on
    model.train()
    model.start_pre_process()
    with torch.no_grad():
        model(hierarchy, features, radii, out_pc)
    model.end_pre_process()
    buffers = {k: v.clone() for k, v in model.state_dict().items() if k.endswith("norm_neigh_dist_") or k.endswith("norm_num_neighs_")}
    # BatchNorm running statistics were touched by the pre-process forward: reset them (the test starts from a fresh model)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.reset_running_stats()

    res = {}
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        m = seg_models.FPNSegUNetMLPGeluRotEqFAUST(1, 20, 0.0)
        reinit_by_name(m)
        m.load_state_dict({**m.state_dict(), **buffers})
        m = m.to(dt)
        for mod_name, mod in m.named_modules():
            for b in ("norm_neigh_dist_", "norm_num_neighs_"):
                key = (mod_name + "." if mod_name else "") + b
                if key in buffers:
                    setattr(mod, b, buffers[key].to(dt))
        m.train()
        for p in list(hierarchy.pcs_) + [out_pc]:
            p.pts_ = p.pts_.to(dt)
            p.local_frames_ = p.local_frames_.to(dt)
        hierarchy.neigh_cache_ = {}
        pred = m(hierarchy, features.to(dt), radii, out_pc)
        loss = torch.nn.functional.cross_entropy(pred, out_labels)
        loss.backward()
        res["logits_" + tag] = pred.detach().numpy()
        res["loss_" + tag] = np.float64(loss.item())
        grads = {n: p.grad for n, p in m.named_parameters()}
        res["grad_names"] = np.array(sorted(grads))
        res["grad_norms_" + tag] = np.array([float(grads[n].double().norm()) for n in sorted(grads)])
        for n in FULL_GRADS:
            res["grad_%s_%s" % (tag, n)] = grads[n].detach().numpy()
        print(tag, "loss", loss.item(), "logits", tuple(pred.shape))
    rel = np.abs(res["logits_f32"] - res["logits_f64"]).max() / np.abs(res["logits_f64"]).max()
    print("reference fp32 vs fp64 logits: %.2e" % rel)

    out = {"pts": pts.numpy(), "batch_ids": batch_ids.numpy(), "features": features.numpy(),
           "out_labels": out_labels.numpy().astype(np.int64), "radii": np.array(radii, np.float64),
           "out_pts": out_pts.numpy().astype(np.float32), "out_batch": out_b.numpy().astype(np.int32),
           "out_frames": out_pc.local_frames_.numpy().astype(np.float32),
           "buffer_names": np.array(sorted(buffers)), "buffer_values": np.array([float(buffers[k]) for k in sorted(buffers)])}
    for l, p in enumerate(hierarchy.pcs_):
        out["pts_%d" % l] = p.pts_.numpy().astype(np.float32)
        out["batch_%d" % l] = p.batch_ids_.numpy().astype(np.int32)
        out["frames_%d" % l] = p.local_frames_.numpy().astype(np.float32)
    for l, s in enumerate(hierarchy.sub_sampled_objs_):
        out["cell_ids_%d" % l] = s.grid_.cell_ids_.numpy().astype(np.int64)
    # every neighbourhood the model asked for (integer parity of the ball query on the box)
    for key, nb in hierarchy.neigh_cache_.items():
        out["nb_" + key] = nb.neighbors_.numpy().astype(np.int32)
        out["ends_" + key] = nb.start_ids_.numpy().astype(np.int32)
    out.update(res)
    np.savez_compressed(os.path.join(OUT, "fpn_faust.npz"), **out)
    print("written", os.path.getsize(os.path.join(OUT, "fpn_faust.npz")) / 1e6, "MB; neighbourhoods:", sorted(hierarchy.neigh_cache_))


if __name__ == "__main__":
    main()
