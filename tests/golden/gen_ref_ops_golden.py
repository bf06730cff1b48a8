"""Runs the UNMODIFIED reference CUDA ops (oracle/_ref/point_cloud_lib_ops*.so, built from
/root/reference by oracle/build_ref.py) on a B200 and freezes their outputs for small seeded inputs:

    gpurun -- python tests/golden/gen_ref_ops_golden.py      # writes gpurun_out/ref_ops_golden.npz
    cp gpurun_out/ref_ops_golden.npz tests/golden/

Pins the integer oracle (oracle/se3_oracle.c): tests/test_oracle_ref_ops.py compares it, on the CPU,
with these vectors.  Rows of the ball query are stored canonicalised (sorted by source index inside a
row) because the reference's in-row order is an atomic race.
"""
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))


def cloud(n, b, seed, scale=(1.0, 1.0, 1.0), shift=(0.0, 0.0, 0.0)):
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(n, 3, generator=g) * torch.tensor(scale) + torch.tensor(shift)
    batch = torch.sort(torch.randint(0, b, (n,), generator=g))[0].to(torch.int32)
    return pts, batch


def main():
    assert glob.glob(os.path.join(ROOT, "oracle", "_ref", "point_cloud_lib_ops*.so")), "build oracle/_ref first"
    import point_cloud_lib_ops as ref
    dev = torch.device("cuda:0")
    out = {}
    cases = {
        "bq_same": (cloud(1500, 3, 10), None, 0.12),
        "bq_cross": (cloud(1200, 2, 11), cloud(500, 2, 12, (1.3, 1.3, 1.3), (-0.15, -0.15, -0.15)), 0.2),
        "bq_flat": (cloud(900, 1, 13, (1.0, 1.0, 0.02)), None, 0.1),
    }
    for name, ((src, bs), dstp, r) in cases.items():
        dst, bd = (src, bs) if dstp is None else dstp
        mn = torch.stack([src[bs == i].min(0)[0] for i in range(int(bs.max()) + 1)]) - 1e-6
        mx = torch.stack([src[bs == i].max(0)[0] for i in range(int(bs.max()) + 1)]) - 1e-6
        nc = torch.max(((mx - mn) / r).to(torch.int32) + 1, dim=0)[0]
        rad = torch.full((3,), r, dtype=torch.float32)
        a = [t.to(dev) for t in (src, dst, bs, bd, mn, nc, rad)]
        nb, ends = ref.ball_query(*a, 0)
        keys = ref.compute_keys(a[0], a[2], a[4], a[5], a[6])
        nb = nb.cpu().numpy()
        nb = nb[np.lexsort((nb[:, 1], nb[:, 0]))]
        out.update({name + "_src": src.numpy(), name + "_dst": dst.numpy(), name + "_bs": bs.numpy(),
                    name + "_bd": bd.numpy(), name + "_min": mn.numpy(), name + "_nc": nc.numpy(),
                    name + "_r": np.float32(r), name + "_nb": nb, name + "_ends": ends.cpu().numpy(),
                    name + "_keys": keys.cpu().numpy()})
        print(name, "E =", nb.shape[0])
    for name, (pts, b), k in (("knn_a", cloud(2000, 3, 20, (1.0, 0.5, 2.0)), 16), ("knn_b", cloud(40, 4, 21), 16)):
        idx = ref.knn_query(pts.to(dev), b.to(dev), k).cpu().numpy()
        out.update({name + "_pts": pts.numpy(), name + "_b": b.numpy(), name + "_idx": idx})
        print(name, idx.shape)
    # feat_basis_proj / grad
    g = torch.Generator().manual_seed(30)
    m, n, c, k = 300, 400, 16, 32
    deg = torch.randint(0, 12, (m,), generator=g)
    rows = torch.repeat_interleave(torch.arange(m), deg)
    cols = torch.randint(0, n, (rows.shape[0],), generator=g)
    nbr = torch.stack((rows, cols), 1).to(torch.int32)
    ends = torch.cumsum(deg, 0).to(torch.int32)
    basis = torch.randn(rows.shape[0], k, generator=g)
    feats = torch.randn(n, c, generator=g)
    grads = torch.randn(m, c, k, generator=g)
    T = ref.feat_basis_proj(basis.to(dev), feats.to(dev), nbr.to(dev), ends.to(dev))
    fg, bg = ref.feat_basis_proj_grad(basis.to(dev), feats.to(dev), nbr.to(dev), ends.to(dev), grads.to(dev))
    out.update({"fbp_basis": basis.numpy(), "fbp_feats": feats.numpy(), "fbp_nbr": nbr.numpy(), "fbp_ends": ends.numpy(),
                "fbp_grads": grads.numpy(), "fbp_T": T.cpu().numpy(), "fbp_fg": fg.cpu().numpy(),
                "fbp_bg": bg.cpu().numpy()})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_ops_golden.npz"), **out)
    print("wrote gpurun_out/ref_ops_golden.npz")


if __name__ == "__main__":
    main()
