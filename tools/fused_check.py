"""Quick GPU check of the fused tcgen05 convolution kernel against the fp64 oracle (small problems first)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import layer_oracle as lo  # noqa: E402

DEV = "cuda:0"


def problem(n, r, f, cin, cout, seed=0):
    from se3conv3d_b200.pc import PointcloudRotEquiv, BQNeighborhood
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv
    pts = torch.rand(n, 3, generator=torch.Generator().manual_seed(seed))
    cfg = {"pca": True, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": False, "n_frames": f}
    pc = PointcloudRotEquiv(pts.to(DEV), torch.zeros(n, dtype=torch.int32, device=DEV), cfg)
    neigh = BQNeighborhood(pc, pc, r)
    torch.manual_seed(2)
    layer = PNEConvLayerRotEquiv(9, cin, cout, 32, "mlp_gelu").to(DEV)
    with torch.no_grad():
        layer.proj_biases_.copy_(0.1 * torch.randn(32))
    layer.norm_neigh_dist_.fill_(1.0 / r)
    layer.norm_num_neighs_.fill_(n / neigh.neighbors_.shape[0])
    layer.precision = 1
    x = torch.randn(n * f, cin, generator=torch.Generator().manual_seed(3)).to(DEV)
    dy = torch.randn(n * f, cout, generator=torch.Generator().manual_seed(4)).to(DEV) / cout ** 0.5
    return pc, neigh, layer, x, dy


def run(n, r, f, cin, cout):
    pc, neigh, layer, x, dy = problem(n, r, f, cin, cout)
    xx = x.clone().requires_grad_(True)
    y = layer(pc, pc, xx, neigh)
    torch.cuda.synchronize()
    y.backward(dy)
    torch.cuda.synchronize()
    c = lambda t: t.detach().cpu().double()
    with torch.no_grad():
        ref = lo.conv_forward_backward(c(x), c(layer.proj_axes_), c(layer.proj_biases_), c(layer.conv_weights_), c(pc.pts_),
                                       c(pc.pts_), c(pc.local_frames_), c(pc.local_frames_), neigh.neighbors_.cpu(),
                                       float(layer.norm_neigh_dist_), float(layer.norm_num_neighs_), c(dy))
    got = (y, xx.grad, layer.conv_weights_.grad, layer.proj_axes_.grad, layer.proj_biases_.grad)
    print("n=%d f=%d %d->%d E=%d" % (n, f, cin, cout, neigh.neighbors_.shape[0]))
    for name, a, b in zip(("y", "dx", "dW", "dA", "dB"), got, ref):
        m = lo.err_metrics(a.detach().cpu().numpy(), b.numpy())
        print("   %-2s max/max %.2e relL2 %.2e p99.9 %.2e" % (name, *m), flush=True)


if __name__ == "__main__":
    print("SE3_FUSED =", os.environ.get("SE3_FUSED", "1"), flush=True)
    for cfg in [(300, 0.3, 2, 32, 32), (2048, 0.16, 2, 32, 32), (2048, 0.16, 1, 32, 32), (3000, 0.12, 2, 16, 32),
                (3000, 0.12, 2, 64, 16), (8192, 0.1, 2, 32, 64), (3000, 0.12, 1, 64, 64), (8192, 0.1, 2, 32, 32)]:
        run(*cfg)
    # timing on the last problem
    pc, neigh, layer, x, dy = problem(40000, 0.055, 2, 32, 32)
    xx = x.clone().requires_grad_(True)
    for _ in range(3):
        y = layer(pc, pc, xx, neigh)
        y.backward(dy)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        y = layer(pc, pc, xx, neigh)
        y.backward(dy)
    torch.cuda.synchronize()
    print("40000 pts E=%d fwd+bwd %.3f ms" % (neigh.neighbors_.shape[0], (time.perf_counter() - t0) * 100))
    from se3conv3d_b200 import _lib

    def once():
        y = layer(pc, pc, xx, neigh)
        y.backward(dy)
    for name, ms, cnt in _lib.profile_kernels(once, reps=5):
        print("  %-50s %.1f us x %d" % (name, ms * 1e3, cnt))
