"""Fused construction of a PointHierarchyRotEquiv and the neighbourhoods a model will ask for
(SURVEY 8 row f3).  One native call (se3_hierarchy_build) replaces the per-object chain of
tasks/SemSeg/train_dfaust_rot.py:108-158 (create_hierarchy) + the lazy create_neighborhood calls of
models/Encoder.py:134-154, Decoder.py:72-80, FPNDecoder.py:104-113: the objects returned here are the
same classes with the same attributes, their tensors are views into one device arena.

    h, out_pc = build_point_hierarchy(pts, batch_ids, ref_frames_cfg, init_subsample, grid_subsamples,
                                      neighborhoods=[(src_level, dst_level, radius), ...], output_cloud=True)
    h.create_neighborhood(1, 1, "ball_query", bq_radius=0.1)   # served from the pre-filled cache
"""
import ctypes as C

import torch

from .._lib import lib, check, ptr, stream, Se3Error, HierDesc, HierResult, HIER_MAX_CLOUDS, HIER_MAX_NEIGH
from .grid import Grid
from .hierarchy import PointHierarchyRotEquiv
from .neighborhood import BQNeighborhood, ConvGeometry
from .pointcloud_rot_equiv import PointcloudRotEquiv
from .subsample import GridSubSample

_arena_hint = {}


class _Views(object):
    """Typed windows into the arena (byte offsets from the native result -> tensors, one op each)."""

    def __init__(self, arena):
        self.f32 = arena.view(torch.float32)
        self.i32 = arena.view(torch.int32)
        self.i64 = arena.view(torch.int64)

    def f(self, off, *shape):
        return self._v(self.f32, off, 4, shape)

    def i(self, off, *shape):
        return self._v(self.i32, off, 4, shape)

    def l(self, off, *shape):
        return self._v(self.i64, off, 8, shape)

    @staticmethod
    def _v(base, off, size, shape):
        stride, acc = [], 1
        for s in reversed(shape):
            stride.append(acc)
            acc *= s
        return torch.as_strided(base, shape, tuple(reversed(stride)), off // size)


def _make_cloud(v, c, cfg, n_frames, n_batches):
    pc = PointcloudRotEquiv.__new__(PointcloudRotEquiv)
    pc.pts_with_grads_ = False
    pc.batch_size_host_ = n_batches
    pc._batch_size = None
    pc.pts_ = v.f(c.pts, c.n, 3)
    pc.batch_ids_ = v.i(c.batch, c.n)
    pc.neigh_cache_ = {}
    pc.local_frames_pca_cache_ = {}
    pc.local_frames_config_ = cfg
    pc.standard_knn_ = False
    pc.ref_frames_pts = None
    pc.n_frames_ = n_frames
    pc.local_frames_ = v.f(c.frames, c.n, n_frames, 9)
    pc._batch_ids_frames = None
    rec = v.f(c.rec, max(c.n * n_frames, 1), 12)
    pc._se3_records = ((id(pc.pts_), id(pc.local_frames_), pc.pts_._version, pc.local_frames_._version), rec)
    return pc


def _make_sampler(v, pc_src, c, cell, rnd=False):
    samp = GridSubSample.__new__(GridSubSample)
    samp.pc_src_ = pc_src
    samp.ids_ = None
    samp.cell_size_ = cell
    samp.rnd_sample_ = rnd
    g = Grid.__new__(Grid)
    g.pointcloud_ = pc_src
    g.cell_size_ = cell
    g.bounding_box_ = None   # the fused builder does not keep per-grid bounding boxes
    g.num_cells_ = None
    g.cell_ids_ = v.l(c.cell_ids, c.n)
    g.sorted_ids_ = v.l(c.sorted_ids, c.n)
    g._sorted_cell_ids = None
    g.num_used_cells_ = int(c.m)
    g.cell_ends_ = v.i(c.cell_ends, int(c.m))
    samp.grid_ = g
    return samp


def build_point_hierarchy(p_pts, p_batch_ids, p_ref_frames_config, p_init_subsample, p_grid_subsamples,
                          neighborhoods=(), output_cloud=False, n_batches=None):
    """Returns (PointHierarchyRotEquiv, output PointcloudRotEquiv or None).  `neighborhoods` lists
    (src_level, dst_level, radius) ball queries; level len(p_grid_subsamples) + 1 is the output cloud."""
    cfg = p_ref_frames_config
    if not cfg["pca"] or cfg["neigh_method"] != "knn":
        raise Se3Error("build_point_hierarchy: only k-NN PCA reference frames are fused (use the per-object path)")
    pts = p_pts.detach().to(torch.float32).contiguous()
    b = p_batch_ids.to(torch.int32).contiguous()
    dev = pts.device
    n = int(pts.shape[0])
    n_pool = len(p_grid_subsamples)
    if n_pool + 2 > HIER_MAX_CLOUDS or len(neighborhoods) > HIER_MAX_NEIGH:
        raise Se3Error("build_point_hierarchy: too many levels / neighbourhoods for one fused call")
    if n_batches is None:
        n_batches = int(b.max()) + 1
    fixed = cfg["fixed_axis"]
    d = HierDesc()
    d.n, d.n_batches, d.n_pool, d.init_cell = n, int(n_batches), n_pool, float(p_init_subsample)
    for i, c in enumerate(p_grid_subsamples):
        d.cells[i] = float(c)
    d.knn_k = int(cfg["neigh_kwargs"]["neigh_k"])
    d.n_frames = int(cfg["n_frames"])
    d.fixed_axis = -1 if (fixed is None or fixed is False or not fixed) else int(fixed)
    d.out_cloud = 1 if output_cloud else 0
    d.n_neigh = len(neighborhoods)
    for i, (s, t, r) in enumerate(neighborhoods):
        d.neigh_src[i], d.neigh_dst[i], d.neigh_radius[i] = int(s), int(t), float(r)
    u = torch.rand((n_pool + 3) * n, device=dev, dtype=torch.float32)
    key = (n, n_pool, len(neighborhoods), bool(output_cloud), str(dev))
    nbytes = _arena_hint.get(key, (64 << 20) + n * 4096)
    res = HierResult()
    L = lib()
    for _ in range(6):
        arena = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        rc = L.se3_hierarchy_build(C.byref(d), ptr(pts), ptr(b), ptr(u), ptr(u[(n_pool + 2) * n:]), ptr(arena),
                                   arena.numel(), C.byref(res), stream())
        if rc != -3:  # SE3_EWORKSPACE
            break
        nbytes = max(2 * nbytes, int(res.arena_used * 1.5))
    check(rc, "se3_hierarchy_build")
    _arena_hint[key] = int(res.arena_used * 1.25) + (4 << 20)

    v = _Views(arena)
    n_frames = d.n_frames
    clouds = [_make_cloud(v, res.clouds[i], cfg, n_frames, n_batches) for i in range(res.n_clouds)]
    h = PointHierarchyRotEquiv.__new__(PointHierarchyRotEquiv)
    h.pcs_ = clouds[:n_pool + 1]
    h.sub_sampled_objs_ = [_make_sampler(v, h.pcs_[l], res.clouds[l], float(p_grid_subsamples[l])) for l in range(n_pool)]
    h.neigh_cache_ = {}
    h.fused_arena_ = arena
    # the raw cloud's init_cell grid (pooling raw features to level 0 / labels of the output cloud)
    h.init_cell_ids_ = v.l(res.raw.cell_ids, n)
    h.init_sorted_ids_ = v.l(res.raw.sorted_ids, n)
    h.init_cell_ends_ = v.i(res.raw.cell_ends, int(res.raw.m))
    out_pc = clouds[n_pool + 1] if output_cloud else None
    if out_pc is not None:
        out_pc.picked_ids_ = v.l(res.out_picked, int(res.clouds[n_pool + 1].n))
    neighs = []
    for i, (s, t, r) in enumerate(neighborhoods):
        nr = res.neigh[i]
        src, dst = clouds[s], clouds[t]
        e = int(nr.e)
        nb = BQNeighborhood.__new__(BQNeighborhood)
        nb.radius_ = r
        nb.max_neighbors_ = 0
        nb.pc_src_, nb.samples_ = src, dst
        nb._neighbors = None
        nb.start_ids_ = v.i(nr.row_ends, int(dst.pts_.shape[0]))
        col_src, edge_dst = v.i(nr.col_src, max(e, 1)), v.i(nr.edge_dst, max(e, 1))
        nb._csr_columns = (edge_dst[:e], col_src[:e])
        geom = ConvGeometry.__new__(ConvGeometry)
        geom.n_in, geom.n_out, geom.n_edges = int(src.pts_.shape[0]), int(dst.pts_.shape[0]), e
        geom.f_in = geom.f_out = n_frames
        geom.pts_in, geom.pts_out = src.pts_, dst.pts_
        geom.frames_in, geom.frames_out = src.local_frames_, dst.local_frames_
        geom.rec_in, geom.rec_out = src._se3_records[1], dst._se3_records[1]
        geom.row_ends, geom.col_src = nb.start_ids_, col_src
        geom.t_row_ends = v.i(nr.t_row_ends, max(geom.n_in, 1))
        geom.t_edge, geom.t_dst = v.i(nr.t_edge, max(e, 1)), v.i(nr.t_dst, max(e, 1))
        nb.conv_geometry_cache_ = {(id(src), id(dst), id(src.local_frames_), id(dst.local_frames_),
                                    nb._neighbors_token()): geom}
        neighs.append(nb)
        if s <= n_pool and t <= n_pool:
            h.neigh_cache_[str(s) + "_" + str(t) + "_ball_query" + str(r)] = nb
    h.fused_neighborhoods_ = neighs
    return h, out_pc
