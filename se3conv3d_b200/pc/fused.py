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
import os
import weakref

import torch

from .._lib import lib, check, ptr, stream, Se3Error, HierDesc, HierResult, HIER_MAX_CLOUDS, HIER_MAX_NEIGH
from .grid import Grid
from .hierarchy import PointHierarchyRotEquiv
from .neighborhood import BQNeighborhood, ConvGeometry
from .pointcloud_rot_equiv import PointcloudRotEquiv
from .subsample import GridSubSample

_arena_hint = {}
_EAGER = bool(os.environ.get("SE3_EAGER_VIEWS"))


class _Views(object):
    """Typed windows into the arena (byte offsets from the native result -> tensors, one op each)."""

    def __init__(self, arena):
        self.arena = arena
        self.f32 = arena.view(torch.float32)
        self.i32 = arena.view(torch.int32)
        self.i64 = arena.view(torch.int64)
        self.base_addr = arena.data_ptr()

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
    n, rec_off = c.n, c.rec

    pts_t, frames_t = pc.pts_, pc.local_frames_   # (the thunks must not refer to their owner: no reference cycles,
                                                  # so dropping a hierarchy returns its arena to the allocator at once)

    def records():
        rec = v.f(rec_off, max(n * n_frames, 1), 12)
        return ((id(pts_t), id(frames_t), pts_t._version, frames_t._version), rec)
    pc._lazy = {"_se3_records": records}
    pc._rec_addr = v.base_addr + rec_off
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
    g._sorted_cell_ids = None
    g.num_used_cells_ = int(c.m)
    n, m, o_ids, o_sorted, o_ends = c.n, int(c.m), c.cell_ids, c.sorted_ids, c.cell_ends
    g._lazy = {"cell_ids_": lambda: v.l(o_ids, n), "sorted_ids_": lambda: v.l(o_sorted, n),
               "cell_ends_": lambda: v.i(o_ends, m)}
    samp.grid_ = g
    return samp


def build_point_hierarchy(p_pts, p_batch_ids, p_ref_frames_config, p_init_subsample, p_grid_subsamples,
                          neighborhoods=(), output_cloud=False, n_batches=None):
    """Returns (PointHierarchyRotEquiv, output PointcloudRotEquiv or None).  `neighborhoods` lists
    (src_level, dst_level, radius) ball queries; level len(p_grid_subsamples) + 1 is the output cloud."""
    cfg = p_ref_frames_config
    sampled = not cfg["pca"]
    if sampled and cfg.get("fixed_axis"):
        raise Se3Error("build_point_hierarchy: sampled frames about a fixed axis are not fused (use the per-object path)")
    if not sampled and cfg["neigh_method"] != "knn":
        raise Se3Error("build_point_hierarchy: PCA frames are fused for k-NN neighbourhoods only (use the per-object path)")
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
    d.knn_k = 0 if sampled else int(cfg["neigh_kwargs"]["neigh_k"])
    d.n_frames = int(cfg["n_frames"])
    d.fixed_axis = -1 if (fixed is None or fixed is False or not fixed) else int(fixed)
    d.out_cloud = 1 if output_cloud else 0
    d.n_neigh = len(neighborhoods)
    for i, (s, t, r) in enumerate(neighborhoods):
        d.neigh_src[i], d.neigh_dst[i], d.neigh_radius[i] = int(s), int(t), float(r)
    if sampled:   # Gaussian quaternion components for every (point, frame) of every cloud, then the cell variates
        nq = (n_pool + 2) * n * int(cfg["n_frames"]) * 4
        u = torch.cat((torch.randn(nq, device=dev, dtype=torch.float32), torch.rand(n, device=dev, dtype=torch.float32)))
    else:
        u = torch.rand((n_pool + 3) * n, device=dev, dtype=torch.float32)
    key = (n, n_pool, len(neighborhoods), bool(output_cloud), str(dev))
    nbytes = _arena_hint.get(key, (64 << 20) + n * 4096)
    res = HierResult()
    L = lib()
    for _ in range(6):
        arena = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        rc = L.se3_hierarchy_build(C.byref(d), ptr(pts), ptr(b), ptr(u), ptr(u[u.numel() - n:]), ptr(arena),
                                   arena.numel(), C.byref(res), stream())
        if rc != -3:  # SE3_EWORKSPACE
            break
        nbytes = max(2 * nbytes, int(res.arena_used * 1.5))
    check(rc, "se3_hierarchy_build")
    # monotone and quantised (32 MB): consecutive builds of similar clouds then ask the caching allocator for the SAME
    # size, so the previous steps' arenas are reused instead of new device allocations (which synchronise)
    want = ((int(res.arena_used * 1.25) + (4 << 20) + (32 << 20) - 1) >> 25) << 25
    _arena_hint[key] = max(want, _arena_hint.get(key, 0)) if rc == 0 and nbytes <= 4 * want else want

    v = _Views(arena)
    n_frames = d.n_frames
    clouds = [_make_cloud(v, res.clouds[i], cfg, n_frames, n_batches) for i in range(res.n_clouds)]
    h = PointHierarchyRotEquiv.__new__(PointHierarchyRotEquiv)
    h.pcs_ = clouds[:n_pool + 1]
    h.sub_sampled_objs_ = [_make_sampler(v, h.pcs_[l], res.clouds[l], float(p_grid_subsamples[l])) for l in range(n_pool)]
    h.neigh_cache_ = {}
    h.fused_arena_ = arena
    # the raw cloud's init_cell grid (pooling raw features to level 0 / labels of the output cloud)
    raw_m, o_ci, o_si, o_ce = int(res.raw.m), res.raw.cell_ids, res.raw.sorted_ids, res.raw.cell_ends
    h._lazy = {"init_cell_ids_": lambda: v.l(o_ci, n), "init_sorted_ids_": lambda: v.l(o_si, n),
               "init_cell_ends_": lambda: v.i(o_ce, raw_m)}
    out_pc = clouds[n_pool + 1] if output_cloud else None
    if out_pc is not None:
        o_pick, n_out_pc = res.out_picked, int(res.clouds[n_pool + 1].n)
        out_pc._lazy["picked_ids_"] = lambda: v.l(o_pick, n_out_pc)
    neighs = []
    base = v.base_addr
    for i, (s, t, r) in enumerate(neighborhoods):
        neighs.append(_make_neighborhood(v, base, res.neigh[i], clouds[s], clouds[t], r, n_frames))
        # weak: the hierarchy owns the neighbourhoods; a strong entry would close a cloud -> neighbourhood -> cloud cycle
        # and the arena would wait for the garbage collector instead of returning to the allocator at once
        clouds[s].__dict__.setdefault("_fused_bq_", {})[(id(clouds[t]), float(r))] = weakref.ref(neighs[-1])
        if s <= n_pool and t <= n_pool:
            h.neigh_cache_[str(s) + "_" + str(t) + "_ball_query" + str(r)] = neighs[-1]
    h.fused_neighborhoods_ = neighs
    if _EAGER:   # A/B aid: materialise every window right away
        objs = clouds + [sm.grid_ for sm in h.sub_sampled_objs_] + neighs + [h]
        objs += [g for nb in neighs for g in nb.conv_geometry_cache_.values()]
        for o in objs:
            for name in list(o.__dict__.get("_lazy", {})):
                getattr(o, name)
    return h, out_pc


def _make_neighborhood(v, base, nr, src, dst, r, n_frames):
    """BQNeighborhood + its ConvGeometry over arena windows.  The conv calls only need device addresses
    (`geom._addr`, read by custom_ops.make_conv_desc); the tensors of the reference contract (`start_ids_`,
    `neighbors_`, the CSR columns, ...) are created when somebody asks for them."""
    e = int(nr.e)
    n_in, n_out = int(src.pts_.shape[0]), int(dst.pts_.shape[0])
    o_re, o_cs, o_ed, o_tre, o_te, o_td = nr.row_ends, nr.col_src, nr.edge_dst, nr.t_row_ends, nr.t_edge, nr.t_dst
    nb = BQNeighborhood.__new__(BQNeighborhood)
    nb._prebuilt_ = True
    nb.radius_ = r
    nb.max_neighbors_ = 0
    nb.pc_src_, nb.samples_ = src, dst
    nb._neighbors = None
    nb._fused_token = base + o_cs
    nb.n_edges_ = e
    nb._lazy = {"start_ids_": lambda: v.i(o_re, n_out),
                "_csr_columns": lambda: (v.i(o_ed, max(e, 1))[:e], v.i(o_cs, max(e, 1))[:e])}
    geom = ConvGeometry.__new__(ConvGeometry)
    geom.n_in, geom.n_out, geom.n_edges = n_in, n_out, e
    geom.f_in = geom.f_out = n_frames
    geom._addr = {"pts_in": src.pts_.data_ptr(), "pts_out": dst.pts_.data_ptr(),
                  "frames_in": src.local_frames_.data_ptr(), "frames_out": dst.local_frames_.data_ptr(),
                  "rec_in": src._rec_addr, "rec_out": dst._rec_addr, "row_ends": base + o_re, "col_src": base + o_cs,
                  "t_row_ends": base + o_tre, "t_edge": base + o_te, "t_dst": base + o_td}
    geom._keep = (src, dst, v)   # the arena outlives the geometry
    geom._lazy = {"pts_in": lambda: src.pts_, "pts_out": lambda: dst.pts_,
                  "frames_in": lambda: src.local_frames_, "frames_out": lambda: dst.local_frames_,
                  "rec_in": lambda: src._se3_records[1], "rec_out": lambda: dst._se3_records[1],
                  "row_ends": lambda: v.i(o_re, n_out), "col_src": lambda: v.i(o_cs, max(e, 1)),
                  "t_row_ends": lambda: v.i(o_tre, max(n_in, 1)), "t_edge": lambda: v.i(o_te, max(e, 1)),
                  "t_dst": lambda: v.i(o_td, max(e, 1))}
    nb.conv_geometry_cache_ = {(id(src), id(dst), id(src.local_frames_), id(dst.local_frames_),
                                nb._neighbors_token()): geom}
    return nb
