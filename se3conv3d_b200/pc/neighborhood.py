from abc import ABC, abstractmethod

import torch

from .._lib import lib, check, ptr, stream, workspace, Se3Error, num_batches, LazyAttrs
from ..custom_ops import BallQuery, KNNQuery


def gather_records(p_pc):
    """[N*F,12] float32 records (point, frame) of a cloud for the tensor-core kernels (se3_pack_records),
    cached on the cloud and rebuilt when its coordinates or frames are replaced."""
    pts, frames = p_pc.pts_, p_pc.local_frames_
    key = (id(pts), id(frames), pts._version, frames._version)
    cached = getattr(p_pc, "_se3_records", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    p32 = pts.detach().to(torch.float32).contiguous()
    f32 = frames.detach().to(torch.float32).contiguous()
    n, f = int(p32.shape[0]), int(f32.shape[1])
    rec = torch.empty((max(n * f, 1), 12), dtype=torch.float32, device=p32.device)
    check(lib().se3_pack_records(ptr(p32), ptr(f32), n, f, ptr(rec), stream()), "se3_pack_records")
    try:
        # the keyed tensors are held by the entry: their ids cannot be recycled while the entry is alive
        p_pc._se3_records = (key, rec, pts, frames)
    except AttributeError:
        pass
    return rec


class ConvGeometry(LazyAttrs):
    """Device-resident record a conv call needs: forward CSR (int32), transposed CSR, coordinates
    and frames.  Built once per (neighbourhood, pc_in, pc_out) and cached on the neighbourhood; it
    replaces the reference's per-call sha256-keyed rot-tensor cache
    (layers/PNEConvLayerRotEquiv.py:71-128) -- the expanded [E*C,9] geometry is never materialised."""

    def __init__(self, p_pc_in, p_pc_out, p_neighborhood):
        nb = p_neighborhood.neighbors_
        if nb.dtype != torch.int64:
            nb = nb.to(torch.int64)
        nb = nb.contiguous()
        dev = nb.device
        self.n_in = int(p_pc_in.pts_.shape[0])
        self.n_out = int(p_pc_out.pts_.shape[0])
        self.n_edges = int(nb.shape[0])
        self.f_in = int(getattr(p_pc_in, "n_frames_", 1))
        self.f_out = int(getattr(p_pc_out, "n_frames_", 1))
        self.pts_in = p_pc_in.pts_.detach().to(torch.float32).contiguous()
        self.pts_out = p_pc_out.pts_.detach().to(torch.float32).contiguous()
        self.frames_in = p_pc_in.local_frames_.detach().to(torch.float32).contiguous()
        self.frames_out = p_pc_out.local_frames_.detach().to(torch.float32).contiguous()
        self.rec_in = gather_records(p_pc_in)
        self.rec_out = self.rec_in if p_pc_out is p_pc_in else gather_records(p_pc_out)
        if getattr(p_neighborhood, "keep_empty_", False):
            raise Se3Error("a keep_empty k-NN neighbourhood is padded with -1 sources (pc/KnnNeighborhood.py:100-135); "
                           "it cannot feed a convolution -- build it with p_keep_empty=False")
        self.row_ends = p_neighborhood.start_ids_.to(torch.int32).contiguous()
        if self.row_ends.shape[0] != self.n_out:
            raise Se3Error("neighbourhood has %d rows but the output cloud has %d points" %
                           (self.row_ends.shape[0], self.n_out))
        L = lib()
        e = max(self.n_edges, 1)
        self.col_src = torch.empty(e, dtype=torch.int32, device=dev)
        self.t_row_ends = torch.empty(max(self.n_in, 1), dtype=torch.int32, device=dev)
        self.t_edge = torch.empty(e, dtype=torch.int32, device=dev)
        self.t_dst = torch.empty(e, dtype=torch.int32, device=dev)
        ws = workspace(L.se3_csr_transpose_workspace_bytes(self.n_edges, self.n_in), dev, 'csr')
        check(L.se3_csr_transpose(ptr(nb), self.n_edges, self.n_in, self.n_out, ptr(ws), ws.numel(),
                                  ptr(self.col_src), ptr(self.t_row_ends), ptr(self.t_edge), ptr(self.t_dst),
                                  stream()), "se3_csr_transpose")


class Neighborhood(LazyAttrs, ABC):
    """Neighbourhood interface: `neighbors_` [E,2] (sample, source), `start_ids_` [M] inclusive
    row ends (pc/Neighborhood.py:7-37)."""

    def __init__(self, p_pc_src, p_samples):
        self.pc_src_ = p_pc_src
        self.samples_ = p_samples
        self.neighbors_ = None
        self.start_ids_ = None
        self.conv_geometry_cache_ = {}
        self.__compute_neighborhood__()

    @abstractmethod
    def __compute_neighborhood__(self):
        pass

    def conv_geometry(self, p_pc_in, p_pc_out):
        """Cached ConvGeometry; keyed on the identity of the clouds and their frame tensors, so a
        cloud whose frames were re-sampled gets a fresh record."""
        fi, fo = getattr(p_pc_in, "local_frames_", None), getattr(p_pc_out, "local_frames_", None)
        key = (id(p_pc_in), id(p_pc_out), id(fi), id(fo), self._neighbors_token())
        hit = self.conv_geometry_cache_.get(key)
        if hit is not None:
            geom = hit
            ver = geom.__dict__.get("_versions")
            if ver is None or ver == self._versions(p_pc_in, p_pc_out):
                return geom
        geom = ConvGeometry(p_pc_in, p_pc_out, self)
        # in-place edits of the coordinates / frames invalidate the record (tensor versions); the keyed objects are
        # held by the entry so that their ids cannot be recycled while it is cached
        geom._versions = self._versions(p_pc_in, p_pc_out)
        geom._keyed = (p_pc_in, p_pc_out, fi, fo)
        self.conv_geometry_cache_ = {key: geom}
        return geom

    @staticmethod
    def _versions(p_pc_in, p_pc_out):
        out = []
        for pc in (p_pc_in, p_pc_out):
            for name in ("pts_", "local_frames_"):
                t = getattr(pc, name, None)
                out.append(t._version if isinstance(t, torch.Tensor) else None)
        return tuple(out)

    def _neighbors_token(self):
        return id(self.neighbors_)

    def __repr__(self):
        return "### Neighbors:\n{}\n### Start indices:\n{}".format(self.neighbors_, self.start_ids_)


class BQNeighborhood(Neighborhood):
    """Ball-query neighbourhood (pc/BQNeighborhood.py:12-64)."""

    def __new__(cls, p_pc_src=None, p_samples=None, p_radius=None, p_max_neighbors=0):
        # A ball query the fused hierarchy builder has already answered (pc.build_point_hierarchy registers its
        # neighbourhoods on the source cloud): the models construct the seg-head neighbourhood directly
        # (models/FPNSegUNet.py:171-175), outside the hierarchy's cache -- hand back the prebuilt object.
        pre = getattr(p_pc_src, "_fused_bq_", None) if p_pc_src is not None else None
        if pre and not p_max_neighbors:
            ref = pre.get((id(p_samples), float(p_radius)))
            hit = ref() if ref is not None else None
            if hit is not None:
                return hit
        return super(BQNeighborhood, cls).__new__(cls)

    def __init__(self, p_pc_src, p_samples, p_radius, p_max_neighbors=0):
        if self.__dict__.get("_prebuilt_"):
            return
        self.radius_ = p_radius
        self.max_neighbors_ = p_max_neighbors
        super(BQNeighborhood, self).__init__(p_pc_src, p_samples)

    def __compute_neighborhood__(self):
        self.neighbors_, self.start_ids_ = BallQuery.apply(
            self.pc_src_.pts_, self.samples_.pts_, self.pc_src_.batch_ids_, self.samples_.batch_ids_, self.radius_,
            self.max_neighbors_, num_batches(self.pc_src_))

    # `neighbors_` [E,2] int64 (sample, source) is the reference contract.  A neighbourhood that came out of
    # the fused hierarchy builder carries the int32 CSR columns instead and materialises the pair list on
    # first access (the conv kernels never need it).
    @property
    def neighbors_(self):
        if self._neighbors is None and getattr(self, "_csr_columns", None) is not None:
            edge_dst, col_src = self._csr_columns
            self._neighbors = torch.stack((edge_dst.to(torch.int64), col_src.to(torch.int64)), dim=1)
        return self._neighbors

    @neighbors_.setter
    def neighbors_(self, v):
        self._neighbors = v

    def _neighbors_token(self):
        if self._neighbors is None and self.__dict__.get("_fused_token") is not None:
            return self._fused_token        # fused builder: the CSR columns are still unmaterialised arena windows
        cols = getattr(self, "_csr_columns", None)
        return id(cols[1]) if (self._neighbors is None and cols is not None) else id(self._neighbors)


class KnnNeighborhood(Neighborhood):
    """k-NN neighbourhood (pc/KnnNeighborhood.py:14-135), k <= 64 as in the reference.  The self-query branch runs on
    the sweep kernel (se3_knn_query); between two different clouds (the reference's torch_cluster.knn branch, used by
    the global-pooling convolutions, pc/KnnNeighborhood.py:78-84) every sample scans the sources of its batch item
    (se3_knn_cross).  `p_standard_knn` (the reference's deterministic torch_cluster path for evaluation) maps onto the
    same exact kernels: both are exact k-NN, ties broken by scan order."""

    def __init__(self, p_pc_src, p_samples, p_k, p_keep_empty=False, p_standard_knn=False):
        self.k_ = p_k
        self.keep_empty_ = p_keep_empty
        self.standard_knn_ = p_standard_knn
        super(KnnNeighborhood, self).__init__(p_pc_src, p_samples)

    def __compute_neighborhood__(self):
        if not 1 <= int(self.k_) <= 64:
            raise Se3Error("KnnNeighborhood: k must be in 1..64 (the reference's custom op has the same limit, "
                           "custom_ops/knn_query/knn_query.cu:135-197)")
        self._pairs = None
        self._ends = None
        if self.pc_src_ is self.samples_:
            # [N,k] int32 table straight from the sweep kernel; the [N*k,2] pair list and the row ends of
            # the reference layout are materialised lazily (the frame construction only needs the table)
            self.knn_table_ = KNNQuery.apply(self.pc_src_.pts_, self.pc_src_.batch_ids_, self.k_)
            return
        src, dst = self.pc_src_, self.samples_
        pts_s = src.pts_.detach().to(torch.float32).contiguous()
        pts_d = dst.pts_.detach().to(torch.float32).contiguous()
        b_s = src.batch_ids_.to(torch.int64)
        b_d = dst.batch_ids_.to(torch.int32).contiguous()
        nb = max(num_batches(src), num_batches(dst))
        ends = torch.searchsorted(b_s.contiguous(), torch.arange(nb, device=b_s.device, dtype=torch.int64),
                                  right=True).to(torch.int32).contiguous()
        m = int(pts_d.shape[0])
        out = torch.empty((m, int(self.k_)), dtype=torch.int32, device=pts_d.device)
        check(lib().se3_knn_cross(ptr(pts_s), ptr(ends), ptr(pts_d), ptr(b_d), m, int(self.k_), ptr(out), stream()),
              "se3_knn_cross")
        self.knn_table_ = out

    def _materialise(self):
        if self._pairs is not None:
            return
        cur = self.knn_table_
        n, dev = cur.shape[0], cur.device
        centers = torch.arange(n, dtype=torch.int32, device=dev).unsqueeze(1).expand(n, self.k_)
        pairs = torch.stack((centers.reshape(-1), cur.reshape(-1)), dim=-1)
        if self.keep_empty_:
            ends = (torch.arange(n, dtype=torch.int32, device=dev) + 1) * self.k_
        else:
            pairs = pairs[pairs[:, 1] >= 0]
            counts = torch.bincount(pairs[:, 0].to(torch.int64), minlength=n)
            ends = torch.cumsum(counts, 0).to(torch.int32)
        self._pairs, self._ends = pairs, ends

    @property
    def neighbors_(self):
        if getattr(self, "knn_table_", None) is None:
            return None
        self._materialise()
        return self._pairs

    @neighbors_.setter
    def neighbors_(self, v):
        self._pairs = v

    @property
    def start_ids_(self):
        if getattr(self, "knn_table_", None) is None:
            return None
        self._materialise()
        return self._ends

    @start_ids_.setter
    def start_ids_(self, v):
        self._ends = v
