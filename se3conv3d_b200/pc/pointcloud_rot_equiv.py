import torch

from .._lib import lib, check, ptr, stream
from .pointcloud import Pointcloud, pool_by_index
from .rotation_functions import (sample_reference_frames, sample_reference_frames_pca,
                                 sample_global_reference_frames_pca)
from .neighborhood import KnnNeighborhood, BQNeighborhood


def _shuffle_and_take(all_frames, n_keep):
    """Uniform random per-point permutation of the candidate frames, keep the first n_keep
    (pc/PointcloudRotEquiv.py:148-168).  One torch.rand draw per point feeds se3_frames_select; the
    reference spends a torch.multinomial call on the same distribution."""
    all_frames = all_frames.contiguous()
    n, n_cand = all_frames.shape[0], all_frames.shape[1]
    u = torch.rand(n, device=all_frames.device, dtype=torch.float32)
    out = torch.empty((n, n_keep, 9), dtype=torch.float32, device=all_frames.device)
    check(lib().se3_frames_select(ptr(all_frames), ptr(u), n, n_cand, int(n_keep), ptr(out), stream()),
          "se3_frames_select")
    return out


class PointcloudRotEquiv(Pointcloud):
    """Point cloud with `n_frames_` local reference frames per point, `local_frames_` [N,F,9]
    (same constructor / attributes as pc/PointcloudRotEquiv.py:13-52).

    p_ref_frames_config keys: pca, neigh_method, neigh_kwargs, fixed_axis, n_frames."""

    def __init__(self, p_pts, p_batch_ids, p_ref_frames_config, ref_frames_pts=None, standard_knn=False, **kwargs):
        bs_host = kwargs.pop("batch_size_host", None)
        grads = kwargs.pop("requires_grad", False)
        super(PointcloudRotEquiv, self).__init__(p_pts, p_batch_ids, requires_grad=grads, batch_size_host=bs_host,
                                                 **kwargs)
        self.neigh_cache_ = {}
        self.local_frames_pca_cache_ = {}
        self.local_frames_config_ = p_ref_frames_config
        self.standard_knn_ = standard_knn
        self.ref_frames_pts = ref_frames_pts
        frames = self.get_local_ref_frames()
        self.n_frames_ = frames.shape[1]
        self.local_frames_ = torch.as_tensor(frames, **kwargs).contiguous()
        self._batch_ids_frames = None  # built on first use (only the pooling helpers read it)
        if self.pts_with_grads_:
            self.local_frames_.requires_grad = True

    @property
    def batch_ids_considering_frames_(self):
        """batch id of every (point, frame) feature row (pc/PointcloudRotEquiv.py:50)."""
        if self._batch_ids_frames is None:
            self._batch_ids_frames = torch.repeat_interleave(self.batch_ids_, self.n_frames_)
        return self._batch_ids_frames

    @batch_ids_considering_frames_.setter
    def batch_ids_considering_frames_(self, v):
        self._batch_ids_frames = v

    def get_ref_frame_neighborhood(self, p_neigh_method, **kwargs):
        key = str(p_neigh_method)
        if p_neigh_method == "knn":
            key += str(kwargs["neigh_k"])
        elif p_neigh_method == "ball_query":
            key += str(kwargs["bq_radius"])
        if key not in self.neigh_cache_:
            if p_neigh_method == "knn":
                # keep_empty: every point gets exactly k entries even in tiny clouds
                self.neigh_cache_[key] = KnnNeighborhood(self, self, kwargs["neigh_k"], p_keep_empty=True,
                                                         p_standard_knn=self.standard_knn_)
            elif p_neigh_method == "ball_query":
                self.neigh_cache_[key] = BQNeighborhood(self, self, kwargs["bq_radius"])
            else:
                raise ValueError("unknown neighbourhood method " + str(p_neigh_method))
        return self.neigh_cache_[key]

    def get_local_ref_frames(self):
        cfg = self.local_frames_config_
        dev = self.pts_.device
        if cfg["pca"]:
            if "se3-all" not in self.local_frames_pca_cache_:
                if self.ref_frames_pts is not None:
                    b = self.pts_.shape[0]  # one point per batch item
                    pts = self.ref_frames_pts.reshape(b, -1, self.ref_frames_pts.shape[-1])
                    cand = sample_global_reference_frames_pca(pts, axis_fixed=cfg["fixed_axis"], device=dev)
                else:
                    neigh = self.get_ref_frame_neighborhood(cfg["neigh_method"], **cfg["neigh_kwargs"])
                    cand = sample_reference_frames_pca(self.pts_, neigh, axis_fixed=cfg["fixed_axis"], device=dev)
                self.local_frames_pca_cache_["se3-all"] = cand
            return _shuffle_and_take(self.local_frames_pca_cache_["se3-all"], cfg["n_frames"])
        n_origins = 1 if self.ref_frames_pts is not None else self.pts_.shape[0]
        return sample_reference_frames(n_origins=n_origins, n_frames=cfg["n_frames"], axis_fixed=cfg["fixed_axis"],
                                       device=dev)

    def to_device(self, p_device):
        super(PointcloudRotEquiv, self).to_device(p_device)
        self.local_frames_ = self.local_frames_.to(p_device)
        self.batch_ids_considering_frames_ = self.batch_ids_considering_frames_.to(p_device)

    def feature_pooling(self, p_in_tensor, p_pooling_method="avg"):
        """Pools the F per-frame feature rows of every point (pc/PointcloudRotEquiv.py:224-251): one kernel each way
        (se3_frame_pool_fwd / _bwd) on CUDA float32 matrices."""
        f = self.local_frames_config_["n_frames"]
        n = self.pts_.shape[0]
        if p_pooling_method not in ("avg", "sum", "max", "min"):
            raise ValueError("unknown pooling method " + str(p_pooling_method))
        if p_in_tensor.is_cuda and p_in_tensor.dim() == 2 and p_in_tensor.dtype == torch.float32:
            from ..custom_ops.functions import FramePool, POOL_MODES
            return FramePool.apply(p_in_tensor, f, POOL_MODES[p_pooling_method])
        x = p_in_tensor.reshape(n, f, *p_in_tensor.shape[1:])
        if p_pooling_method == "avg":
            return x.mean(dim=1)
        if p_pooling_method == "sum":
            return x.sum(dim=1)
        if p_pooling_method == "max":
            return x.max(dim=1)[0]
        return x.min(dim=1)[0]

    def _item_rows(self, frames):
        """(inclusive row ends per batch item int32 [B], item of every row int32) for rows = points x frames."""
        key = "_item_rows_%d" % frames
        hit = self.__dict__.get(key)
        if hit is None:
            b = self.batch_ids_.to(torch.int64)
            nb = int(getattr(self, "batch_size_host_", None) or self.batch_size_)
            ends = (torch.searchsorted(b.contiguous(), torch.arange(nb, device=b.device), right=True) * frames).to(torch.int32)
            rows = torch.repeat_interleave(self.batch_ids_.to(torch.int32), frames) if frames > 1 else self.batch_ids_.to(torch.int32)
            hit = (ends.contiguous(), rows.contiguous())
            self.__dict__[key] = hit
        return hit

    def _batch_pool(self, x, frames, method):
        if method in ("avg", "sum") and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32:
            from ..custom_ops.functions import BatchPool, POOL_MODES
            ends, rows = self._item_rows(frames)
            return BatchPool.apply(x, ends, rows, POOL_MODES[method])
        ids = (self.batch_ids_considering_frames_ if frames > 1 else self.batch_ids_).to(torch.int64)
        return pool_by_index(x, ids, method)

    def global_pooling_specific_feature_pooling(self, p_in_tensor, p_global_pooling_method="avg",
                                                p_feature_pooling_method="avg"):
        pooled = self.feature_pooling(p_in_tensor, p_pooling_method=p_feature_pooling_method)
        return self._batch_pool(pooled, 1, p_global_pooling_method)

    def global_pooling(self, p_in_tensor, p_pooling_method="avg"):
        return self._batch_pool(p_in_tensor, self.n_frames_, p_pooling_method)

    def global_upsample(self, p_in_tensor):
        return torch.index_select(p_in_tensor, 0, self.batch_ids_considering_frames_.to(torch.int64))

    def __repr__(self):
        return "### Points:\n{}\n### Local Ref Frames:\n{}\n### Batch Ids:\n{}\n### Batch Size:\n{}".format(
            self.pts_, self.local_frames_, self.batch_ids_, self.batch_size_)
