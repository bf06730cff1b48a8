from .pointcloud import Pointcloud
from .pointcloud_rot_equiv import PointcloudRotEquiv
from .subsample import GridSubSample
from .neighborhood import KnnNeighborhood, BQNeighborhood
from .._lib import Se3Error, num_batches, LazyAttrs


class PointHierarchy(LazyAttrs):
    """Hierarchy of progressively sub-sampled clouds: `pcs_`, `sub_sampled_objs_`, `neigh_cache_`
    (pc/PointHierarchy.py:10-92; kwargs `grid_radii`, and the misspelt `neihg_k`, are API)."""

    def __init__(self, p_point_cloud, p_num_sub_samples, p_subsample_method="grid_avg", **kwargs):
        self.sub_sampled_objs_ = []
        self.pcs_ = [p_point_cloud]
        for level in range(p_num_sub_samples):
            new_pc, samp = self.__create_sub_sample__(self.pcs_[-1], p_subsample_method, level, **kwargs)
            self.sub_sampled_objs_.append(samp)
            self.pcs_.append(new_pc)
        self.neigh_cache_ = {}

    @staticmethod
    def __make_sampler__(p_point_cloud, p_samp_method, p_id, **kwargs):
        if p_samp_method == "grid_avg":
            return GridSubSample(p_point_cloud, kwargs["grid_radii"][p_id], False)
        if p_samp_method == "grid_rnd":
            return GridSubSample(p_point_cloud, kwargs["grid_radii"][p_id], True)
        if p_samp_method == "fps":
            raise Se3Error("farthest-point sub-sampling (torch_cluster.fps) is outside the B200 hot path; "
                           "no shipped config uses it")
        raise ValueError("unknown sub-sample method " + str(p_samp_method))

    def __create_sub_sample__(self, p_point_cloud, p_samp_method, p_id, **kwargs):
        samp = self.__make_sampler__(p_point_cloud, p_samp_method, p_id, **kwargs)
        new_pts = samp.__subsample_tensor__(p_point_cloud.pts_, "avg")
        new_batch_ids = samp.__subsample_tensor__(p_point_cloud.batch_ids_, "max")
        new_pc = Pointcloud(new_pts, new_batch_ids)
        new_pc.batch_size_host_ = num_batches(p_point_cloud)
        return new_pc, samp

    def create_neighborhood(self, p_pc_src_id, p_pc_dest_id, p_neigh_method, **kwargs):
        key = str(p_pc_src_id) + "_" + str(p_pc_dest_id) + "_" + p_neigh_method
        if p_neigh_method == "knn":
            key += str(kwargs["neihg_k"])
        elif p_neigh_method == "ball_query":
            key += str(kwargs["bq_radius"])
        if key not in self.neigh_cache_:
            src, dst = self.pcs_[p_pc_src_id], self.pcs_[p_pc_dest_id]
            if p_neigh_method == "knn":
                self.neigh_cache_[key] = KnnNeighborhood(src, dst, kwargs["neihg_k"])
            elif p_neigh_method == "ball_query":
                self.neigh_cache_[key] = BQNeighborhood(src, dst, kwargs["bq_radius"])
            else:
                raise ValueError("unknown neighbourhood method " + str(p_neigh_method))
        return self.neigh_cache_[key]

    def clear_neigh_cache(self):
        self.neigh_cache_ = {}

    def pool_tensor(self, p_tensor, p_pc_src_id, p_pc_dest_id, p_pool_method):
        assert p_pc_dest_id - p_pc_src_id == 1
        return self.sub_sampled_objs_[p_pc_src_id].__subsample_tensor__(p_tensor, p_pool_method)

    def upsample_tensor(self, p_tensor, p_pc_src_id, p_pc_dest_id):
        assert p_pc_src_id - p_pc_dest_id == 1
        return self.sub_sampled_objs_[p_pc_dest_id].__upsample_tensor__(p_tensor)


class PointHierarchyRotEquiv(PointHierarchy):
    """Hierarchy whose every level is a PointcloudRotEquiv with freshly built frames
    (pc/PointHierarchyRotEquiv.py:7-44)."""

    def __create_sub_sample__(self, p_point_cloud, p_samp_method, p_id, **kwargs):
        samp = self.__make_sampler__(p_point_cloud, p_samp_method, p_id, **kwargs)
        new_pts = samp.__subsample_tensor__(p_point_cloud.pts_, "avg")
        new_batch_ids = samp.__subsample_tensor__(p_point_cloud.batch_ids_, "max")
        return PointcloudRotEquiv(new_pts, new_batch_ids, p_point_cloud.local_frames_config_,
                                  batch_size_host=num_batches(p_point_cloud)), samp
