"""Point-cloud structures with the reference's API (point_cloud_lib/point_cloud_lib/pc/__init__.py),
re-backed by the se3conv3d_b200 kernels."""
from .rotation_functions import (all_index_combinations, random_rotate, sample_reference_frames, get_relative_rot,
                                 change_points_to_local_frame, change_direction_to_local_frame, random_rotation,
                                 random_rotations, sample_global_reference_frames_pca, sample_reference_frames_pca,
                                 matrix_to_rotation_6d, quaternion_to_matrix, matrix_to_quaternion)
from .pointcloud import Pointcloud
from .grid import BoundingBox, Grid
from .neighborhood import Neighborhood, BQNeighborhood, KnnNeighborhood, ConvGeometry
from .subsample import SubSample, GridSubSample
from .pointcloud_rot_equiv import PointcloudRotEquiv
from .hierarchy import PointHierarchy, PointHierarchyRotEquiv
from .fused import build_point_hierarchy
