from .pc_sam import PointCloudSAM, PointSAM, build_point_sam  # noqa: F401
