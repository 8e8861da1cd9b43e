"""`simple_knn._C` namespace of the reference extension (submodules/simple-knn/ext.cpp:15-17)."""
from gs_icp_slam_b200.knn import distCUDA2

__all__ = ["distCUDA2"]
