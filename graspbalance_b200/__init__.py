"""graspbalance_b200 -- B200-native (sm_100a) point-cloud operators for GraspBalance, behind the reference's own API.

Python modules mirror the reference files they replace:

    graspbalance_b200._ext                  <- pointnet2._ext          (PointNet/_ext_src, pybind module A)
    graspbalance_b200.pointnet2_batch_cuda  <- pointnet2_batch_cuda    (pointnet2_batch/src, pybind module B)
    graspbalance_b200.knn_C                 <- KNN._C                  (KNN/Pytorch_CUDA_KNN, pybind module C)
    graspbalance_b200.pointnet2_utils       <- PointNet/pointnet2_utils.py
    graspbalance_b200.group / subsample / upsampling  <- ModifiedNetTools/{group,subsample,upsampling}.py
    graspbalance_b200.knn_modules           <- KNN/knn_modules.py
    graspbalance_b200.collision_detector    <- collision_detector.py
    graspbalance_b200.modules               <- TrainModel/modules.py: GraspWidthGrouping (fused 4-depth grasp crop)

All compute goes through libgbops.so (C ABI in include/gbops.h).  There is no CPU, Triton or PyTorch fallback.
"""
import sys
import types

__version__ = "0.1.0"


def install_as_reference_modules():
    """Register this package's native-module replacements under the names the reference's own Python files import
    (`pointnet2._ext`, `pointnet2_batch_cuda`, `KNN._C`), so that the unmodified reference sources
    (PointNet/pointnet2_utils.py, ModifiedNetTools/*.py, KNN/knn_modules.py) run on the B200 kernels."""
    from . import _ext, knn_C, pointnet2_batch_cuda
    pkg = types.ModuleType("pointnet2")
    pkg._ext = _ext
    pkg.__path__ = []
    sys.modules["pointnet2"] = pkg
    sys.modules["pointnet2._ext"] = _ext
    sys.modules["pointnet2_batch_cuda"] = pointnet2_batch_cuda
    knn_pkg = types.ModuleType("KNN")
    knn_pkg._C = knn_C
    knn_pkg.__path__ = []
    sys.modules["KNN"] = knn_pkg
    sys.modules["KNN._C"] = knn_C
