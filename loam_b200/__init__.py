"""loam_b200 — B200-native (sm_100a) LOAM feature extraction + feature registration.

Python mirror of the reference's `loam_python` module (python/loam_bindings.cpp): same class,
function and argument names, backed by hand-written CUDA kernels through the C-ABI in
include/loamgpu.h.  There is no CPU path in this package.
"""
from .api import (FeatureExtractionParams, LidarParams, LoamFeatures, Pose3d, Quaterniond, RegistrationDetail,
                  RegistrationIterationInfo, RegistrationParams, RegistrationTerminationType, computeCurvature,
                  computeValidPoints, extractFeatures, extractFeatureIndices, registerFeatures, odometry,
                  get_context, release_context, get_multi_context, LocalMap, extractFeaturesDewarped)

CONVERGED = RegistrationTerminationType.CONVERGED
MAX_ITER = RegistrationTerminationType.MAX_ITER
INSUFFICIENT_ASSOCIATIONS = RegistrationTerminationType.INSUFFICIENT_ASSOCIATIONS

__all__ = [
    "LidarParams", "FeatureExtractionParams", "RegistrationParams", "Pose3d", "Quaterniond", "LoamFeatures",
    "RegistrationDetail", "RegistrationIterationInfo", "RegistrationTerminationType", "extractFeatures",
    "computeCurvature", "computeValidPoints", "registerFeatures", "extractFeatureIndices", "odometry",
    "extractFeaturesDewarped",
    "get_context", "release_context", "get_multi_context", "LocalMap", "CONVERGED", "MAX_ITER", "INSUFFICIENT_ASSOCIATIONS",
]
