"""wildcat_slam_b200 — B200-native (sm_100a) implementation of Wildcat-SLAM's sliding-window odometry hot path:
surfel extraction, surfel correspondence, the robust window solve and the B-spline correction spreading, behind
the reference's src/odometry entry points (odometry.py) over an extern "C" ABI (include/wildcat_b200.h).

Importing the package never touches CUDA; the first Context does, and fails loudly without a device."""
from . import types  # noqa: F401

__all__ = ["types", "abi", "odometry", "synthetic", "build"]
