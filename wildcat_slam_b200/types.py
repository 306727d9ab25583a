"""POD types of the C ABI (include/wildcat_b200.h) as numpy dtypes / ctypes structures.

Field-for-field mirrors of the reference's value types:
  POINT48  <- hilti_ros::Point            src/common/common.h:12-28
  SURFEL   <- Surfel                      src/odometry/surfel.h:35-127
  SAMPLE   <- SampleState                 src/odometry/surfel.h:9-23
  IMU      <- ImuState                    src/odometry/surfel.h:25-33
  CORR     <- SurfelCorrespondence        src/odometry/surfel.h:124-127 (as indices)
Sizes are asserted against the header's static sizes by tests/test_abi.py.
"""
import ctypes as C

import numpy as np

POINT48 = np.dtype(
    {
        "names": ["x", "y", "z", "pad", "intensity", "time", "ring"],
        "formats": ["<f4", "<f4", "<f4", "<f4", "<f4", "<f8", "<u2"],
        "offsets": [0, 4, 8, 12, 16, 24, 32],
        "itemsize": 48,
    }
)
SURFEL = np.dtype(
    [
        ("timestamp", "<f8"),
        ("resolution", "<f8"),
        ("plane_std_deviation", "<f8"),
        ("rot", "<f8", (4,)),  # x, y, z, w
        ("pos", "<f8", (3,)),
        ("center", "<f8", (3,)),
        ("covariance", "<f8", (9,)),
        ("norm", "<f8", (3,)),
        ("is_in_body_frame", "<i4"),
        ("_pad", "<i4"),
    ]
)
CORR = np.dtype([("s1", "<i4"), ("s2", "<i4")])
SAMPLE = np.dtype(
    [
        ("timestamp", "<f8"),
        ("data_cor", "<f8", (12,)),  # rot_cor, pos_cor, bg, ba
        ("grav", "<f8", (3,)),
        ("rot", "<f8", (4,)),
        ("pos", "<f8", (3,)),
    ]
)
IMU = np.dtype(
    [("timestamp", "<f8"), ("pos", "<f8", (3,)), ("rot", "<f8", (4,)), ("acc", "<f8", (3,)), ("gyr", "<f8", (3,))]
)
MARKER = np.dtype([("position", "<f8", (3,)), ("orientation", "<f8", (4,)), ("scale", "<f8", (3,)), ("color", "<f4", (4,))])
ASSIGN = np.dtype([("vx", "<i4"), ("vy", "<i4"), ("vz", "<i4"), ("leaf", "<i4")])

assert POINT48.itemsize == 48 and SURFEL.itemsize == 208 and SAMPLE.itemsize == 184 and IMU.itemsize == 112

WC_MAX_ITER_LOG = 128

# wc_status
WC_OK, WC_EINVAL, WC_EINVAL_TIME_ORDER, WC_EOUT_OF_SPAN, WC_ETOO_FEW_TARGETS = 0, 1, 2, 3, 4
WC_ECAPACITY, WC_ECUDA, WC_ECOMM, WC_ENUMERIC = 5, 6, 7, 8
WC_JAC_REFERENCE_OVERWRITE, WC_JAC_EXACT = 0, 1
WC_PREC_F64, WC_PREC_MIXED, WC_PREC_F32 = 0, 1, 2
TERMINATION = ["NO_CONVERGENCE", "FUNCTION_TOL", "GRADIENT_TOL", "PARAMETER_TOL", "MIN_RADIUS", "FAILURE"]


class Params(C.Structure):
    """wc_params; defaults = the reference's compile-time constants (see wc_default_params)."""

    _fields_ = [
        ("voxel_size", C.c_float),
        ("max_layer", C.c_int32),
        ("layer_point_size", C.c_int32 * 3),
        ("cluster_min_points", C.c_int32),
        ("planer_threshold", C.c_float),
        ("min_plane_likeness", C.c_double),
        ("cluster_time_gap", C.c_double),
        ("view_point", C.c_double * 3),
        ("center_dist_threshold", C.c_double),
        ("angular_dist_threshold", C.c_double),
        ("surfel_dist_threshold", C.c_double),
        ("knn_candidates", C.c_int32),
        ("_pad0", C.c_int32),
        ("time_diff_threshold", C.c_double),
        ("cauchy_a", C.c_double),
        ("weight_floor", C.c_double),
        ("imu_rate", C.c_double),
        ("weight_gyr", C.c_double),
        ("weight_acc", C.c_double),
        ("weight_bg", C.c_double),
        ("weight_ba", C.c_double),
        ("max_points", C.c_int64),
        ("max_surfels", C.c_int64),
        ("max_corrs", C.c_int64),
        ("max_samples", C.c_int32),
        ("max_imu_states", C.c_int32),
    ]


class SolveOpts(C.Structure):
    _fields_ = [
        ("max_num_iterations", C.c_int32),
        ("jacobian_mode", C.c_int32),
        ("fix_first_position", C.c_int32),
        ("use_imu_factors", C.c_int32),
        ("initial_trust_region_radius", C.c_double),
        ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double),
        ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("precision", C.c_int32),
        ("_pad", C.c_int32),
    ]


class SolveSummary(C.Structure):
    _fields_ = [
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
        ("num_iterations", C.c_int32),
        ("num_successful_steps", C.c_int32),
        ("num_unsuccessful_steps", C.c_int32),
        ("termination", C.c_int32),
        ("num_residual_blocks_sld", C.c_int32),
        ("num_residual_blocks_fix", C.c_int32),
        ("num_residual_blocks_imu", C.c_int32),
        ("num_linearizations", C.c_int32),
        ("iter_cost", C.c_double * WC_MAX_ITER_LOG),
        ("iter_radius", C.c_double * WC_MAX_ITER_LOG),
        ("iter_accepted", C.c_int8 * WC_MAX_ITER_LOG),
        ("gpu_ms_total", C.c_double),
        ("gpu_ms_linearize", C.c_double),
    ]


class PassStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("ms_total", "ms_extract", "ms_extract_keys", "ms_extract_emit", "ms_match", "ms_pack",
                                          "ms_solve")] + [(n, C.c_int64) for n in ("n_surfels", "n_sld_corr", "n_fix_corr", "n_launches")]


class Pc2Layout(C.Structure):
    """wc_pc2_layout: where the fields of hilti_ros::Point sit inside one PointCloud2 point (-1: absent)."""
    _fields_ = [("point_step", C.c_uint32), ("off_x", C.c_int32), ("off_y", C.c_int32), ("off_z", C.c_int32),
                ("off_intensity", C.c_int32), ("off_time", C.c_int32), ("off_ring", C.c_int32)]


class SweepFilter(C.Structure):
    """wc_sweep_filter: lidar -> IMU extrinsic, range limits and blind box (lio_config.h:18-30)."""

    _fields_ = [("ext_q", C.c_double * 4), ("ext_t", C.c_double * 3), ("min_range", C.c_double), ("max_range", C.c_double),
                ("blind_box_min", C.c_double * 3), ("blind_box_max", C.c_double * 3)]


def default_sweep_filter() -> SweepFilter:
    """Python-side copy of wc_default_sweep_filter (lio_config.h:18-30); the quaternion is Eigen's conversion of the
    rotation matrix [[-5.32125e-08, -1, 0], [-1, -5.32125e-08, 0], [0, 0, -1]] (trace <= 0 branch, i = 0)."""
    import math

    f = SweepFilter()
    t = math.sqrt(-5.32125e-08 - -5.32125e-08 - -1.0 + 1.0)
    f.ext_q[:] = [0.5 * t, (-1.0 + -1.0) * (0.5 / t), 0.0, 0.0]
    f.ext_t[:] = [-0.001, -0.00855, 0.055]
    f.min_range, f.max_range = 0.3, 120.0
    f.blind_box_min[:] = [-0.8, -0.5, -0.4]
    f.blind_box_max[:] = [0.3, 0.5, 0.4]
    return f


def default_params() -> Params:
    """Python-side copy of wc_default_params (the library's own is checked against this in tests).

    surfel_extraction.cc:327,24,33; knn_surfel_matcher.h:37-41; cost_functor.h:24,112;
    lidar_odometry.cc:270,309; lio_config.h:10-14,32,42-45.
    """
    import math

    p = Params()
    p.voxel_size = 0.8
    p.max_layer = 2
    p.layer_point_size[:] = [20, 20, 20]
    p.cluster_min_points = 20
    p.planer_threshold = 0.01
    p.min_plane_likeness = 0.1
    p.cluster_time_gap = 0.05
    p.view_point[:] = [0.0, 0.0, 0.0]
    p.center_dist_threshold = 1.0
    p.angular_dist_threshold = 5.0 * math.pi / 180.0
    p.surfel_dist_threshold = 0.1
    p.knn_candidates = 10
    p.time_diff_threshold = 0.06
    p.cauchy_a = 0.4
    p.weight_floor = math.pow(0.05 / 6, 2)
    p.imu_rate = 200.0
    gnd, and_, grw, arw, w = 0.00015198973532354657, 0.006308226052016165, 0.00011673723527962174, 2.664506559330434e-06, 0.01
    p.weight_gyr = 1 / (gnd * math.sqrt(p.imu_rate)) * w
    p.weight_acc = 1 / (and_ * math.sqrt(p.imu_rate)) * w
    p.weight_bg = 1 / (grw / math.sqrt(p.imu_rate)) * w
    p.weight_ba = 1 / (arw / math.sqrt(p.imu_rate)) * w
    p.max_points = 1 << 21
    p.max_surfels = 1 << 18
    p.max_corrs = 1 << 19
    p.max_samples = 128
    p.max_imu_states = 8192
    return p


def default_solve_opts() -> SolveOpts:
    """ceres::Solver::Options as left by lidar_odometry.cc:551-554."""
    o = SolveOpts()
    o.max_num_iterations = 100
    o.jacobian_mode = WC_JAC_REFERENCE_OVERWRITE
    o.fix_first_position = 1
    o.use_imu_factors = 1
    o.initial_trust_region_radius = 1e4
    o.max_trust_region_radius = 1e16
    o.min_trust_region_radius = 1e-32
    o.min_relative_decrease = 1e-3
    o.min_lm_diagonal = 1e-6
    o.max_lm_diagonal = 1e32
    o.function_tolerance = 1e-6
    o.gradient_tolerance = 1e-10
    o.parameter_tolerance = 1e-8
    return o


def ptr(a):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)
