// ORACLE — TEST INFRASTRUCTURE ONLY (see wc_math.h / wc_oracle.h).
// Surfel extraction, surfel pose update, undistortion and correspondence search restated from
//   src/odometry/surfel_extraction.{h,cc}, src/odometry/surfel.h, src/odometry/knn_surfel_matcher.{h,cc},
//   src/odometry/lidar_odometry.cc:143-170.
// PARITY UNPINNED for BuildSurfels and Match (no reference test touches them).
#include <cstdint>
#include <deque>
#include <memory>
#include <set>
#include <unordered_map>
#include <vector>

#include "wc_math.h"
#include "wc_oracle.h"

using namespace wco;

namespace {

// surfel_extraction.h:22-25
struct PointWithCov {
  double timestamp;
  V3     pw;
};

// surfel_extraction.h:55-81
struct VoxelLoc {
  int32_t x, y, z;
  VoxelLoc(const V3& pos, double resolution) {
    x = (int32_t)std::floor(pos.x / resolution);
    y = (int32_t)std::floor(pos.y / resolution);
    z = (int32_t)std::floor(pos.z / resolution);
  }
  bool operator==(const VoxelLoc& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct VoxelLocHash {
  size_t operator()(const VoxelLoc& s) const {
    const int64_t HASH_P = 116101, MAX_N = 10000000000LL;
    return (size_t)(((((int64_t)s.z * HASH_P) % MAX_N + s.y) * HASH_P) % MAX_N + s.x);
  }
};

struct Ctx {
  const wc_params*              prm;
  std::vector<wc_surfel>*       out;
  std::vector<wco_surfel_info>* info;
  double                        rel_margin;
  int64_t                       near_threshold = 0;
  void                          Margin(double lam0, double likeness, double thr, double min_like) {
    if (std::fabs(lam0 - thr) <= rel_margin * thr || std::fabs(likeness - min_like) <= rel_margin * min_like) ++near_threshold;
  }
};

struct Plane {
  bool is_plane = false;
};

// surfel_extraction.cc:12-65
void ClusterSurfels(Ctx& cx, const std::vector<PointWithCov>& points, double resolution, const V3& view_point,
                    double planer_threshold, double min_plane_likeness, int layer) {
  std::vector<std::vector<PointWithCov>> cluster_points;
  cluster_points.push_back({points[0]});
  for (size_t i = 1; i < points.size(); ++i) {
    if (points[i].timestamp - cluster_points.back().back().timestamp > cx.prm->cluster_time_gap) {
      cluster_points.push_back({points[i]});
    } else {
      cluster_points.back().push_back(points[i]);
    }
  }
  for (auto& cluster : cluster_points) {
    if ((int)cluster.size() < cx.prm->cluster_min_points) continue;
    V3     center;
    M3     covariance;
    double timestamp   = 0;
    int    points_size = (int)cluster.size();
    for (const auto& pv : cluster) {
      covariance = covariance + outer(pv.pw, pv.pw);
      center     = center + pv.pw;
      timestamp += pv.timestamp;
    }
    center     = center / points_size;
    timestamp  = timestamp / points_size;
    covariance = covariance / points_size - outer(center, center);

    double evals[3];
    M3     evecs;
    SymEig3(covariance, evals, evecs);
    double plane_likeness = 2 * (evals[1] - evals[0]) / (evals[0] + evals[1] + evals[2]);
    cx.Margin(evals[0], plane_likeness, planer_threshold, min_plane_likeness);
    if (evals[0] > planer_threshold || plane_likeness < min_plane_likeness) continue;

    V3 nrm = evecs.col(0);
    if (dot(nrm, center - view_point) < 0) nrm = -nrm;
    wc_surfel s;
    std::memset(&s, 0, sizeof(s));
    s.timestamp           = timestamp;
    s.resolution          = resolution;
    s.plane_std_deviation = std::sqrt(evals[0]);
    s.rot[3]              = 1.0;  // Quaterniond rot{1,0,0,0}, surfel.h:114
    center.store(s.center);
    covariance.store(s.covariance);
    nrm.store(s.norm);
    s.is_in_body_frame = 0;
    cx.out->push_back(s);
    if (cx.info) {
      wco_surfel_info si;
      si.n_points = points_size, si.layer = layer;
      si.evals[0] = evals[0], si.evals[1] = evals[1], si.evals[2] = evals[2];
      si.likeness = plane_likeness;
      cx.info->push_back(si);
    }
  }
}

// surfel_extraction.h:83-123, .cc:69-184,304-314
struct OctoTree {
  std::vector<PointWithCov> temp_points_;
  Plane                     plane_;
  int                       layer_;
  std::unique_ptr<OctoTree> leaves_[8];
  double                    voxel_center_[3];
  float                     quarter_length_;
  Ctx*                      cx_;

  OctoTree(Ctx* cx, int layer) : layer_(layer), cx_(cx) {}
  int Threshold() const { return cx_->prm->layer_point_size[layer_ < 3 ? layer_ : 2]; }

  void InitPlane(const std::vector<PointWithCov>& points, Plane* plane) {
    M3  covariance;
    V3  center;
    int n = (int)points.size();
    for (const auto& pv : points) {
      covariance = covariance + outer(pv.pw, pv.pw);
      center     = center + pv.pw;
    }
    center     = center / n;
    covariance = covariance / n - outer(center, center);
    double evals[3];
    M3     evecs;
    SymEig3(covariance, evals, evecs);
    double plane_likeness = 2 * (evals[1] - evals[0]) / (evals[0] + evals[1] + evals[2]);
    double thr            = (double)cx_->prm->planer_threshold;  // float member compared in double
    cx_->Margin(evals[0], plane_likeness, thr, cx_->prm->min_plane_likeness);
    plane->is_plane = (evals[0] < thr && plane_likeness > cx_->prm->min_plane_likeness);
  }

  void InitOctoTree() {
    if ((int)temp_points_.size() > Threshold()) {
      InitPlane(temp_points_, &plane_);
      CutOctoTree();  // forced split whether planar or not (.cc:131-138)
    }
  }

  void CutOctoTree() {
    if (layer_ >= cx_->prm->max_layer) return;
    for (size_t i = 0; i < temp_points_.size(); i++) {
      int xyz[3] = {0, 0, 0};
      if (temp_points_[i].pw.x > voxel_center_[0]) xyz[0] = 1;
      if (temp_points_[i].pw.y > voxel_center_[1]) xyz[1] = 1;
      if (temp_points_[i].pw.z > voxel_center_[2]) xyz[2] = 1;
      int leafnum = 4 * xyz[0] + 2 * xyz[1] + xyz[2];
      if (!leaves_[leafnum]) {
        leaves_[leafnum].reset(new OctoTree(cx_, layer_ + 1));
        // int * float -> float, double + float -> double (.cc:163-166)
        leaves_[leafnum]->voxel_center_[0] = voxel_center_[0] + (float)((2 * xyz[0] - 1) * quarter_length_);
        leaves_[leafnum]->voxel_center_[1] = voxel_center_[1] + (float)((2 * xyz[1] - 1) * quarter_length_);
        leaves_[leafnum]->voxel_center_[2] = voxel_center_[2] + (float)((2 * xyz[2] - 1) * quarter_length_);
        leaves_[leafnum]->quarter_length_  = quarter_length_ / 2;
      }
      leaves_[leafnum]->temp_points_.push_back(temp_points_[i]);
    }
    for (int i = 0; i < 8; i++) {
      if (leaves_[i] && (int)leaves_[i]->temp_points_.size() > leaves_[i]->Threshold()) {
        InitPlane(leaves_[i]->temp_points_, &leaves_[i]->plane_);
        if (!leaves_[i]->plane_.is_plane) leaves_[i]->CutOctoTree();
      }
    }
  }

  void ExtractSurfelInfo() {
    if (plane_.is_plane) {
      V3 view(cx_->prm->view_point);
      ClusterSurfels(*cx_, temp_points_, (double)(float)(quarter_length_ * 4), view, (double)cx_->prm->planer_threshold,
                     cx_->prm->min_plane_likeness, layer_);
    }
    for (auto& leaf : leaves_)
      if (leaf) leaf->ExtractSurfelInfo();
  }
};

// total order used for the final sort: the reference sorts by timestamp only (surfel_extraction.cc:334,
// std::sort, ties unspecified — Q5); ties are broken by (resolution descending, center) on both sides of
// the parity check.
bool SurfelLess(const wc_surfel& a, const wc_surfel& b) {
  if (a.timestamp != b.timestamp) return a.timestamp < b.timestamp;
  if (a.resolution != b.resolution) return a.resolution > b.resolution;
  for (int k = 0; k < 3; ++k)
    if (a.center[k] != b.center[k]) return a.center[k] < b.center[k];
  return false;
}

}  // namespace

extern "C" int64_t wco_build_surfels(const wc_params* prm, const wc_point48* pts, int64_t n, wc_surfel* out,
                                     int64_t cap, wc_point_assign* assign, wco_surfel_info* info_out,
                                     double rel_margin, int64_t* near_threshold) {
  // surfel_extraction.cc:317-324
  std::vector<PointWithCov> points;
  points.reserve(n);
  for (int64_t i = 0; i < n; ++i) {
    PointWithCov np;
    np.timestamp = pts[i].time;
    np.pw        = V3((double)pts[i].x, (double)pts[i].y, (double)pts[i].z);
    points.push_back(np);
  }
  std::vector<wc_surfel>       surfels;
  std::vector<wco_surfel_info> infos;
  Ctx                          cx;
  cx.prm = prm, cx.out = &surfels, cx.info = info_out ? &infos : nullptr, cx.rel_margin = rel_margin;

  // BuildVoxelMap, surfel_extraction.cc:186-220.  voxel_size is a float promoted to double (Q2).
  const float                                                            voxel_size = prm->voxel_size;
  std::unordered_map<VoxelLoc, std::unique_ptr<OctoTree>, VoxelLocHash> feat_map;
  for (int64_t i = 0; i < n; ++i) {
    const PointWithCov& p_v = points[i];
    VoxelLoc            position(p_v.pw, (double)voxel_size);
    auto                it = feat_map.find(position);
    if (it == feat_map.end()) {
      std::unique_ptr<OctoTree> t(new OctoTree(&cx, 0));
      t->quarter_length_  = voxel_size / 4;
      t->voxel_center_[0] = (0.5 + position.x) * voxel_size;
      t->voxel_center_[1] = (0.5 + position.y) * voxel_size;
      t->voxel_center_[2] = (0.5 + position.z) * voxel_size;
      it                  = feat_map.emplace(position, std::move(t)).first;
    }
    it->second->temp_points_.push_back(p_v);
    if (assign) {
      // the child codes CutOctoTree (.cc:148-166) would assign on the way down, for every point
      const OctoTree& r = *it->second;
      int             code[2];
      double          c[3] = {r.voxel_center_[0], r.voxel_center_[1], r.voxel_center_[2]};
      float           q    = r.quarter_length_;
      for (int l = 0; l < 2; ++l) {
        int b[3] = {p_v.pw.x > c[0], p_v.pw.y > c[1], p_v.pw.z > c[2]};
        code[l]  = 4 * b[0] + 2 * b[1] + b[2];
        for (int k = 0; k < 3; ++k) c[k] = c[k] + (float)((2 * b[k] - 1) * q);
        q = q / 2;
      }
      assign[i].vx = position.x, assign[i].vy = position.y, assign[i].vz = position.z;
      assign[i].leaf = 8 * code[0] + code[1];
    }
  }
  for (auto& e : feat_map) e.second->InitOctoTree();
  for (auto& e : feat_map) e.second->ExtractSurfelInfo();

  std::vector<size_t> order(surfels.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return SurfelLess(surfels[a], surfels[b]); });
  if (near_threshold) *near_threshold = cx.near_threshold;
  if ((int64_t)surfels.size() > cap) return -1;
  for (size_t i = 0; i < order.size(); ++i) {
    out[i] = surfels[order[i]];
    if (info_out) info_out[i] = infos[order[i]];
  }
  return (int64_t)surfels.size();
}

// ---------------------------------------------------------------------------------------------------
// lidar_odometry.cc:143-170; surfel.h:48-58
// ---------------------------------------------------------------------------------------------------
namespace {
int64_t ImuLowerBound(const wc_imu_state* imu, int64_t n_imu, double t) {
  int64_t lo = 0, hi = n_imu;  // first idx with imu[idx].timestamp >= t
  while (lo < hi) {
    int64_t mid = (lo + hi) / 2;
    if (imu[mid].timestamp < t) lo = mid + 1; else hi = mid;
  }
  return lo;
}
}  // namespace

extern "C" int wco_update_surfel_poses(const wc_imu_state* imu, int64_t n_imu, wc_surfel* surfels, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    wc_surfel& s   = surfels[i];
    int64_t    idx = ImuLowerBound(imu, n_imu, s.timestamp);
    if (idx == 0 || idx == n_imu) return WC_EOUT_OF_SPAN;  // CHECK at :164
    double factor = (s.timestamp - imu[idx - 1].timestamp) / (imu[idx].timestamp - imu[idx - 1].timestamp);
    V3     pos    = V3(imu[idx - 1].pos) * (1 - factor) + V3(imu[idx].pos) * factor;
    Q4     rot    = Slerp(Q4::FromCoeffs(imu[idx - 1].rot), factor, Q4::FromCoeffs(imu[idx].rot));
    pos.store(s.pos);
    rot.storeCoeffs(s.rot);
    if (!s.is_in_body_frame) {
      s.is_in_body_frame = 1;
      Q4 rc              = rot.conjugate();
      V3 c               = rc * (V3(s.center) - pos);
      V3 nn              = rc * V3(s.norm);
      M3 cov             = (ToMatrix(rc) * M3::FromRowMajor(s.covariance)) * ToMatrix(rot);
      c.store(s.center), nn.store(s.norm), cov.store(s.covariance);
    }
  }
  return WC_OK;
}

extern "C" int wco_undistort_sweep(const wc_imu_state* imu, int64_t n_imu, const wc_point48* in, int64_t n,
                                   wc_point48* out) {
  for (int64_t i = 0; i < n; ++i) {
    int64_t idx = ImuLowerBound(imu, n_imu, in[i].time);
    if (!(idx >= 1 && idx < n_imu)) return WC_EOUT_OF_SPAN;
    double factor = (in[i].time - imu[idx - 1].timestamp) / (imu[idx].timestamp - imu[idx - 1].timestamp);
    V3     pos    = V3(imu[idx - 1].pos) * (1 - factor) + V3(imu[idx].pos) * factor;
    Q4     rot    = Slerp(Q4::FromCoeffs(imu[idx - 1].rot), factor, Q4::FromCoeffs(imu[idx].rot));
    V3     p      = rot * V3((double)in[i].x, (double)in[i].y, (double)in[i].z) + pos;
    out[i]        = in[i];
    out[i].x = (float)p.x, out[i].y = (float)p.y, out[i].z = (float)p.z;
  }
  return WC_OK;
}

// AddLidarScan's per-point loop, lidar_odometry.cc:489-496: extrinsic, time-order CHECK, range / blind-box filter.
// Returns the number of kept points, or a negative wc_status.
extern "C" int64_t wco_filter_points(const wc_sweep_filter* f, const wc_point48* in, int64_t n, wc_point48* out) {
  const Q4 q = Q4::FromCoeffs(f->ext_q);
  const V3 t(f->ext_t);
  int64_t  k = 0;
  for (int64_t i = 0; i < n; ++i) {
    wc_point48 pt = in[i];
    const V3   p  = q * V3((double)pt.x, (double)pt.y, (double)pt.z) + t;  // Rigid3d * point (rigid_transform.h), :490
    pt.x = (float)p.x, pt.y = (float)p.y, pt.z = (float)p.z;
    if (k > 0 && pt.time < out[k - 1].time) return -(int64_t)WC_EINVAL_TIME_ORDER;  // CHECK :491: vs. points_buff_.back()
    const float  n2 = pt.x * pt.x + pt.y * pt.y + pt.z * pt.z;                         // Vector3f::norm(), :492
    const double nr = (double)sqrtf(n2);
    if (nr < f->min_range || nr > f->max_range) continue;
    const double x = pt.x, y = pt.y, z = pt.z;  // AlignedBox<double,3>::contains
    if (x >= f->blind_box_min[0] && x <= f->blind_box_max[0] && y >= f->blind_box_min[1] && y <= f->blind_box_max[1] &&
        z >= f->blind_box_min[2] && z <= f->blind_box_max[2])
      continue;
    out[k++] = pt;
  }
  return k;
}

// ---------------------------------------------------------------------------------------------------
// knn_surfel_matcher.cc
// ---------------------------------------------------------------------------------------------------
namespace {

struct SurfelView {
  V3     center_w, norm_w;
  double timestamp;
};
SurfelView View(const wc_surfel& s) {
  Q4         q = Q4::FromCoeffs(s.rot);
  SurfelView v;
  v.center_w  = q * V3(s.center) + V3(s.pos);  // surfel.h:67-69
  v.norm_w    = q * V3(s.norm);                 // surfel.h:78-80
  v.timestamp = s.timestamp;
  return v;
}

// flann::L2_Simple<double>: sequential sum of squared differences
inline double L2Simple(const double* a, const double* b) {
  double r = 0;
  for (int i = 0; i < 6; ++i) {
    double d = a[i] - b[i];
    r += d * d;
  }
  return r;
}

struct Cand {
  double d;
  int    i;
  bool   operator<(const Cand& o) const { return d < o.d || (d == o.d && i < o.i); }
};

// bounded sorted top-k (ascending by (distance, index))
struct TopK {
  int               k;
  std::vector<Cand> v;
  explicit TopK(int k_) : k(k_) { v.reserve(k_ + 1); }
  double Worst() const { return (int)v.size() < k ? INFINITY : v.back().d; }
  void   Push(double d, int i) {
    Cand c{d, i};
    if ((int)v.size() == k && !(c < v.back())) return;
    auto it = std::upper_bound(v.begin(), v.end(), c);
    v.insert(it, c);
    if ((int)v.size() > k) v.pop_back();
  }
};

// Exact kd-tree (what flann::KDTreeSingleIndex with SearchParams(-1, 0) computes: the exact k nearest in
// squared L2; leaf size 15 as in knn_surfel_matcher.cc:71).  Tie order differs from FLANN's unspecified one:
// we order ties by index, in both search modes.
struct KdTree6 {
  struct Node {
    int    lo, hi;  // point range in idx_
    int    dim;     // -1 leaf
    double split;
    int    left, right;
    double bmin[6], bmax[6];
  };
  const double*     pts_;
  std::vector<int>  idx_;
  std::vector<Node> nodes_;
  void              Build(const double* pts, int n) {
    pts_ = pts;
    idx_.resize(n);
    for (int i = 0; i < n; ++i) idx_[i] = i;
    nodes_.clear();
    nodes_.reserve(2 * (n / 8 + 1));
    if (n > 0) BuildRec(0, n);
  }
  int BuildRec(int lo, int hi) {
    Node nd;
    nd.lo = lo, nd.hi = hi, nd.dim = -1, nd.left = nd.right = -1, nd.split = 0;
    for (int d = 0; d < 6; ++d) nd.bmin[d] = INFINITY, nd.bmax[d] = -INFINITY;
    for (int i = lo; i < hi; ++i)
      for (int d = 0; d < 6; ++d) {
        double v   = pts_[6 * (size_t)idx_[i] + d];
        nd.bmin[d] = std::min(nd.bmin[d], v), nd.bmax[d] = std::max(nd.bmax[d], v);
      }
    int id = (int)nodes_.size();
    nodes_.push_back(nd);
    if (hi - lo > 15) {
      int    best = 0;
      double span = -1;
      for (int d = 0; d < 6; ++d)
        if (nd.bmax[d] - nd.bmin[d] > span) span = nd.bmax[d] - nd.bmin[d], best = d;
      if (span > 0) {
        int mid = (lo + hi) / 2;
        std::nth_element(idx_.begin() + lo, idx_.begin() + mid, idx_.begin() + hi,
                         [&](int a, int b) { return pts_[6 * (size_t)a + best] < pts_[6 * (size_t)b + best]; });
        nodes_[id].dim   = best;
        nodes_[id].split = pts_[6 * (size_t)idx_[mid] + best];
        int l            = BuildRec(lo, mid);
        int r            = BuildRec(mid, hi);
        nodes_[id].left = l, nodes_[id].right = r;
      }
    }
    return id;
  }
  double BoxDist(const Node& nd, const double* q) const {
    double r = 0;
    for (int d = 0; d < 6; ++d) {
      double e = q[d] < nd.bmin[d] ? nd.bmin[d] - q[d] : (q[d] > nd.bmax[d] ? q[d] - nd.bmax[d] : 0.0);
      r += e * e;
    }
    return r;
  }
  void Search(int id, const double* q, TopK& top) const {
    const Node& nd = nodes_[id];
    // conservative pruning: a box whose lower bound (shrunk by a relative slack for the different
    // rounding of the bound) exceeds the current worst cannot hold a better candidate
    if (BoxDist(nd, q) * (1.0 - 1e-12) > top.Worst()) return;
    if (nd.dim < 0) {
      for (int i = nd.lo; i < nd.hi; ++i) top.Push(L2Simple(q, pts_ + 6 * (size_t)idx_[i]), idx_[i]);
      return;
    }
    if (q[nd.dim] < nd.split) {
      Search(nd.left, q, top), Search(nd.right, q, top);
    } else {
      Search(nd.right, q, top), Search(nd.left, q, top);
    }
  }
};

void ToVector(const wc_params* prm, const SurfelView& v, double* f) {
  // knn_surfel_matcher.cc:91-98
  V3 c = v.center_w / prm->center_dist_threshold;
  V3 m = v.norm_w / prm->angular_dist_threshold;
  f[0] = c.x, f[1] = c.y, f[2] = c.z, f[3] = m.x, f[4] = m.y, f[5] = m.z;
}

}  // namespace

extern "C" void wco_knn6(const double* query6, int64_t nq, const double* target6, int64_t nt, int k, int use_kdtree,
                         int32_t* out_idx, double* out_dist2) {
  KdTree6 tree;
  if (use_kdtree) tree.Build(target6, (int)nt);
  for (int64_t i = 0; i < nq; ++i) {
    TopK top(k);
    if (use_kdtree) {
      if (nt > 0) tree.Search(0, query6 + 6 * i, top);
    } else {
      for (int64_t j = 0; j < nt; ++j) top.Push(L2Simple(query6 + 6 * i, target6 + 6 * j), (int)j);
    }
    for (int j = 0; j < k; ++j) {
      out_idx[i * k + j]   = j < (int)top.v.size() ? top.v[j].i : -1;
      out_dist2[i * k + j] = j < (int)top.v.size() ? top.v[j].d : INFINITY;
    }
  }
}

extern "C" int64_t wco_match(const wc_params* prm, const wc_surfel* query, int64_t nq, const wc_surfel* target,
                             int64_t nt, int self_match, int use_kdtree, wc_corr_idx* out, uint8_t* first_is_target) {
  if (nt == 0) return 0;  // knn_surfel_matcher.cc:18-20
  const int           k = prm->knn_candidates;
  std::vector<double> tf(6 * (size_t)nt), qf(6);
  std::vector<SurfelView> tv(nt);
  for (int64_t j = 0; j < nt; ++j) tv[j] = View(target[j]), ToVector(prm, tv[j], &tf[6 * (size_t)j]);
  KdTree6 tree;
  if (use_kdtree) tree.Build(tf.data(), (int)nt);
  std::set<std::pair<int64_t, int64_t>> surfel_pairs;  // (query idx, target idx); same index space iff self_match
  int64_t                              n_out = 0;
  for (int64_t i = 0; i < nq; ++i) {
    SurfelView sv = View(query[i]);
    ToVector(prm, sv, qf.data());
    TopK top(k);
    if (use_kdtree) tree.Search(0, qf.data(), top);
    else
      for (int64_t j = 0; j < nt; ++j) top.Push(L2Simple(qf.data(), &tf[6 * (size_t)j]), (int)j);
    for (const Cand& c : top.v) {  // Q8: with < k targets the reference reads garbage; we stop at nt
      const SurfelView& nv = tv[c.i];
      if (std::fabs(nv.timestamp - sv.timestamp) < prm->time_diff_threshold) continue;
      // Surfel::AngularDistance, surfel.h:105-107 (acos un-clamped: NaN compares false, Q9)
      if (std::acos(dot(sv.norm_w, nv.norm_w)) > prm->angular_dist_threshold) continue;
      if (std::fabs(dot(sv.norm_w, sv.center_w - nv.center_w)) > prm->surfel_dist_threshold) continue;
      if (self_match) {
        if (surfel_pairs.count({i, c.i}) || surfel_pairs.count({c.i, i})) continue;
        surfel_pairs.insert({i, c.i});
      }
      bool query_first = sv.timestamp < nv.timestamp;
      if (self_match) {
        out[n_out].s1 = (int32_t)(query_first ? i : c.i);
        out[n_out].s2 = (int32_t)(query_first ? c.i : i);
      } else {
        // s1 = earlier surfel.  For the fixed-window matcher the target (fixed) surfel is the older one in
        // the reference's use; first_is_target records which array s1 indexes.
        out[n_out].s1 = (int32_t)(query_first ? i : c.i);
        out[n_out].s2 = (int32_t)(query_first ? c.i : i);
      }
      if (first_is_target) first_is_target[n_out] = query_first ? 0 : 1;
      ++n_out;
      break;
    }
  }
  return n_out;
}
