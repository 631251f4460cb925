/* TEST INFRASTRUCTURE ONLY — CPU oracle for the LOAM hot path (plain C99).
 * See loam_oracle.h for the scope statement and the parity-pinning status.
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off, no -march: IEEE fp64, no FMA,
 * matching the reference's baseline x86-64 build).
 */
#include "loam_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================== */
/*                               FEATURES                                   */
/* ======================================================================== */

/* common.h:81-86 : sqrt(x*x + y*y + z*z), left-associated */
static double point_range(const double* p) { return sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]); }

/* features-inl.h:53-87 */
int orc_curvature(const double* xyz, uint64_t n, const orc_lidar_params* lp, const orc_fe_params* fe, double* curv) {
  const uint64_t R = lp->scan_lines, P = lp->points_per_line, N = fe->neighbor_points;
  if (n != R * P) return 1; /* common.h:104-113 */
  for (uint64_t line = 0; line < R; line++) {
    for (uint64_t j = 0; j < P; j++) {
      const uint64_t idx = line * P + j;
      if (j < N || j >= P - N) { /* unsigned wrap of P - N kept (size_t arithmetic, features-inl.h:66-67) */
        curv[idx] = -1;
      } else {
        double d[3];
        for (int a = 0; a < 3; a++) d[a] = -(2.0 * (double)N) * xyz[3 * idx + a];
        for (uint64_t k = 1; k <= N; k++)
          for (int a = 0; a < 3; a++) d[a] = d[a] + xyz[3 * (idx - k) + a] + xyz[3 * (idx + k) + a];
        curv[idx] = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      }
    }
  }
  return 0;
}

/* features-inl.h:90-124 ; features.cpp:20-68 */
int orc_valid_mask(const double* xyz, uint64_t n, const orc_lidar_params* lp, const orc_fe_params* fe, uint8_t* mask) {
  const uint64_t R = lp->scan_lines, P = lp->points_per_line, N = fe->neighbor_points;
  if (n != R * P) return 1;
  memset(mask, 1, n);
  for (uint64_t line = 0; line < R; line++) {
    for (uint64_t j = 0; j < P; j++) {
      const uint64_t idx = line * P + j;
      /* CHECK 1 (features.cpp:20-27) */
      if (j < N || j >= P - N) {
        mask[idx] = 0;
        continue;
      }
      const double r = point_range(xyz + 3 * idx);
      const double rn = point_range(xyz + 3 * (idx + 1));
      const double rp = point_range(xyz + 3 * (idx - 1));
      /* CHECK 2 (features.cpp:30-41) */
      if (r < lp->min_range || r > lp->max_range) {
        mask[idx] = 0;
        for (uint64_t k = 1; k <= N; k++) {
          mask[idx + k] = 0;
          mask[idx - k] = 0;
        }
        continue;
      }
      /* CHECK 3 (features.cpp:44-54) */
      if (rn - r > fe->occlusion_thresh) {
        for (uint64_t k = 1; k <= N; k++) mask[idx + k] = 0;
        continue;
      } else if (r - rn > fe->occlusion_thresh) {
        for (uint64_t k = 0; k < N; k++) mask[idx - k] = 0;
        continue;
      }
      /* CHECK 4 (features.cpp:57-68) */
      const double diff_next = fabs(rp - r);
      const double diff_prev = fabs(rn - r);
      if (diff_next > fe->parallel_thresh * r && diff_prev > fe->parallel_thresh * r) mask[idx] = 0;
    }
  }
  return 0;
}

typedef struct {
  double c;
  uint64_t idx;
} curv_rec;

static int curv_cmp(const void* a, const void* b) {
  const curv_rec* x = (const curv_rec*)a;
  const curv_rec* y = (const curv_rec*)b;
  if (x->c < y->c) return -1;
  if (x->c > y->c) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx); /* documented tie-break: ascending index */
}

/* features-inl.h:11-50 (orchestration), :137-157 (edge walk), :160-180 (planar walk) */
int orc_extract(const double* xyz, uint64_t n, const orc_lidar_params* lp, const orc_fe_params* fe,
                uint32_t* edge_idx, uint64_t* n_edge, uint32_t* planar_idx, uint64_t* n_planar, uint64_t* n_ties) {
  const uint64_t R = lp->scan_lines, P = lp->points_per_line, N = fe->neighbor_points, S = fe->number_sectors;
  if (n != R * P) return 1;
  *n_edge = 0;
  *n_planar = 0;
  if (n_ties) *n_ties = 0;
  if (n == 0) return 0;
  if (S == 0) return 2; /* reference divides by zero (UB) */
  const uint64_t pps = P / S;
  double* curv = (double*)malloc(sizeof(double) * n);
  uint8_t* mask = (uint8_t*)malloc(n);
  curv_rec* rec = (curv_rec*)malloc(sizeof(curv_rec) * (P ? P : 1));
  orc_curvature(xyz, n, lp, fe, curv);
  orc_valid_mask(xyz, n, lp, fe, mask);
  for (uint64_t line = 0; line < R; line++) {
    for (uint64_t s = 0; s < S; s++) {
      const uint64_t start = line * P + s * pps;
      const uint64_t end = (s == S - 1) ? (line + 1) * P : start + pps;
      const uint64_t m = end - start;
      for (uint64_t i = 0; i < m; i++) {
        rec[i].c = curv[start + i];
        rec[i].idx = start + i;
      }
      qsort(rec, m, sizeof(curv_rec), curv_cmp); /* features-inl.h:38 */
      if (n_ties) {
        for (uint64_t i = 1; i < m; i++) {
          if (rec[i].c == rec[i - 1].c && mask[rec[i].idx] && mask[rec[i - 1].idx] &&
              (rec[i].c > fe->edge_feat_threshold || rec[i].c < fe->planar_feat_threshold))
            (*n_ties)++;
        }
      }
      /* edge walk: largest curvature first (features-inl.h:143-156) */
      uint64_t cnt = 0;
      for (uint64_t i = m; i > 0; i--) {
        const curv_rec c = rec[i - 1];
        if (mask[c.idx] && c.c > fe->edge_feat_threshold) {
          edge_idx[(*n_edge)++] = (uint32_t)c.idx;
          for (uint64_t k = 0; k < N; k++) {
            mask[c.idx + k] = 0;
            mask[c.idx - k] = 0;
          }
          cnt++;
        }
        if (cnt > fe->max_edge_feats_per_sector) break; /* emits up to max+1: reference behaviour */
      }
      /* planar walk: smallest curvature first (features-inl.h:166-179) */
      cnt = 0;
      for (uint64_t i = 0; i < m; i++) {
        const curv_rec c = rec[i];
        if (mask[c.idx] && c.c < fe->planar_feat_threshold) {
          planar_idx[(*n_planar)++] = (uint32_t)c.idx;
          for (uint64_t k = 0; k < N; k++) {
            mask[c.idx + k] = 0;
            mask[c.idx - k] = 0;
          }
          cnt++;
        }
        if (cnt > fe->max_planar_feats_per_sector) break;
      }
    }
  }
  free(curv);
  free(mask);
  free(rec);
  return 0;
}

/* ======================================================================== */
/*                         POSE / GEOMETRY                                  */
/* ======================================================================== */

static void cross3(const double* a, const double* b, double* o) {
  const double x = a[1] * b[2] - a[2] * b[1];
  const double y = a[2] * b[0] - a[0] * b[2];
  const double z = a[0] * b[1] - a[1] * b[0];
  o[0] = x;
  o[1] = y;
  o[2] = z;
}
static double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

/* Eigen quaternion * vector (QuaternionBase::_transformVector): uv = 2 (u x v); v + w uv + u x uv.
 * q = (x, y, z, w); no normalisation. */
static void quat_rotate(const double* q, const double* v, double* o) {
  double uv[3], uuv[3];
  cross3(q, v, uv);
  uv[0] += uv[0];
  uv[1] += uv[1];
  uv[2] += uv[2];
  cross3(q, uv, uuv);
  o[0] = v[0] + q[3] * uv[0] + uuv[0];
  o[1] = v[1] + q[3] * uv[1] + uuv[1];
  o[2] = v[2] + q[3] * uv[2] + uuv[2];
}

/* Eigen quaternion product a*b, (x,y,z,w) storage */
static void quat_mul(const double* a, const double* b, double* o) {
  const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  o[0] = x;
  o[1] = y;
  o[2] = z;
  o[3] = w;
}

/* geometry.cpp:21 */
void orc_pose_act(const double* pose, const double* pt, double* out) {
  double r[3];
  quat_rotate(pose, pt, r);
  out[0] = r[0] + pose[4];
  out[1] = r[1] + pose[5];
  out[2] = r[2] + pose[6];
}

/* De-warp (motion compensation) of an organised scan — an EXTENSION: the reference leaves it to its caller
 * (README.md:63), so this restatement is the definition the CUDA path is checked against, not a pinned behaviour.
 * Column c of every ring is measured at fraction s = c / P of the sweep; the sensor pose at that instant, relative
 * to the start of the sweep, is interp(Identity, start_T_end, s) with the rotation interpolated by normalised
 * linear interpolation of the quaternion (hemisphere of Identity) and the translation linearly.  out = T(s) * p,
 * i.e. every point expressed in the frame of the sweep start.  Only IEEE +,-,*,/ and sqrt: bit-reproducible. */
void orc_dewarp(const double* xyz, uint64_t n, uint64_t points_per_line, const double* start_T_end, double* out) {
  double q[4] = {start_T_end[0], start_T_end[1], start_T_end[2], start_T_end[3]};
  if (q[3] < 0.0) {
    q[0] = -q[0];
    q[1] = -q[1];
    q[2] = -q[2];
    q[3] = -q[3];
  }
  for (uint64_t i = 0; i < n; i++) {
    const double s = (double)(i % points_per_line) / (double)points_per_line;
    double pose[7];
    pose[0] = s * q[0];
    pose[1] = s * q[1];
    pose[2] = s * q[2];
    pose[3] = (1.0 - s) + s * q[3];
    const double nn = sqrt(((pose[0] * pose[0] + pose[1] * pose[1]) + pose[2] * pose[2]) + pose[3] * pose[3]);
    pose[0] = pose[0] / nn;
    pose[1] = pose[1] / nn;
    pose[2] = pose[2] / nn;
    pose[3] = pose[3] / nn;
    pose[4] = s * start_T_end[4];
    pose[5] = s * start_T_end[5];
    pose[6] = s * start_T_end[6];
    orc_pose_act(pose, xyz + 3 * i, out + 3 * i);
  }
}

/* geometry.cpp:16-18 */
void orc_pose_compose(const double* p1, const double* p2, double* out) {
  double q[4], r[3];
  quat_mul(p1, p2, q);
  quat_rotate(p1, p2 + 4, r);
  out[0] = q[0];
  out[1] = q[1];
  out[2] = q[2];
  out[3] = q[3];
  out[4] = p1[4] + r[0];
  out[5] = p1[5] + r[1];
  out[6] = p1[6] + r[2];
}

/* geometry.cpp:10-13 ; Eigen inverse = conjugate / squaredNorm */
void orc_pose_inverse(const double* p, double* out) {
  const double n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3];
  double qi[4] = {-p[0] / n2, -p[1] / n2, -p[2] / n2, p[3] / n2};
  double nt[3] = {-p[4], -p[5], -p[6]}, r[3];
  quat_rotate(qi, nt, r);
  out[0] = qi[0];
  out[1] = qi[1];
  out[2] = qi[2];
  out[3] = qi[3];
  out[4] = r[0];
  out[5] = r[1];
  out[6] = r[2];
}

/* Eigen angularDistance: d = q1 * conj(q2); 2 atan2(|d.vec|, |d.w|) */
double orc_quat_angular_distance(const double* q1, const double* q2) {
  double c[4] = {-q2[0], -q2[1], -q2[2], q2[3]}, d[4];
  quat_mul(q1, c, d);
  return 2.0 * atan2(norm3(d), fabs(d[3]));
}

/* geometry-inl.h:21-27 */
double orc_point_to_line(const double* p, const double* a, const double* b) {
  double d1[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
  double d2[3] = {p[0] - b[0], p[1] - b[1], p[2] - b[2]};
  double ab[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
  double c[3];
  cross3(d1, d2, c);
  return norm3(c) / norm3(ab);
}

/* geometry-inl.h:30-33 */
double orc_point_to_plane(const double* p, const double* n, double d) {
  return fabs(n[0] * p[0] + n[1] * p[1] + n[2] * p[2] - d);
}

/* Cyclic Jacobi eigen-decomposition of a symmetric 3x3 (stands in for Eigen's
 * SelfAdjointEigenSolver<Matrix3d>, geometry.cpp:49 — any backward-stable
 * solver agrees to ~1e-15; eigenvector sign is irrelevant to the line). */
static void jacobi_eig3(double A[3][3], double V[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; sweep++) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    const double tr = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
    if (off <= 1e-20 * tr) break;
    for (int p = 0; p < 2; p++) {
      for (int q = p + 1; q < 3; q++) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0);
        const double s = t * c;
        A[p][p] = A[p][p] - t * apq;
        A[q][q] = A[q][q] + t * apq;
        A[p][q] = 0.0;
        A[q][p] = 0.0;
        const int r = 3 - p - q;
        const double arp = A[r][p], arq = A[r][q];
        A[r][p] = c * arp - s * arq;
        A[p][r] = A[r][p];
        A[r][q] = s * arp + c * arq;
        A[q][r] = A[r][q];
        for (int k = 0; k < 3; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
    }
  }
}

/* geometry.cpp:42-59 */
double orc_fit_line(const double* pts, uint32_t K, double* a, double* b) {
  double center[3] = {0, 0, 0};
  for (uint32_t k = 0; k < K; k++)
    for (int i = 0; i < 3; i++) center[i] += pts[3 * k + i];
  for (int i = 0; i < 3; i++) center[i] /= (double)K;
  double S[3][3] = {{0}}, V[3][3];
  for (uint32_t k = 0; k < K; k++) {
    double c[3] = {pts[3 * k] - center[0], pts[3 * k + 1] - center[1], pts[3 * k + 2] - center[2]};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) S[i][j] += c[i] * c[j];
  }
  jacobi_eig3(S, V);
  int big = 0;
  if (S[1][1] > S[big][big]) big = 1;
  if (S[2][2] > S[big][big]) big = 2;
  for (int i = 0; i < 3; i++) {
    a[i] = center[i] + 0.1 * V[i][big];
    b[i] = center[i] - 0.1 * V[i][big];
  }
  /* geometry.cpp:55-56: the ratio is computed and discarded, so the condition number stays DBL_MAX */
  return DBL_MAX;
}

/* Least squares min |A x - b| for a K x 3 matrix (K <= 16) by column-pivoted
 * Householder QR, following Eigen's ColPivHouseholderQR (geometry.cpp:67). */
static void colpiv_qr_solve3(const double* pts, uint32_t K, double* x) {
  double A[16][3], c[16];
  int perm[3] = {0, 1, 2};
  if (K > 16) K = 16;
  for (uint32_t k = 0; k < K; k++) {
    A[k][0] = pts[3 * k];
    A[k][1] = pts[3 * k + 1];
    A[k][2] = pts[3 * k + 2];
    c[k] = 1.0;
  }
  const int size = K < 3 ? (int)K : 3;
  double maxnorm = 0.0;
  for (int j = 0; j < 3; j++) {
    double s = 0;
    for (uint32_t k = 0; k < K; k++) s += A[k][j] * A[k][j];
    s = sqrt(s);
    if (s > maxnorm) maxnorm = s;
  }
  /* Eigen ColPivHouseholderQR::computeInPlace: abs2(maxnorm * eps) / rows — ONE division by the row count */
  const double th = maxnorm * DBL_EPSILON;
  const double threshold_helper = th * th / (double)K;
  int nonzero_pivots = size;
  for (int k = 0; k < size; k++) {
    /* pivot: remaining column with the biggest norm (recomputed directly) */
    int big = k;
    double bigsq = -1.0;
    for (int j = k; j < 3; j++) {
      double s = 0;
      for (uint32_t r = k; r < K; r++) s += A[r][j] * A[r][j];
      if (s > bigsq) {
        bigsq = s;
        big = j;
      }
    }
    if (nonzero_pivots == size && bigsq < threshold_helper * (double)(K - k)) nonzero_pivots = k;
    if (big != k) {
      for (uint32_t r = 0; r < K; r++) {
        double t = A[r][k];
        A[r][k] = A[r][big];
        A[r][big] = t;
      }
      int t = perm[k];
      perm[k] = perm[big];
      perm[big] = t;
    }
    /* Householder (Eigen makeHouseholder) */
    double tail = 0;
    for (uint32_t r = k + 1; r < K; r++) tail += A[r][k] * A[r][k];
    const double c0 = A[k][k];
    double tau, beta;
    if (tail <= DBL_MIN) {
      tau = 0;
      beta = c0;
      for (uint32_t r = k + 1; r < K; r++) A[r][k] = 0;
    } else {
      beta = sqrt(c0 * c0 + tail);
      if (c0 >= 0) beta = -beta;
      for (uint32_t r = k + 1; r < K; r++) A[r][k] = A[r][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    A[k][k] = beta;
    /* apply H = I - tau v v^T (v = [1; essential]) to the trailing columns and the rhs */
    for (int j = k + 1; j < 3; j++) {
      double w = A[k][j];
      for (uint32_t r = k + 1; r < K; r++) w += A[r][k] * A[r][j];
      w *= tau;
      A[k][j] -= w;
      for (uint32_t r = k + 1; r < K; r++) A[r][j] -= w * A[r][k];
    }
    {
      double w = c[k];
      for (uint32_t r = k + 1; r < K; r++) w += A[r][k] * c[r];
      w *= tau;
      c[k] -= w;
      for (uint32_t r = k + 1; r < K; r++) c[r] -= w * A[r][k];
    }
  }
  double y[3] = {0, 0, 0};
  for (int i = nonzero_pivots - 1; i >= 0; i--) {
    double s = c[i];
    for (int j = i + 1; j < nonzero_pivots; j++) s -= A[i][j] * y[j];
    y[i] = s / A[i][i];
  }
  x[0] = x[1] = x[2] = 0;
  for (int i = 0; i < nonzero_pivots; i++) x[perm[i]] = y[i];
}

/* geometry.cpp:62-73 */
double orc_fit_plane(const double* pts, uint32_t K, double* normal, double* d) {
  double abc[3];
  colpiv_qr_solve3(pts, K, abc);
  const double nrm = norm3(abc);
  normal[0] = abc[0] / nrm;
  normal[1] = abc[1] / nrm;
  normal[2] = abc[2] / nrm;
  *d = 1.0 / nrm;
  double sum = 0;
  for (uint32_t k = 0; k < K; k++)
    sum += (pts[3 * k] * normal[0] + pts[3 * k + 1] * normal[1] + pts[3 * k + 2] * normal[2]) - *d;
  return sum / (double)K; /* signed mean: reference behaviour */
}

/* ======================================================================== */
/*                         KD-TREE (nanoflann-style)                        */
/* ======================================================================== */

typedef struct {
  int32_t left, right; /* -1,-1 => leaf */
  uint32_t lo, hi;     /* leaf range in ind[] */
  int divfeat;
  double divlow, divhigh;
} kdnode;

struct orc_kdtree {
  const double* pts;
  uint64_t n;
  uint32_t* ind;
  kdnode* nodes;
  size_t n_nodes, cap;
  double bb_lo[3], bb_hi[3];
};

#define KD_LEAF 20 /* registration-inl.h:21,23 */

typedef struct {
  uint32_t k, count;
  uint32_t* idx;
  double* d2;
} knn_set;

static double knn_worst(const knn_set* s) { return s->count < s->k ? DBL_MAX : s->d2[s->k - 1]; }

/* sorted insertion by (d2, idx) ascending — documented tie-break */
static void knn_add(knn_set* s, double d2, uint32_t idx) {
  if (s->count == s->k) {
    const double wd = s->d2[s->k - 1];
    if (!(d2 < wd || (d2 == wd && idx < s->idx[s->k - 1]))) return;
  }
  uint32_t i = s->count < s->k ? s->count : s->k - 1;
  while (i > 0 && (s->d2[i - 1] > d2 || (s->d2[i - 1] == d2 && s->idx[i - 1] > idx))) {
    s->d2[i] = s->d2[i - 1];
    s->idx[i] = s->idx[i - 1];
    i--;
  }
  s->d2[i] = d2;
  s->idx[i] = idx;
  if (s->count < s->k) s->count++;
}

/* nanoflann L2_Simple_Adaptor::evalMetric: diff = query - point; ((0 + d0^2) + d1^2) + d2^2 */
static double sqdist(const double* q, const double* p) {
  const double d0 = q[0] - p[0], d1 = q[1] - p[1], d2 = q[2] - p[2];
  return d0 * d0 + d1 * d1 + d2 * d2;
}

static int32_t kd_new_node(orc_kdtree* t) {
  if (t->n_nodes == t->cap) {
    t->cap = t->cap ? t->cap * 2 : 64;
    t->nodes = (kdnode*)realloc(t->nodes, t->cap * sizeof(kdnode));
  }
  return (int32_t)t->n_nodes++;
}

static int32_t kd_divide(orc_kdtree* t, uint32_t left, uint32_t right, double* lo, double* hi) {
  const int32_t me = kd_new_node(t);
  const uint32_t count = right - left;
  if (count <= KD_LEAF) {
    t->nodes[me].left = t->nodes[me].right = -1;
    t->nodes[me].lo = left;
    t->nodes[me].hi = right;
    for (int a = 0; a < 3; a++) {
      lo[a] = hi[a] = t->pts[3 * t->ind[left] + a];
    }
    for (uint32_t i = left + 1; i < right; i++)
      for (int a = 0; a < 3; a++) {
        const double v = t->pts[3 * t->ind[i] + a];
        if (v < lo[a]) lo[a] = v;
        if (v > hi[a]) hi[a] = v;
      }
    return me;
  }
  /* middleSplit_ */
  const double EPS = 0.00001;
  double max_span = hi[0] - lo[0];
  for (int a = 1; a < 3; a++)
    if (hi[a] - lo[a] > max_span) max_span = hi[a] - lo[a];
  double max_spread = -1;
  int cutfeat = 0;
  double mn = 0, mx = 0;
  for (int a = 0; a < 3; a++) {
    if (hi[a] - lo[a] > (1 - EPS) * max_span) {
      double emin = t->pts[3 * t->ind[left] + a], emax = emin;
      for (uint32_t i = left + 1; i < right; i++) {
        const double v = t->pts[3 * t->ind[i] + a];
        if (v < emin) emin = v;
        if (v > emax) emax = v;
      }
      if (emax - emin > max_spread) {
        cutfeat = a;
        max_spread = emax - emin;
        mn = emin;
        mx = emax;
      }
    }
  }
  const double split_val = (lo[cutfeat] + hi[cutfeat]) / 2;
  double cutval = split_val < mn ? mn : (split_val > mx ? mx : split_val);
  /* planeSplit: [left,lim1) < cutval ; [lim1,lim2) == cutval ; [lim2,right) > cutval */
  uint32_t* ind = t->ind + left;
  uint32_t l = 0, r = count - 1;
  for (;;) {
    while (l <= r && t->pts[3 * ind[l] + cutfeat] < cutval) l++;
    while (r && l <= r && t->pts[3 * ind[r] + cutfeat] >= cutval) r--;
    if (l > r || !r) break;
    uint32_t tmp = ind[l];
    ind[l] = ind[r];
    ind[r] = tmp;
    l++;
    r--;
  }
  const uint32_t lim1 = l;
  r = count - 1;
  for (;;) {
    while (l <= r && t->pts[3 * ind[l] + cutfeat] <= cutval) l++;
    while (r && l <= r && t->pts[3 * ind[r] + cutfeat] > cutval) r--;
    if (l > r || !r) break;
    uint32_t tmp = ind[l];
    ind[l] = ind[r];
    ind[r] = tmp;
    l++;
    r--;
  }
  const uint32_t lim2 = l;
  uint32_t split;
  if (lim1 > count / 2)
    split = lim1;
  else if (lim2 < count / 2)
    split = lim2;
  else
    split = count / 2;
  if (split == 0 || split == count) split = count / 2; /* guard for degenerate duplicates */

  double llo[3], lhi[3], rlo[3], rhi[3];
  memcpy(llo, lo, sizeof llo);
  memcpy(lhi, hi, sizeof lhi);
  memcpy(rlo, lo, sizeof rlo);
  memcpy(rhi, hi, sizeof rhi);
  lhi[cutfeat] = cutval;
  rlo[cutfeat] = cutval;
  const int32_t lc = kd_divide(t, left, left + split, llo, lhi);
  const int32_t rc = kd_divide(t, left + split, right, rlo, rhi);
  t->nodes[me].left = lc;
  t->nodes[me].right = rc;
  t->nodes[me].divfeat = cutfeat;
  t->nodes[me].divlow = lhi[cutfeat];
  t->nodes[me].divhigh = rlo[cutfeat];
  for (int a = 0; a < 3; a++) {
    lo[a] = llo[a] < rlo[a] ? llo[a] : rlo[a];
    hi[a] = lhi[a] > rhi[a] ? lhi[a] : rhi[a];
  }
  return me;
}

orc_kdtree* orc_kdtree_build(const double* pts, uint64_t n) {
  orc_kdtree* t = (orc_kdtree*)calloc(1, sizeof(orc_kdtree));
  t->pts = pts;
  t->n = n;
  if (n == 0) return t;
  t->ind = (uint32_t*)malloc(sizeof(uint32_t) * n);
  for (uint64_t i = 0; i < n; i++) t->ind[i] = (uint32_t)i;
  for (int a = 0; a < 3; a++) t->bb_lo[a] = t->bb_hi[a] = pts[a];
  for (uint64_t i = 1; i < n; i++)
    for (int a = 0; a < 3; a++) {
      const double v = pts[3 * i + a];
      if (v < t->bb_lo[a]) t->bb_lo[a] = v;
      if (v > t->bb_hi[a]) t->bb_hi[a] = v;
    }
  double lo[3], hi[3];
  memcpy(lo, t->bb_lo, sizeof lo);
  memcpy(hi, t->bb_hi, sizeof hi);
  kd_divide(t, 0, (uint32_t)n, lo, hi);
  return t;
}

void orc_kdtree_free(orc_kdtree* t) {
  if (!t) return;
  free(t->ind);
  free(t->nodes);
  free(t);
}

static void kd_search(const orc_kdtree* t, int32_t ni, const double* q, knn_set* rs, double mindistsq, double* dists) {
  const kdnode* nd = &t->nodes[ni];
  if (nd->left < 0) {
    for (uint32_t i = nd->lo; i < nd->hi; i++) {
      const uint32_t id = t->ind[i];
      const double d = sqdist(q, t->pts + 3 * id);
      if (d <= knn_worst(rs)) knn_add(rs, d, id);
    }
    return;
  }
  const int f = nd->divfeat;
  const double val = q[f];
  const double diff1 = val - nd->divlow, diff2 = val - nd->divhigh;
  int32_t best, other;
  double cut;
  if (diff1 + diff2 < 0) {
    best = nd->left;
    other = nd->right;
    cut = diff2 * diff2;
  } else {
    best = nd->right;
    other = nd->left;
    cut = diff1 * diff1;
  }
  kd_search(t, best, q, rs, mindistsq, dists);
  const double dst = dists[f];
  mindistsq = mindistsq + cut - dst;
  dists[f] = cut;
  if (mindistsq <= knn_worst(rs)) kd_search(t, other, q, rs, mindistsq, dists);
  dists[f] = dst;
}

/* kdtree.cpp:10-28 : k nearest (unbounded), then keep those with sqrt(d2) < max_dist (strict) */
static uint32_t radius_filter(knn_set* rs, double max_dist) {
  uint32_t m = 0;
  for (uint32_t i = 0; i < rs->count; i++) {
    if (max_dist <= 0 || sqrt(rs->d2[i]) < max_dist) {
      rs->idx[m] = rs->idx[i];
      rs->d2[m] = rs->d2[i];
      m++;
    }
  }
  return m;
}

uint32_t orc_kdtree_knn(const orc_kdtree* t, const double* q, uint32_t k, double max_dist, uint32_t* idx_out,
                        double* d2_out) {
  double d2loc[64];
  knn_set rs = {k, 0, idx_out, d2_out ? d2_out : d2loc};
  if (k == 0 || k > 64 || t->n == 0) return 0;
  double dists[3] = {0, 0, 0}, mind = 0;
  for (int a = 0; a < 3; a++) {
    if (q[a] < t->bb_lo[a]) dists[a] = (q[a] - t->bb_lo[a]) * (q[a] - t->bb_lo[a]);
    if (q[a] > t->bb_hi[a]) dists[a] = (q[a] - t->bb_hi[a]) * (q[a] - t->bb_hi[a]);
    mind += dists[a];
  }
  kd_search(t, 0, q, &rs, mind, dists);
  return radius_filter(&rs, max_dist);
}

uint32_t orc_knn_brute(const double* pts, uint64_t n, const double* q, uint32_t k, double max_dist, uint32_t* idx_out,
                       double* d2_out) {
  double d2loc[64];
  knn_set rs = {k, 0, idx_out, d2_out ? d2_out : d2loc};
  if (k == 0 || k > 64) return 0;
  for (uint64_t i = 0; i < n; i++) {
    const double d = sqdist(q, pts + 3 * i);
    if (d <= knn_worst(&rs)) knn_add(&rs, d, (uint32_t)i);
  }
  return radius_filter(&rs, max_dist);
}

/* ======================================================================== */
/*           REGISTRATION: residuals, Ceres-2.2.0-style LM, ICF loop        */
/* ======================================================================== */

typedef struct {
  int is_plane;
  double p[3];  /* source point already transformed by the current estimate (registration.cpp:52-57) */
  double a[3];  /* edge: line point a ; plane: normal */
  double b[3];  /* edge: line point b ; plane: b[0] = d */
} resblock;

/* Value + 1x7 ambient Jacobian of one residual at x = (qx qy qz qw tx ty tz).
 * registration-inl.h:92-117 ; geometry-inl.h:21-33 ; derivative == Ceres autodiff of the same expression. */
static double residual_eval(const resblock* rb, const double* x, double* J7) {
  const double* u = x;
  const double w = x[3];
  double uv[3], uuv[3], pt[3];
  cross3(u, rb->p, uv);
  uv[0] += uv[0];
  uv[1] += uv[1];
  uv[2] += uv[2];
  cross3(u, uv, uuv);
  for (int i = 0; i < 3; i++) pt[i] = rb->p[i] + w * uv[i] + uuv[i] + x[4 + i];
  double r, g[3];
  if (!rb->is_plane) {
    double d1[3], d2[3], ab[3], c[3];
    for (int i = 0; i < 3; i++) {
      d1[i] = pt[i] - rb->a[i];
      d2[i] = pt[i] - rb->b[i];
      ab[i] = rb->a[i] - rb->b[i];
    }
    cross3(d1, d2, c);
    const double num = norm3(c), den = norm3(ab);
    r = num / den;
    if (J7) {
      if (num > 0) {
        double ch[3] = {c[0] / num, c[1] / num, c[2] / num};
        cross3(ab, ch, g);
        g[0] /= den;
        g[1] /= den;
        g[2] /= den;
      } else { /* Jet sqrt at 0 is NaN in Ceres (measure-zero input); use 0 */
        g[0] = g[1] = g[2] = 0;
      }
    }
  } else {
    const double s = rb->a[0] * pt[0] + rb->a[1] * pt[1] + rb->a[2] * pt[2] - rb->b[0];
    r = fabs(s);
    if (J7) {
      const double sg = copysign(1.0, s); /* ceres Jet abs */
      g[0] = sg * rb->a[0];
      g[1] = sg * rb->a[1];
      g[2] = sg * rb->a[2];
    }
  }
  if (J7) {
    /* d pt / d u_i = w * 2(e_i x p) + e_i x uv + u x 2(e_i x p) ; d pt / d w = uv ; d pt / d t = I */
    for (int i = 0; i < 3; i++) {
      double e[3] = {0, 0, 0}, Ai[3], t1[3], t2[3];
      e[i] = 1.0;
      cross3(e, rb->p, Ai);
      Ai[0] += Ai[0];
      Ai[1] += Ai[1];
      Ai[2] += Ai[2];
      cross3(e, uv, t1);
      cross3(u, Ai, t2);
      J7[i] = g[0] * (w * Ai[0] + t1[0] + t2[0]) + g[1] * (w * Ai[1] + t1[1] + t2[1]) +
              g[2] * (w * Ai[2] + t1[2] + t2[2]);
    }
    J7[3] = g[0] * uv[0] + g[1] * uv[1] + g[2] * uv[2];
    J7[4] = g[0];
    J7[5] = g[1];
    J7[6] = g[2];
  }
  return r;
}

/* ceres::QuaternionManifold (w-first convention) applied literally to Eigen's
 * (x,y,z,w) memory: slot 0 plays "w".  SURVEY §8a-notes. */
static void manifold_plus(const double* x, const double* delta, double* out) {
  const double nd = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  if (nd == 0.0) {
    for (int i = 0; i < 4; i++) out[i] = x[i];
  } else {
    const double sbd = sin(nd) / nd;
    const double q0 = cos(nd), q1 = sbd * delta[0], q2 = sbd * delta[1], q3 = sbd * delta[2];
    out[0] = q0 * x[0] - q1 * x[1] - q2 * x[2] - q3 * x[3];
    out[1] = q0 * x[1] + q1 * x[0] + q2 * x[3] - q3 * x[2];
    out[2] = q0 * x[2] - q1 * x[3] + q2 * x[0] + q3 * x[1];
    out[3] = q0 * x[3] + q1 * x[2] - q2 * x[1] + q3 * x[0];
  }
  out[4] = x[4] + delta[3];
  out[5] = x[5] + delta[4];
  out[6] = x[6] + delta[5];
}

/* 4x3 PlusJacobian of the (mis-ordered) quaternion manifold; rows = memory slots */
static void manifold_plus_jacobian(const double* m, double PJ[4][3]) {
  PJ[0][0] = -m[1]; PJ[0][1] = -m[2]; PJ[0][2] = -m[3];
  PJ[1][0] =  m[0]; PJ[1][1] =  m[3]; PJ[1][2] = -m[2];
  PJ[2][0] = -m[3]; PJ[2][1] =  m[0]; PJ[2][2] =  m[1];
  PJ[3][0] =  m[2]; PJ[3][1] = -m[1]; PJ[3][2] =  m[0];
}

/* ceres::HuberLoss(1.0) + Corrector (rho'' <= 0 branch => scale by sqrt(rho')) */
static void huber(double s, double* rho0, double* sqrt_rho1) {
  if (s > 1.0) {
    const double r = sqrt(s);
    double rho1 = 1.0 / r;
    if (rho1 < DBL_MIN) rho1 = DBL_MIN;
    *rho0 = 2.0 * r - 1.0;
    *sqrt_rho1 = sqrt(rho1);
  } else {
    *rho0 = s;
    *sqrt_rho1 = 1.0;
  }
}

/* Evaluate the whole problem at x.  If J != NULL: fills corrected residuals r[M],
 * corrected tangent Jacobian J[M][6] (unscaled) and gradient g[6] = J^T r. */
static double problem_eval(const resblock* rb, size_t M, const double* x, double* r, double* J, double* g) {
  double cost = 0;
  double PJ[4][3];
  if (J) {
    manifold_plus_jacobian(x, PJ);
    for (int j = 0; j < 6; j++) g[j] = 0;
  }
  for (size_t i = 0; i < M; i++) {
    double J7[7];
    const double res = residual_eval(&rb[i], x, J ? J7 : NULL);
    double rho0, sr1;
    huber(res * res, &rho0, &sr1);
    cost += 0.5 * rho0;
    if (J) {
      double* Ji = J + 6 * i;
      for (int j = 0; j < 3; j++) {
        double acc = 0;
        for (int k = 0; k < 4; k++) acc += J7[k] * PJ[k][j];
        Ji[j] = acc * sr1;
      }
      for (int j = 0; j < 3; j++) Ji[3 + j] = J7[4 + j] * sr1;
      r[i] = res * sr1;
      for (int j = 0; j < 6; j++) g[j] += Ji[j] * r[i];
    }
  }
  return cost;
}

/* min | [Js ; diag(D)] y - [r ; 0] |  by unpivoted Householder QR (Ceres DENSE_QR -> Eigen HouseholderQR). */
static int dense_qr_solve(const double* Js, const double* r, size_t M, const double* D, double* y, double* work) {
  const size_t rows = M + 6;
  double* A = work;            /* rows x 6, row-major */
  double* c = work + rows * 6; /* rows */
  memcpy(A, Js, sizeof(double) * M * 6);
  memset(A + M * 6, 0, sizeof(double) * 36);
  for (int j = 0; j < 6; j++) A[(M + j) * 6 + j] = D[j];
  memcpy(c, r, sizeof(double) * M);
  for (int j = 0; j < 6; j++) c[M + j] = 0;
  for (int k = 0; k < 6; k++) {
    double tail = 0;
    for (size_t i = k + 1; i < rows; i++) tail += A[i * 6 + k] * A[i * 6 + k];
    const double c0 = A[k * 6 + k];
    double tau, beta;
    if (tail <= DBL_MIN) {
      tau = 0;
      beta = c0;
      for (size_t i = k + 1; i < rows; i++) A[i * 6 + k] = 0;
    } else {
      beta = sqrt(c0 * c0 + tail);
      if (c0 >= 0) beta = -beta;
      const double inv = c0 - beta;
      for (size_t i = k + 1; i < rows; i++) A[i * 6 + k] /= inv;
      tau = (beta - c0) / beta;
    }
    A[k * 6 + k] = beta;
    for (int j = k + 1; j < 6; j++) {
      double w = A[k * 6 + j];
      for (size_t i = k + 1; i < rows; i++) w += A[i * 6 + k] * A[i * 6 + j];
      w *= tau;
      A[k * 6 + j] -= w;
      for (size_t i = k + 1; i < rows; i++) A[i * 6 + j] -= w * A[i * 6 + k];
    }
    double w = c[k];
    for (size_t i = k + 1; i < rows; i++) w += A[i * 6 + k] * c[i];
    w *= tau;
    c[k] -= w;
    for (size_t i = k + 1; i < rows; i++) c[i] -= w * A[i * 6 + k];
  }
  for (int i = 5; i >= 0; i--) {
    double s = c[i];
    for (int j = i + 1; j < 6; j++) s -= A[i * 6 + j] * y[j];
    y[i] = s / A[i * 6 + i];
  }
  for (int i = 0; i < 6; i++)
    if (!isfinite(y[i])) return 1;
  return 0;
}

static double norm7(const double* x) {
  double s = 0;
  for (int i = 0; i < 7; i++) s += x[i] * x[i];
  return sqrt(s);
}

/* Restatement of ceres::Solve with TRUST_REGION / LEVENBERG_MARQUARDT / DENSE_QR,
 * max_num_iterations = 4, all other options at their Ceres 2.2.0 defaults
 * (registration-inl.h:51-56).  x (7) is updated in place. */
static void lm_solve(const resblock* rb, size_t M, double* x_user, int armed_flag, uint32_t* n_iters_out,
                     double* cost_out) {
  const int max_num_iterations = 4;
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  const double min_relative_decrease = 1e-3, min_radius = 1e-32, max_radius = 1e16;
  const double min_diag = 1e-6, max_diag = 1e32;
  double radius = 1e4, decrease_factor = 2.0;
  int reuse_diagonal = 0;

  double* r = (double*)malloc(sizeof(double) * (M + 1));
  double* J = (double*)malloc(sizeof(double) * (M + 1) * 6);
  double* work = (double*)malloc(sizeof(double) * (M + 6) * 7);
  double x[7], g[6], scale[6], diag[6], cand[7];
  memcpy(x, x_user, sizeof x);
  double x_norm = norm7(x);

  /* IterationZero */
  double x_cost = problem_eval(rb, M, x, r, J, g);
  if (cost_out) cost_out[0] = x_cost;
  for (int j = 0; j < 6; j++) {
    double s = 0;
    for (size_t i = 0; i < M; i++) s += J[6 * i + j] * J[6 * i + j];
    scale[j] = 1.0 / (1.0 + sqrt(s));
  }
  for (size_t i = 0; i < M; i++)
    for (int j = 0; j < 6; j++) J[6 * i + j] *= scale[j];
  double ng[6], proj[7];
  for (int j = 0; j < 6; j++) ng[j] = -g[j];
  manifold_plus(x, ng, proj);
  double gmax = 0;
  for (int i = 0; i < 7; i++) gmax = fmax(gmax, fabs(x[i] - proj[i]));

  int iteration = 0, step_successful = 1, armed = 0;
  for (;;) {
    /* FinalizeIterationAndCheckIfMinimizerCanContinue */
    if (iteration >= max_num_iterations) break;
    if (step_successful && gmax <= gradient_tolerance) break;
    if (radius <= min_radius) break;
    iteration++;
    /* LevenbergMarquardtStrategy::ComputeStep */
    if (!reuse_diagonal) {
      for (int j = 0; j < 6; j++) {
        double s = 0;
        for (size_t i = 0; i < M; i++) s += J[6 * i + j] * J[6 * i + j];
        diag[j] = fmin(fmax(s, min_diag), max_diag);
      }
    }
    double D[6], y[6], step[6];
    for (int j = 0; j < 6; j++) D[j] = sqrt(diag[j] / radius);
    const int fail = dense_qr_solve(J, r, M, D, y, work);
    reuse_diagonal = 1;
    int valid = 0;
    double model_cost_change = 0;
    if (!fail) {
      for (int j = 0; j < 6; j++) step[j] = -y[j];
      /* model_cost_change = -(J step)^T (r + J step / 2) */
      for (size_t i = 0; i < M; i++) {
        double m = 0;
        for (int j = 0; j < 6; j++) m += J[6 * i + j] * step[j];
        model_cost_change += m * (r[i] + m / 2.0);
      }
      model_cost_change = -model_cost_change;
      valid = model_cost_change > 0.0;
    }
    if (!valid) { /* HandleInvalidStep */
      radius *= 0.5;
      reuse_diagonal = 1;
      step_successful = 0;
      continue;
    }
    double delta[6];
    for (int j = 0; j < 6; j++) delta[j] = step[j] * scale[j];
    manifold_plus(x, delta, cand);
    const double cand_cost = problem_eval(rb, M, cand, NULL, NULL, NULL);
    if (armed || !armed_flag) {
      double dn = 0;
      for (int i = 0; i < 7; i++) dn += (x[i] - cand[i]) * (x[i] - cand[i]);
      if (sqrt(dn) <= parameter_tolerance * (x_norm + parameter_tolerance)) break;
      if (fabs(x_cost - cand_cost) <= function_tolerance * x_cost) break;
    }
    const double rel = (x_cost - cand_cost) / model_cost_change;
    if (rel > min_relative_decrease) { /* HandleSuccessfulStep */
      memcpy(x, cand, sizeof x);
      x_norm = norm7(x);
      x_cost = problem_eval(rb, M, x, r, J, g);
      for (size_t i = 0; i < M; i++)
        for (int j = 0; j < 6; j++) J[6 * i + j] *= scale[j];
      for (int j = 0; j < 6; j++) ng[j] = -g[j];
      manifold_plus(x, ng, proj);
      gmax = 0;
      for (int i = 0; i < 7; i++) gmax = fmax(gmax, fabs(x[i] - proj[i]));
      step_successful = 1;
      const double q = 2.0 * rel - 1.0;
      radius = radius / fmax(1.0 / 3.0, 1.0 - q * q * q);
      radius = fmin(max_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = 0;
      armed = 1;
    } else {
      step_successful = 0;
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = 1;
    }
  }
  memcpy(x_user, x, sizeof x); /* x always holds the last accepted point (monotonic steps) */
  if (n_iters_out) *n_iters_out = (uint32_t)iteration;
  if (cost_out) cost_out[1] = x_cost;
  free(r);
  free(J);
  free(work);
}

/* ------------------------------------------------------------------------ test hooks
 * (tests/test_oracle_jacobians.py: the analytic derivatives and the LM step sequence are checked against torch
 * float64 autograd of the literal functor expressions and against an independent QR-based model of the solver) */
static resblock* blocks_from_flat(const int32_t* is_plane, const double* P, const double* A, const double* B, size_t M) {
  resblock* rb = (resblock*)malloc(sizeof(resblock) * (M + 1));
  for (size_t i = 0; i < M; i++) {
    rb[i].is_plane = is_plane[i];
    memcpy(rb[i].p, P + 3 * i, 3 * sizeof(double));
    memcpy(rb[i].a, A + 3 * i, 3 * sizeof(double));
    memcpy(rb[i].b, B + 3 * i, 3 * sizeof(double));
  }
  return rb;
}

/* one residual: value and 1x7 ambient Jacobian (x y z w tx ty tz); B = line point b, or (d, -, -) for a plane */
double orc_residual_eval(int32_t is_plane, const double* p, const double* a, const double* b, const double* x,
                         double* J7) {
  resblock rb;
  rb.is_plane = is_plane;
  memcpy(rb.p, p, sizeof rb.p);
  memcpy(rb.a, a, sizeof rb.a);
  memcpy(rb.b, b, sizeof rb.b);
  return residual_eval(&rb, x, J7);
}

/* whole problem at x: cost; optional corrected residuals r[M], corrected tangent Jacobian J[M][6], gradient g[6] */
double orc_problem_eval(const int32_t* is_plane, const double* P, const double* A, const double* B, uint64_t M,
                        const double* x, double* r, double* J, double* g) {
  resblock* rb = blocks_from_flat(is_plane, P, A, B, (size_t)M);
  double* rr = r ? r : (J ? (double*)malloc(sizeof(double) * (M + 1)) : NULL);
  double gg[6];
  const double c = problem_eval(rb, (size_t)M, x, J ? rr : NULL, J, J ? (g ? g : gg) : NULL);
  if (rr && rr != r) free(rr);
  free(rb);
  return c;
}

void orc_manifold_plus(const double* x, const double* delta, double* out) { manifold_plus(x, delta, out); }

/* the restated ceres::Solve on explicit residual blocks; x (7) in/out; returns LM iterations recorded */
uint32_t orc_lm_solve(const int32_t* is_plane, const double* P, const double* A, const double* B, uint64_t M,
                      double* x, int armed_flag, double* cost2) {
  resblock* rb = blocks_from_flat(is_plane, P, A, B, (size_t)M);
  uint32_t it = 0;
  lm_solve(rb, (size_t)M, x, armed_flag, &it, cost2);
  free(rb);
  return it;
}

typedef uint32_t (*knn_fn)(const void* ctx, const double* pts, uint64_t n, const double* q, uint32_t k, double md,
                           uint32_t* idx);
static uint32_t knn_tree(const void* ctx, const double* pts, uint64_t n, const double* q, uint32_t k, double md,
                         uint32_t* idx) {
  (void)pts;
  (void)n;
  return orc_kdtree_knn((const orc_kdtree*)ctx, q, k, md, idx, NULL);
}
static uint32_t knn_bf(const void* ctx, const double* pts, uint64_t n, const double* q, uint32_t k, double md,
                       uint32_t* idx) {
  (void)ctx;
  return orc_knn_brute(pts, n, q, k, md, idx, NULL);
}

/* registration-inl.h:11-78 */
int orc_register(const double* src_edge, uint64_t n_se, const double* src_planar, uint64_t n_sp,
                 const double* tgt_edge, uint64_t n_te, const double* tgt_planar, uint64_t n_tp,
                 const double* init_pose, const orc_reg_params* rp, double* out_pose, orc_detail* detail,
                 int use_kdtree, int armed_flag) {
  if (rp->num_edge_neighbors > 64 || rp->num_plane_neighbors > 16) return 3;
  orc_kdtree* te = use_kdtree ? orc_kdtree_build(tgt_edge, n_te) : NULL;
  orc_kdtree* tp = use_kdtree ? orc_kdtree_build(tgt_planar, n_tp) : NULL;
  knn_fn knn = use_kdtree ? knn_tree : knn_bf;
  resblock* rb = (resblock*)malloc(sizeof(resblock) * (n_se + n_sp + 1));
  double est[7];
  memcpy(est, init_pose, sizeof est);
  int termination = 1; /* MAX_ITER */
  uint32_t n_info = 0;
  for (uint64_t iter = 0; iter < rp->max_iterations; iter++) {
    size_t M = 0;
    uint32_t ne = 0, np = 0;
    uint32_t* ea = detail && detail->edge_assoc && iter < detail->max_iters_cap
                       ? detail->edge_assoc + (size_t)iter * detail->n_src_edge * 2
                       : NULL;
    uint32_t* pa = detail && detail->plane_assoc && iter < detail->max_iters_cap
                       ? detail->plane_assoc + (size_t)iter * detail->n_src_planar * 2
                       : NULL;
    /* associateEdges (registration.cpp:23-62) */
    for (uint64_t i = 0; i < n_se; i++) {
      double pt[3], nb[64 * 3];
      uint32_t idx[64];
      orc_pose_act(est, src_edge + 3 * i, pt);
      const uint32_t m = knn(te, tgt_edge, n_te, pt, (uint32_t)rp->num_edge_neighbors, rp->max_edge_neighbor_dist, idx);
      if (m < rp->min_line_fit_points) continue;
      for (uint32_t k = 0; k < m; k++) memcpy(nb + 3 * k, tgt_edge + 3 * (size_t)idx[k], 3 * sizeof(double));
      resblock* b = &rb[M];
      const double cond = orc_fit_line(nb, m, b->a, b->b);
      if (cond < rp->min_line_condition_number) continue; /* never fires (reference bug kept) */
      b->is_plane = 0;
      memcpy(b->p, pt, sizeof pt);
      M++;
      if (ea) {
        ea[2 * ne] = (uint32_t)i;
        ea[2 * ne + 1] = idx[0];
      }
      ne++;
    }
    /* associatePlanes (registration.cpp:65-103) */
    for (uint64_t i = 0; i < n_sp; i++) {
      double pt[3], nb[16 * 3];
      uint32_t idx[16];
      orc_pose_act(est, src_planar + 3 * i, pt);
      const uint32_t m =
          knn(tp, tgt_planar, n_tp, pt, (uint32_t)rp->num_plane_neighbors, rp->max_plane_neighbor_dist, idx);
      if (m < rp->min_plane_fit_points) continue;
      for (uint32_t k = 0; k < m; k++) memcpy(nb + 3 * k, tgt_planar + 3 * (size_t)idx[k], 3 * sizeof(double));
      resblock* b = &rb[M];
      double d;
      const double avg = orc_fit_plane(nb, m, b->a, &d);
      if (avg > rp->max_avg_point_plane_dist) continue;
      b->is_plane = 1;
      b->b[0] = d;
      memcpy(b->p, pt, sizeof pt);
      M++;
      if (pa) {
        pa[2 * np] = (uint32_t)i;
        pa[2 * np + 1] = idx[0];
      }
      np++;
    }
    if ((uint64_t)ne + np < rp->min_associations) { /* registration-inl.h:45-48 */
      termination = 2;
      break;
    }
    double upd[7] = {0, 0, 0, 1, 0, 0, 0};
    uint32_t lm_it = 0;
    double lm_c[2] = {0, 0};
    if (M > 0) lm_solve(rb, M, upd, armed_flag, &lm_it, lm_c);
    if (detail && iter < detail->max_iters_cap) { /* registration-inl.h:59-61 */
      if (detail->iter_est) memcpy(detail->iter_est + 7 * iter, est, sizeof est);
      if (detail->iter_update) memcpy(detail->iter_update + 7 * iter, upd, sizeof upd);
      if (detail->n_edge_assoc) detail->n_edge_assoc[iter] = ne;
      if (detail->n_plane_assoc) detail->n_plane_assoc[iter] = np;
      if (detail->lm_iters) detail->lm_iters[iter] = lm_it;
      if (detail->lm_cost) memcpy(detail->lm_cost + 2 * iter, lm_c, sizeof lm_c);
    }
    n_info++;
    double nxt[7];
    orc_pose_compose(upd, est, nxt); /* left composition, registration-inl.h:65 */
    memcpy(est, nxt, sizeof est);
    const double ident[4] = {0, 0, 0, 1};
    const double ang = orc_quat_angular_distance(upd, ident);
    const double pos = norm3(upd + 4);
    if (ang < rp->rotation_convergence_thresh && pos < rp->position_convergence_thresh) {
      termination = 0;
      break;
    }
  }
  if (detail) {
    detail->termination = termination;
    detail->n_iters = n_info;
  }
  memcpy(out_pose, est, sizeof est);
  orc_kdtree_free(te);
  orc_kdtree_free(tp);
  free(rb);
  return 0;
}
