// Sim(3) pose-graph Levenberg-Marquardt step on the GPU.
//
// Replaces, for one LM iteration of OnlineSLAM.pose_graph_optimize (vista_slam/slam.py:108-140):
//   * PoseGraphOpt.forward (vista_slam/pose_graph.py:100-149): r_e = Log(T_e X_i^-1 X_j) in R^7 for every edge with at
//     least one optimised endpoint (get_related_edge_idxs, :150-154), optimised / fixed node split (:73-98);
//   * the autograd Jacobian PyPose builds for it (left perturbation X <- Exp(delta) X, LieTensor.add_):
//     dr/d delta_j = J_l^-1(r) Ad_{T_e X_i^-1},  dr/d delta_i = -dr/d delta_j;
//   * pp.optim.LM.step: A = J^T W J (W = diag(conf_e), 7 per edge), diagonal clamped to [min, max] and scaled by
//     (1 + damping), dense Cholesky solve of A d = -J^T W r (the reference's solver is a dense torch Cholesky),
//     X <- Exp(d) X, loss = r^T W r before and after.
// Everything is fp64 inside (node / measurement tensors are fp32 like the reference's), deterministic (no atomics: every
// output element has one owner thread and a fixed summation order), and there is no host synchronisation: the caller reads
// the two losses to accept / reject the step, exactly where pp.optim.LM does.
//
// Kernels: pg_edge_kernel (one thread per edge: group algebra in closed form, J_l via phi_1(ad) by scaling-and-squaring,
// 7x7 Gauss-Jordan), pg_assemble_kernel (one CTA per optimised node: its block row of A and of g), pg_damp_kernel,
// blocked right-looking Cholesky (diag / trsm / syrk kernels, 32-wide panels), pg_solve_kernel (forward + backward
// substitution, one CTA), pg_update_kernel, pg_loss_kernel.  The problem is small (<= a few thousand edges, A of order
// 7 * optimised nodes): this is latency-bound work; the design goal is zero host round trips and determinism.
#include <math.h>

#include "common.cuh"
#include "host_util.h"
#include "ops.h"

namespace sta {

namespace {

struct Sim3d {
  double t[3];
  double q[4];  // x, y, z, w
  double s;
};

__device__ __forceinline__ void quat_mul(const double* a, const double* b, double* o) {
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  o[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
}
__device__ __forceinline__ void quat_rot(const double* q, const double* v, double* o) {
  // v + 2 w (u x v) + 2 u x (u x v)
  const double cx = q[1] * v[2] - q[2] * v[1], cy = q[2] * v[0] - q[0] * v[2], cz = q[0] * v[1] - q[1] * v[0];
  const double dx = q[1] * cz - q[2] * cy, dy = q[2] * cx - q[0] * cz, dz = q[0] * cy - q[1] * cx;
  o[0] = v[0] + 2.0 * (q[3] * cx + dx);
  o[1] = v[1] + 2.0 * (q[3] * cy + dy);
  o[2] = v[2] + 2.0 * (q[3] * cz + dz);
}
__device__ __forceinline__ void quat_to_rot(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
__device__ __forceinline__ Sim3d sim3_load(const float* p) {
  Sim3d X;
  X.t[0] = p[0]; X.t[1] = p[1]; X.t[2] = p[2];
  double n = 0.0;
  for (int i = 0; i < 4; ++i) { X.q[i] = p[3 + i]; n += X.q[i] * X.q[i]; }
  n = 1.0 / sqrt(n);
  for (int i = 0; i < 4; ++i) X.q[i] *= n;
  X.s = p[7];
  return X;
}
__device__ __forceinline__ Sim3d sim3_mul(const Sim3d& A, const Sim3d& B) {  // x -> s R x + t composition
  Sim3d C;
  double rt[3];
  quat_rot(A.q, B.t, rt);
  for (int i = 0; i < 3; ++i) C.t[i] = A.s * rt[i] + A.t[i];
  quat_mul(A.q, B.q, C.q);
  C.s = A.s * B.s;
  return C;
}
__device__ __forceinline__ Sim3d sim3_inv(const Sim3d& X) {
  Sim3d Y;
  Y.q[0] = -X.q[0]; Y.q[1] = -X.q[1]; Y.q[2] = -X.q[2]; Y.q[3] = X.q[3];
  Y.s = 1.0 / X.s;
  double rt[3];
  quat_rot(Y.q, X.t, rt);
  for (int i = 0; i < 3; ++i) Y.t[i] = -Y.s * rt[i];
  return Y;
}

// W = C I + A Phi + B Phi^2 = int_0^1 exp(u (sigma I + Phi)) du  (translation part of the Sim(3) exponential)
__device__ void w_coeffs(double sigma, double theta, double& C, double& A, double& B) {
  // int_0^1 e^{u sigma} u^k du = sum_n sigma^n / (n! (n + k + 1)): series below |sigma| = 1e-2 (the closed forms cancel)
  const double es = exp(sigma);
  const bool small_s = fabs(sigma) < 1e-2;
  const double s1 = sigma, s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2;
  C = small_s ? 1.0 + s1 / 2 + s2 / 6 + s3 / 24 + s4 / 120 : expm1(sigma) / sigma;
  if (theta < 1e-4) {  // theta^2 corrections (<= 4e-10) are far below the fp32 resolution of the inputs
    if (small_s) {
      A = 0.5 + s1 / 3 + s2 / 8 + s3 / 30 + s4 / 144;
      B = 1.0 / 6 + s1 / 8 + s2 / 20 + s3 / 72 + s4 / 336;
    } else {
      A = (es * (sigma - 1.0) + 1.0) / s2;
      B = (es * (s2 - 2.0 * sigma + 2.0) - 2.0) / (2.0 * s3);
    }
  } else {
    const double den = s2 + theta * theta, sn = sin(theta), cs = cos(theta);
    A = (es * (sigma * sn - theta * cs) + theta) / (theta * den);
    B = (C - (es * (sigma * cs + theta * sn) - sigma) / den) / (theta * theta);
  }
}
__device__ void w_matrix(const double* phi, double sigma, double* W) {
  const double th = sqrt(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
  double C, A, B;
  w_coeffs(sigma, th, C, A, B);
  const double x = phi[0], y = phi[1], z = phi[2];
  // Phi = [[0,-z,y],[z,0,-x],[-y,x,0]];  Phi^2 = phi phi^T - |phi|^2 I
  const double t2 = th * th;
  W[0] = C + B * (x * x - t2); W[1] = -A * z + B * x * y;   W[2] = A * y + B * x * z;
  W[3] = A * z + B * x * y;    W[4] = C + B * (y * y - t2); W[5] = -A * x + B * y * z;
  W[6] = -A * y + B * x * z;   W[7] = A * x + B * y * z;    W[8] = C + B * (z * z - t2);
}
__device__ void sim3_log(const Sim3d& X, double* xi) {
  double q[4] = {X.q[0], X.q[1], X.q[2], X.q[3]};
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  const double k = (n < 1e-12) ? 2.0 : 2.0 * atan2(n, q[3]) / n;
  xi[3] = k * q[0]; xi[4] = k * q[1]; xi[5] = k * q[2];
  xi[6] = log(X.s);
  double W[9];
  w_matrix(xi + 3, xi[6], W);
  // tau = W^-1 t (3x3 inverse by cofactors)
  const double c0 = W[4] * W[8] - W[5] * W[7], c1 = W[5] * W[6] - W[3] * W[8], c2 = W[3] * W[7] - W[4] * W[6];
  const double det = W[0] * c0 + W[1] * c1 + W[2] * c2, id = 1.0 / det;
  const double* t = X.t;
  xi[0] = id * (c0 * t[0] + (W[2] * W[7] - W[1] * W[8]) * t[1] + (W[1] * W[5] - W[2] * W[4]) * t[2]);
  xi[1] = id * (c1 * t[0] + (W[0] * W[8] - W[2] * W[6]) * t[1] + (W[2] * W[3] - W[0] * W[5]) * t[2]);
  xi[2] = id * (c2 * t[0] + (W[1] * W[6] - W[0] * W[7]) * t[1] + (W[0] * W[4] - W[1] * W[3]) * t[2]);
}
__device__ Sim3d sim3_exp(const double* xi) {
  Sim3d X;
  const double* phi = xi + 3;
  const double th = sqrt(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
  const double k = (th < 1e-8) ? 0.5 - th * th / 48.0 : sin(0.5 * th) / th;
  X.q[0] = k * phi[0]; X.q[1] = k * phi[1]; X.q[2] = k * phi[2]; X.q[3] = cos(0.5 * th);
  X.s = exp(xi[6]);
  double W[9];
  w_matrix(phi, xi[6], W);
  for (int i = 0; i < 3; ++i) X.t[i] = W[3 * i] * xi[0] + W[3 * i + 1] * xi[1] + W[3 * i + 2] * xi[2];
  return X;
}

// 7x7 helpers (row-major)
__device__ void mat7_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 7; ++j) {
      double s = 0.0;
      for (int k = 0; k < 7; ++k) s = fma(A[7 * i + k], B[7 * k + j], s);
      C[7 * i + j] = s;
    }
}
// Ad_X, tangent order [tau | phi | sigma]
__device__ void sim3_adj(const Sim3d& X, double* A) {
  double R[9];
  quat_to_rot(X.q, R);
  for (int i = 0; i < 49; ++i) A[i] = 0.0;
  const double* t = X.t;
  const double H[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      A[7 * i + j] = X.s * R[3 * i + j];
      A[7 * i + 3 + j] = H[3 * i] * R[j] + H[3 * i + 1] * R[3 + j] + H[3 * i + 2] * R[6 + j];
      A[7 * (3 + i) + 3 + j] = R[3 * i + j];
    }
  for (int i = 0; i < 3; ++i) A[7 * i + 6] = -t[i];
  A[48] = 1.0;
}
// ad_xi
__device__ void sim3_ad(const double* xi, double* a) {
  for (int i = 0; i < 49; ++i) a[i] = 0.0;
  const double* tau = xi;
  const double* phi = xi + 3;
  const double Hp[9] = {0, -phi[2], phi[1], phi[2], 0, -phi[0], -phi[1], phi[0], 0};
  const double Ht[9] = {0, -tau[2], tau[1], tau[2], 0, -tau[0], -tau[1], tau[0], 0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      a[7 * i + j] = Hp[3 * i + j] + (i == j ? xi[6] : 0.0);
      a[7 * i + 3 + j] = Ht[3 * i + j];
      a[7 * (3 + i) + 3 + j] = Hp[3 * i + j];
    }
  for (int i = 0; i < 3; ++i) a[7 * i + 6] = -tau[i];
}
// P = phi_1(M) = sum_n M^n / (n+1)!  by scaling and squaring: phi_1(2M) = (exp(M) + I) phi_1(M) / 2, exp(2M) = exp(M)^2
__device__ void phi1_7(const double* M, double* P) {
  double nrm = 0.0;
  for (int i = 0; i < 7; ++i) {
    double s = 0.0;
    for (int j = 0; j < 7; ++j) s += fabs(M[7 * i + j]);
    nrm = fmax(nrm, s);
  }
  int k = 0;
  double sc = 1.0;
  while (nrm * sc > 0.5 && k < 30) { sc *= 0.5; ++k; }
  double Ms[49], T[49], U[49], E[49];
  for (int i = 0; i < 49; ++i) { Ms[i] = M[i] * sc; P[i] = 0.0; T[i] = 0.0; }
  for (int i = 0; i < 7; ++i) { P[8 * i] = 1.0; T[8 * i] = 1.0; }
  for (int n = 1; n <= 14; ++n) {  // T = Ms^n / (n+1)!
    mat7_mul(T, Ms, U);
    const double f = 1.0 / (n + 1);
    for (int i = 0; i < 49; ++i) { T[i] = U[i] * f; P[i] += T[i]; }
  }
  mat7_mul(Ms, P, E);  // exp(Ms) = I + Ms phi_1(Ms)
  for (int i = 0; i < 7; ++i) E[8 * i] += 1.0;
  for (int s = 0; s < k; ++s) {
    for (int i = 0; i < 49; ++i) T[i] = 0.5 * E[i];
    for (int i = 0; i < 7; ++i) T[8 * i] += 0.5;
    mat7_mul(T, P, U);
    for (int i = 0; i < 49; ++i) P[i] = U[i];
    mat7_mul(E, E, U);
    for (int i = 0; i < 49; ++i) E[i] = U[i];
  }
}
// in-place Gauss-Jordan inverse with partial pivoting; returns false if singular
__device__ bool mat7_inv(double* A, double* Inv) {
  for (int i = 0; i < 49; ++i) Inv[i] = 0.0;
  for (int i = 0; i < 7; ++i) Inv[8 * i] = 1.0;
  for (int c = 0; c < 7; ++c) {
    int piv = c;
    double best = fabs(A[7 * c + c]);
    for (int r = c + 1; r < 7; ++r)
      if (fabs(A[7 * r + c]) > best) { best = fabs(A[7 * r + c]); piv = r; }
    if (best < 1e-300) return false;
    if (piv != c)
      for (int j = 0; j < 7; ++j) {
        double t = A[7 * c + j]; A[7 * c + j] = A[7 * piv + j]; A[7 * piv + j] = t;
        t = Inv[7 * c + j]; Inv[7 * c + j] = Inv[7 * piv + j]; Inv[7 * piv + j] = t;
      }
    const double d = 1.0 / A[7 * c + c];
    for (int j = 0; j < 7; ++j) { A[7 * c + j] *= d; Inv[7 * c + j] *= d; }
    for (int r = 0; r < 7; ++r) {
      if (r == c) continue;
      const double f = A[7 * r + c];
      if (f == 0.0) continue;
      for (int j = 0; j < 7; ++j) { A[7 * r + j] -= f * A[7 * c + j]; Inv[7 * r + j] -= f * Inv[7 * c + j]; }
    }
  }
  return true;
}

// ---------------------------------------------------------------------------
// one thread per edge: residual (and, if J != null, dr/d delta_j; dr/d delta_i = -that)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
pg_edge_kernel(const float* __restrict__ nodes, const long long* __restrict__ edges, const float* __restrict__ meas,
               const int* __restrict__ opt_map, int E, double* __restrict__ r_out, double* __restrict__ J_out,
               int* __restrict__ related) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const long long i = edges[2 * e], j = edges[2 * e + 1];
  const bool rel = opt_map[i] >= 0 || opt_map[j] >= 0;  // pose_graph.py:150-154
  related[e] = rel ? 1 : 0;
  if (!rel) {
    for (int k = 0; k < 7; ++k) r_out[7 * e + k] = 0.0;
    return;
  }
  const Sim3d T = sim3_load(meas + 8 * e), Xi = sim3_load(nodes + 8 * i), Xj = sim3_load(nodes + 8 * j);
  const Sim3d A = sim3_mul(T, sim3_inv(Xi));
  const Sim3d Err = sim3_mul(A, Xj);
  double r[7];
  sim3_log(Err, r);
  for (int k = 0; k < 7; ++k) r_out[7 * e + k] = r[k];
  if (J_out == nullptr) return;
  double ad[49], Jl[49], Jinv[49], Ad[49], J[49];
  sim3_ad(r, ad);
  phi1_7(ad, Jl);
  if (!mat7_inv(Jl, Jinv)) {
    for (int k = 0; k < 49; ++k) Jinv[k] = (k % 8 == 0) ? 1.0 : 0.0;
  }
  sim3_adj(A, Ad);
  mat7_mul(Jinv, Ad, J);
  for (int k = 0; k < 49; ++k) J_out[49 * e + k] = J[k];
}

// ---------------------------------------------------------------------------
// one CTA per optimised node a: block row a of A = J^T W J and of g = J^T W r.  64 threads: thread (p, q) = entry of
// the 7x7 blocks (49), threads 49..55 = the 7 entries of g.  Edges are visited in index order: fixed summation order.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
pg_assemble_kernel(const long long* __restrict__ edges, const float* __restrict__ weights, const int* __restrict__ opt_map,
                   const int* __restrict__ related, const double* __restrict__ r, const double* __restrict__ J, int E, int n7,
                   double* __restrict__ A, double* __restrict__ g) {
  const int a = blockIdx.x;
  const int t = threadIdx.x;
  const int p = t / 7, q = t - 7 * p;
  double diag = 0.0, gv = 0.0;
  for (int e = 0; e < E; ++e) {
    if (!related[e]) continue;
    const int oi = opt_map[edges[2 * e]], oj = opt_map[edges[2 * e + 1]];
    if (oi != a && oj != a) continue;
    const double* Je = J + 49ll * e;
    const float* w = weights + 7 * e;
    for (int side = 0; side < 2; ++side) {  // side 0: a is the first endpoint (J_a = -J), side 1: the second (J_a = +J)
      if ((side == 0 ? oi : oj) != a) continue;
      const int b = side == 0 ? oj : oi;
      if (t < 49) {
        double s = 0.0;
        for (int k = 0; k < 7; ++k) s = fma(Je[7 * k + p] * static_cast<double>(w[k]), Je[7 * k + q], s);
        diag += s;                                             // J_a^T W J_a
        if (b >= 0) A[(7ll * a + p) * n7 + 7 * b + q] -= s;     // J_a^T W J_b with J_b = -J_a (single owner thread)
      } else if (t < 56) {
        const int pp = t - 49;
        double s = 0.0;
        for (int k = 0; k < 7; ++k) s = fma(Je[7 * k + pp] * static_cast<double>(w[k]), r[7ll * e + k], s);
        gv += side == 0 ? -s : s;
      }
    }
  }
  if (t < 49) A[(7ll * a + p) * n7 + 7 * a + q] += diag;
  else if (t < 56) g[7 * a + (t - 49)] = gv;
}

// LM damping of pp.optim.LM.step: A.diagonal().clamp_(min, max); A.diagonal() += damping * A.diagonal()
__global__ void pg_damp_kernel(double* A, int n7, double dmin, double dmax, double damping) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n7) return;
  double d = A[static_cast<long long>(i) * n7 + i];
  d = fmin(fmax(d, dmin), dmax);
  A[static_cast<long long>(i) * n7 + i] = d * (1.0 + damping);
}

// ---------------------------------------------------------------------------
// dense blocked right-looking Cholesky (lower), fp64, 32-wide panels.  info[3] is set to 0 on a non-positive pivot.
// ---------------------------------------------------------------------------
constexpr int NB = 32;
__global__ void __launch_bounds__(NB * NB) chol_diag_kernel(double* A, int n, int k0, double* info) {
  __shared__ double S[NB][NB + 1];
  const int nb = min(NB, n - k0);
  const int r = threadIdx.y, c = threadIdx.x;
  if (r < nb && c < nb) S[r][c] = A[static_cast<long long>(k0 + r) * n + k0 + c];
  __syncthreads();
  for (int k = 0; k < nb; ++k) {
    if (r == k && c == k) {
      const double d = S[k][k];
      if (!(d > 0.0)) { info[3] = 0.0; S[k][k] = 1.0; } else S[k][k] = sqrt(d);
    }
    __syncthreads();
    if (c == k && r > k && r < nb) S[r][k] /= S[k][k];
    __syncthreads();
    if (r > k && r < nb && c > k && c <= r) S[r][c] -= S[r][k] * S[c][k];
    __syncthreads();
  }
  if (r < nb && c < nb && c <= r) A[static_cast<long long>(k0 + r) * n + k0 + c] = S[r][c];
}
// rows below the panel: X L_kk^T = A_ik  (one thread per row)
__global__ void __launch_bounds__(NB) chol_trsm_kernel(double* A, int n, int k0) {
  __shared__ double L[NB][NB + 1];
  const int nb = min(NB, n - k0);
  for (int idx = threadIdx.x; idx < NB * NB; idx += NB) {
    const int r = idx / NB, c = idx % NB;
    L[r][c] = (r < nb && c <= r) ? A[static_cast<long long>(k0 + r) * n + k0 + c] : 0.0;
  }
  __syncthreads();
  const int row = k0 + nb + blockIdx.x * NB + threadIdx.x;
  if (row >= n) return;
  double x[NB];
  double* ar = A + static_cast<long long>(row) * n + k0;
  for (int c = 0; c < nb; ++c) x[c] = ar[c];
  for (int c = 0; c < nb; ++c) {
    double s = x[c];
    for (int cc = 0; cc < c; ++cc) s -= x[cc] * L[c][cc];
    x[c] = s / L[c][c];
  }
  for (int c = 0; c < nb; ++c) ar[c] = x[c];
}
// trailing update: A_ij -= L_ik L_jk^T for tiles j <= i below / right of the panel
__global__ void __launch_bounds__(NB * NB) chol_syrk_kernel(double* A, int n, int k0) {
  if (blockIdx.x > blockIdx.y) return;  // (bj, bi) with bj <= bi
  __shared__ double Li[NB][NB + 1], Lj[NB][NB + 1];
  const int nb = min(NB, n - k0);
  const int base = k0 + nb;
  const int i = base + blockIdx.y * NB + threadIdx.y, j = base + blockIdx.x * NB + threadIdx.x;
  const int li = base + blockIdx.y * NB + threadIdx.y, lj = base + blockIdx.x * NB + threadIdx.y;
  Li[threadIdx.y][threadIdx.x] = (li < n && threadIdx.x < nb) ? A[static_cast<long long>(li) * n + k0 + threadIdx.x] : 0.0;
  Lj[threadIdx.y][threadIdx.x] = (lj < n && threadIdx.x < nb) ? A[static_cast<long long>(lj) * n + k0 + threadIdx.x] : 0.0;
  __syncthreads();
  if (i >= n || j >= n || j > i) return;
  double s = 0.0;
  for (int c = 0; c < NB; ++c) s = fma(Li[threadIdx.y][c], Lj[threadIdx.x][c], s);
  A[static_cast<long long>(i) * n + j] -= s;
}
// x = -(L L^T)^-1 g, one CTA: blocked forward then backward substitution with the right-hand side in shared memory
__global__ void __launch_bounds__(1024) pg_solve_kernel(const double* __restrict__ A, const double* __restrict__ g, int n,
                                                        double* __restrict__ x) {
  extern __shared__ double b[];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < n; i += nt) b[i] = -g[i];
  __syncthreads();
  for (int k0 = 0; k0 < n; k0 += NB) {  // L y = b
    const int nb = min(NB, n - k0);
    if (tid == 0) {
      for (int c = 0; c < nb; ++c) {
        double s = b[k0 + c];
        const double* lr = A + static_cast<long long>(k0 + c) * n + k0;
        for (int cc = 0; cc < c; ++cc) s -= lr[cc] * b[k0 + cc];
        b[k0 + c] = s / lr[c];
      }
    }
    __syncthreads();
    for (int i = k0 + nb + tid; i < n; i += nt) {
      const double* lr = A + static_cast<long long>(i) * n + k0;
      double s = b[i];
      for (int c = 0; c < nb; ++c) s -= lr[c] * b[k0 + c];
      b[i] = s;
    }
    __syncthreads();
  }
  const int nblk = (n + NB - 1) / NB;
  for (int kb = nblk - 1; kb >= 0; --kb) {  // L^T x = y
    const int k0 = kb * NB, nb = min(NB, n - k0);
    if (tid == 0) {
      for (int c = nb - 1; c >= 0; --c) {
        double s = b[k0 + c];
        for (int cc = c + 1; cc < nb; ++cc) s -= A[static_cast<long long>(k0 + cc) * n + k0 + c] * b[k0 + cc];
        b[k0 + c] = s / A[static_cast<long long>(k0 + c) * n + k0 + c];
      }
    }
    __syncthreads();
    for (int i = tid; i < k0; i += nt) {  // b_i -= sum_c L[k0 + c][i] x[k0 + c]
      double s = b[i];
      for (int c = 0; c < nb; ++c) s -= A[static_cast<long long>(k0 + c) * n + i] * b[k0 + c];
      b[i] = s;
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt) x[i] = b[i];
}

// nodes_out = Exp(delta_a) * node for optimised nodes (LieTensor.add_), copy for fixed ones
__global__ void pg_update_kernel(const float* __restrict__ nodes, const int* __restrict__ opt_map, const double* __restrict__ delta,
                                 int num_nodes, float* __restrict__ out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= num_nodes) return;
  const int a = opt_map[v];
  if (a < 0) {
    for (int k = 0; k < 8; ++k) out[8 * v + k] = nodes[8 * v + k];
    return;
  }
  const Sim3d X = sim3_load(nodes + 8 * v);
  const Sim3d Y = sim3_mul(sim3_exp(delta + 7 * a), X);
  const double qn = 1.0 / sqrt(Y.q[0] * Y.q[0] + Y.q[1] * Y.q[1] + Y.q[2] * Y.q[2] + Y.q[3] * Y.q[3]);
  for (int k = 0; k < 3; ++k) out[8 * v + k] = static_cast<float>(Y.t[k]);
  for (int k = 0; k < 4; ++k) out[8 * v + 3 + k] = static_cast<float>(Y.q[k] * qn);
  out[8 * v + 7] = static_cast<float>(Y.s);
}

// loss = sum_e r_e^T W_e r_e (related edges), |delta|_2 -> info; fixed-order reduction in one CTA
__global__ void __launch_bounds__(256) pg_loss_kernel(const double* __restrict__ r, const float* __restrict__ weights,
                                                      const int* __restrict__ related, int E, const double* __restrict__ delta,
                                                      int n7, double* __restrict__ info, int slot) {
  __shared__ double red[256];
  double s = 0.0;
  for (int e = threadIdx.x; e < E; e += 256) {
    if (!related[e]) continue;
    for (int k = 0; k < 7; ++k) s += r[7ll * e + k] * r[7ll * e + k] * static_cast<double>(weights[7 * e + k]);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) info[slot] = red[0];
  if (delta != nullptr) {
    __syncthreads();
    double d = 0.0;
    for (int i = threadIdx.x; i < n7; i += 256) d += delta[i] * delta[i];
    red[threadIdx.x] = d;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) info[2] = sqrt(red[0]);
  }
}

__global__ void pg_opt_map_kernel(const long long* __restrict__ opt_idx, int num_opt, int num_nodes, int* __restrict__ opt_map,
                                  int phase) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (phase == 0) {
    if (i < num_nodes) opt_map[i] = -1;
  } else if (i < num_opt) {
    const long long v = opt_idx[i];
    if (v >= 0 && v < num_nodes) opt_map[v] = i;  // pose_graph.py:90-91
  }
}

struct PgScratch {
  double *r, *J, *A, *g, *delta, *r2, *info;
  int *related, *opt_map;
};
size_t carve(char* base, int Nn, int E, int n_opt, PgScratch* s) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = (off + 255) & ~static_cast<size_t>(255);
    char* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  };
  const size_t n7 = 7ull * n_opt;
  char* r = take(7ull * E * 8);
  char* J = take(49ull * E * 8);
  char* A = take(n7 * n7 * 8);
  char* g = take(n7 * 8);
  char* d = take(n7 * 8);
  char* r2 = take(7ull * E * 8);
  char* info = take(4 * 8);
  char* rel = take(4ull * E);
  char* om = take(4ull * Nn);
  if (s) {
    s->r = reinterpret_cast<double*>(r); s->J = reinterpret_cast<double*>(J); s->A = reinterpret_cast<double*>(A);
    s->g = reinterpret_cast<double*>(g); s->delta = reinterpret_cast<double*>(d); s->r2 = reinterpret_cast<double*>(r2);
    s->info = reinterpret_cast<double*>(info); s->related = reinterpret_cast<int*>(rel); s->opt_map = reinterpret_cast<int*>(om);
  }
  return off + 256;
}

}  // namespace

size_t pose_graph_scratch_bytes(int num_nodes, int num_edges, int num_opt) {
  if (num_nodes < 0 || num_edges < 0 || num_opt < 0) return 0;
  return carve(nullptr, num_nodes, num_edges, num_opt, nullptr);
}

int launch_pose_graph_lm_step(const float* nodes, int num_nodes, const long long* edges, const float* meas, const float* weights,
                              int num_edges, const long long* opt_idx, int num_opt, double damping, double dmin, double dmax,
                              float* nodes_out, double* info_out, void* scratch, cudaStream_t st) {
  STA_REQUIRE(nodes && edges && meas && weights && opt_idx && nodes_out && info_out && scratch, "null pointer");
  STA_REQUIRE(num_nodes > 0 && num_edges > 0 && num_opt > 0 && num_opt <= num_nodes, "empty problem");
  STA_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 7) == 0, "scratch must be 8-byte aligned");
  const int n7 = 7 * num_opt;
  STA_REQUIRE(static_cast<size_t>(n7) * 8 <= 200 * 1024, "more than 3657 optimised nodes: the one-CTA solve needs the RHS in shared memory");
  PgScratch s;
  carve(static_cast<char*>(scratch), num_nodes, num_edges, num_opt, &s);
  const int E = num_edges;
  pg_opt_map_kernel<<<(num_nodes + 255) / 256, 256, 0, st>>>(opt_idx, num_opt, num_nodes, s.opt_map, 0);
  pg_opt_map_kernel<<<(num_opt + 255) / 256, 256, 0, st>>>(opt_idx, num_opt, num_nodes, s.opt_map, 1);
  STA_CHECK_CUDA(cudaMemsetAsync(s.A, 0, static_cast<size_t>(n7) * n7 * sizeof(double), st));
  const double one = 1.0;
  STA_CHECK_CUDA(cudaMemcpyAsync(s.info + 3, &one, sizeof(double), cudaMemcpyHostToDevice, st));  // Cholesky-ok flag
  pg_edge_kernel<<<(E + 63) / 64, 64, 0, st>>>(nodes, edges, meas, s.opt_map, E, s.r, s.J, s.related);
  pg_assemble_kernel<<<num_opt, 64, 0, st>>>(edges, weights, s.opt_map, s.related, s.r, s.J, E, n7, s.A, s.g);
  pg_loss_kernel<<<1, 256, 0, st>>>(s.r, weights, s.related, E, nullptr, 0, s.info, 0);
  pg_damp_kernel<<<(n7 + 255) / 256, 256, 0, st>>>(s.A, n7, dmin, dmax, damping);
  for (int k0 = 0; k0 < n7; k0 += NB) {
    chol_diag_kernel<<<1, dim3(NB, NB), 0, st>>>(s.A, n7, k0, s.info);
    const int rest = n7 - k0 - NB;
    if (rest > 0) {
      const int nt = (rest + NB - 1) / NB;
      chol_trsm_kernel<<<nt, NB, 0, st>>>(s.A, n7, k0);
      chol_syrk_kernel<<<dim3(nt, nt), dim3(NB, NB), 0, st>>>(s.A, n7, k0);
    }
  }
  static PerDeviceOnce once;
  STA_CHECK_CUDA(once.run([&] {
    return cudaFuncSetAttribute(pg_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }));
  pg_solve_kernel<<<1, 1024, static_cast<size_t>(n7) * sizeof(double), st>>>(s.A, s.g, n7, s.delta);
  pg_update_kernel<<<(num_nodes + 127) / 128, 128, 0, st>>>(nodes, s.opt_map, s.delta, num_nodes, nodes_out);
  pg_edge_kernel<<<(E + 63) / 64, 64, 0, st>>>(nodes_out, edges, meas, s.opt_map, E, s.r2, nullptr, s.related);
  pg_loss_kernel<<<1, 256, 0, st>>>(s.r2, weights, s.related, E, s.delta, n7, s.info, 1);
  STA_CHECK_CUDA(cudaMemcpyAsync(info_out, s.info, 4 * sizeof(double), cudaMemcpyDeviceToDevice, st));
  STA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sta
