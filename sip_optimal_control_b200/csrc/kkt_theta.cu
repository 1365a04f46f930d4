// See kkt_theta.cuh.  Small generic kernels (any tree, any dims, p <= kMaxThetaDim); the
// heavy part of the theta path -- p stagewise solves per factor -- runs on the LQR kernels.
#include "kkt_theta.cuh"

#include "generic_kernels.cuh"

namespace sipoc {
namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ int64_t problem() {
  return static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
}

// One thread per (problem, theta column): column j of J, every row written once
// (helpers.cpp:190-240).
__global__ void __launch_bounds__(kThreads)
theta_jacobian_kernel(DevTables t, KktThetaModel m, double *J, int64_t batch, int64_t ld) {
  const int64_t b = problem();
  if (b >= batch) return;
  const int p = t.theta_dim, j = blockIdx.y;
  const size_t L = static_cast<size_t>(ld);
  GVec col{J + static_cast<size_t>(j) * t.kkt_dim * L + b, L};
  GCVec nhxt{m.node_hxt + b, L}, njct{m.node_jct + b, L}, njgt{m.node_jgt + b, L};
  GCVec ehxt{m.edge_hxt + b, L}, ehut{m.edge_hut + b, L}, edyn{m.edge_dynt + b, L},
      ejct{m.edge_jct + b, L}, ejgt{m.edge_jgt + b, L};
  const int xd = t.x_dim, yd = t.y_dim;
  // rows no block owns: the root's dynamics row and theta itself
  for (int r = 0; r < t.n[t.root]; ++r) col(xd + t.y_dyn[t.root] + r) = 0.0;
  for (int r = 0; r < p; ++r) col(t.sx_dim + r) = 0.0;
  for (int i = 0; i < t.N; ++i) {
    const int n = t.n[i], c = t.node_c[i], g = t.node_g[i];
    for (int r = 0; r < n; ++r) col(t.x_state[i] + r) = nhxt(t.n_off[i] * p + j * n + r);
    for (int r = 0; r < c; ++r)
      col(xd + t.y_node_c[i] + r) = njct(t.node_c_off[i] * p + j * c + r);
    for (int r = 0; r < g; ++r)
      col(xd + yd + t.z_node[i] + r) = njgt(t.node_g_off[i] * p + j * g + r);
  }
  for (int e = 0; e < t.E; ++e) {
    const int par = t.parents[e], ch = t.children[e];
    const int np = t.n[par], nc = t.n[ch], mm = t.m[e], c = t.edge_c[e], g = t.edge_g[e];
    for (int r = 0; r < np; ++r) col(t.x_state[par] + r) += ehxt(t.pn_off[e] * p + j * np + r);
    for (int r = 0; r < mm; ++r) col(t.x_control[e] + r) = ehut(t.m_off[e] * p + j * mm + r);
    for (int r = 0; r < nc; ++r) col(xd + t.y_dyn[ch] + r) = edyn(t.cn_off[e] * p + j * nc + r);
    for (int r = 0; r < c; ++r)
      col(xd + t.y_edge_c[e] + r) = ejct(t.edge_c_off[e] * p + j * c + r);
    for (int r = 0; r < g; ++r)
      col(xd + yd + t.z_edge[e] + r) = ejgt(t.edge_g_off[e] * p + j * g + r);
  }
}

// One thread per (problem, lower entry (i, j)) of S (helpers.cpp:389-401).
__global__ void __launch_bounds__(kThreads)
theta_schur_kernel(DevTables t, KktThetaModel m, const double *r1, const double *J,
                   const double *KinvJ, double *S, int64_t batch, int64_t ld) {
  const int64_t b = problem();
  if (b >= batch) return;
  const int p = t.theta_dim, i = blockIdx.y % p, j = blockIdx.y / p;
  if (i < j) return;
  const size_t L = static_cast<size_t>(ld);
  GCVec nhtt{m.node_htt + b, L}, ehtt{m.edge_htt + b, L}, R1{r1 + b, L};
  double s = 0.0;
  for (int k = 0; k < t.N; ++k) s += nhtt(k * p * p + j * p + i);
  for (int e = 0; e < t.E; ++e) s += ehtt(e * p * p + j * p + i);
  if (i == j) s += R1(t.sx_dim + i);
  GCVec ji{J + static_cast<size_t>(i) * t.kkt_dim * L + b, L},
      kj{KinvJ + static_cast<size_t>(j) * t.kkt_dim * L + b, L};
  double dot = 0.0, dot2 = 0.0;
  auto span = [&](int lo, int hi) {
    int r = lo;
    for (; r + 1 < hi; r += 2) {
      dot += ji(r) * kj(r);
      dot2 += ji(r + 1) * kj(r + 1);
    }
    if (r < hi) dot += ji(r) * kj(r);
  };
  span(0, t.sx_dim);            // the theta rows hold no stagewise entry
  span(t.x_dim, t.kkt_dim);
  S[(static_cast<size_t>(j) * p + i) * L + b] = s - (dot + dot2);
}

// One thread per problem: lower Cholesky of S in place (helpers.cpp:403-407).
__global__ void __launch_bounds__(kThreads)
theta_chol_kernel(int p, double *S, int *ok, int64_t batch, int64_t ld) {
  const int64_t b = problem();
  if (b >= batch) return;
  GVec A{S + b, static_cast<size_t>(ld)};
  bool good = true;
  for (int j = 0; j < p; ++j) {
    double d = A(j * p + j);
    for (int k = 0; k < j; ++k) d -= A(k * p + j) * A(k * p + j);
    if (!(d > 0.0)) good = false;
    const double l = sqrt(d);
    A(j * p + j) = l;
    for (int i = j + 1; i < p; ++i) {
      double v = A(j * p + i);
      for (int k = 0; k < j; ++k) v -= A(k * p + i) * A(k * p + j);
      A(j * p + i) = v / l;
    }
  }
  if (!good && ok != nullptr) ok[b] = 0;
}

// t_i = b_theta,i - J(:, i)' sol over the stagewise rows (helpers.cpp:923-929).
__global__ void __launch_bounds__(kThreads)
theta_rhs_kernel(DevTables t, const double *bvec, const double *J, const double *sol, double *tvec,
                 int64_t batch, int64_t ld) {
  const int64_t b = problem();
  if (b >= batch) return;
  const int i = blockIdx.y;
  const size_t L = static_cast<size_t>(ld);
  GCVec ji{J + static_cast<size_t>(i) * t.kkt_dim * L + b, L}, x{sol + b, L}, rhs{bvec + b, L};
  double dot = 0.0, dot2 = 0.0;
  auto span = [&](int lo, int hi) {
    int r = lo;
    for (; r + 1 < hi; r += 2) {
      dot += ji(r) * x(r);
      dot2 += ji(r + 1) * x(r + 1);
    }
    if (r < hi) dot += ji(r) * x(r);
  };
  span(0, t.sx_dim);
  span(t.x_dim, t.kkt_dim);
  tvec[static_cast<size_t>(i) * L + b] = rhs(t.sx_dim + i) - (dot + dot2);
}

// theta = S^-1 t by the two triangular solves (helpers.cpp:931-937), in place in tvec.
__global__ void __launch_bounds__(kThreads)
theta_backsolve_kernel(int p, const double *S, double *tvec, int64_t batch, int64_t ld) {
  const int64_t b = problem();
  if (b >= batch) return;
  const size_t L = static_cast<size_t>(ld);
  GCVec A{S + b, L};
  GVec v{tvec + b, L};
  for (int i = 0; i < p; ++i) {
    double s = v(i);
    for (int k = 0; k < i; ++k) s -= A(k * p + i) * v(k);
    v(i) = s / A(i * p + i);
  }
  for (int i = p - 1; i >= 0; --i) {
    double s = v(i);
    for (int k = i + 1; k < p; ++k) s -= A(i * p + k) * v(k);
    v(i) = s / A(i * p + i);
  }
}

// sol -= (Kinv J) theta on the stagewise rows, theta rows <- theta (helpers.cpp:939-950).
// One thread per (problem, chunk of rows).
constexpr int kRowChunk = 32;
__global__ void __launch_bounds__(kThreads)
theta_update_kernel(DevTables t, const double *KinvJ, const double *tvec, double *sol, int64_t batch,
                    int64_t ld) {
  const int64_t b = problem();
  if (b >= batch) return;
  const int p = t.theta_dim;
  const size_t L = static_cast<size_t>(ld);
  double th[kMaxThetaDim];
  for (int j = 0; j < p; ++j) th[j] = tvec[static_cast<size_t>(j) * L + b];
  GVec x{sol + b, L};
  const int r0 = blockIdx.y * kRowChunk, r1 = min(r0 + kRowChunk, t.kkt_dim);
  for (int r = r0; r < r1; ++r) {
    if (r >= t.sx_dim && r < t.x_dim) {
      x(r) = th[r - t.sx_dim];
      continue;
    }
    double s = 0.0;
    for (int j = 0; j < p; ++j)
      s += __ldg(KinvJ + (static_cast<size_t>(j) * t.kkt_dim + r) * L + b) * th[j];
    x(r) -= s;
  }
}

// theta terms of y += K x, one thread per problem walking nodes then edges
// (theta branches of helpers.cpp:1019-1368).
__global__ void __launch_bounds__(kThreads)
theta_apply_kernel(DevTables t, KktThetaModel m, unsigned parts, const double *r1,
                   const double *in_x, const double *in_y, const double *in_z, double *out_x,
                   double *out_y, double *out_z, int64_t batch, int64_t ld) {
  const int64_t b = problem();
  if (b >= batch) return;
  const int p = t.theta_dim;
  const size_t L = static_cast<size_t>(ld);
  const bool pH = parts & kKktH, pC = parts & kKktC, pCT = parts & kKktCT, pG = parts & kKktG,
             pGT = parts & kKktGT, pR = parts & kKktReg;
  GCVec X{in_x + b, L}, Y{in_y + b, L}, Z{in_z + b, L};
  GVec OX{out_x + b, L}, OY{out_y + b, L}, OZ{out_z + b, L};
  GCVec nhxt{m.node_hxt + b, L}, njct{m.node_jct + b, L}, njgt{m.node_jgt + b, L},
      nhtt{m.node_htt + b, L};
  GCVec ehxt{m.edge_hxt + b, L}, ehut{m.edge_hut + b, L}, edyn{m.edge_dynt + b, L},
      ejct{m.edge_jct + b, L}, ejgt{m.edge_jgt + b, L}, ehtt{m.edge_htt + b, L};
  double th[kMaxThetaDim], acc[kMaxThetaDim];
  const bool need_theta = pH || pC || pG || pR;
  for (int j = 0; j < p; ++j) {
    th[j] = need_theta ? X(t.sx_dim + j) : 0.0;
    acc[j] = 0.0;
  }
  // rows x p block times theta into `out` rows, and its transpose times `in` rows into acc
  auto forward = [&](const GCVec &blk, int base, int rows, const GVec &out, int row0) {
    for (int r = 0; r < rows; ++r) {
      double s = 0.0;
      for (int j = 0; j < p; ++j) s += blk(base + j * rows + r) * th[j];
      out(row0 + r) += s;
    }
  };
  auto adjoint = [&](const GCVec &blk, int base, int rows, const GCVec &in, int row0) {
    for (int j = 0; j < p; ++j) {
      double s = 0.0;
      for (int r = 0; r < rows; ++r) s += blk(base + j * rows + r) * in(row0 + r);
      acc[j] += s;
    }
  };
  auto hessian = [&](const GCVec &blk, int base) {
    for (int i = 0; i < p; ++i) {
      double s = 0.0;
      for (int j = 0; j < p; ++j) s += blk(base + j * p + i) * th[j];
      acc[i] += s;
    }
  };
  for (int i = 0; i < t.N; ++i) {
    const int n = t.n[i], c = t.node_c[i], g = t.node_g[i];
    if (pH) {
      forward(nhxt, t.n_off[i] * p, n, OX, t.x_state[i]);
      adjoint(nhxt, t.n_off[i] * p, n, X, t.x_state[i]);
      hessian(nhtt, i * p * p);
    }
    if (pC) forward(njct, t.node_c_off[i] * p, c, OY, t.y_node_c[i]);
    if (pCT) adjoint(njct, t.node_c_off[i] * p, c, Y, t.y_node_c[i]);
    if (pG) forward(njgt, t.node_g_off[i] * p, g, OZ, t.z_node[i]);
    if (pGT) adjoint(njgt, t.node_g_off[i] * p, g, Z, t.z_node[i]);
  }
  for (int e = 0; e < t.E; ++e) {
    const int par = t.parents[e], ch = t.children[e];
    const int np = t.n[par], nc = t.n[ch], mm = t.m[e], c = t.edge_c[e], g = t.edge_g[e];
    if (pH) {
      forward(ehxt, t.pn_off[e] * p, np, OX, t.x_state[par]);
      adjoint(ehxt, t.pn_off[e] * p, np, X, t.x_state[par]);
      forward(ehut, t.m_off[e] * p, mm, OX, t.x_control[e]);
      adjoint(ehut, t.m_off[e] * p, mm, X, t.x_control[e]);
      hessian(ehtt, e * p * p);
    }
    if (pC) {
      forward(edyn, t.cn_off[e] * p, nc, OY, t.y_dyn[ch]);
      forward(ejct, t.edge_c_off[e] * p, c, OY, t.y_edge_c[e]);
    }
    if (pCT) {
      adjoint(edyn, t.cn_off[e] * p, nc, Y, t.y_dyn[ch]);
      adjoint(ejct, t.edge_c_off[e] * p, c, Y, t.y_edge_c[e]);
    }
    if (pG) forward(ejgt, t.edge_g_off[e] * p, g, OZ, t.z_edge[e]);
    if (pGT) adjoint(ejgt, t.edge_g_off[e] * p, g, Z, t.z_edge[e]);
  }
  if (pR) {
    GCVec R1{r1 + b, L};
    for (int j = 0; j < p; ++j) acc[j] += R1(t.sx_dim + j) * th[j];
  }
  if (pH || pCT || pGT || pR)
    for (int j = 0; j < p; ++j) OX(t.sx_dim + j) += acc[j];
}

dim3 grid_for(int64_t batch, int y) {
  return dim3(static_cast<unsigned>((batch + kThreads - 1) / kThreads),
              static_cast<unsigned>(y > 0 ? y : 1));
}

}  // namespace

void launch_theta_jacobian(const DevTables &t, const KktThetaModel &m, double *J, int64_t batch,
                           int64_t ld, cudaStream_t s) {
  theta_jacobian_kernel<<<grid_for(batch, t.theta_dim), kThreads, 0, s>>>(t, m, J, batch, ld);
}

void launch_theta_schur(const DevTables &t, const KktThetaModel &m, const double *r1,
                        const double *J, const double *KinvJ, double *S, int *ok, int64_t batch,
                        int64_t ld, cudaStream_t s) {
  const int p = t.theta_dim;
  theta_schur_kernel<<<grid_for(batch, p * p), kThreads, 0, s>>>(t, m, r1, J, KinvJ, S, batch, ld);
  theta_chol_kernel<<<grid_for(batch, 1), kThreads, 0, s>>>(p, S, ok, batch, ld);
}

void launch_theta_solve(const DevTables &t, const double *b, const double *J, const double *KinvJ,
                        const double *S, double *tvec, double *sol, int64_t batch, int64_t ld,
                        cudaStream_t s) {
  const int p = t.theta_dim;
  theta_rhs_kernel<<<grid_for(batch, p), kThreads, 0, s>>>(t, b, J, sol, tvec, batch, ld);
  theta_backsolve_kernel<<<grid_for(batch, 1), kThreads, 0, s>>>(p, S, tvec, batch, ld);
  theta_update_kernel<<<grid_for(batch, (t.kkt_dim + kRowChunk - 1) / kRowChunk), kThreads, 0, s>>>(
      t, KinvJ, tvec, sol, batch, ld);
}

void launch_theta_apply(const DevTables &t, const KktThetaModel &m, unsigned parts, const double *r1,
                        const double *in_x, const double *in_y, const double *in_z, double *out_x,
                        double *out_y, double *out_z, int64_t batch, int64_t ld, cudaStream_t s) {
  theta_apply_kernel<<<grid_for(batch, 1), kThreads, 0, s>>>(t, m, parts, r1, in_x, in_y, in_z,
                                                             out_x, out_y, out_z, batch, ld);
}

}  // namespace sipoc
