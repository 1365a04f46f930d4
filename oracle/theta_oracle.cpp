// ORACLE — TEST INFRASTRUCTURE ONLY.  See theta_oracle.hpp.
#include "theta_oracle.hpp"

#include <cmath>

namespace sipoc_oracle {

namespace {
// prefix sums used to address the theta blocks
struct ThetaOffsets {
  std::vector<int> n_off, nc_off, ng_off, pn_off, m_off, cn_off, ec_off, eg_off;
  ThetaOffsets(const KktLayout &K, const Tree &t) {
    const int E = t.num_edges, N = E + 1;
    n_off.assign(N + 1, 0);
    nc_off.assign(N + 1, 0);
    ng_off.assign(N + 1, 0);
    for (int i = 0; i < N; ++i) {
      n_off[i + 1] = n_off[i] + K.lqr.n[i];
      nc_off[i + 1] = nc_off[i] + K.node_c[i];
      ng_off[i + 1] = ng_off[i] + K.node_g[i];
    }
    pn_off.assign(E + 1, 0);
    m_off.assign(E + 1, 0);
    cn_off.assign(E + 1, 0);
    ec_off.assign(E + 1, 0);
    eg_off.assign(E + 1, 0);
    for (int e = 0; e < E; ++e) {
      pn_off[e + 1] = pn_off[e] + K.lqr.n[t.parents[e]];
      m_off[e + 1] = m_off[e] + K.lqr.m[e];
      cn_off[e + 1] = cn_off[e] + K.lqr.n[t.children[e]];
      ec_off[e + 1] = ec_off[e] + K.edge_c[e];
      eg_off[e + 1] = eg_off[e] + K.edge_g[e];
    }
  }
};
}  // namespace

void theta_sizes(const KktLayout &K, const Tree &t, int p, long long out[10]) {
  const ThetaOffsets o(K, t);
  const int E = t.num_edges, N = E + 1;
  out[0] = 1LL * o.n_off[N] * p;
  out[1] = 1LL * o.nc_off[N] * p;
  out[2] = 1LL * o.ng_off[N] * p;
  out[3] = 1LL * N * p * p;
  out[4] = 1LL * o.pn_off[E] * p;
  out[5] = 1LL * o.m_off[E] * p;
  out[6] = 1LL * o.cn_off[E] * p;
  out[7] = 1LL * o.ec_off[E] * p;
  out[8] = 1LL * o.eg_off[E] * p;
  out[9] = 1LL * E * p * p;
}

// helpers.cpp:190-240.  J is [stagewise_kkt_dim x p] in the STAGEWISE layout [x_s | y | z].
static void form_theta_jacobian(const Tree &t, const KktLayout &K, int p, const ThetaModel &tm,
                                std::vector<double> &J) {
  const ThetaOffsets o(K, t);
  const int E = t.num_edges, N = E + 1, sx = K.x_dim, kd = K.kkt_dim;
  J.assign(static_cast<size_t>(kd) * p, 0.0);
  auto put = [&](int row0, const double *blk, int rows, bool add) {
    for (int j = 0; j < p; ++j)
      for (int r = 0; r < rows; ++r) {
        double &dst = J[static_cast<size_t>(j) * kd + row0 + r];
        dst = (add ? dst : 0.0) + blk[j * rows + r];
      }
  };
  for (int i = 0; i < N; ++i) {
    put(K.x_state[i], tm.node_hxt + o.n_off[i] * p, K.lqr.n[i], false);
    put(sx + K.y_node_c[i], tm.node_jct + o.nc_off[i] * p, K.node_c[i], false);
    put(sx + K.y_dim + K.z_node[i], tm.node_jgt + o.ng_off[i] * p, K.node_g[i], false);
  }
  for (int e = 0; e < E; ++e) {
    const int par = t.parents[e], ch = t.children[e];
    put(K.x_state[par], tm.edge_hxt + o.pn_off[e] * p, K.lqr.n[par], true);
    put(K.x_control[e], tm.edge_hut + o.m_off[e] * p, K.lqr.m[e], false);
    put(sx + K.y_dyn[ch], tm.edge_dynt + o.cn_off[e] * p, K.lqr.n[ch], false);
    put(sx + K.y_edge_c[e], tm.edge_jct + o.ec_off[e] * p, K.edge_c[e], false);
    put(sx + K.y_dim + K.z_edge[e], tm.edge_jgt + o.eg_off[e] * p, K.edge_g[e], false);
  }
}

// helpers.cpp:242-407.
bool theta_factor(const CompiledTree &ct, const Tree &t, const KktLayout &K, int p,
                  const KktModel &m, const ThetaModel &tm, const double *w, const double *r1,
                  const double *r2, const double *r3, ThetaWorkspace &ws) {
  // r1's stagewise entries come first, theta's last (types.cpp:24-64): the stagewise
  // reduction reads r1[0 .. sx).
  if (!kkt_factor(ct, K, m, w, r1, r2, r3, ws.kkt, nullptr)) return false;
  const int E = t.num_edges, N = E + 1, sx = K.x_dim, kd = K.kkt_dim;
  form_theta_jacobian(t, K, p, tm, ws.J);
  ws.KinvJ.assign(static_cast<size_t>(kd) * p, 0.0);
  for (int j = 0; j < p; ++j)  // :387, column by column
    kkt_solve(ct, K, m, ws.J.data() + static_cast<size_t>(j) * kd,
              ws.KinvJ.data() + static_cast<size_t>(j) * kd, ws.kkt);
  ws.S.assign(static_cast<size_t>(p) * p, 0.0);
  for (int i = 0; i < N; ++i)  // :391-394
    for (int q = 0; q < p * p; ++q) ws.S[q] += tm.node_htt[i * p * p + q];
  for (int e = 0; e < E; ++e)  // :395-398
    for (int q = 0; q < p * p; ++q) ws.S[q] += tm.edge_htt[e * p * p + q];
  for (int i = 0; i < p; ++i) ws.S[i * p + i] += r1[sx + i];  // :399-400
  for (int j = 0; j < p; ++j)                                  // :401
    for (int i = 0; i < p; ++i) {
      double dot = 0.0;
      for (int r = 0; r < kd; ++r)
        dot += ws.J[static_cast<size_t>(i) * kd + r] * ws.KinvJ[static_cast<size_t>(j) * kd + r];
      ws.S[j * p + i] -= dot;
    }
  // :403-407, Eigen::LLT on the lower triangle: fails on a pivot <= 0.
  ws.L = ws.S;
  for (int j = 0; j < p; ++j) {
    double d = ws.L[j * p + j];
    for (int k = 0; k < j; ++k) d -= ws.L[k * p + j] * ws.L[k * p + j];
    if (!(d > 0.0)) return false;
    const double l = std::sqrt(d);
    ws.L[j * p + j] = l;
    for (int i = j + 1; i < p; ++i) {
      double v = ws.L[j * p + i];
      for (int k = 0; k < j; ++k) v -= ws.L[k * p + i] * ws.L[k * p + j];
      ws.L[j * p + i] = v / l;
    }
  }
  return true;
}

// helpers.cpp:896-951.
void theta_solve(const CompiledTree &ct, const KktLayout &K, int p, const KktModel &m,
                 const double *b, double *sol, ThetaWorkspace &ws) {
  const int sx = K.x_dim, kd = K.kkt_dim, yz = K.y_dim + K.z_dim;
  ws.rhs.assign(kd, 0.0);
  ws.sol.assign(kd, 0.0);
  for (int r = 0; r < sx; ++r) ws.rhs[r] = b[r];
  for (int r = 0; r < yz; ++r) ws.rhs[sx + r] = b[sx + p + r];
  kkt_solve(ct, K, m, ws.rhs.data(), ws.sol.data(), ws.kkt);  // :921
  ws.t.assign(p, 0.0);
  for (int i = 0; i < p; ++i) {  // :927-929
    double dot = 0.0;
    for (int r = 0; r < kd; ++r) dot += ws.J[static_cast<size_t>(i) * kd + r] * ws.sol[r];
    ws.t[i] = b[sx + i] - dot;
  }
  for (int i = 0; i < p; ++i) {  // :933-934
    double s = ws.t[i];
    for (int k = 0; k < i; ++k) s -= ws.L[k * p + i] * ws.t[k];
    ws.t[i] = s / ws.L[i * p + i];
  }
  for (int i = p - 1; i >= 0; --i) {  // :935-937
    double s = ws.t[i];
    for (int k = i + 1; k < p; ++k) s -= ws.L[i * p + k] * ws.t[k];
    ws.t[i] = s / ws.L[i * p + i];
  }
  for (int r = 0; r < kd; ++r) {  // :939-942
    double s = 0.0;
    for (int j = 0; j < p; ++j) s += ws.KinvJ[static_cast<size_t>(j) * kd + r] * ws.t[j];
    ws.sol[r] -= s;
  }
  for (int r = 0; r < sx; ++r) sol[r] = ws.sol[r];  // :944-950
  for (int i = 0; i < p; ++i) sol[sx + i] = ws.t[i];
  for (int r = 0; r < yz; ++r) sol[sx + p + r] = ws.sol[sx + r];
}

// add_Kx_to_y (helpers.cpp:953-977) with the theta branches of :1019-1368.
void theta_apply(const CompiledTree &ct, const Tree &t, const KktLayout &K, int p,
                 const KktModel &m, const ThetaModel &tm, const double *w, const double *r1,
                 const double *r2, const double *r3, const double *x, double *y) {
  const ThetaOffsets o(K, t);
  const int E = t.num_edges, N = E + 1, sx = K.x_dim, xd = sx + p;
  const double *x_y = x + xd, *x_z = x_y + K.y_dim;
  double *y_y = y + xd, *y_z = y_y + K.y_dim;
  // stagewise blocks and the regularization of the stagewise rows
  kkt_apply(ct, K, m, w, r1, r2, r3, x, x_y, x_z, y, y_y, y_z);
  const double *th = x + sx;
  double *yt = y + sx;
  auto fwd = [&](const double *blk, int rows, double *out) {  // out += blk theta
    for (int j = 0; j < p; ++j)
      for (int r = 0; r < rows; ++r) out[r] += blk[j * rows + r] * th[j];
  };
  auto adj = [&](const double *blk, int rows, const double *in) {  // y_theta += blk' in
    for (int j = 0; j < p; ++j)
      for (int r = 0; r < rows; ++r) yt[j] += blk[j * rows + r] * in[r];
  };
  for (int i = 0; i < N; ++i) {
    const int n = K.lqr.n[i];
    fwd(tm.node_hxt + o.n_off[i] * p, n, y + K.x_state[i]);                 // H, :1029-1041
    adj(tm.node_hxt + o.n_off[i] * p, n, x + K.x_state[i]);
    fwd(tm.node_htt + i * p * p, p, yt);
    fwd(tm.node_jct + o.nc_off[i] * p, K.node_c[i], y_y + K.y_node_c[i]);   // C
    adj(tm.node_jct + o.nc_off[i] * p, K.node_c[i], x_y + K.y_node_c[i]);   // C'
    fwd(tm.node_jgt + o.ng_off[i] * p, K.node_g[i], y_z + K.z_node[i]);     // G
    adj(tm.node_jgt + o.ng_off[i] * p, K.node_g[i], x_z + K.z_node[i]);     // G'
  }
  for (int e = 0; e < E; ++e) {
    const int par = t.parents[e], ch = t.children[e];
    fwd(tm.edge_hxt + o.pn_off[e] * p, K.lqr.n[par], y + K.x_state[par]);   // H, :1043-1067
    adj(tm.edge_hxt + o.pn_off[e] * p, K.lqr.n[par], x + K.x_state[par]);
    fwd(tm.edge_hut + o.m_off[e] * p, K.lqr.m[e], y + K.x_control[e]);
    adj(tm.edge_hut + o.m_off[e] * p, K.lqr.m[e], x + K.x_control[e]);
    fwd(tm.edge_htt + e * p * p, p, yt);
    fwd(tm.edge_dynt + o.cn_off[e] * p, K.lqr.n[ch], y_y + K.y_dyn[ch]);    // C
    adj(tm.edge_dynt + o.cn_off[e] * p, K.lqr.n[ch], x_y + K.y_dyn[ch]);    // C'
    fwd(tm.edge_jct + o.ec_off[e] * p, K.edge_c[e], y_y + K.y_edge_c[e]);
    adj(tm.edge_jct + o.ec_off[e] * p, K.edge_c[e], x_y + K.y_edge_c[e]);
    fwd(tm.edge_jgt + o.eg_off[e] * p, K.edge_g[e], y_z + K.z_edge[e]);     // G
    adj(tm.edge_jgt + o.eg_off[e] * p, K.edge_g[e], x_z + K.z_edge[e]);     // G'
  }
  for (int i = 0; i < p; ++i) yt[i] += r1[sx + i] * th[i];  // :966-968 on the theta rows
}

}  // namespace sipoc_oracle
