// Batched model-callback scatter (sip_optimal_control.cpp:44-123).
//
// Pure data movement: per problem ~ (x_dim + y_dim + z_dim) doubles read and written once.
// Threads run along the batch (the unit-stride dimension of the engine layout, so every
// access of a warp is one 256-byte row segment); blockIdx.y picks a node, an edge, or the
// one serial item (the objective sum and the theta gradient, whose terms the reference adds
// in node order, then edge order -- kept, so the sums are bit-identical).
#include "model_scatter.cuh"

namespace sipoc {
namespace {

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads)
model_scatter_kernel(DevTables t, ModelValues v, const double *__restrict__ x,
                     const double *__restrict__ x0, bool new_x, double *__restrict__ f,
                     double *__restrict__ grad, double *__restrict__ c, double *__restrict__ g,
                     int64_t batch, int64_t ld) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (b >= batch) return;
  const size_t L = static_cast<size_t>(ld);
  auto at = [&](const double *a, int flat) { return __ldg(a + static_cast<size_t>(flat) * L + b); };
  auto put = [&](double *a, int flat, double val) { a[static_cast<size_t>(flat) * L + b] = val; };
  const int item = blockIdx.y, p = t.theta_dim;
  if (item < t.N) {
    if (!new_x) return;
    const int node = item, n = t.n[node];
    // gradient rows of the node's state: its own df_dx, then df_dx of every edge leaving it,
    // in edge order (:55-62, :67-76)
    for (int row = 0; row < n; ++row) {
      double acc = at(v.node_df_dx, t.n_off[node] + row);
      for (int ce = t.child_offsets[node]; ce < t.child_offsets[node + 1]; ++ce)
        acc += at(v.edge_df_dx, t.pn_off[t.child_edges[ce]] + row);
      put(grad, t.x_state[node] + row, acc);
    }
    if (node == t.root)  // initial-state residual (:88-94)
      for (int row = 0; row < n; ++row)
        put(c, t.y_dyn[node] + row, at(x0, row) - at(x, t.x_state[node] + row));
    for (int r = 0; r < t.node_c[node]; ++r)  // :95-99
      put(c, t.y_node_c[node] + r, at(v.node_c, t.node_c_off[node] + r));
    for (int r = 0; r < t.node_g[node]; ++r)  // :112-116
      put(g, t.z_node[node] + r, at(v.node_g, t.node_g_off[node] + r));
  } else if (item < t.N + t.E) {
    if (!new_x) return;
    const int e = item - t.N, child = t.children[e];
    for (int row = 0; row < t.m[e]; ++row)  // :77-80
      put(grad, t.x_control[e] + row, at(v.edge_df_du, t.m_off[e] + row));
    for (int r = 0; r < t.n[child]; ++r)  // :100-104
      put(c, t.y_dyn[child] + r, at(v.edge_dyn_res, t.cn_off[e] + r));
    for (int r = 0; r < t.edge_c[e]; ++r)  // :105-107
      put(c, t.y_edge_c[e] + r, at(v.edge_c, t.edge_c_off[e] + r));
    for (int r = 0; r < t.edge_g[e]; ++r)  // :117-121
      put(g, t.z_edge[e] + r, at(v.edge_g, t.edge_g_off[e] + r));
  } else {
    double acc = 0.0;  // :44-50
    for (int node = 0; node < t.N; ++node) acc += at(v.node_f, node);
    for (int e = 0; e < t.E; ++e) acc += at(v.edge_f, e);
    f[b] = acc;
    if (!new_x) return;
    for (int row = 0; row < p; ++row) {  // :63-65, :81-83
      double a = 0.0;
      for (int node = 0; node < t.N; ++node) a += at(v.node_df_dtheta, node * p + row);
      for (int e = 0; e < t.E; ++e) a += at(v.edge_df_dtheta, e * p + row);
      put(grad, t.sx_dim + row, a);
    }
  }
}

}  // namespace

int launch_model_scatter(const DevTables &t, const ModelValues &v, const double *x,
                         const double *initial_state, bool new_x, double *f, double *gradient_f,
                         double *c, double *g, int64_t batch, int64_t ld, cudaStream_t s) {
  const dim3 grid(static_cast<unsigned>((batch + kThreads - 1) / kThreads),
                  static_cast<unsigned>(t.N + t.E + 1));
  model_scatter_kernel<<<grid, kThreads, 0, s>>>(t, v, x, initial_state, new_x, f, gradient_f, c,
                                                 g, batch, ld);
  return 1;
}

}  // namespace sipoc
