// See types.hpp.  Memory bookkeeping of the reference's model / workspace views and the
// structural checks; nothing here computes.
//
// Every view type lists its blocks ONCE (the *_blocks functions below, in the reference's
// arena order); reserve, free, mem_assign and num_bytes are four policies applied to that
// list, so the three allocation modes cannot drift apart.
#include "types.hpp"

#include <algorithm>
#include <cmath>
#include <limits>

#include "../../include/sipoc.h"

namespace sip::optimal_control {

namespace {

struct Allocate {
  void operator()(double *&b, int count) const { b = new double[std::max(count, 1)](); }
  void operator()(int *&b, int count) const { b = new int[std::max(count, 1)](); }
  void operator()(double **&b, int count) const { b = new double *[std::max(count, 1)](); }
};
struct Release {
  template <class T>
  void operator()(T *&b, int) const {
    delete[] b;
    b = nullptr;
  }
};
struct Carve {
  unsigned char *cursor;
  template <class T>
  void operator()(T *&b, int count) {
    b = reinterpret_cast<T *>(cursor);
    cursor += static_cast<size_t>(count) * sizeof(T);
  }
};
struct Count {
  long long bytes = 0;
  template <class T>
  void operator()(T *&, int count) {
    bytes += static_cast<long long>(count) * static_cast<long long>(sizeof(T));
  }
};

template <class Fn>
void node_blocks(NodeModelCallbackOutput &o, int n, int c, int g, int p, Fn &&fn) {
  fn(o.df_dx, n);
  fn(o.df_dtheta, p);
  fn(o.c, c);
  fn(o.dc_dx, c * n);
  fn(o.dc_dtheta, c * p);
  fn(o.g, g);
  fn(o.dg_dx, g * n);
  fn(o.dg_dtheta, g * p);
  fn(o.d2L_dx2, n * n);
  fn(o.d2L_dxdtheta, n * p);
  fn(o.d2L_dtheta2, p * p);
}

template <class Fn>
void edge_blocks(EdgeModelCallbackOutput &o, int np, int nc, int m, int c, int g, int p, Fn &&fn) {
  fn(o.df_dx, np);
  fn(o.df_du, m);
  fn(o.df_dtheta, p);
  fn(o.dyn_res, nc);
  fn(o.ddyn_dx, nc * np);
  fn(o.ddyn_du, nc * m);
  fn(o.ddyn_dtheta, nc * p);
  fn(o.c, c);
  fn(o.dc_dx, c * np);
  fn(o.dc_du, c * m);
  fn(o.dc_dtheta, c * p);
  fn(o.g, g);
  fn(o.dg_dx, g * np);
  fn(o.dg_du, g * m);
  fn(o.dg_dtheta, g * p);
  fn(o.d2L_dx2, np * np);
  fn(o.d2L_dxdu, np * m);
  fn(o.d2L_du2, m * m);
  fn(o.d2L_dxdtheta, np * p);
  fn(o.d2L_dudtheta, m * p);
  fn(o.d2L_dtheta2, p * p);
}

template <class Fn>
void output_blocks(ModelCallbackOutput &mco, const Dimensions &d, const Topology &t, Fn &&fn) {
  const int p = d.theta_dim;
  for (int i = 0; i < t.num_nodes(); ++i)
    node_blocks(mco.nodes[i], d.get_state_dim(i), d.get_node_c_dim(i), d.get_node_g_dim(i), p, fn);
  for (int e = 0; e < t.num_edges; ++e)
    edge_blocks(mco.edges[e], d.get_state_dim(t.edge_parents[e]),
                d.get_state_dim(t.edge_children[e]), d.get_control_dim(e), d.get_edge_c_dim(e),
                d.get_edge_g_dim(e), p, fn);
}

int align_up(long long bytes) {
  constexpr long long a = alignof(std::max_align_t);
  return static_cast<int>((bytes + a - 1) / a * a);
}

sipoc_structure describe(const Dimensions &d, const Topology &t) {
  sipoc_structure s{};
  s.num_edges = t.num_edges;
  s.root = t.root;
  s.edge_parents = t.edge_parents;
  s.edge_children = t.edge_children;
  s.state_dims = d.state_dims;
  s.control_dims = d.control_dims;
  s.node_c_dims = d.node_c_dims;
  s.node_g_dims = d.node_g_dims;
  s.edge_c_dims = d.edge_c_dims;
  s.edge_g_dims = d.edge_g_dims;
  s.theta_dim = d.theta_dim;
  s.batch = 1;
  s.device = -1;
  return s;
}

}  // namespace

// ---- ModelCallbackInput (types.cpp:136-158) ---------------------------------------------
void ModelCallbackInput::reserve(const Topology &topology) {
  theta = nullptr;
  nodes = new NodeModelCallbackInput[topology.num_nodes()]();
  edges = new EdgeModelCallbackInput[std::max(topology.num_edges, 1)]();
}
void ModelCallbackInput::free() {
  delete[] nodes;
  delete[] edges;
  nodes = nullptr;
  edges = nullptr;
}
auto ModelCallbackInput::mem_assign(const Topology &topology, unsigned char *mem_ptr) -> int {
  theta = nullptr;
  nodes = reinterpret_cast<NodeModelCallbackInput *>(mem_ptr);
  edges = reinterpret_cast<EdgeModelCallbackInput *>(nodes + topology.num_nodes());
  return num_bytes(topology.num_edges);
}

// ---- ModelCallbackOutput (types.cpp:160-383) --------------------------------------------
void ModelCallbackOutput::reserve(const Dimensions &d, const Topology &t) {
  nodes = new NodeModelCallbackOutput[t.num_nodes()]();
  edges = new EdgeModelCallbackOutput[std::max(t.num_edges, 1)]();
  output_blocks(*this, d, t, Allocate{});
}
void ModelCallbackOutput::free(const Topology &t) {
  if (nodes != nullptr && edges != nullptr) {
    // sizes are irrelevant to delete[]: walk with a dimension-free description
    const std::vector<int> ones(static_cast<size_t>(t.num_edges) + 1, 1);
    const Dimensions d{0, ones.data(), ones.data(), nullptr, nullptr, nullptr, nullptr};
    output_blocks(*this, d, t, Release{});
  }
  delete[] nodes;
  delete[] edges;
  nodes = nullptr;
  edges = nullptr;
}
auto ModelCallbackOutput::mem_assign(const Dimensions &d, const Topology &t,
                                     unsigned char *mem_ptr) -> int {
  nodes = reinterpret_cast<NodeModelCallbackOutput *>(mem_ptr);
  edges = reinterpret_cast<EdgeModelCallbackOutput *>(nodes + t.num_nodes());
  Carve carve{reinterpret_cast<unsigned char *>(edges + t.num_edges)};
  output_blocks(*this, d, t, carve);
  return static_cast<int>(carve.cursor - mem_ptr);
}
auto ModelCallbackOutput::num_bytes(const Dimensions &d, const Topology &t) -> int {
  long long doubles = 0;
  const int p = d.theta_dim;
  for (int i = 0; i < t.num_nodes(); ++i)
    doubles += node_output_doubles(d.get_state_dim(i), d.get_node_c_dim(i), d.get_node_g_dim(i), p);
  for (int e = 0; e < t.num_edges; ++e)
    doubles += edge_output_doubles(d.get_state_dim(t.edge_parents[e]),
                                   d.get_state_dim(t.edge_children[e]), d.get_control_dim(e),
                                   d.get_edge_c_dim(e), d.get_edge_g_dim(e), p);
  return static_cast<int>(t.num_nodes() * sizeof(NodeModelCallbackOutput) +
                          t.num_edges * sizeof(EdgeModelCallbackOutput) +
                          doubles * sizeof(double));
}

// ---- Input ----------------------------------------------------------------------------
auto Input::num_bound_sides() const -> int {
  // one side per finite bound (the quantity sip::num_bound_sides reports)
  const int x_dim = dimensions.get_x_dim(topology.num_edges);
  int sides = 0;
  for (int i = 0; i < x_dim; ++i) {
    if (lower_bounds != nullptr && std::isfinite(lower_bounds[i])) ++sides;
    if (upper_bounds != nullptr && std::isfinite(upper_bounds[i])) ++sides;
  }
  return sides;
}

// The engine's structural checks are the reference's (types.cpp:68-134): dimensions first,
// then the tree.  No device is involved.
auto validate_input(const Dimensions &dimensions, const Topology &topology)
    -> InputValidationStatus {
  const sipoc_structure s = describe(dimensions, topology);
  switch (sipoc_validate(&s)) {
    case SIPOC_OK: return InputValidationStatus::SUCCESS;
    case SIPOC_INVALID_TOPOLOGY: return InputValidationStatus::INVALID_TOPOLOGY;
    default: return InputValidationStatus::INVALID_DIMENSIONS;
  }
}

// ---- Workspace::RegularizedLQRData (types.cpp:385-600) -----------------------------------
namespace {
// Pointer tables, then per-node and per-edge blocks, then the theta arrays and the scratch.
template <class Fn>
void lqr_data_tables(Workspace::RegularizedLQRData &r, int E, Fn &&fn) {
  const int N = E + 1;
  fn(r.node_mod_w_inv, N);
  fn(r.edge_mod_w_inv, E);
  fn(r.Q_mod, N);
  fn(r.M_mod, E);
  fn(r.R_mod, E);
  fn(r.q_mod, N);
  fn(r.r_mod, E);
  fn(r.c_mod, N);
  fn(r.dyn_r2, N);
  fn(r.node_c_r2_inv, N);
  fn(r.edge_c_r2_inv, E);
}
template <class Fn>
void lqr_data_blocks(Workspace::RegularizedLQRData &r, const Dimensions &d, int E, Fn &&fn) {
  const int N = E + 1, nmax = d.max_state_dim(N), p = d.theta_dim;
  for (int i = 0; i < N; ++i) {
    const int n = d.get_state_dim(i);
    fn(r.node_mod_w_inv[i], d.get_node_g_dim(i));
    fn(r.Q_mod[i], n * n);
    fn(r.q_mod[i], n);
    fn(r.c_mod[i], n);
    fn(r.dyn_r2[i], n);
    fn(r.node_c_r2_inv[i], d.get_node_c_dim(i));
  }
  for (int e = 0; e < E; ++e) {
    const int m = d.get_control_dim(e);
    fn(r.edge_mod_w_inv[e], d.get_edge_g_dim(e));
    fn(r.M_mod[e], nmax * m);  // sized for the largest state, as in the reference
    fn(r.R_mod[e], m * m);
    fn(r.r_mod[e], m);
    fn(r.edge_c_r2_inv[e], d.get_edge_c_dim(e));
  }
  const int kkt = d.get_stagewise_kkt_dim(E);
  fn(r.theta_jacobian, p > 0 ? kkt * p : 0);
  fn(r.theta_solution, p > 0 ? kkt * p : 0);
  fn(r.theta_schur, p * p);
  fn(r.theta_schur_factor, p * p);
  fn(r.theta_rhs, p);
  fn(r.theta_stagewise_rhs, p > 0 ? kkt : 0);
  fn(r.stagewise_scratch, 2 * nmax * (p > 0 ? p : 1));
}
}  // namespace

void Workspace::RegularizedLQRData::reserve(const Dimensions &d, int E) {
  lqr_data_tables(*this, E, Allocate{});
  lqr_data_blocks(*this, d, E, Allocate{});
}
void Workspace::RegularizedLQRData::free(int E) {
  if (Q_mod == nullptr) return;
  const std::vector<int> ones(static_cast<size_t>(E) + 1, 1);
  const Dimensions d{0, ones.data(), ones.data(), nullptr, nullptr, nullptr, nullptr};
  lqr_data_blocks(*this, d, E, Release{});
  lqr_data_tables(*this, E, Release{});
}
auto Workspace::RegularizedLQRData::mem_assign(const Dimensions &d, int E,
                                               unsigned char *mem_ptr) -> int {
  Carve carve{mem_ptr};
  lqr_data_tables(*this, E, carve);
  lqr_data_blocks(*this, d, E, carve);
  return static_cast<int>(carve.cursor - mem_ptr);
}
auto Workspace::RegularizedLQRData::num_bytes(const Dimensions &d, int E) -> int {
  // lqr_data_blocks dereferences the tables, so the arena is measured on a scratch copy
  RegularizedLQRData probe{};
  std::vector<double *> slots(static_cast<size_t>(11) * (E + 1), nullptr);
  double **cursor = slots.data();
  lqr_data_tables(probe, E, [&](double **&table, int count) {
    table = cursor;
    cursor += count;
  });
  Count count;
  lqr_data_blocks(probe, d, E, count);
  return static_cast<int>(count.bytes + (6LL * (E + 1) + 5LL * E) * sizeof(double *));
}

// ---- Workspace (types.cpp:24-64, 602-757) -------------------------------------------------
namespace {
template <class Fn>
void workspace_vectors(Workspace &w, const Dimensions &d, int E, Fn &&fn) {
  const int N = E + 1;
  fn(w.gradient_f, d.get_x_dim(E));
  fn(w.c, d.get_y_dim(E));
  fn(w.g, d.get_z_dim(E));
  fn(w.x_state_offsets, N);
  fn(w.x_control_offsets, E);
  fn(w.y_dyn_offsets, N);
  fn(w.y_node_c_offsets, N);
  fn(w.y_edge_c_offsets, E);
  fn(w.z_node_offsets, N);
  fn(w.z_edge_offsets, E);
  fn(w.ddyn_dx, E);
  fn(w.ddyn_du, E);
}
}  // namespace

void populate_workspace_metadata(Workspace &w, const Dimensions &d, const Topology &t) {
  const int E = t.num_edges, N = E + 1;
  w.stagewise_x_dim = d.get_stagewise_x_dim(E);
  w.x_dim = d.get_x_dim(E);
  w.y_dim = d.get_y_dim(E);
  w.z_dim = d.get_z_dim(E);
  w.stagewise_kkt_dim = d.get_stagewise_kkt_dim(E);
  // x: state i then control i, by index; y: per node dynamics then node equalities, edge
  // equalities after all nodes; z: node inequalities then edge inequalities.
  int x = 0, y = 0, z = 0;
  for (int i = 0; i < N; ++i) {
    w.x_state_offsets[i] = x;
    x += i < E ? d.get_state_dim(i) : 0;
    if (i < E) {
      w.x_control_offsets[i] = x;
      x += d.get_control_dim(i);
    }
    w.y_dyn_offsets[i] = y;
    y += d.get_state_dim(i);
    w.y_node_c_offsets[i] = y;
    y += d.get_node_c_dim(i);
    w.z_node_offsets[i] = z;
    z += d.get_node_g_dim(i);
  }
  for (int e = 0; e < E; ++e) {
    w.y_edge_c_offsets[e] = y;
    y += d.get_edge_c_dim(e);
    w.z_edge_offsets[e] = z;
    z += d.get_edge_g_dim(e);
    w.ddyn_dx[e] = w.model_callback_output.edges[e].ddyn_dx;
    w.ddyn_du[e] = w.model_callback_output.edges[e].ddyn_du;
  }
}

void Workspace::reserve(const Dimensions &d, const Topology &t) {
  model_callback_input.reserve(t);
  model_callback_output.reserve(d, t);
  workspace_vectors(*this, d, t.num_edges, Allocate{});
  populate_workspace_metadata(*this, d, t);
  lqr_workspace.reserve(d, t);
  lqr_output.reserve(t.num_edges);
  regularized_lqr_data.reserve(d, t.num_edges);
}

void Workspace::free(const Topology &t) {
  delete staging;
  staging = nullptr;
  model_callback_input.free();
  model_callback_output.free(t);
  const std::vector<int> ones(static_cast<size_t>(t.num_edges) + 1, 1);
  const Dimensions d{0, ones.data(), ones.data(), nullptr, nullptr, nullptr, nullptr};
  workspace_vectors(*this, d, t.num_edges, Release{});
  lqr_workspace.free(t.num_edges);
  lqr_output.free();
  regularized_lqr_data.free(t.num_edges);
}

auto Workspace::mem_assign(const Dimensions &d, const Topology &t, unsigned char *mem_ptr) -> int {
  const int E = t.num_edges;
  int used = model_callback_input.mem_assign(t, mem_ptr);
  used += model_callback_output.mem_assign(d, t, mem_ptr + used);
  Carve carve{mem_ptr + used};
  workspace_vectors(*this, d, E, carve);
  used = align_up(carve.cursor - mem_ptr);
  populate_workspace_metadata(*this, d, t);
  used = align_up(used + lqr_workspace.mem_assign(d, t, mem_ptr + used));
  used = align_up(used + lqr_output.mem_assign(E, mem_ptr + used));
  used = align_up(used + regularized_lqr_data.mem_assign(d, E, mem_ptr + used));
  return used;
}

auto Workspace::num_bytes(const Dimensions &d, const Topology &t) -> int {
  const int E = t.num_edges;
  long long total = ModelCallbackInput::num_bytes(E) + ModelCallbackOutput::num_bytes(d, t);
  total += static_cast<long long>(d.get_x_dim(E) + d.get_y_dim(E) + d.get_z_dim(E)) * sizeof(double);
  total += (7LL * E + 4) * sizeof(int) + 2LL * E * sizeof(double *);
  total = align_up(total);
  total = align_up(total + LQR::Workspace::num_bytes(d, t));
  total = align_up(total + LQR::Output::num_bytes(E));
  total = align_up(total + RegularizedLQRData::num_bytes(d, E));
  return static_cast<int>(total);
}

#ifdef SIPOC_HAVE_SIP
void Workspace::reserve(const Dimensions &d, const Topology &t, int num_bound_sides,
                        const sip::Settings &settings) {
  reserve(d, t);
  const int E = t.num_edges;
  sip_workspace.reserve(d.get_x_dim(E), d.get_z_dim(E), d.get_y_dim(E), num_bound_sides, settings);
}
auto Workspace::mem_assign(const Dimensions &d, const Topology &t, int num_bound_sides,
                           const sip::Settings &settings, unsigned char *mem_ptr) -> int {
  const int E = t.num_edges;
  int used = mem_assign(d, t, mem_ptr);
  used += sip_workspace.mem_assign(d.get_x_dim(E), d.get_z_dim(E), d.get_y_dim(E),
                                   num_bound_sides, settings, mem_ptr + used);
  return used;
}
auto Workspace::num_bytes(const Dimensions &d, const Topology &t, int num_bound_sides,
                          const sip::Settings &settings) -> int {
  const int E = t.num_edges;
  return num_bytes(d, t) + sip::Workspace::num_bytes(d.get_x_dim(E), d.get_z_dim(E),
                                                     d.get_y_dim(E), num_bound_sides, settings);
}
#endif

}  // namespace sip::optimal_control
