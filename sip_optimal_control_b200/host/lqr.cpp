// See lqr.hpp.  Nothing here computes: the classes gather the caller's per-node /
// per-edge blocks into the flat arrays of the C ABI and forward the calls.
#include "lqr.hpp"

#include <algorithm>

#include "../../include/sipoc.h"

namespace sip::optimal_control {

// ---- Topology (lqr.cpp:12-60) ------------------------------------------------------
int Topology::num_nodes() const { return num_edges + 1; }

void Topology::reserve(int edges) {
  free();
  num_edges = edges;
  owned_parents_ = new int[std::max(edges, 1)]();
  owned_children_ = new int[std::max(edges, 1)]();
  edge_parents = owned_parents_;
  edge_children = owned_children_;
}

void Topology::free() {
  delete[] owned_parents_;
  delete[] owned_children_;
  owned_parents_ = owned_children_ = nullptr;
  edge_parents = edge_children = nullptr;
}

void Topology::set_chain() {
  if (owned_parents_ == nullptr) reserve(num_edges);
  root = 0;
  for (int e = 0; e < num_edges; ++e) {
    owned_parents_[e] = e;
    owned_children_[e] = e + 1;
  }
}

void Topology::set_tree(int new_root, const int *parents, const int *children) {
  if (owned_parents_ == nullptr) reserve(num_edges);
  root = new_root;
  std::copy(parents, parents + num_edges, owned_parents_);
  std::copy(children, children + num_edges, owned_children_);
}

// ---- Dimensions (lqr.cpp:62-180) -----------------------------------------------------
void Dimensions::reserve(int num_edges) {
  free();
  const int counts[6] = {num_edges + 1, num_edges, num_edges + 1, num_edges + 1, num_edges,
                         num_edges};
  for (int i = 0; i < 6; ++i) owned_[i] = new int[std::max(counts[i], 1)]();
  state_dims = owned_[0];
  control_dims = owned_[1];
  node_c_dims = owned_[2];
  node_g_dims = owned_[3];
  edge_c_dims = owned_[4];
  edge_g_dims = owned_[5];
}

void Dimensions::free() {
  for (int *&p : owned_) {
    delete[] p;
    p = nullptr;
  }
  state_dims = control_dims = node_c_dims = node_g_dims = edge_c_dims = edge_g_dims = nullptr;
}

void Dimensions::set_uniform(int num_edges, int state_dim, int control_dim, int node_c_dim,
                             int node_g_dim, int edge_c_dim, int edge_g_dim, int theta) {
  if (owned_[0] == nullptr) reserve(num_edges);
  theta_dim = theta;
  std::fill(owned_[0], owned_[0] + num_edges + 1, state_dim);
  std::fill(owned_[1], owned_[1] + num_edges, control_dim);
  std::fill(owned_[2], owned_[2] + num_edges + 1, node_c_dim);
  std::fill(owned_[3], owned_[3] + num_edges + 1, node_g_dim);
  std::fill(owned_[4], owned_[4] + num_edges, edge_c_dim);
  std::fill(owned_[5], owned_[5] + num_edges, edge_g_dim);
}

int Dimensions::get_schur_dim() const { return theta_dim; }
int Dimensions::get_state_dim(int node) const { return state_dims[node]; }
int Dimensions::get_control_dim(int edge) const { return control_dims[edge]; }
// A null constraint-dimension array means "all zero" (lqr.cpp:98-112).
int Dimensions::get_node_c_dim(int node) const { return node_c_dims ? node_c_dims[node] : 0; }
int Dimensions::get_node_g_dim(int node) const { return node_g_dims ? node_g_dims[node] : 0; }
int Dimensions::get_edge_c_dim(int edge) const { return edge_c_dims ? edge_c_dims[edge] : 0; }
int Dimensions::get_edge_g_dim(int edge) const { return edge_g_dims ? edge_g_dims[edge] : 0; }

int Dimensions::get_stagewise_x_dim(int num_edges) const {
  int total = state_dims[num_edges];
  for (int e = 0; e < num_edges; ++e) total += state_dims[e] + control_dims[e];
  return total;
}
int Dimensions::get_x_dim(int num_edges) const { return get_stagewise_x_dim(num_edges) + theta_dim; }
int Dimensions::get_y_dim(int num_edges) const {
  int total = 0;
  for (int i = 0; i <= num_edges; ++i) total += state_dims[i] + get_node_c_dim(i);
  for (int e = 0; e < num_edges; ++e) total += get_edge_c_dim(e);
  return total;
}
int Dimensions::get_z_dim(int num_edges) const {
  int total = 0;
  for (int i = 0; i <= num_edges; ++i) total += get_node_g_dim(i);
  for (int e = 0; e < num_edges; ++e) total += get_edge_g_dim(e);
  return total;
}
int Dimensions::get_stagewise_kkt_dim(int num_edges) const {
  return get_stagewise_x_dim(num_edges) + get_y_dim(num_edges) + get_z_dim(num_edges);
}

// ---- LQR::Output / Workspace ---------------------------------------------------------
void LQR::Output::reserve(int num_edges) {
  x = new double *[num_edges + 1]();
  u = new double *[std::max(num_edges, 1)]();
  y = new double *[num_edges + 1]();
}
void LQR::Output::free() {
  delete[] x;
  delete[] u;
  delete[] y;
  x = u = y = nullptr;
}

void LQR::Workspace::reserve(int, int, int) {}  // sized when the engine is created
void LQR::Workspace::reserve(const Dimensions &, const Topology &) {}
void LQR::Workspace::free(int) {
  if (engine != nullptr) sipoc_destroy(engine);
  engine = nullptr;
  for (auto &v : in) v.clear();
  for (auto &v : out) v.clear();
}

namespace {

sipoc_structure describe(const Dimensions &d, const Topology &t, int64_t batch, int device) {
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
  s.batch = batch;
  s.device = device;
  s.flags = 0;
  return s;
}

// Blocks of one array, in node / edge order, appended to `flat`.
void gather(double **table, int count, const std::vector<int> &elems, std::vector<double> &flat) {
  size_t o = 0;
  for (int i = 0; i < count; ++i) {
    if (elems[i] == 0) continue;
    std::copy(table[i], table[i] + elems[i], flat.begin() + o);
    o += elems[i];
  }
}
void scatter(const std::vector<double> &flat, double **table, int count,
             const std::vector<int> &elems) {
  size_t o = 0;
  for (int i = 0; i < count; ++i) {
    if (elems[i] == 0) continue;
    std::copy(flat.begin() + o, flat.begin() + o + elems[i], table[i]);
    o += elems[i];
  }
}

struct BlockSizes {
  std::vector<int> nn, n, nm, mm, m, a, b;
  BlockSizes(const Dimensions &d, const Topology &t) {
    const int E = t.num_edges;
    for (int i = 0; i <= E; ++i) {
      nn.push_back(d.state_dims[i] * d.state_dims[i]);
      n.push_back(d.state_dims[i]);
    }
    for (int e = 0; e < E; ++e) {
      const int np = d.state_dims[t.edge_parents[e]], nc = d.state_dims[t.edge_children[e]],
                mm_ = d.control_dims[e];
      nm.push_back(np * mm_);
      mm.push_back(mm_ * mm_);
      m.push_back(mm_);
      a.push_back(nc * np);
      b.push_back(nc * mm_);
    }
  }
};

}  // namespace

// ---- LQR (lqr.cpp:635-871) -------------------------------------------------------------
LQR::LQR(const Input &data, Workspace &workspace)
    : input_(data), workspace_(workspace), traversal_status_(compile_topology()) {}

auto LQR::compile_topology() -> FactorStatus {
  if (workspace_.engine != nullptr) {
    sipoc_destroy(workspace_.engine);
    workspace_.engine = nullptr;
  }
  const sipoc_structure s = describe(input_.dimensions, input_.topology, 1, -1);
  const sipoc_error rc = sipoc_create(&s, &workspace_.engine);
  if (rc != SIPOC_OK) {
    workspace_.engine = nullptr;
    // An invalid tree is latched and reported by every later factor (lqr.cpp:646-648).
    return FactorStatus::INVALID_TOPOLOGY;
  }
  sipoc_lqr_sizes z{};
  sipoc_lqr_get_sizes(workspace_.engine, &z);
  const int64_t sizes[9] = {z.Q, z.M, z.R, z.q, z.r, z.A, z.B, z.c, z.delta};
  for (int i = 0; i < 9; ++i) workspace_.in[i].assign(static_cast<size_t>(sizes[i]), 0.0);
  workspace_.out[0].assign(static_cast<size_t>(z.x), 0.0);
  workspace_.out[1].assign(static_cast<size_t>(z.u), 0.0);
  workspace_.out[2].assign(static_cast<size_t>(z.y), 0.0);
  return FactorStatus::SUCCESS;
}

LQR::FactorStatus LQR::factor_with_status() {
  if (traversal_status_ != FactorStatus::SUCCESS) return traversal_status_;
  const Topology &t = input_.topology;
  const BlockSizes bs(input_.dimensions, t);
  const int N = t.num_nodes(), E = t.num_edges;
  auto &w = workspace_;
  gather(input_.Q, N, bs.nn, w.in[0]);
  gather(input_.M, E, bs.nm, w.in[1]);
  gather(input_.R, E, bs.mm, w.in[2]);
  gather(input_.A, E, bs.a, w.in[5]);
  gather(input_.B, E, bs.b, w.in[6]);
  gather(input_.delta, N, bs.n, w.in[8]);
  const sipoc_lqr_input in{w.in[0].data(), w.in[1].data(), w.in[2].data(), nullptr, nullptr,
                           w.in[5].data(), w.in[6].data(), nullptr,        w.in[8].data()};
  int status = 0;
  if (sipoc_lqr_factor_host(w.engine, &in, &status) != SIPOC_OK)
    return FactorStatus::INVALID_TOPOLOGY;
  return static_cast<FactorStatus>(status);
}

bool LQR::factor() { return factor_with_status() == FactorStatus::SUCCESS; }

void LQR::solve(Output &output) {
  const Topology &t = input_.topology;
  const BlockSizes bs(input_.dimensions, t);
  const int N = t.num_nodes(), E = t.num_edges;
  auto &w = workspace_;
  gather(input_.q, N, bs.n, w.in[3]);
  gather(input_.r, E, bs.m, w.in[4]);
  gather(input_.c, N, bs.n, w.in[7]);
  const sipoc_lqr_input in{nullptr, nullptr, nullptr, w.in[3].data(), w.in[4].data(),
                           nullptr, nullptr, w.in[7].data(), nullptr};
  const sipoc_lqr_output out{w.out[0].data(), w.out[1].data(), w.out[2].data()};
  if (sipoc_lqr_solve_host(w.engine, &in, &out) != SIPOC_OK) return;
  scatter(w.out[0], output.x, N, bs.n);
  scatter(w.out[1], output.u, E, bs.m);
  scatter(w.out[2], output.y, N, bs.n);
}

// ---- BatchedLQR --------------------------------------------------------------------------
BatchedLQR::BatchedLQR(const Dimensions &d, const Topology &t, int64_t batch, int device) {
  const sipoc_structure s = describe(d, t, batch, device);
  if (sipoc_create(&s, &engine_) != SIPOC_OK) {
    engine_ = nullptr;
    status_ = LQR::FactorStatus::INVALID_TOPOLOGY;
  }
}
BatchedLQR::~BatchedLQR() {
  if (engine_ != nullptr) sipoc_destroy(engine_);
}
int64_t BatchedLQR::batch_stride() const { return sipoc_batch_stride(engine_); }

namespace {
sipoc_lqr_input to_abi(const BatchedLQR::DeviceInput &i) {
  return sipoc_lqr_input{i.Q, i.M, i.R, i.q, i.r, i.A, i.B, i.c, i.delta};
}
}  // namespace

bool BatchedLQR::factor_with_status(const DeviceInput &in, int *status, void *stream) {
  const sipoc_lqr_input a = to_abi(in);
  return engine_ != nullptr && sipoc_lqr_factor(engine_, &a, status, stream) == SIPOC_OK;
}
bool BatchedLQR::solve(const DeviceInput &in, const DeviceOutput &out, void *stream) {
  const sipoc_lqr_input a = to_abi(in);
  const sipoc_lqr_output o{out.x, out.u, out.y};
  return engine_ != nullptr && sipoc_lqr_solve(engine_, &a, &o, stream) == SIPOC_OK;
}
bool BatchedLQR::factor_solve(const DeviceInput &in, const DeviceOutput &out, int *status,
                              void *stream) {
  const sipoc_lqr_input a = to_abi(in);
  const sipoc_lqr_output o{out.x, out.u, out.y};
  return engine_ != nullptr &&
         sipoc_lqr_factor_solve(engine_, &a, &o, status, stream) == SIPOC_OK;
}

}  // namespace sip::optimal_control
