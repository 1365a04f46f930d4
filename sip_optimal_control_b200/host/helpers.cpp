// See helpers.hpp / types.hpp.  Marshalling only; every number is computed on the GPU.
#include "helpers.hpp"

#include <algorithm>

#include "../../include/sipoc.h"

namespace sip::optimal_control {

namespace {

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

double *block(int count) { return new double[std::max(count, 1)](); }

}  // namespace

// ---- ModelCallbackOutput (types.cpp:136-383, dynamic allocation mode only) -------------
void ModelCallbackOutput::reserve(const Dimensions &d, const Topology &t) {
  const int N = t.num_nodes(), E = t.num_edges;
  nodes = new NodeModelCallbackOutput[N]();
  edges = new EdgeModelCallbackOutput[std::max(E, 1)]();
  for (int i = 0; i < N; ++i) {
    const int n = d.get_state_dim(i), c = d.get_node_c_dim(i), g = d.get_node_g_dim(i);
    nodes[i] = NodeModelCallbackOutput{0.0,          block(n), block(c),    block(c * n),
                                       block(g),     block(g * n), block(n * n)};
  }
  for (int e = 0; e < E; ++e) {
    const int np = d.get_state_dim(t.edge_parents[e]), nc = d.get_state_dim(t.edge_children[e]);
    const int m = d.get_control_dim(e), c = d.get_edge_c_dim(e), g = d.get_edge_g_dim(e);
    edges[e] = EdgeModelCallbackOutput{0.0,           block(np),     block(m),      block(nc),
                                       block(nc * np), block(nc * m), block(c),      block(c * np),
                                       block(c * m),   block(g),      block(g * np), block(g * m),
                                       block(np * np), block(np * m), block(m * m)};
  }
}

void ModelCallbackOutput::free(const Topology &t) {
  if (nodes != nullptr) {
    for (int i = 0; i < t.num_nodes(); ++i) {
      NodeModelCallbackOutput &o = nodes[i];
      for (double *p : {o.df_dx, o.c, o.dc_dx, o.g, o.dg_dx, o.d2L_dx2}) delete[] p;
    }
  }
  if (edges != nullptr) {
    for (int e = 0; e < t.num_edges; ++e) {
      EdgeModelCallbackOutput &o = edges[e];
      for (double *p : {o.df_dx, o.df_du, o.dyn_res, o.ddyn_dx, o.ddyn_du, o.c, o.dc_dx, o.dc_du,
                        o.g, o.dg_dx, o.dg_du, o.d2L_dx2, o.d2L_dxdu, o.d2L_du2})
        delete[] p;
    }
  }
  delete[] nodes;
  delete[] edges;
  nodes = nullptr;
  edges = nullptr;
}

// The engine validates on creation with the reference's rules (types.cpp:68-134).
auto validate_input(const Dimensions &dimensions, const Topology &topology)
    -> InputValidationStatus {
  sipoc_engine *probe = nullptr;
  const sipoc_structure s = describe(dimensions, topology);
  const sipoc_error rc = sipoc_create(&s, &probe);
  if (probe != nullptr) sipoc_destroy(probe);
  if (rc == SIPOC_INVALID_DIMENSIONS) return InputValidationStatus::INVALID_DIMENSIONS;
  if (rc == SIPOC_INVALID_TOPOLOGY) return InputValidationStatus::INVALID_TOPOLOGY;
  return InputValidationStatus::SUCCESS;
}

void Workspace::reserve(const Dimensions &dimensions, const Topology &topology) {
  model_callback_output.reserve(dimensions, topology);
}
void Workspace::free(const Topology &topology) {
  model_callback_output.free(topology);
  lqr_workspace.free();
}

// ---- CallbackProvider (helpers.cpp:11-26, 242-370, 749-900, 953-977) -----------------------
CallbackProvider::CallbackProvider(const Input &input, Workspace &workspace)
    : input_(input), workspace_(workspace), input_is_valid_(false) {
  sipoc_engine *&engine = workspace_.lqr_workspace.engine;
  if (engine != nullptr) sipoc_destroy(engine);
  engine = nullptr;
  const sipoc_structure s = describe(input.dimensions, input.topology);
  input_is_valid_ = sipoc_create(&s, &engine) == SIPOC_OK;
  if (!input_is_valid_) {
    engine = nullptr;
    return;
  }
  sipoc_kkt_sizes z{};
  sipoc_kkt_get_sizes(engine, &z);
  const int64_t sizes[12] = {z.node_hxx, z.node_jc,  z.node_jg,  z.edge_hxx,
                             z.edge_hxu, z.edge_huu, z.edge_A,   z.edge_B,
                             z.edge_jcx, z.edge_jcu, z.edge_jgx, z.edge_jgu};
  for (int i = 0; i < 12; ++i)
    workspace_.model[i].assign(static_cast<size_t>(std::max<int64_t>(sizes[i], 1)), 0.0);
  for (auto &v : workspace_.vec) v.assign(static_cast<size_t>(std::max<int64_t>(z.kkt_dim, 1)), 0.0);
}

// Per-node / per-edge blocks -> the flat arrays of sipoc_kkt_model.
void CallbackProvider::gather_model() {
  const Dimensions &d = input_.dimensions;
  const Topology &t = input_.topology;
  const ModelCallbackOutput &mco = workspace_.model_callback_output;
  size_t o[12] = {0};
  auto put = [&](int which, const double *src, int count) {
    std::copy(src, src + count, workspace_.model[which].begin() + o[which]);
    o[which] += count;
  };
  for (int i = 0; i < t.num_nodes(); ++i) {
    const int n = d.get_state_dim(i);
    put(0, mco.nodes[i].d2L_dx2, n * n);
    put(1, mco.nodes[i].dc_dx, d.get_node_c_dim(i) * n);
    put(2, mco.nodes[i].dg_dx, d.get_node_g_dim(i) * n);
  }
  for (int e = 0; e < t.num_edges; ++e) {
    const int np = d.get_state_dim(t.edge_parents[e]), nc = d.get_state_dim(t.edge_children[e]);
    const int m = d.get_control_dim(e), c = d.get_edge_c_dim(e), g = d.get_edge_g_dim(e);
    const EdgeModelCallbackOutput &eo = mco.edges[e];
    put(3, eo.d2L_dx2, np * np);
    put(4, eo.d2L_dxdu, np * m);
    put(5, eo.d2L_du2, m * m);
    put(6, eo.ddyn_dx, nc * np);
    put(7, eo.ddyn_du, nc * m);
    put(8, eo.dc_dx, c * np);
    put(9, eo.dc_du, c * m);
    put(10, eo.dg_dx, g * np);
    put(11, eo.dg_du, g * m);
  }
}

bool CallbackProvider::factor(const double *w, const double *r1, const double *r2,
                              const double *r3) {
  if (!input_is_valid_) return false;  // helpers.cpp:244-246
  gather_model();
  auto &mm = workspace_.model;
  const sipoc_kkt_model model{mm[0].data(), mm[1].data(), mm[2].data(),  mm[3].data(),
                              mm[4].data(), mm[5].data(), mm[6].data(),  mm[7].data(),
                              mm[8].data(), mm[9].data(), mm[10].data(), mm[11].data()};
  // Zero-length regularization vectors still need a valid pointer.
  static const double none = 0.0;
  int ok = 0;
  const sipoc_error rc = sipoc_kkt_factor_host(workspace_.lqr_workspace.engine, &model,
                                               w ? w : &none, r1 ? r1 : &none, r2 ? r2 : &none,
                                               r3 ? r3 : &none, &ok);
  return rc == SIPOC_OK && ok != 0;
}

void CallbackProvider::solve(const double *b, double *sol) {
  sipoc_kkt_solve_host(workspace_.lqr_workspace.engine, b, sol);
}

void CallbackProvider::add_Kx_to_y(const double *w, const double *r1, const double *r2,
                                   const double *r3, const double *x_x, const double *x_y,
                                   const double *x_z, double *y_x, double *y_y, double *y_z) {
  const int E = input_.topology.num_edges;
  const int xd = input_.dimensions.get_x_dim(E), yd = input_.dimensions.get_y_dim(E),
            zd = input_.dimensions.get_z_dim(E);
  std::vector<double> &x = workspace_.vec[0], &y = workspace_.vec[1];
  std::copy(x_x, x_x + xd, x.begin());
  std::copy(x_y, x_y + yd, x.begin() + xd);
  std::copy(x_z, x_z + zd, x.begin() + xd + yd);
  std::copy(y_x, y_x + xd, y.begin());
  std::copy(y_y, y_y + yd, y.begin() + xd);
  std::copy(y_z, y_z + zd, y.begin() + xd + yd);
  static const double none = 0.0;
  sipoc_kkt_apply_host(workspace_.lqr_workspace.engine, w ? w : &none, r1, r2, r3 ? r3 : &none,
                       x.data(), y.data());
  std::copy(y.begin(), y.begin() + xd, y_x);
  std::copy(y.begin() + xd, y.begin() + xd + yd, y_y);
  std::copy(y.begin() + xd + yd, y.begin() + xd + yd + zd, y_z);
}

void CallbackProvider::add_Hx_to_y(const double *x, double *y) {
  sipoc_kkt_apply_block_host(workspace_.lqr_workspace.engine, SIPOC_KKT_BLOCK_H, x, y);
}
void CallbackProvider::add_Cx_to_y(const double *x, double *y) {
  sipoc_kkt_apply_block_host(workspace_.lqr_workspace.engine, SIPOC_KKT_BLOCK_C, x, y);
}
void CallbackProvider::add_CTx_to_y(const double *x, double *y) {
  sipoc_kkt_apply_block_host(workspace_.lqr_workspace.engine, SIPOC_KKT_BLOCK_CT, x, y);
}
void CallbackProvider::add_Gx_to_y(const double *x, double *y) {
  sipoc_kkt_apply_block_host(workspace_.lqr_workspace.engine, SIPOC_KKT_BLOCK_G, x, y);
}
void CallbackProvider::add_GTx_to_y(const double *x, double *y) {
  sipoc_kkt_apply_block_host(workspace_.lqr_workspace.engine, SIPOC_KKT_BLOCK_GT, x, y);
}

}  // namespace sip::optimal_control
