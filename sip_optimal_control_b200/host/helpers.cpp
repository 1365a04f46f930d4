// See helpers.hpp.  Marshalling only; every number is computed on the GPU.
#include "helpers.hpp"

#include <algorithm>

#include "../../include/sipoc.h"
#include "device_state.hpp"

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

}  // namespace

// ---- CallbackProvider (helpers.cpp:11-26, 190-407, 749-951, 953-1368) ----------------------
CallbackProvider::CallbackProvider(const Input &input, Workspace &workspace)
    : input_(input), workspace_(workspace), input_is_valid_(false) {
  LQR::Workspace &lw = workspace_.lqr_workspace;
  lw.release_device();
  delete workspace_.staging;
  workspace_.staging = nullptr;
  auto *dev = new LQR::DeviceState();
  const sipoc_structure s = describe(input.dimensions, input.topology);
  last_error_ = sipoc_create(&s, &dev->engine);
  input_is_valid_ = last_error_ == SIPOC_OK;
  if (!input_is_valid_) {
    dev->engine = nullptr;
    delete dev;
    return;
  }
  lw.device = dev;
  sipoc_kkt_sizes z{};
  sipoc_kkt_get_sizes(dev->engine, &z);
  auto *st = new Workspace::Staging();
  const int64_t sizes[12] = {z.node_hxx, z.node_jc,  z.node_jg,  z.edge_hxx,
                             z.edge_hxu, z.edge_huu, z.edge_A,   z.edge_B,
                             z.edge_jcx, z.edge_jcu, z.edge_jgx, z.edge_jgu};
  const int64_t tsizes[10] = {z.node_hxt, z.node_jct,  z.node_jgt, z.node_htt, z.edge_hxt,
                              z.edge_hut, z.edge_dynt, z.edge_jct, z.edge_jgt, z.edge_htt};
  for (int i = 0; i < 12; ++i)
    st->model[i].assign(static_cast<size_t>(std::max<int64_t>(sizes[i], 1)), 0.0);
  for (int i = 0; i < 10; ++i)
    st->theta[i].assign(static_cast<size_t>(std::max<int64_t>(tsizes[i], 1)), 0.0);
  for (auto &v : st->vec) v.assign(static_cast<size_t>(std::max<int64_t>(z.kkt_dim, 1)), 0.0);
  workspace_.staging = st;
  // the compiled topology, for callers that read it from the workspace (helpers.cpp:218-219)
  if (lw.child_offsets != nullptr) {
    const int E = input.topology.num_edges;
    sipoc_get_topology(dev->engine, lw.child_offsets, lw.child_edges, lw.preorder_nodes,
                       lw.postorder_nodes);
    std::copy(input.topology.edge_parents, input.topology.edge_parents + E, lw.edge_parents);
    std::copy(input.topology.edge_children, input.topology.edge_children + E, lw.edge_children);
  }
}

// Per-node / per-edge blocks -> the flat arrays of sipoc_kkt_model (+ theta).
void CallbackProvider::gather_model() {
  const Dimensions &d = input_.dimensions;
  const Topology &t = input_.topology;
  const ModelCallbackOutput &mco = workspace_.model_callback_output;
  Workspace::Staging &st = *workspace_.staging;
  const int p = d.theta_dim;
  size_t o[12] = {0}, ot[10] = {0};
  auto put = [&](int which, const double *src, int count) {
    std::copy(src, src + count, st.model[which].begin() + o[which]);
    o[which] += count;
  };
  auto putt = [&](int which, const double *src, int count) {
    if (count == 0) return;
    std::copy(src, src + count, st.theta[which].begin() + ot[which]);
    ot[which] += count;
  };
  for (int i = 0; i < t.num_nodes(); ++i) {
    const int n = d.get_state_dim(i), c = d.get_node_c_dim(i), g = d.get_node_g_dim(i);
    const NodeModelCallbackOutput &no = mco.nodes[i];
    put(0, no.d2L_dx2, n * n);
    put(1, no.dc_dx, c * n);
    put(2, no.dg_dx, g * n);
    putt(0, no.d2L_dxdtheta, n * p);
    putt(1, no.dc_dtheta, c * p);
    putt(2, no.dg_dtheta, g * p);
    putt(3, no.d2L_dtheta2, p * p);
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
    putt(4, eo.d2L_dxdtheta, np * p);
    putt(5, eo.d2L_dudtheta, m * p);
    putt(6, eo.ddyn_dtheta, nc * p);
    putt(7, eo.dc_dtheta, c * p);
    putt(8, eo.dg_dtheta, g * p);
    putt(9, eo.d2L_dtheta2, p * p);
  }
}

namespace {
struct ModelView {
  sipoc_kkt_theta_model theta;
  sipoc_kkt_model model;
  explicit ModelView(Workspace::Staging &st, int p) {
    auto &m = st.model;
    auto &t = st.theta;
    theta = sipoc_kkt_theta_model{t[0].data(), t[1].data(), t[2].data(), t[3].data(), t[4].data(),
                                  t[5].data(), t[6].data(), t[7].data(), t[8].data(), t[9].data()};
    model = sipoc_kkt_model{m[0].data(), m[1].data(), m[2].data(),  m[3].data(),
                            m[4].data(), m[5].data(), m[6].data(),  m[7].data(),
                            m[8].data(), m[9].data(), m[10].data(), m[11].data(),
                            p > 0 ? &theta : nullptr};
  }
  ModelView(const ModelView &) = delete;
};
}  // namespace

bool CallbackProvider::factor(const double *w, const double *r1, const double *r2,
                              const double *r3) {
  if (!input_is_valid_) return false;  // helpers.cpp:244-246
  gather_model();
  const ModelView view(*workspace_.staging, input_.dimensions.theta_dim);
  // Zero-length regularization vectors still need a valid pointer.
  static const double none = 0.0;
  int ok = 0;
  last_error_ = sipoc_kkt_factor_host(engine_of(workspace_.lqr_workspace), &view.model,
                                      w ? w : &none, r1 ? r1 : &none, r2 ? r2 : &none,
                                      r3 ? r3 : &none, &ok);
  model_is_current_ = false;  // the next operator call reads the workspace's model again
  return last_error_ == SIPOC_OK && ok != 0;
}

void CallbackProvider::solve(const double *b, double *sol) {
  if (!input_is_valid_) return;
  last_error_ = sipoc_kkt_solve_host(engine_of(workspace_.lqr_workspace), b, sol);
}

// The operators read the current model_callback_output (helpers.cpp:1161-1183).
bool CallbackProvider::upload_model() {
  if (!input_is_valid_) return false;
  if (model_is_current_) return true;
  gather_model();
  const ModelView view(*workspace_.staging, input_.dimensions.theta_dim);
  last_error_ = sipoc_kkt_set_model_host(engine_of(workspace_.lqr_workspace), &view.model);
  return last_error_ == SIPOC_OK;
}

void CallbackProvider::add_Kx_to_y(const double *w, const double *r1, const double *r2,
                                   const double *r3, const double *x_x, const double *x_y,
                                   const double *x_z, double *y_x, double *y_y, double *y_z) {
  if (!upload_model()) return;
  const int E = input_.topology.num_edges;
  const int xd = input_.dimensions.get_x_dim(E), yd = input_.dimensions.get_y_dim(E),
            zd = input_.dimensions.get_z_dim(E);
  std::vector<double> &x = workspace_.staging->vec[0], &y = workspace_.staging->vec[1];
  std::copy(x_x, x_x + xd, x.begin());
  std::copy(x_y, x_y + yd, x.begin() + xd);
  std::copy(x_z, x_z + zd, x.begin() + xd + yd);
  std::copy(y_x, y_x + xd, y.begin());
  std::copy(y_y, y_y + yd, y.begin() + xd);
  std::copy(y_z, y_z + zd, y.begin() + xd + yd);
  static const double none = 0.0;
  last_error_ = sipoc_kkt_apply_host(engine_of(workspace_.lqr_workspace), w ? w : &none, r1,
                                     r2, r3 ? r3 : &none, x.data(), y.data());
  if (last_error_ != SIPOC_OK) return;
  std::copy(y.begin(), y.begin() + xd, y_x);
  std::copy(y.begin() + xd, y.begin() + xd + yd, y_y);
  std::copy(y.begin() + xd + yd, y.begin() + xd + yd + zd, y_z);
}

void CallbackProvider::apply_block(int block, const double *x, double *y) {
  if (!upload_model()) return;
  last_error_ = sipoc_kkt_apply_block_host(engine_of(workspace_.lqr_workspace), block, x, y);
}
void CallbackProvider::add_Hx_to_y(const double *x, double *y) { apply_block(SIPOC_KKT_BLOCK_H, x, y); }
void CallbackProvider::add_Cx_to_y(const double *x, double *y) { apply_block(SIPOC_KKT_BLOCK_C, x, y); }
void CallbackProvider::add_CTx_to_y(const double *x, double *y) { apply_block(SIPOC_KKT_BLOCK_CT, x, y); }
void CallbackProvider::add_Gx_to_y(const double *x, double *y) { apply_block(SIPOC_KKT_BLOCK_G, x, y); }
void CallbackProvider::add_GTx_to_y(const double *x, double *y) { apply_block(SIPOC_KKT_BLOCK_GT, x, y); }

}  // namespace sip::optimal_control
