// See sip_optimal_control.hpp.  Marshalling only; the scatter runs on the GPU.
#include "sip_optimal_control.hpp"

#include <algorithm>
#include <vector>

#include "../../include/sipoc.h"
#include "device_state.hpp"

namespace sip::optimal_control {

int evaluate_model(const Input &input, Workspace &workspace, const ModelEvaluationPoint &point) {
  const Dimensions &d = input.dimensions;
  const Topology &t = input.topology;
  LQR::DeviceState *dev = workspace.lqr_workspace.device;
  if (dev == nullptr || dev->engine == nullptr) return SIPOC_INVALID_ARGUMENT;
  const int E = t.num_edges, N = t.num_nodes(), p = d.theta_dim;

  // the views the user's model reads (sip_optimal_control.cpp:15-42)
  workspace.model_callback_input.theta = point.x + workspace.stagewise_x_dim;
  for (int node = 0; node < N; ++node)
    workspace.model_callback_input.nodes[node] = NodeModelCallbackInput{
        node, point.x + workspace.x_state_offsets[node],
        point.y + workspace.y_node_c_offsets[node], point.z + workspace.z_node_offsets[node]};
  for (int e = 0; e < E; ++e) {
    const int parent = t.edge_parents[e], child = t.edge_children[e];
    workspace.model_callback_input.edges[e] = EdgeModelCallbackInput{
        e,
        parent,
        child,
        point.x + workspace.x_state_offsets[parent],
        point.x + workspace.x_control_offsets[e],
        point.x + workspace.x_state_offsets[child],
        point.y + workspace.y_dyn_offsets[child],
        point.y + workspace.y_edge_c_offsets[e],
        point.z + workspace.z_edge_offsets[e]};
  }
  input.model_callback(workspace.model_callback_input, workspace.model_callback_output);  // :43

  // per-node / per-edge values -> the twelve flat arrays of sipoc_model_values
  sipoc_model_value_sizes_t sz{};
  sipoc_model_value_sizes(dev->engine, &sz);
  const int64_t sizes[12] = {sz.node_f,      sz.node_df_dx, sz.node_df_dtheta, sz.node_c,
                             sz.node_g,      sz.edge_f,     sz.edge_df_dx,     sz.edge_df_du,
                             sz.edge_df_dtheta, sz.edge_dyn_res, sz.edge_c,    sz.edge_g};
  std::vector<double> flat[12];
  size_t at[12] = {0};
  for (int i = 0; i < 12; ++i) flat[i].assign(static_cast<size_t>(std::max<int64_t>(sizes[i], 1)), 0.0);
  auto put = [&](int which, const double *src, int count) {
    if (count == 0) return;
    std::copy(src, src + count, flat[which].begin() + at[which]);
    at[which] += count;
  };
  const ModelCallbackOutput &mco = workspace.model_callback_output;
  for (int node = 0; node < N; ++node) {
    const NodeModelCallbackOutput &o = mco.nodes[node];
    put(0, &o.f, 1);
    put(1, o.df_dx, d.get_state_dim(node));
    put(2, o.df_dtheta, p);
    put(3, o.c, d.get_node_c_dim(node));
    put(4, o.g, d.get_node_g_dim(node));
  }
  for (int e = 0; e < E; ++e) {
    const EdgeModelCallbackOutput &o = mco.edges[e];
    put(5, &o.f, 1);
    put(6, o.df_dx, d.get_state_dim(t.edge_parents[e]));
    put(7, o.df_du, d.get_control_dim(e));
    put(8, o.df_dtheta, p);
    put(9, o.dyn_res, d.get_state_dim(t.edge_children[e]));
    put(10, o.c, d.get_edge_c_dim(e));
    put(11, o.g, d.get_edge_g_dim(e));
  }
  const sipoc_model_values v{flat[0].data(), flat[1].data(), flat[2].data(),  flat[3].data(),
                             flat[4].data(), flat[5].data(), flat[6].data(),  flat[7].data(),
                             flat[8].data(), flat[9].data(), flat[10].data(), flat[11].data()};
  // zero-sized outputs still need an address
  double none = 0.0;
  auto or_none = [&](double *ptr, int size) { return size > 0 ? ptr : &none; };
  const double *x0 = d.get_state_dim(t.root) > 0 ? input.initial_state : &none;
  return sipoc_model_scatter_host(dev->engine, &v, point.x, x0, point.new_x ? 1 : 0,
                                  &workspace.f, or_none(workspace.gradient_f, workspace.x_dim),
                                  or_none(workspace.c, workspace.y_dim),
                                  or_none(workspace.g, workspace.z_dim));
}

}  // namespace sip::optimal_control
