// The reference's types.hpp surface for the Newton-KKT linear-solve path (same struct and
// member names; types.hpp:12-126 model callback input / output views, :128-160 Input and
// validate_input, :162-320 Workspace), with the same three allocation modes and the same
// arena byte counts (reserve / free, mem_assign, num_bytes).
//
// What differs, and why:
//  * the members that belong to the un-vendored SIP outer loop (sip::Workspace, sip::Settings,
//    Input::timeout_callback, Input::residual_scaling) exist only when "sip/types.hpp" is on the
//    include path; Workspace::num_bytes / mem_assign / reserve therefore come in a form without
//    the sip term (and, under SIPOC_HAVE_SIP, in the reference's full signature);
//  * RegularizedLQRData keeps the reference's host fields and sizes, but the reduction writes
//    its results (Q_mod, M_mod, ...) on the device: the host arrays are not populated;
//  * Workspace::staging holds the flat copies handed to the C ABI.
#pragma once

#include <cstddef>
#include <functional>
#include <vector>

#include "lqr.hpp"

#if defined(__has_include)
#if __has_include("sip/types.hpp")
#include "sip/types.hpp"
#define SIPOC_HAVE_SIP 1
#endif
#endif

namespace sip::optimal_control {

struct NodeModelCallbackInput {  // types.hpp:12-17
  int node;
  const double *state;
  const double *equality_constraint_multipliers;
  const double *inequality_constraint_multipliers;
};

struct EdgeModelCallbackInput {  // types.hpp:19-29
  int edge;
  int parent;
  int child;
  const double *parent_state;
  const double *control;
  const double *child_state;
  const double *costate;
  const double *equality_constraint_multipliers;
  const double *inequality_constraint_multipliers;
};

struct ModelCallbackInput {  // types.hpp:31-44
  const double *theta;
  NodeModelCallbackInput *nodes;
  EdgeModelCallbackInput *edges;

  void reserve(const Topology &topology);
  void free();
  auto mem_assign(const Topology &topology, unsigned char *mem_ptr) -> int;
  static constexpr auto num_bytes(const int num_edges) -> int {
    return static_cast<int>((num_edges + 1) * sizeof(NodeModelCallbackInput) +
                            num_edges * sizeof(EdgeModelCallbackInput));
  }
};

// Model outputs of one node: value, gradients, equality / inequality residuals with their
// Jacobians w.r.t. the state and theta (column-major, rows = constraints), Hessian blocks.
struct NodeModelCallbackOutput {  // types.hpp:48-61
  double f;
  double *df_dx;
  double *df_dtheta;
  double *c;
  double *dc_dx;
  double *dc_dtheta;
  double *g;
  double *dg_dx;
  double *dg_dtheta;
  double *d2L_dx2;
  double *d2L_dxdtheta;
  double *d2L_dtheta2;
};

// Model outputs of one edge (parent state x, control u, theta; the dynamics residual also
// depends on the child state through a fixed -I).
struct EdgeModelCallbackOutput {  // types.hpp:66-89
  double f;
  double *df_dx;
  double *df_du;
  double *df_dtheta;
  double *dyn_res;
  double *ddyn_dx;
  double *ddyn_du;
  double *ddyn_dtheta;
  double *c;
  double *dc_dx;
  double *dc_du;
  double *dc_dtheta;
  double *g;
  double *dg_dx;
  double *dg_du;
  double *dg_dtheta;
  double *d2L_dx2;
  double *d2L_dxdu;
  double *d2L_du2;
  double *d2L_dxdtheta;
  double *d2L_dudtheta;
  double *d2L_dtheta2;
};

// Doubles behind one node's / one edge's output views (n = own / parent state dim).
constexpr int node_output_doubles(int n, int c, int g, int p) {
  return n + p + c * (1 + n + p) + g * (1 + n + p) + n * n + n * p + p * p;
}
constexpr int edge_output_doubles(int n_parent, int n_child, int m, int c, int g, int p) {
  return (n_parent + m + p) + n_child * (1 + n_parent + m + p) + c * (1 + n_parent + m + p) +
         g * (1 + n_parent + m + p) + n_parent * n_parent + n_parent * m + m * m +
         n_parent * p + m * p + p * p;
}

struct ModelCallbackOutput {  // types.hpp:91-126
  NodeModelCallbackOutput *nodes;
  EdgeModelCallbackOutput *edges;

  void reserve(const Dimensions &dimensions, const Topology &topology);
  void free(const Topology &topology);
  auto mem_assign(const Dimensions &dimensions, const Topology &topology,
                  unsigned char *mem_ptr) -> int;
  static constexpr auto num_bytes(int state_dim, int control_dim, int num_edges, int node_c_dim,
                                  int node_g_dim, int edge_c_dim, int edge_g_dim,
                                  int theta_dim = 0) -> int {
    const int N = num_edges + 1;
    return static_cast<int>(N * sizeof(NodeModelCallbackOutput) +
                            num_edges * sizeof(EdgeModelCallbackOutput)) +
           (N * node_output_doubles(state_dim, node_c_dim, node_g_dim, theta_dim) +
            num_edges * edge_output_doubles(state_dim, state_dim, control_dim, edge_c_dim,
                                            edge_g_dim, theta_dim)) *
               static_cast<int>(sizeof(double));
  }
  static auto num_bytes(const Dimensions &dimensions, const Topology &topology) -> int;
};

struct Input {  // types.hpp:128-155
  using ModelCallback = std::function<void(const ModelCallbackInput &, ModelCallbackOutput &)>;

  Dimensions dimensions;
  Topology topology;
  const double *initial_state = nullptr;  // [state dim of the root]
  ModelCallback model_callback;
#ifdef SIPOC_HAVE_SIP
  ::sip::Input::TimeoutCallback timeout_callback;
#else
  std::function<bool()> timeout_callback;
#endif
  // Bounds in the flattened primal ordering [x_0, u_0, ..., x_E, theta]; null = none.
  const double *lower_bounds = nullptr;
  const double *upper_bounds = nullptr;
#ifdef SIPOC_HAVE_SIP
  ::sip::Input::ResidualScaling residual_scaling;
#endif

  auto num_bound_sides() const -> int;
};

enum class InputValidationStatus {  // types.hpp:157-161
  SUCCESS = 0,
  INVALID_DIMENSIONS = 1,
  INVALID_TOPOLOGY = 2,
};

auto validate_input(const Dimensions &dimensions, const Topology &topology)
    -> InputValidationStatus;

struct Workspace {  // types.hpp:162-320
  struct RegularizedLQRData {
    double **node_mod_w_inv;
    double **edge_mod_w_inv;
    double **Q_mod;
    double **M_mod;
    double **R_mod;
    double **q_mod;
    double **r_mod;
    double **c_mod;
    double **dyn_r2;
    double **node_c_r2_inv;
    double **edge_c_r2_inv;
    double *theta_jacobian;
    double *theta_solution;
    double *theta_schur;
    double *theta_schur_factor;
    double *theta_rhs;
    double *theta_stagewise_rhs;
    double *stagewise_scratch;

    void reserve(const Dimensions &dimensions, int num_edges);
    void free(int num_edges);
    auto mem_assign(const Dimensions &dimensions, int num_edges, unsigned char *mem_ptr) -> int;

    // Uniform dims.  Pointer tables: mod_w_inv and c_r2_inv 2T + 1 each, Q_mod / q_mod /
    // c_mod / dyn_r2 T + 1 each, M_mod / R_mod / r_mod T each.
    static constexpr auto num_bytes(int state_dim, int control_dim, int num_edges,
                                    int node_c_dim, int node_g_dim, int edge_c_dim,
                                    int edge_g_dim, int theta_dim = 0) -> int {
      const int T = num_edges, N = num_edges + 1, n = state_dim, m = control_dim, p = theta_dim;
      const int kkt = T * (n + m) + n + (node_c_dim + n) * N + edge_c_dim * T + node_g_dim * N +
                      edge_g_dim * T;
      const int pointers = 2 * (2 * T + 1) + 4 * N + 3 * T;
      const int doubles = N * node_g_dim + T * edge_g_dim + N * node_c_dim + T * edge_c_dim +
                          N * n * n + T * n * m + T * m * m + 3 * N * n + T * m +
                          (p > 0 ? 2 * kkt * p + 2 * p * p + p + kkt : 0) + 2 * n * (p > 0 ? p : 1);
      return pointers * static_cast<int>(sizeof(double *)) +
             doubles * static_cast<int>(sizeof(double));
    }
    static auto num_bytes(const Dimensions &dimensions, int num_edges) -> int;
  };

  // The allocation modes without the SIP outer loop's own workspace ...
  void reserve(const Dimensions &dimensions, const Topology &topology);
  void free(const Topology &topology);
  auto mem_assign(const Dimensions &dimensions, const Topology &topology, unsigned char *mem_ptr)
      -> int;
  static auto num_bytes(const Dimensions &dimensions, const Topology &topology) -> int;
#ifdef SIPOC_HAVE_SIP
  // ... and with it, in the reference's signatures (types.hpp:239-289).
  void reserve(const Dimensions &dimensions, const Topology &topology, int num_bound_sides,
               const sip::Settings &settings);
  auto mem_assign(const Dimensions &dimensions, const Topology &topology, int num_bound_sides,
                  const sip::Settings &settings, unsigned char *mem_ptr) -> int;
  static auto num_bytes(const Dimensions &dimensions, const Topology &topology,
                        int num_bound_sides, const sip::Settings &settings) -> int;
#endif

  ModelCallbackInput model_callback_input;
  ModelCallbackOutput model_callback_output;

  double f;
  double *gradient_f;
  double *c;
  double *g;
  int stagewise_x_dim;
  int x_dim;
  int y_dim;
  int z_dim;
  int stagewise_kkt_dim;
  int *x_state_offsets;
  int *x_control_offsets;
  int *y_dyn_offsets;
  int *y_node_c_offsets;
  int *y_edge_c_offsets;
  int *z_node_offsets;
  int *z_edge_offsets;
  double **ddyn_dx;
  double **ddyn_du;

  LQR::Workspace lqr_workspace;
  LQR::Output lqr_output;
  RegularizedLQRData regularized_lqr_data;
#ifdef SIPOC_HAVE_SIP
  sip::Workspace sip_workspace;
#endif

  // Not in the reference: flat host copies of the model blocks handed to the C ABI
  // (12 stagewise + 10 theta arrays) and two KKT vectors; heap, sized by CallbackProvider.
  struct Staging {
    std::vector<double> model[12], theta[10], vec[2];
    bool uploaded = false;
  };
  Staging *staging = nullptr;
};

// Offsets, dimensions and the dynamics-Jacobian tables of `workspace` (types.cpp:24-64);
// called by reserve / mem_assign.
void populate_workspace_metadata(Workspace &workspace, const Dimensions &dimensions,
                                 const Topology &topology);

}  // namespace sip::optimal_control
