// The part of the reference's types.hpp the Newton-KKT linear-solve path reads
// (theta_dim == 0): per-node / per-edge model outputs, Input, Workspace, validate_input.
// Same struct and member names as the reference (types.hpp:48-126, 128-160, 162-320);
// the members that only the SIP outer loop or the theta (Schur) path use are omitted.
#pragma once

#include <vector>

#include "lqr.hpp"

namespace sip::optimal_control {

// Model outputs of one node: value, gradient, equality / inequality residuals with their
// state Jacobians (column-major, rows = constraints), and the Lagrangian Hessian block.
struct NodeModelCallbackOutput {  // types.hpp:48-61 without the theta blocks
  double f;
  double *df_dx;                 // [n]
  double *c, *dc_dx;             // [c], [c x n]
  double *g, *dg_dx;             // [g], [g x n]
  double *d2L_dx2;               // [n x n]
};

// Model outputs of one edge (parent state x, control u, child state): as above plus the
// dynamics residual and its Jacobians (rows = child state) and the x-u / u-u Hessian blocks.
struct EdgeModelCallbackOutput {  // types.hpp:66-89 without the theta blocks
  double f;
  double *df_dx, *df_du;                      // [n_parent], [m]
  double *dyn_res, *ddyn_dx, *ddyn_du;        // [n_child], [n_child x n_parent], [n_child x m]
  double *c, *dc_dx, *dc_du;                  // [c], [c x n_parent], [c x m]
  double *g, *dg_dx, *dg_du;                  // [g], [g x n_parent], [g x m]
  double *d2L_dx2, *d2L_dxdu, *d2L_du2;       // [n_parent x n_parent], [n_parent x m], [m x m]
};

struct ModelCallbackOutput {  // types.hpp:91-126
  NodeModelCallbackOutput *nodes = nullptr;
  EdgeModelCallbackOutput *edges = nullptr;

  void reserve(const Dimensions &dimensions, const Topology &topology);
  void free(const Topology &topology);
};

struct Input {  // types.hpp:128-155, the structure part
  Dimensions dimensions;
  Topology topology;
};

enum class InputValidationStatus {  // types.hpp:157-161
  SUCCESS = 0,
  INVALID_DIMENSIONS = 1,
  INVALID_TOPOLOGY = 2,
};

auto validate_input(const Dimensions &dimensions, const Topology &topology)
    -> InputValidationStatus;

struct Workspace {  // types.hpp:162-320; RegularizedLQRData lives on the device
  ModelCallbackOutput model_callback_output;
  LQR::Workspace lqr_workspace;

  void reserve(const Dimensions &dimensions, const Topology &topology);
  void free(const Topology &topology);

  // Flat host staging of the model blocks and vectors handed to the C ABI.
  std::vector<double> model[12];
  std::vector<double> vec[2];
};

}  // namespace sip::optimal_control
