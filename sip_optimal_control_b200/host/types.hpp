// The part of the reference's types.hpp the Newton-KKT linear-solve path reads
// (theta_dim == 0): per-node / per-edge model outputs, Input, Workspace, validate_input.
// Same struct and member names as the reference (types.hpp:48-126, 128-160, 162-320);
// the members that only the SIP outer loop or the theta (Schur) path use are omitted.
#pragma once

#include <vector>

#include "lqr.hpp"

namespace sip::optimal_control {

struct NodeModelCallbackOutput {  // types.hpp:48-61 without the theta blocks
  double f;
  double *df_dx;
  double *c;
  double *dc_dx;
  double *g;
  double *dg_dx;
  double *d2L_dx2;
};

struct EdgeModelCallbackOutput {  // types.hpp:66-89 without the theta blocks
  double f;
  double *df_dx;
  double *df_du;
  double *dyn_res;
  double *ddyn_dx;
  double *ddyn_du;
  double *c;
  double *dc_dx;
  double *dc_du;
  double *g;
  double *dg_dx;
  double *dg_du;
  double *d2L_dx2;
  double *d2L_dxdu;
  double *d2L_du2;
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
