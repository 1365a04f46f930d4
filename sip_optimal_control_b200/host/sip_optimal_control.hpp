// Host-side stand-in for the model-evaluation half of the reference's top-level solve
// (sip_optimal_control.hpp:8-9, sip_optimal_control.cpp:13-127).
//
// The reference's `solve` hands SIP thirteen callbacks; eight of them are CallbackProvider's
// (helpers.hpp) and four are plain reads of the workspace (get_f / get_grad_f / get_c / get_g,
// sip_optimal_control.cpp:167-177).  The one that computes is `model_callback`: it aims the
// per-node / per-edge input views at SIP's (x, y, z), runs the user's model, and scatters the
// node / edge values into workspace.f / gradient_f / c / g.  evaluate_model is that lambda
// with the scatter done on the GPU (sipoc_model_scatter_host).  The interior-point loop
// itself lives in the un-vendored `sip` library and is not part of this path.
#pragma once

#include "types.hpp"

namespace sip::optimal_control {

// The fields of sip::ModelCallbackInput the lambda reads (sip_optimal_control.cpp:15-16, 52).
struct ModelEvaluationPoint {
  const double *x;  // [x_dim]   = [x_0, u_0, ..., x_E, theta]
  const double *y;  // [y_dim]
  const double *z;  // [z_dim]
  bool new_x;
};

// Needs the engine a CallbackProvider constructed on `workspace` holds.  Returns 0
// (SIPOC_OK) or the sipoc_error of the failing call; workspace.f is always written,
// gradient_f / c / g only when point.new_x (sip_optimal_control.cpp:52).
int evaluate_model(const Input &input, Workspace &workspace, const ModelEvaluationPoint &point);

}  // namespace sip::optimal_control
