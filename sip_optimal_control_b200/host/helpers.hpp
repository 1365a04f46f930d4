// Host-side stand-in for the reference's CallbackProvider (helpers.hpp:7-33): the Newton-KKT
// linear-solve callbacks SIP invokes once per interior-point iteration, theta (Schur)
// variables included.  Every method forwards to one C-ABI entry point of libsipoc.so
// (include/sipoc.h); the arithmetic runs on the GPU, this class only gathers / scatters the
// caller's per-block pointers.
//
//   method                      C ABI call                       reference
//   --------------------------  -------------------------------  ------------------------
//   factor                      sipoc_kkt_factor_host            helpers.cpp:190-407
//   solve                       sipoc_kkt_solve_host             helpers.cpp:749-951
//   add_Kx_to_y                 sipoc_kkt_apply_host             helpers.cpp:953-977
//   add_{H,C,CT,G,GT}x_to_y     sipoc_kkt_apply_block_host       helpers.cpp:979-1368
//
// The operator methods read the CURRENT workspace.model_callback_output, like the reference
// (helpers.cpp:1161-1183): the model is gathered and uploaded again on every call unless the
// caller has declared it unchanged (model_unchanged()).  The reference's void methods cannot
// report a failing call; the last error is latched and readable through last_error().
#pragma once

#include "types.hpp"

namespace sip::optimal_control {

class CallbackProvider {
 public:
  // Creates the engine handle for the structure in `input` (batch of one) inside
  // `workspace`; an invalid structure is latched and every later factor returns false.
  CallbackProvider(const Input &input, Workspace &workspace);

  // KKT -> LQR reduction with the diagonal weights (w, r3 indexed like z; r1 like x; r2
  // like y), the regularized LQR factorization and, with theta_dim > 0, the Schur
  // complement on theta.  False when a weight is not positive or a factorization fails.
  bool factor(const double *w, const double *r1, const double *r2, const double *r3);

  // sol = K^-1 b on flat [x | y | z] vectors, against the last successful factor.
  void solve(const double *b, double *sol);

  // (y_x, y_y, y_z) += K(w, r1, r2, r3) (x_x, x_y, x_z).
  void add_Kx_to_y(const double *w, const double *r1, const double *r2, const double *r3,
                   const double *x_x, const double *x_y, const double *x_z, double *y_x,
                   double *y_y, double *y_z);

  // One block of the operator at a time, y += B x, without the regularization terms:
  //   H : x-vector -> x-vector        C : x-vector -> y-vector      CT: y-vector -> x-vector
  //   G : x-vector -> z-vector        GT: z-vector -> x-vector
  void add_Hx_to_y(const double *x, double *y);
  void add_Cx_to_y(const double *x, double *y);
  void add_CTx_to_y(const double *x, double *y);
  void add_Gx_to_y(const double *x, double *y);
  void add_GTx_to_y(const double *x, double *y);

  // Not in the reference.  model_unchanged(): the model blocks have not changed since the
  // last upload (skips the gather + upload of the next operator calls; any factor resets it).
  void model_unchanged() { model_is_current_ = true; }
  int last_error() const { return last_error_; }

 private:
  void gather_model();       // per-block model pointers -> the flat arrays the C ABI takes
  bool upload_model();       // gather + sipoc_kkt_set_model_host unless declared current
  void apply_block(int block, const double *x, double *y);

  const Input &input_;
  Workspace &workspace_;
  bool input_is_valid_;
  bool model_is_current_ = false;
  int last_error_ = 0;
};

}  // namespace sip::optimal_control
