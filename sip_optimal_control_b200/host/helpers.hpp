// Host-side stand-in for the reference's CallbackProvider (helpers.hpp:7-33), theta_dim == 0:
// the Newton-KKT linear-solve callbacks SIP invokes once per interior-point iteration.  Every
// method forwards to one C-ABI entry point of libsipoc.so (include/sipoc.h); the arithmetic
// runs on the GPU, this class only gathers / scatters the caller's per-block pointers.
//
//   method                      C ABI call                       reference
//   --------------------------  -------------------------------  ------------------------
//   factor                      sipoc_kkt_factor_host            helpers.cpp:242-370
//   solve                       sipoc_kkt_solve_host             helpers.cpp:749-900
//   add_Kx_to_y                 sipoc_kkt_apply_host             helpers.cpp:953-977
//   add_{H,C,CT,G,GT}x_to_y     sipoc_kkt_apply_block_host       helpers.cpp:979-1368
#pragma once

#include "types.hpp"

namespace sip::optimal_control {

class CallbackProvider {
 public:
  // Creates the engine handle for the structure in `input` (batch of one) inside
  // `workspace`; an invalid structure is latched and every later factor returns false.
  CallbackProvider(const Input &input, Workspace &workspace);

  // KKT -> LQR reduction with the diagonal weights (w, r3 indexed like z; r1 like x; r2
  // like y) and the regularized LQR factorization.  False when a weight is not positive
  // or the factorization fails.
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

 private:
  void gather_model();  // per-block model pointers -> the flat arrays the C ABI takes

  const Input &input_;
  Workspace &workspace_;
  bool input_is_valid_;
};

}  // namespace sip::optimal_control
