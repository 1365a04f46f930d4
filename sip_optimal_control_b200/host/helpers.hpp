// The reference's CallbackProvider (helpers.hpp:7-33) over the C ABI, theta_dim == 0:
// the Newton-KKT linear-solve callbacks SIP invokes once per interior-point iteration.
#pragma once

#include "types.hpp"

namespace sip::optimal_control {

class CallbackProvider {
 public:
  CallbackProvider(const Input &input, Workspace &workspace);

  bool factor(const double *w, const double *r1, const double *r2, const double *r3);
  void solve(const double *b, double *sol);
  void add_Kx_to_y(const double *w, const double *r1, const double *r2, const double *r3,
                   const double *x_x, const double *x_y, const double *x_z, double *y_x,
                   double *y_y, double *y_z);
  // y += B x for one block of the operator (helpers.hpp:17-21 of the reference).
  void add_Hx_to_y(const double *x, double *y);
  void add_Cx_to_y(const double *x, double *y);
  void add_CTx_to_y(const double *x, double *y);
  void add_Gx_to_y(const double *x, double *y);
  void add_GTx_to_y(const double *x, double *y);

 private:
  void gather_model();

  const Input &input_;
  Workspace &workspace_;
  bool input_is_valid_;
};

}  // namespace sip::optimal_control
