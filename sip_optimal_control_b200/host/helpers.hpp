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

 private:
  void gather_model();

  const Input &input_;
  Workspace &workspace_;
  bool input_is_valid_;
};

}  // namespace sip::optimal_control
