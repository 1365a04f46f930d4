// Batched scatter of one model evaluation into the flat objective / gradient / constraint
// vectors the interior-point loop reads (the model_callback lambda of
// sip_optimal_control.cpp:13-127, after the user's callback has run).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "structure.hpp"

namespace sipoc {

// Values of one model evaluation, engine layout [flat][ld]; flat concatenates the per-node
// (per-edge) vectors in index order (NodeModelCallbackOutput / EdgeModelCallbackOutput,
// types.hpp:48-89: f, df_dx, df_dtheta, c, g; f, df_dx, df_du, df_dtheta, dyn_res, c, g).
struct ModelValues {
  const double *node_f, *node_df_dx, *node_df_dtheta, *node_c, *node_g;
  const double *edge_f, *edge_df_dx, *edge_df_du, *edge_df_dtheta, *edge_dyn_res, *edge_c,
      *edge_g;
};

// f [ld]; gradient_f [x_dim][ld]; c [y_dim][ld]; g [z_dim][ld]; x [x_dim][ld];
// initial_state [n_root][ld].  new_x == false computes f only.  Returns the launch count.
int launch_model_scatter(const DevTables &t, const ModelValues &v, const double *x,
                         const double *initial_state, bool new_x, double *f, double *gradient_f,
                         double *c, double *g, int64_t batch, int64_t ld, cudaStream_t s);

}  // namespace sipoc
