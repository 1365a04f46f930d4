// Host- and device-side description of ONE problem structure (tree + dims)
// shared by every problem of a batch.  Index tables only; no numerics here.
//
// Follows the reference's structural code: Topology / Dimensions
// (lqr.hpp:5-64), compile_topology_data (lqr.cpp:563-631), validate_input
// (types.cpp:68-134) and populate_workspace_metadata (types.cpp:24-64).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/sipoc.h"

namespace sipoc {

// Largest theta_dim (Schur / global variables) the theta kernels keep in registers.
constexpr int kMaxThetaDim = 32;

// Device-visible index tables.  All pointers point into one int32 device
// allocation owned by the engine.
struct DevTables {
  int E, N, root;
  int max_n, max_m;
  const int *parents, *children;          // [E]
  const int *n;                           // [N] state dims
  const int *m;                           // [E] control dims
  const int *child_offsets;               // [N+1]
  const int *child_edges;                 // [E]
  const int *preorder, *postorder;        // [N]
  const int *in_edge;                     // [N] incoming edge or -1
  const int *nn_off, *n_off;              // [N+1]
  const int *nm_off, *mm_off, *m_off, *a_off, *b_off, *w_off, *k_off;  // [E+1]
  const int *hxx_edge_off;                // [E+1]
  // Newton-KKT
  const int *node_c, *node_g;             // [N]
  const int *edge_c, *edge_g;             // [E]
  const int *x_state, *y_dyn, *y_node_c, *z_node;   // [N]
  const int *x_control, *y_edge_c, *z_edge;         // [E]
  const int *jc_node_off, *jg_node_off;   // [N+1]
  const int *jcx_off, *jcu_off, *jgx_off, *jgu_off;  // [E+1]
  const int *node_c_off, *node_g_off;     // [N+1] prefix sums of node_c / node_g
  const int *edge_c_off, *edge_g_off;     // [E+1]
  const int *pn_off, *cn_off;             // [E+1] prefix sums of the parent / child state dims
  // x = [x_0, u_0, ..., x_E, theta]: x_dim counts theta, sx_dim does not (types.cpp:24-64).
  int x_dim, y_dim, z_dim, kkt_dim;
  int theta_dim, sx_dim;
};

struct HostStructure {
  int E = 0, N = 1, root = 0;
  std::vector<int> parents, children, n, m, node_c, node_g, edge_c, edge_g;
  std::vector<int> child_offsets, child_edges, preorder, postorder, in_edge;
  std::vector<int> nn_off, n_off, nm_off, mm_off, m_off, a_off, b_off, w_off, k_off,
      hxx_edge_off;
  std::vector<int> x_state, x_control, y_dyn, y_node_c, y_edge_c, z_node, z_edge;
  std::vector<int> jc_node_off, jg_node_off, jcx_off, jcu_off, jgx_off, jgu_off;
  std::vector<int> node_c_off, node_g_off, edge_c_off, edge_g_off, pn_off, cn_off;
  int x_dim = 0, y_dim = 0, z_dim = 0, kkt_dim = 0, theta_dim = 0, sx_dim = 0;
  int max_n = 0, max_m = 0;
  bool is_chain = false;    // parent[e] == e, child[e] == e + 1, root == 0
  bool is_uniform = false;  // all n equal, all m equal
  bool has_constraints = false;

  // Validates (types.cpp:68-134, lqr.cpp:563-631) and fills every table.
  sipoc_error build(const sipoc_structure &s, std::string &err);

  // Serialises all tables into one int vector and returns, via `fix`, a
  // DevTables whose pointers are OFFSETS (as pointers from nullptr) to be
  // rebased onto the device allocation.
  std::vector<int> serialise(DevTables &fix) const;
};

}  // namespace sipoc
