// C++ restatement of the reference's own unit tests for the hot path, written against
// the reference's class names (Topology, Dimensions, LQR, CallbackProvider) and run
// against the GPU engine through the host mirror in this directory.  gtest is not in
// this image, so the harness is a plain main() with CHECK macros; tests/test_host_cpp.py
// builds and runs it under `pytest -m gpu`.
//
// Reference tests restated (file:line in joaospinto/sip_optimal_control):
//   LQRFactor.ReportsSuccess / BoolFactorWrapsStatusApi           tests/lqr_test.cpp:188-204
//   LQRFactor.ReportsInvalidDelta / F / G failure                 tests/lqr_test.cpp:206-227
//   LQRSolve.SolvesNonuniformDiagonalDeltaProblem                 tests/lqr_test.cpp:229-263
//   LQRSolve.SolvesBranchingTreeProblem                           tests/lqr_test.cpp:300-335, 411-429
//   LQRTopology.ReusesCompiledTopologyAcrossFactorAndSolveCalls   tests/lqr_test.cpp:431-450
//   LQRFactor.RejectsInvalidTreeTopology                          tests/lqr_test.cpp:452-464
//   CallbackProvider.SolvesChainWithNodeAndEdgeConstraints        tests/variable_dimensions_test.cpp:77-181, 265-290
//   CallbackProvider.SolvesBranchedSystemWithZeroDimensionalRoot  tests/variable_dimensions_test.cpp:316-336
//   InputValidation.AcceptsSeparateNodeAndEdgeDimensions          tests/variable_dimensions_test.cpp:183-224
//   Workspace.UniformStaticAndDynamicMemorySizesMatch             tests/variable_dimensions_test.cpp:226-263
//   CallbackProvider on a mem_assign'ed arena (size == num_bytes) tests/variable_dimensions_test.cpp:281-287
//   CallbackProvider.SolvesBranchedSystemWithSchurVariables       tests/variable_dimensions_test.cpp:338-363
// (Workspace::num_bytes / mem_assign are used in their form without the un-vendored SIP outer
// loop's own workspace term: sip::Workspace and sip::Settings do not exist in this image.)
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "helpers.hpp"
#include "sip_optimal_control.hpp"

using namespace sip::optimal_control;

static int g_failures = 0;
#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) {                                                           \
      std::printf("  CHECK failed: %s  (%s:%d)\n", #cond, __FILE__, __LINE__); \
      ++g_failures;                                                          \
    }                                                                        \
  } while (0)

// A problem owning its blocks (column-major) plus the pointer tables LQR::Input wants.
struct Problem {
  Topology topology;
  Dimensions dimensions;
  std::vector<std::vector<double>> Q, M, R, q, r, A, B, c, delta, x, u, y;
  std::vector<double *> pQ, pM, pR, pq, pr, pA, pB, pc, pdelta, px, pu, py;

  Problem(int E, int root, const std::vector<int> &parents, const std::vector<int> &children,
          const std::vector<int> &n, const std::vector<int> &m) {
    topology.num_edges = E;
    topology.reserve(E);
    topology.set_tree(root, parents.data(), children.data());
    dimensions.reserve(E);
    int *sd = const_cast<int *>(dimensions.state_dims);
    int *cd = const_cast<int *>(dimensions.control_dims);
    for (int i = 0; i <= E; ++i) sd[i] = n[i];
    for (int e = 0; e < E; ++e) cd[e] = m[e];
    for (int i = 0; i <= E; ++i) {
      Q.emplace_back(n[i] * n[i], 0.0);
      q.emplace_back(n[i], 0.0);
      c.emplace_back(n[i], 0.0);
      delta.emplace_back(n[i], 1.0);
      x.emplace_back(n[i], 0.0);
      y.emplace_back(n[i], 0.0);
    }
    for (int e = 0; e < E; ++e) {
      const int np = n[parents[e]], nc = n[children[e]];
      M.emplace_back(np * m[e], 0.0);
      R.emplace_back(m[e] * m[e], 0.0);
      r.emplace_back(m[e], 0.0);
      A.emplace_back(nc * np, 0.0);
      B.emplace_back(nc * m[e], 0.0);
      u.emplace_back(m[e], 0.0);
    }
    link();
  }
  void link() {
    auto tab = [](std::vector<std::vector<double>> &blocks, std::vector<double *> &ptrs) {
      ptrs.clear();
      for (auto &b : blocks) ptrs.push_back(b.data());
      if (ptrs.empty()) ptrs.push_back(nullptr);
    };
    tab(Q, pQ); tab(M, pM); tab(R, pR); tab(q, pq); tab(r, pr); tab(A, pA); tab(B, pB);
    tab(c, pc); tab(delta, pdelta); tab(x, px); tab(u, pu); tab(y, py);
  }
  LQR::Input input() {
    return LQR::Input{pQ.data(), pM.data(), pR.data(), pq.data(), pr.data(), pA.data(),
                      pB.data(), pc.data(), pdelta.data(), dimensions, topology};
  }
  LQR::Output output() { return LQR::Output{px.data(), pu.data(), py.data()}; }
  int n(int node) const { return dimensions.get_state_dim(node); }
  int m(int edge) const { return dimensions.get_control_dim(edge); }
  ~Problem() {
    topology.free();
    dimensions.free();
  }
};

static void set(std::vector<double> &blk, int rows, std::initializer_list<double> row_major) {
  int k = 0;  // Eigen's comma initialiser fills row by row
  const int cols = static_cast<int>(blk.size()) / (rows > 0 ? rows : 1);
  for (double v : row_major) {
    blk[(k % cols) * rows + (k / cols)] = v;
    ++k;
  }
}
static void diag(std::vector<double> &blk, int n, std::initializer_list<double> d) {
  std::fill(blk.begin(), blk.end(), 0.0);
  int i = 0;
  for (double v : d) {
    blk[i * n + i] = v;
    ++i;
  }
}

// KKT residual 2-norm of the LQR system on a tree (tests/lqr_test.cpp:152-186, 371-409).
static double residual_norm(const Problem &p) {
  const Topology &t = p.topology;
  double sq = 0.0;
  for (int node = 0; node < t.num_nodes(); ++node) {
    const int n = p.n(node);
    for (int i = 0; i < n; ++i) {
      double s = p.q[node][i] - p.y[node][i];
      for (int j = 0; j < n; ++j) s += p.Q[node][(i >= j ? j * n + i : i * n + j)] * p.x[node][j];
      for (int e = 0; e < t.num_edges; ++e) {
        if (t.edge_parents[e] != node) continue;
        const int child = t.edge_children[e], nc = p.n(child);
        for (int a = 0; a < p.m(e); ++a) s += p.M[e][a * n + i] * p.u[e][a];
        for (int k = 0; k < nc; ++k) s += p.A[e][i * nc + k] * p.y[child][k];
      }
      sq += s * s;
    }
  }
  for (int e = 0; e < t.num_edges; ++e) {
    const int pa = t.edge_parents[e], ch = t.edge_children[e];
    const int n = p.n(pa), nc = p.n(ch), m = p.m(e);
    for (int a = 0; a < m; ++a) {
      double s = p.r[e][a];
      for (int j = 0; j < m; ++j) s += p.R[e][(a >= j ? j * m + a : a * m + j)] * p.u[e][j];
      for (int i = 0; i < n; ++i) s += p.M[e][a * n + i] * p.x[pa][i];
      for (int k = 0; k < nc; ++k) s += p.B[e][a * nc + k] * p.y[ch][k];
      sq += s * s;
    }
    for (int k = 0; k < nc; ++k) {
      double s = p.c[ch][k] - p.x[ch][k] - p.delta[ch][k] * p.y[ch][k];
      for (int i = 0; i < n; ++i) s += p.A[e][i * nc + k] * p.x[pa][i];
      for (int a = 0; a < m; ++a) s += p.B[e][a * nc + k] * p.u[e][a];
      sq += s * s;
    }
  }
  const int root = t.root;
  for (int i = 0; i < p.n(root); ++i) {
    const double s = -p.x[root][i] - p.delta[root][i] * p.y[root][i] + p.c[root][i];
    sq += s * s;
  }
  return std::sqrt(sq);
}

// tests/lqr_test.cpp:45-75: Q = I, R = I, A = I, B = ones, M = 0, delta = 1.
static Problem identity_chain(int n, int m, int T) {
  std::vector<int> parents(T), children(T);
  for (int e = 0; e < T; ++e) {
    parents[e] = e;
    children[e] = e + 1;
  }
  Problem p(T, 0, parents, children, std::vector<int>(T + 1, n), std::vector<int>(T, m));
  for (int i = 0; i <= T; ++i)
    for (int d = 0; d < n; ++d) p.Q[i][d * n + d] = 1.0;
  for (int e = 0; e < T; ++e) {
    for (int d = 0; d < m; ++d) p.R[e][d * m + d] = 1.0;
    for (int d = 0; d < n; ++d) p.A[e][d * n + d] = 1.0;
    std::fill(p.B[e].begin(), p.B[e].end(), 1.0);
  }
  return p;
}

static void test_factor_status_api() {
  {
    Problem p = identity_chain(2, 1, 2);
    LQR::Input in = p.input();
    LQR::Workspace ws;
    ws.reserve(p.dimensions, p.topology);
    LQR lqr(in, ws);
    CHECK(lqr.factor_with_status() == LQR::FactorStatus::SUCCESS);
    CHECK(lqr.factor());
    ws.free(p.topology.num_edges);
  }
  {  // delta[T][0] = 0 -> INVALID_DELTA
    Problem p = identity_chain(2, 1, 2);
    p.delta[2][0] = 0.0;
    LQR::Input in = p.input();
    LQR::Workspace ws;
    ws.reserve(p.dimensions, p.topology);
    LQR lqr(in, ws);
    CHECK(lqr.factor_with_status() == LQR::FactorStatus::INVALID_DELTA);
    CHECK(!lqr.factor());
    ws.free(p.topology.num_edges);
  }
  {  // n = m = T = 1, Q[T] = -2 -> F failure
    Problem p = identity_chain(1, 1, 1);
    p.Q[1][0] = -2.0;
    LQR::Input in = p.input();
    LQR::Workspace ws;
    ws.reserve(p.dimensions, p.topology);
    LQR lqr(in, ws);
    CHECK(lqr.factor_with_status() == LQR::FactorStatus::F_FACTORIZATION_FAILURE);
    ws.free(p.topology.num_edges);
  }
  {  // Q[T] = 0, R = -1 -> G failure
    Problem p = identity_chain(1, 1, 1);
    p.Q[1][0] = 0.0;
    p.R[0][0] = -1.0;
    LQR::Input in = p.input();
    LQR::Workspace ws;
    ws.reserve(p.dimensions, p.topology);
    LQR lqr(in, ws);
    CHECK(lqr.factor_with_status() == LQR::FactorStatus::G_FACTORIZATION_FAILURE);
    ws.free(p.topology.num_edges);
  }
}

static void test_nonuniform_delta_chain() {  // tests/lqr_test.cpp:229-263
  const int n = 3, m = 2, T = 3;
  Problem p = identity_chain(n, m, T);
  for (int i = 0; i < T; ++i) {
    set(p.A[i], n, {1.0 + 0.02 * i, 0.03, -0.01, -0.02, 0.95 + 0.01 * i, 0.04, 0.01, -0.03,
                    1.02 - 0.01 * i});
    set(p.B[i], n, {0.2, -0.1, 0.05, 0.15, -0.1, 0.08});
    diag(p.Q[i], n, {1.0 + 0.1 * i, 1.4 + 0.05 * i, 1.8 + 0.03 * i});
    diag(p.R[i], m, {1.2 + 0.1 * i, 1.6 + 0.07 * i});
    p.q[i] = {0.2 + 0.01 * i, -0.1 + 0.02 * i, 0.05 - 0.03 * i};
    p.r[i] = {-0.2 + 0.03 * i, 0.1 - 0.01 * i};
    p.c[i] = {0.03 + 0.01 * i, -0.04 + 0.02 * i, 0.02 - 0.01 * i};
    p.delta[i] = {0.03 + 0.01 * i, 0.11 + 0.02 * i, 0.19 + 0.03 * i};
  }
  diag(p.Q[T], n, {1.3, 1.7, 2.1});
  p.q[T] = {0.06, -0.08, 0.12};
  p.c[T] = {-0.02, 0.05, -0.01};
  p.delta[T] = {0.07, 0.17, 0.29};
  p.link();
  LQR::Input in = p.input();
  // the workspace on a caller arena (lqr.hpp:142-186): sized by num_bytes, carved by mem_assign
  LQR::Workspace ws;
  std::vector<unsigned char> arena(LQR::Workspace::num_bytes(p.dimensions, p.topology));
  CHECK(LQR::Workspace::num_bytes(p.dimensions, p.topology) == LQR::Workspace::num_bytes(n, m, T));
  CHECK(ws.mem_assign(p.dimensions, p.topology, arena.data()) == static_cast<int>(arena.size()));
  LQR lqr(in, ws);
  CHECK(ws.postorder_nodes[0] == T && ws.preorder_nodes[0] == 0);  // compiled into the arena
  CHECK(lqr.factor());
  LQR::Output out = p.output();
  lqr.solve(out);
  const double res = residual_norm(p);
  std::printf("  nonuniform-delta chain residual %.3e\n", res);
  CHECK(res < 1e-12);
  // ReusesCompiledTopologyAcrossFactorAndSolveCalls: a second round gives the same answer.
  const std::vector<double> x1 = p.x[1];
  CHECK(lqr.factor());
  lqr.solve(out);
  CHECK(p.x[1] == x1);
  ws.release_device();  // an arena user drops the GPU engine explicitly
}

static void test_branching_tree() {  // tests/lqr_test.cpp:300-335, 411-429
  Problem p(2, 0, {0, 0}, {1, 2}, {2, 2, 2}, {1, 1});
  set(p.Q[0], 2, {2.0, 0.1, 0.1, 1.5});
  set(p.Q[1], 2, {1.3, 0.2, 0.2, 1.7});
  set(p.Q[2], 2, {1.8, -0.1, -0.1, 1.4});
  set(p.M[0], 2, {0.2, -0.1});
  set(p.M[1], 2, {-0.15, 0.05});
  p.R[0] = {1.6};
  p.R[1] = {1.9};
  set(p.A[0], 2, {1.0, 0.2, 0.0, 0.9});
  set(p.A[1], 2, {0.8, -0.1, 0.3, 1.1});
  set(p.B[0], 2, {0.4, 0.2});
  set(p.B[1], 2, {-0.1, 0.5});
  p.q[0] = {0.3, -0.2};  p.q[1] = {-0.1, 0.4};  p.q[2] = {0.2, 0.1};
  p.r[0] = {-0.3};       p.r[1] = {0.25};
  p.c[0] = {0.1, -0.2};  p.c[1] = {-0.05, 0.1}; p.c[2] = {0.2, 0.15};
  p.delta[0] = {0.7, 0.9}; p.delta[1] = {0.8, 1.1}; p.delta[2] = {1.0, 0.6};
  p.link();
  LQR::Input in = p.input();
  LQR::Workspace ws;
  ws.reserve(p.dimensions, p.topology);
  LQR lqr(in, ws);
  CHECK(lqr.factor_with_status() == LQR::FactorStatus::SUCCESS);
  LQR::Output out = p.output();
  lqr.solve(out);
  const double res = residual_norm(p);
  std::printf("  branching tree residual %.3e\n", res);
  CHECK(res < 1e-12);
  ws.free(p.topology.num_edges);
}

static void test_rejects_invalid_topology() {  // tests/lqr_test.cpp:452-464
  Problem p(2, 0, {0, 0}, {1, 1}, {2, 2, 2}, {1, 1});  // two edges into node 1
  LQR::Input in = p.input();
  LQR::Workspace ws;
  ws.reserve(p.dimensions, p.topology);
  LQR lqr(in, ws);
  CHECK(lqr.factor_with_status() == LQR::FactorStatus::INVALID_TOPOLOGY);
  CHECK(!lqr.factor());
  ws.free(p.topology.num_edges);
}

// ---- Newton-KKT (tests/variable_dimensions_test.cpp) ------------------------------------
static void fill_sequence(double *v, int size, double scale) {  // :46-50
  for (int i = 0; i < size; ++i) v[i] = scale * (i + 1);
}
static void identity_scaled(double *v, int n, double s) {
  for (int i = 0; i < n * n; ++i) v[i] = 0.0;
  for (int i = 0; i < n; ++i) v[i * n + i] = s;
}

// initialize_model (:77-133) + expect_kkt_solve (:135-181) on a workspace the caller set up.
static double kkt_solve_residual(const Input &input, Workspace &workspace, double theta_diagonal) {
  const Dimensions &d = input.dimensions;
  const Topology &t = input.topology;
  const int E = t.num_edges, p = d.theta_dim;
  auto &mco = workspace.model_callback_output;
  for (int node = 0; node <= E; ++node) {
    const int n = d.get_state_dim(node), c = d.get_node_c_dim(node), g = d.get_node_g_dim(node);
    auto &o = mco.nodes[node];
    fill_sequence(o.dc_dx, c * n, 0.013 * (node + 1));
    fill_sequence(o.dc_dtheta, c * p, 0.001 * (node + 1));
    fill_sequence(o.dg_dx, g * n, -0.011 * (node + 1));
    fill_sequence(o.dg_dtheta, g * p, -0.0007 * (node + 1));
    identity_scaled(o.d2L_dx2, n, 2.5 + 0.2 * node);
    fill_sequence(o.d2L_dxdtheta, n * p, 0.0005 * (node + 1));
    identity_scaled(o.d2L_dtheta2, p, theta_diagonal);
  }
  for (int e = 0; e < E; ++e) {
    const int np = d.get_state_dim(t.edge_parents[e]), nch = d.get_state_dim(t.edge_children[e]);
    const int mm = d.get_control_dim(e), c = d.get_edge_c_dim(e), g = d.get_edge_g_dim(e);
    auto &o = mco.edges[e];
    fill_sequence(o.ddyn_dx, nch * np, 0.025 + 0.004 * e);
    fill_sequence(o.ddyn_du, nch * mm, -0.031 - 0.003 * e);
    fill_sequence(o.ddyn_dtheta, nch * p, 0.0009 * (e + 1));
    fill_sequence(o.dc_dx, c * np, 0.017 * (e + 1));
    fill_sequence(o.dc_du, c * mm, 0.019 * (e + 1));
    fill_sequence(o.dc_dtheta, c * p, 0.0008 * (e + 1));
    fill_sequence(o.dg_dx, g * np, -0.014 * (e + 1));
    fill_sequence(o.dg_du, g * mm, 0.016 * (e + 1));
    fill_sequence(o.dg_dtheta, g * p, -0.0006 * (e + 1));
    identity_scaled(o.d2L_dx2, np, 0.3 + 0.05 * e);
    fill_sequence(o.d2L_dxdu, np * mm, 0.009 * (e + 1));
    identity_scaled(o.d2L_du2, mm, 3.0 + 0.2 * e);
    fill_sequence(o.d2L_dxdtheta, np * p, 0.0004 * (e + 1));
    fill_sequence(o.d2L_dudtheta, mm * p, -0.0003 * (e + 1));
    identity_scaled(o.d2L_dtheta2, p, theta_diagonal);
  }

  CallbackProvider callback_provider(input, workspace);
  const int x_dim = d.get_x_dim(E), y_dim = d.get_y_dim(E), z_dim = d.get_z_dim(E),
            kkt_dim = x_dim + y_dim + z_dim;
  CHECK(workspace.x_dim == x_dim && workspace.stagewise_kkt_dim == kkt_dim - p);
  std::vector<double> w(z_dim + 1, 1.3), r2(y_dim + 1, 0.9), r3(z_dim + 1, 0.4), r1(x_dim + 1);
  fill_sequence(r1.data(), x_dim, 0.03);
  for (double &v : r1) v += 0.2;
  {  // the operator works before any factor: it reads the workspace's current model
    std::vector<double> ones(x_dim + 1, 1.0), hx(x_dim + 1, 0.0);
    callback_provider.add_Hx_to_y(ones.data(), hx.data());
    CHECK(callback_provider.last_error() == 0);
    double total = 0.0;
    for (int i = 0; i < x_dim; ++i) total += std::fabs(hx[i]);
    CHECK(total > 0.0);
  }
  CHECK(callback_provider.factor(w.data(), r1.data(), r2.data(), r3.data()));
  std::vector<double> rhs(kkt_dim), solution(kkt_dim, 0.0);
  fill_sequence(rhs.data(), kkt_dim, 0.01);
  callback_provider.solve(rhs.data(), solution.data());
  CHECK(callback_provider.last_error() == 0);
  std::vector<double> px(x_dim + 1, 0.0), py(y_dim + 1, 0.0), pz(z_dim + 1, 0.0);
  callback_provider.add_Kx_to_y(w.data(), r1.data(), r2.data(), r3.data(), solution.data(),
                                solution.data() + x_dim, solution.data() + x_dim + y_dim,
                                px.data(), py.data(), pz.data());
  {  // add_Kx_to_y is the sum of its five blocks and the diagonal terms (helpers.cpp:953-977)
    const double *sx = solution.data(), *sy = sx + x_dim, *sz = sy + y_dim;
    std::vector<double> qx(x_dim + 1, 0.0), qy(y_dim + 1, 0.0), qz(z_dim + 1, 0.0);
    callback_provider.model_unchanged();  // five operator calls on one upload
    callback_provider.add_Hx_to_y(sx, qx.data());
    callback_provider.add_Cx_to_y(sx, qy.data());
    callback_provider.add_CTx_to_y(sy, qx.data());
    callback_provider.add_Gx_to_y(sx, qz.data());
    callback_provider.add_GTx_to_y(sz, qx.data());
    double worst = 0.0;
    for (int i = 0; i < x_dim; ++i) worst = std::max(worst, std::fabs(qx[i] + r1[i] * sx[i] - px[i]));
    for (int i = 0; i < y_dim; ++i) worst = std::max(worst, std::fabs(qy[i] - r2[i] * sy[i] - py[i]));
    for (int i = 0; i < z_dim; ++i)
      worst = std::max(worst, std::fabs(qz[i] - (w[i] + r3[i]) * sz[i] - pz[i]));
    CHECK(worst < 1e-12);
  }
  double sq = 0.0;
  for (int i = 0; i < x_dim; ++i) sq += (px[i] - rhs[i]) * (px[i] - rhs[i]);
  for (int i = 0; i < y_dim; ++i) sq += (py[i] - rhs[x_dim + i]) * (py[i] - rhs[x_dim + i]);
  for (int i = 0; i < z_dim; ++i)
    sq += (pz[i] - rhs[x_dim + y_dim + i]) * (pz[i] - rhs[x_dim + y_dim + i]);
  return std::sqrt(sq);
}

struct ChainTopology {
  std::array<int, 2> parent = {0, 1};
  std::array<int, 2> child = {1, 2};
};
struct BranchTopology {
  std::array<int, 2> parent = {0, 0};
  std::array<int, 2> child = {1, 2};
};

// `arena`: the workspace on caller memory (mem_assign) instead of reserve / free.
static double kkt_case(int theta_dim, const std::array<int, 2> &parent,
                       const std::array<int, 2> &child, const std::array<int, 3> &state_dims,
                       const std::array<int, 2> &control_dims, const std::array<int, 3> &node_c,
                       const std::array<int, 3> &node_g, const std::array<int, 2> &edge_c,
                       const std::array<int, 2> &edge_g, bool arena, double theta_diagonal = 0.0) {
  // aggregate initialisation, as the reference's tests write it
  Input input{
      .dimensions = {theta_dim, state_dims.data(), control_dims.data(), node_c.data(),
                     node_g.data(), edge_c.data(), edge_g.data()},
      .topology = {2, 0, parent.data(), child.data()},
      .model_callback = [](const ModelCallbackInput &, ModelCallbackOutput &) {},
      .timeout_callback = []() { return false; },
  };
  CHECK(validate_input(input.dimensions, input.topology) == InputValidationStatus::SUCCESS);
  CHECK(input.num_bound_sides() == 0);
  Workspace workspace;
  std::vector<unsigned char> memory;
  if (arena) {
    memory.resize(Workspace::num_bytes(input.dimensions, input.topology));
    CHECK(workspace.mem_assign(input.dimensions, input.topology, memory.data()) ==
          static_cast<int>(memory.size()));
  } else {
    workspace.reserve(input.dimensions, input.topology);
  }
  const double res = kkt_solve_residual(input, workspace, theta_diagonal);
  if (arena) {
    workspace.lqr_workspace.release_device();
    delete workspace.staging;
  } else {
    workspace.free(input.topology);
  }
  return res;
}

static void test_callback_provider() {
  const ChainTopology chain_t;
  const BranchTopology branch_t;
  const double chain = kkt_case(0, chain_t.parent, chain_t.child, {2, 1, 3}, {1, 2}, {1, 0, 2},
                                {0, 2, 1}, {1, 2}, {2, 1}, /*arena=*/true);  // :265-290
  std::printf("  KKT chain (arena workspace) residual %.3e\n", chain);
  CHECK(chain < 1e-9);
  const double siblings = kkt_case(0, branch_t.parent, branch_t.child, {2, 1, 3}, {1, 2},
                                   {1, 0, 1}, {1, 1, 0}, {2, 1}, {1, 2}, false);  // :292-314
  std::printf("  KKT sibling-edges residual %.3e\n", siblings);
  CHECK(siblings < 1e-9);
  const double zero_root = kkt_case(0, branch_t.parent, branch_t.child, {0, 1, 3}, {1, 2},
                                    {0, 0, 0}, {0, 0, 0}, {0, 0}, {0, 0}, false);  // :316-336
  std::printf("  KKT zero-dimensional-root residual %.3e\n", zero_root);
  CHECK(zero_root < 1e-9);
  const double schur = kkt_case(2, branch_t.parent, branch_t.child, {2, 1, 3}, {1, 2}, {1, 0, 1},
                                {0, 1, 1}, {1, 2}, {2, 1}, true, 6.0);  // :338-363
  std::printf("  KKT Schur-variables (theta_dim 2, arena) residual %.3e\n", schur);
  CHECK(schur < 1e-8);
}

static void test_input_validation() {  // :183-224
  const std::array<int, 3> state_dims = {2, 1, 3};
  const std::array<int, 2> control_dims = {1, 2};
  const std::array<int, 3> node_c_dims = {0, 1, 0}, node_g_dims = {1, 0, 2};
  const std::array<int, 2> edge_c_dims = {2, 1}, edge_g_dims = {1, 3};
  const Dimensions dimensions{2, state_dims.data(), control_dims.data(), node_c_dims.data(),
                              node_g_dims.data(), edge_c_dims.data(), edge_g_dims.data()};
  const ChainTopology chain_topology;
  const Topology chain{2, 0, chain_topology.parent.data(), chain_topology.child.data()};
  CHECK(validate_input(dimensions, chain) == InputValidationStatus::SUCCESS);
  const BranchTopology tree_topology;
  const Topology tree{2, 0, tree_topology.parent.data(), tree_topology.child.data()};
  CHECK(validate_input(dimensions, tree) == InputValidationStatus::SUCCESS);
  const std::array<int, 2> dag_parent = {0, 1}, dag_child = {2, 2};
  const Topology dag{2, 0, dag_parent.data(), dag_child.data()};
  CHECK(validate_input(dimensions, dag) == InputValidationStatus::INVALID_TOPOLOGY);
  const std::array<int, 2> negative_edge_c_dims = {-1, 1};
  const Dimensions invalid{2, state_dims.data(), control_dims.data(), node_c_dims.data(),
                           node_g_dims.data(), negative_edge_c_dims.data(), edge_g_dims.data()};
  CHECK(validate_input(invalid, tree) == InputValidationStatus::INVALID_DIMENSIONS);
  // reserve / set_uniform / mem_assign of the two structure types (lqr.hpp:12-18, 35-44)
  Dimensions d;
  d.reserve(2);
  d.set_uniform(2, 2, 1, 0, 0, 0, 0);
  CHECK(d.max_state_dim(3) == 2 && d.get_stagewise_kkt_dim(2) == 14);
  std::array<unsigned char, Topology::num_bytes(2) + Dimensions::num_bytes(2)> arena{};
  Topology t2;
  CHECK(t2.mem_assign(2, arena.data()) == Topology::num_bytes(2));
  t2.set_chain();
  Dimensions d2;
  CHECK(d2.mem_assign(2, arena.data() + Topology::num_bytes(2)) == Dimensions::num_bytes(2));
  d2.set_uniform(2, 2, 1, 0, 0, 0, 0);
  CHECK(validate_input(d2, t2) == InputValidationStatus::SUCCESS);
  d.free();
}

static void test_memory_sizes() {  // :226-263, and the values themselves
  constexpr int E = 2, n = 2, m = 1, nc = 1, ng = 2, ec = 3, eg = 1, p = 2;
  const std::array<int, 3> state_dims = {n, n, n}, node_c = {nc, nc, nc}, node_g = {ng, ng, ng};
  const std::array<int, 2> control_dims = {m, m}, edge_c = {ec, ec}, edge_g = {eg, eg};
  const Dimensions dimensions{p, state_dims.data(), control_dims.data(), node_c.data(),
                              node_g.data(), edge_c.data(), edge_g.data()};
  const BranchTopology tree;
  const Topology topology{E, 0, tree.parent.data(), tree.child.data()};
  CHECK(ModelCallbackOutput::num_bytes(n, m, E, nc, ng, ec, eg, p) ==
        ModelCallbackOutput::num_bytes(dimensions, topology));
  CHECK(Workspace::RegularizedLQRData::num_bytes(n, m, E, nc, ng, ec, eg, p) ==
        Workspace::RegularizedLQRData::num_bytes(dimensions, E));
  CHECK(LQR::Workspace::num_bytes(n, m, E) == LQR::Workspace::num_bytes(dimensions, topology));
  // The reference's constexpr formulas evaluated by hand at these dims (types.hpp:104-123,
  // 193-235; lqr.hpp:16-18, 38-40, 104-106, 146-184) on an LP64 target:
  static_assert(Topology::num_bytes(E) == 16 && Dimensions::num_bytes(E) == 60);
  static_assert(LQR::Output::num_bytes(E) == 64);
  static_assert(LQR::Workspace::num_bytes(n, m, E) == 820);
  static_assert(ModelCallbackOutput::num_bytes(n, m, E, nc, ng, ec, eg, p) ==
                3 * 96 + 2 * 176 + (3 * 31 + 2 * 58) * 8);
  static_assert(Workspace::RegularizedLQRData::num_bytes(n, m, E, nc, ng, ec, eg, p) == 2048);
  static_assert(ModelCallbackInput::num_bytes(E) == 3 * 32 + 2 * 64);
  // mem_assign consumes exactly num_bytes, for every view type
  std::vector<unsigned char> arena(Workspace::num_bytes(dimensions, topology));
  Workspace workspace;
  CHECK(workspace.mem_assign(dimensions, topology, arena.data()) == static_cast<int>(arena.size()));
  CHECK(workspace.x_state_offsets[2] == 2 * (n + m) && workspace.z_edge_offsets[1] == 3 * ng + eg);
  CHECK(workspace.ddyn_dx[1] == workspace.model_callback_output.edges[1].ddyn_dx);
  LQR::Output out;
  std::vector<unsigned char> small(LQR::Output::num_bytes(E));
  CHECK(out.mem_assign(E, small.data()) == LQR::Output::num_bytes(E));
}

// The model_callback lambda of sip_optimal_control.cpp:13-127 (evaluate_model) against the
// reference's own loops restated here, on the sibling-edges structure with theta.
static void test_model_evaluation() {
  const BranchTopology bt;
  const std::array<int, 3> state_dims = {2, 1, 3}, node_c = {1, 0, 1}, node_g = {0, 1, 1};
  const std::array<int, 2> control_dims = {1, 2}, edge_c = {1, 2}, edge_g = {2, 1};
  const int p = 2;
  const std::array<double, 2> x_init = {0.25, -0.5};
  int calls = 0;
  Input input{
      .dimensions = {p, state_dims.data(), control_dims.data(), node_c.data(), node_g.data(),
                     edge_c.data(), edge_g.data()},
      .topology = {2, 0, bt.parent.data(), bt.child.data()},
      .initial_state = x_init.data(),
      .model_callback = {},
      .timeout_callback = []() { return false; },
  };
  // values that depend on the views, so that wrong view pointers show up in the result
  input.model_callback = [&](const ModelCallbackInput &in, ModelCallbackOutput &out) {
    ++calls;
    const Dimensions &d = input.dimensions;
    for (int node = 0; node < 3; ++node) {
      auto &o = out.nodes[node];
      const int n = d.get_state_dim(node);
      o.f = 0.5 + node;
      for (int i = 0; i < n; ++i) {
        o.f += in.nodes[node].state[i] * in.nodes[node].state[i];
        o.df_dx[i] = 2.0 * in.nodes[node].state[i] + 0.01 * node;
      }
      for (int i = 0; i < p; ++i) o.df_dtheta[i] = in.theta[i] * (node + 1);
      for (int i = 0; i < d.get_node_c_dim(node); ++i)
        o.c[i] = 1.5 * node + i + in.nodes[node].equality_constraint_multipliers[i];
      for (int i = 0; i < d.get_node_g_dim(node); ++i)
        o.g[i] = -2.0 * node - i + in.nodes[node].inequality_constraint_multipliers[i];
    }
    for (int e = 0; e < 2; ++e) {
      auto &o = out.edges[e];
      const int np = d.get_state_dim(in.edges[e].parent), nc = d.get_state_dim(in.edges[e].child);
      o.f = 0.125 * (e + 1);
      for (int i = 0; i < np; ++i) o.df_dx[i] = 0.3 * in.edges[e].parent_state[i] - e;
      for (int i = 0; i < d.get_control_dim(e); ++i) {
        o.f += in.edges[e].control[i];
        o.df_du[i] = 1.0 + 0.5 * in.edges[e].control[i];
      }
      for (int i = 0; i < p; ++i) o.df_dtheta[i] = -in.theta[i] * (e + 2);
      for (int i = 0; i < nc; ++i)
        o.dyn_res[i] = in.edges[e].child_state[i] - 0.9 * in.edges[e].costate[i];
      for (int i = 0; i < d.get_edge_c_dim(e); ++i)
        o.c[i] = 7.0 + e + i + in.edges[e].equality_constraint_multipliers[i];
      for (int i = 0; i < d.get_edge_g_dim(e); ++i)
        o.g[i] = -7.0 - e - i + in.edges[e].inequality_constraint_multipliers[i];
    }
  };
  Workspace workspace;
  workspace.reserve(input.dimensions, input.topology);
  CallbackProvider provider(input, workspace);
  std::vector<double> x(workspace.x_dim), y(workspace.y_dim), z(workspace.z_dim);
  fill_sequence(x.data(), workspace.x_dim, 0.07);
  fill_sequence(y.data(), workspace.y_dim, -0.03);
  fill_sequence(z.data(), workspace.z_dim, 0.011);
  std::fill_n(workspace.gradient_f, workspace.x_dim, -1.0);
  std::fill_n(workspace.c, workspace.y_dim, -1.0);
  std::fill_n(workspace.g, workspace.z_dim, -1.0);
  CHECK(evaluate_model(input, workspace, {x.data(), y.data(), z.data(), true}) == 0);
  CHECK(calls == 1);

  // sip_optimal_control.cpp:44-123 on the same callback output
  const Dimensions &d = input.dimensions;
  const Topology &t = input.topology;
  const auto &mco = workspace.model_callback_output;
  double f = 0.0;
  for (int node = 0; node < 3; ++node) f += mco.nodes[node].f;
  for (int e = 0; e < 2; ++e) f += mco.edges[e].f;
  std::vector<double> grad(workspace.x_dim, 0.0), c(workspace.y_dim, 0.0), g(workspace.z_dim, 0.0);
  for (int node = 0; node < 3; ++node) {
    for (int r = 0; r < d.get_state_dim(node); ++r)
      grad[workspace.x_state_offsets[node] + r] += mco.nodes[node].df_dx[r];
    for (int r = 0; r < p; ++r) grad[workspace.stagewise_x_dim + r] += mco.nodes[node].df_dtheta[r];
  }
  for (int e = 0; e < 2; ++e) {
    const int parent = t.edge_parents[e];
    for (int r = 0; r < d.get_state_dim(parent); ++r)
      grad[workspace.x_state_offsets[parent] + r] += mco.edges[e].df_dx[r];
    for (int r = 0; r < d.get_control_dim(e); ++r)
      grad[workspace.x_control_offsets[e] + r] += mco.edges[e].df_du[r];
    for (int r = 0; r < p; ++r) grad[workspace.stagewise_x_dim + r] += mco.edges[e].df_dtheta[r];
  }
  for (int r = 0; r < d.get_state_dim(t.root); ++r)
    c[workspace.y_dyn_offsets[t.root] + r] = x_init[r] - x[workspace.x_state_offsets[t.root] + r];
  for (int node = 0; node < 3; ++node)
    std::copy_n(mco.nodes[node].c, d.get_node_c_dim(node), c.data() + workspace.y_node_c_offsets[node]);
  for (int e = 0; e < 2; ++e) {
    const int child = t.edge_children[e];
    std::copy_n(mco.edges[e].dyn_res, d.get_state_dim(child), c.data() + workspace.y_dyn_offsets[child]);
    std::copy_n(mco.edges[e].c, d.get_edge_c_dim(e), c.data() + workspace.y_edge_c_offsets[e]);
  }
  for (int node = 0; node < 3; ++node)
    std::copy_n(mco.nodes[node].g, d.get_node_g_dim(node), g.data() + workspace.z_node_offsets[node]);
  for (int e = 0; e < 2; ++e)
    std::copy_n(mco.edges[e].g, d.get_edge_g_dim(e), g.data() + workspace.z_edge_offsets[e]);

  CHECK(workspace.f == f);
  CHECK(std::equal(grad.begin(), grad.end(), workspace.gradient_f));
  CHECK(std::equal(c.begin(), c.end(), workspace.c));
  CHECK(std::equal(g.begin(), g.end(), workspace.g));
  // new_x == false: only the objective is refreshed (sip_optimal_control.cpp:52)
  workspace.gradient_f[0] = 123.0;
  workspace.f = 0.0;
  CHECK(evaluate_model(input, workspace, {x.data(), y.data(), z.data(), false}) == 0);
  CHECK(workspace.f == f && workspace.gradient_f[0] == 123.0);
  workspace.free(input.topology);
}

int main() {
  const std::pair<const char *, std::function<void()>> tests[] = {
      {"LQRFactor status API", test_factor_status_api},
      {"LQRSolve nonuniform-delta chain + topology reuse", test_nonuniform_delta_chain},
      {"LQRSolve branching tree", test_branching_tree},
      {"LQRFactor rejects invalid topology", test_rejects_invalid_topology},
      {"CallbackProvider KKT cases", test_callback_provider},
      {"InputValidation + structure allocation modes", test_input_validation},
      {"Workspace memory sizes (static == dynamic == reference values)", test_memory_sizes},
      {"Model evaluation scatter (sip_optimal_control.cpp:13-127)", test_model_evaluation},
  };
  for (const auto &t : tests) {
    const int before = g_failures;
    std::printf("[ RUN  ] %s\n", t.first);
    t.second();
    std::printf("[ %s ] %s\n", g_failures == before ? " OK " : "FAIL", t.first);
  }
  std::printf("%d check(s) failed\n", g_failures);
  return g_failures == 0 ? 0 : 1;
}
