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
//   InputValidation (DAG and negative dimension rejected)         tests/variable_dimensions_test.cpp:183-224
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "helpers.hpp"

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
    LQR lqr(in, ws);
    CHECK(lqr.factor_with_status() == LQR::FactorStatus::SUCCESS);
    CHECK(lqr.factor());
    ws.free();
  }
  {  // delta[T][0] = 0 -> INVALID_DELTA
    Problem p = identity_chain(2, 1, 2);
    p.delta[2][0] = 0.0;
    LQR::Input in = p.input();
    LQR::Workspace ws;
    LQR lqr(in, ws);
    CHECK(lqr.factor_with_status() == LQR::FactorStatus::INVALID_DELTA);
    CHECK(!lqr.factor());
    ws.free();
  }
  {  // n = m = T = 1, Q[T] = -2 -> F failure
    Problem p = identity_chain(1, 1, 1);
    p.Q[1][0] = -2.0;
    LQR::Input in = p.input();
    LQR::Workspace ws;
    LQR lqr(in, ws);
    CHECK(lqr.factor_with_status() == LQR::FactorStatus::F_FACTORIZATION_FAILURE);
    ws.free();
  }
  {  // Q[T] = 0, R = -1 -> G failure
    Problem p = identity_chain(1, 1, 1);
    p.Q[1][0] = 0.0;
    p.R[0][0] = -1.0;
    LQR::Input in = p.input();
    LQR::Workspace ws;
    LQR lqr(in, ws);
    CHECK(lqr.factor_with_status() == LQR::FactorStatus::G_FACTORIZATION_FAILURE);
    ws.free();
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
  LQR::Workspace ws;
  LQR lqr(in, ws);
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
  ws.free();
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
  LQR lqr(in, ws);
  CHECK(lqr.factor_with_status() == LQR::FactorStatus::SUCCESS);
  LQR::Output out = p.output();
  lqr.solve(out);
  const double res = residual_norm(p);
  std::printf("  branching tree residual %.3e\n", res);
  CHECK(res < 1e-12);
  ws.free();
}

static void test_rejects_invalid_topology() {  // tests/lqr_test.cpp:452-464
  Problem p(2, 0, {0, 0}, {1, 1}, {2, 2, 2}, {1, 1});  // two edges into node 1
  LQR::Input in = p.input();
  LQR::Workspace ws;
  LQR lqr(in, ws);
  CHECK(lqr.factor_with_status() == LQR::FactorStatus::INVALID_TOPOLOGY);
  CHECK(!lqr.factor());
  ws.free();
}

// ---- Newton-KKT (tests/variable_dimensions_test.cpp) ------------------------------------
static void fill_sequence(double *v, int size, double scale) {  // :46-50
  for (int i = 0; i < size; ++i) v[i] = scale * (i + 1);
}
static void identity_scaled(double *v, int n, double s) {
  for (int i = 0; i < n * n; ++i) v[i] = 0.0;
  for (int i = 0; i < n; ++i) v[i * n + i] = s;
}

static double kkt_case(int E, const std::vector<int> &parents, const std::vector<int> &children,
                       const std::vector<int> &n, const std::vector<int> &m,
                       const std::vector<int> &nc, const std::vector<int> &ng,
                       const std::vector<int> &ec, const std::vector<int> &eg) {
  Input input;
  input.topology.num_edges = E;
  input.topology.reserve(E);
  input.topology.set_tree(0, parents.data(), children.data());
  input.dimensions.reserve(E);
  auto fill = [](const int *dst, const std::vector<int> &src) {
    for (size_t i = 0; i < src.size(); ++i) const_cast<int *>(dst)[i] = src[i];
  };
  fill(input.dimensions.state_dims, n);
  fill(input.dimensions.control_dims, m);
  fill(input.dimensions.node_c_dims, nc);
  fill(input.dimensions.node_g_dims, ng);
  fill(input.dimensions.edge_c_dims, ec);
  fill(input.dimensions.edge_g_dims, eg);
  CHECK(validate_input(input.dimensions, input.topology) == InputValidationStatus::SUCCESS);

  Workspace workspace;
  workspace.reserve(input.dimensions, input.topology);
  auto &mco = workspace.model_callback_output;
  for (int node = 0; node <= E; ++node) {  // initialize_model, :77-133
    const int nn = n[node];
    fill_sequence(mco.nodes[node].dc_dx, nc[node] * nn, 0.013 * (node + 1));
    fill_sequence(mco.nodes[node].dg_dx, ng[node] * nn, -0.011 * (node + 1));
    identity_scaled(mco.nodes[node].d2L_dx2, nn, 2.5 + 0.2 * node);
  }
  for (int e = 0; e < E; ++e) {
    const int np = n[parents[e]], nch = n[children[e]], mm = m[e];
    fill_sequence(mco.edges[e].ddyn_dx, nch * np, 0.025 + 0.004 * e);
    fill_sequence(mco.edges[e].ddyn_du, nch * mm, -0.031 - 0.003 * e);
    fill_sequence(mco.edges[e].dc_dx, ec[e] * np, 0.017 * (e + 1));
    fill_sequence(mco.edges[e].dc_du, ec[e] * mm, 0.019 * (e + 1));
    fill_sequence(mco.edges[e].dg_dx, eg[e] * np, -0.014 * (e + 1));
    fill_sequence(mco.edges[e].dg_du, eg[e] * mm, 0.016 * (e + 1));
    identity_scaled(mco.edges[e].d2L_dx2, np, 0.3 + 0.05 * e);
    fill_sequence(mco.edges[e].d2L_dxdu, np * mm, 0.009 * (e + 1));
    identity_scaled(mco.edges[e].d2L_du2, mm, 3.0 + 0.2 * e);
  }

  // expect_kkt_solve, :135-181
  CallbackProvider callback_provider(input, workspace);
  const int x_dim = input.dimensions.get_x_dim(E), y_dim = input.dimensions.get_y_dim(E),
            z_dim = input.dimensions.get_z_dim(E), kkt_dim = x_dim + y_dim + z_dim;
  std::vector<double> w(z_dim + 1, 1.3), r2(y_dim + 1, 0.9), r3(z_dim + 1, 0.4), r1(x_dim + 1);
  fill_sequence(r1.data(), x_dim, 0.03);
  for (double &v : r1) v += 0.2;
  CHECK(callback_provider.factor(w.data(), r1.data(), r2.data(), r3.data()));
  std::vector<double> rhs(kkt_dim), solution(kkt_dim, 0.0);
  fill_sequence(rhs.data(), kkt_dim, 0.01);
  callback_provider.solve(rhs.data(), solution.data());
  std::vector<double> px(x_dim + 1, 0.0), py(y_dim + 1, 0.0), pz(z_dim + 1, 0.0);
  callback_provider.add_Kx_to_y(w.data(), r1.data(), r2.data(), r3.data(), solution.data(),
                                solution.data() + x_dim, solution.data() + x_dim + y_dim,
                                px.data(), py.data(), pz.data());
  {  // add_Kx_to_y is the sum of its five blocks and the diagonal terms (helpers.cpp:953-977)
    const double *sx = solution.data(), *sy = sx + x_dim, *sz = sy + y_dim;
    std::vector<double> qx(x_dim + 1, 0.0), qy(y_dim + 1, 0.0), qz(z_dim + 1, 0.0);
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
  workspace.free(input.topology);
  input.topology.free();
  input.dimensions.free();
  return std::sqrt(sq);
}

static void test_callback_provider() {
  const double chain = kkt_case(2, {0, 1}, {1, 2}, {2, 1, 3}, {1, 2}, {1, 0, 2}, {0, 2, 1},
                                {1, 2}, {2, 1});  // :265-290
  std::printf("  KKT chain residual %.3e\n", chain);
  CHECK(chain < 1e-9);
  const double siblings = kkt_case(2, {0, 0}, {1, 2}, {2, 1, 3}, {1, 2}, {1, 0, 1}, {1, 1, 0},
                                   {2, 1}, {1, 2});  // :292-314
  std::printf("  KKT sibling-edges residual %.3e\n", siblings);
  CHECK(siblings < 1e-9);
  const double zero_root = kkt_case(2, {0, 0}, {1, 2}, {0, 1, 3}, {1, 2}, {0, 0, 0}, {0, 0, 0},
                                    {0, 0}, {0, 0});  // :316-336
  std::printf("  KKT zero-dimensional-root residual %.3e\n", zero_root);
  CHECK(zero_root < 1e-9);
}

static void test_input_validation() {  // :183-224
  Topology dag;
  dag.num_edges = 2;
  dag.reserve(2);
  const int parents[2] = {0, 1}, children[2] = {1, 1};
  dag.set_tree(0, parents, children);
  Dimensions d;
  d.set_uniform(2, 2, 1, 0, 0, 0, 0);
  CHECK(validate_input(d, dag) == InputValidationStatus::INVALID_TOPOLOGY);
  dag.set_chain();
  CHECK(validate_input(d, dag) == InputValidationStatus::SUCCESS);
  const_cast<int *>(d.state_dims)[1] = -1;
  CHECK(validate_input(d, dag) == InputValidationStatus::INVALID_DIMENSIONS);
  dag.free();
  d.free();
}

int main() {
  const std::pair<const char *, std::function<void()>> tests[] = {
      {"LQRFactor status API", test_factor_status_api},
      {"LQRSolve nonuniform-delta chain + topology reuse", test_nonuniform_delta_chain},
      {"LQRSolve branching tree", test_branching_tree},
      {"LQRFactor rejects invalid topology", test_rejects_invalid_topology},
      {"CallbackProvider KKT cases", test_callback_provider},
      {"InputValidation", test_input_validation},
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
