#include "structure.hpp"

#include <algorithm>

namespace sipoc {

namespace {
int dim_or_zero(const int *d, int i) { return d == nullptr ? 0 : d[i]; }

std::vector<int> prefix(const std::vector<int> &v) {
  std::vector<int> o(v.size() + 1, 0);
  for (size_t i = 0; i < v.size(); ++i) o[i + 1] = o[i] + v[i];
  return o;
}
}  // namespace

sipoc_error HostStructure::build(const sipoc_structure &s, std::string &err) {
  E = s.num_edges;
  N = E + 1;
  root = s.root;
  // Dimension checks first, then topology (types.cpp:72-97).
  if (E < 0 || s.theta_dim < 0) {
    err = "negative num_edges or theta_dim";
    return SIPOC_INVALID_DIMENSIONS;
  }
  if (s.theta_dim > kMaxThetaDim) {
    err = "theta_dim above the engine's limit (32)";
    return SIPOC_UNSUPPORTED;
  }
  theta_dim = s.theta_dim;
  if (s.state_dims == nullptr || (E > 0 && s.control_dims == nullptr)) {
    err = "state_dims / control_dims must not be NULL";
    return SIPOC_INVALID_ARGUMENT;
  }
  n.assign(s.state_dims, s.state_dims + N);
  m.assign(s.control_dims, s.control_dims + E);
  node_c.resize(N);
  node_g.resize(N);
  edge_c.resize(E);
  edge_g.resize(E);
  for (int i = 0; i < N; ++i) {
    node_c[i] = dim_or_zero(s.node_c_dims, i);
    node_g[i] = dim_or_zero(s.node_g_dims, i);
    if (n[i] < 0 || node_c[i] < 0 || node_g[i] < 0) {
      err = "negative node dimension";
      return SIPOC_INVALID_DIMENSIONS;
    }
  }
  for (int e = 0; e < E; ++e) {
    edge_c[e] = dim_or_zero(s.edge_c_dims, e);
    edge_g[e] = dim_or_zero(s.edge_g_dims, e);
    if (m[e] < 0 || edge_c[e] < 0 || edge_g[e] < 0) {
      err = "negative edge dimension";
      return SIPOC_INVALID_DIMENSIONS;
    }
  }

  if (E > 0 && (s.edge_parents == nullptr || s.edge_children == nullptr)) {
    err = "edge_parents / edge_children must not be NULL";
    return SIPOC_INVALID_TOPOLOGY;
  }
  if (root < 0 || root >= N) {
    err = "root out of range";
    return SIPOC_INVALID_TOPOLOGY;
  }
  parents.assign(s.edge_parents, s.edge_parents + E);
  children.assign(s.edge_children, s.edge_children + E);

  // CSR children in edge order, then an explicit-stack DFS (lqr.cpp:576-628).
  child_offsets.assign(N + 1, 0);
  in_edge.assign(N, -1);
  for (int e = 0; e < E; ++e) {
    const int p = parents[e], c = children[e];
    if (p < 0 || p >= N || c < 0 || c >= N || p == c) {
      err = "edge endpoint out of range or self loop";
      return SIPOC_INVALID_TOPOLOGY;
    }
    if (c == root || in_edge[c] != -1) {  // in-degree rule, types.cpp:108-118
      err = "a node has more than one incoming edge (or the root has one)";
      return SIPOC_INVALID_TOPOLOGY;
    }
    in_edge[c] = e;
    ++child_offsets[p + 1];
  }
  for (int i = 0; i < N; ++i) child_offsets[i + 1] += child_offsets[i];
  child_edges.assign(E, 0);
  {
    std::vector<int> cur(child_offsets.begin(), child_offsets.end() - 1);
    for (int e = 0; e < E; ++e) child_edges[cur[parents[e]]++] = e;
  }
  preorder.clear();
  preorder.reserve(N);
  {
    std::vector<int> stack{root};
    std::vector<char> seen(N, 0);
    while (!stack.empty()) {
      const int node = stack.back();
      stack.pop_back();
      if (seen[node]) {
        err = "topology is not a tree (node reached twice)";
        return SIPOC_INVALID_TOPOLOGY;
      }
      seen[node] = 1;
      preorder.push_back(node);
      for (int ci = child_offsets[node + 1] - 1; ci >= child_offsets[node]; --ci)
        stack.push_back(children[child_edges[ci]]);
    }
    if (static_cast<int>(preorder.size()) != N) {
      err = "topology is disconnected or cyclic";
      return SIPOC_INVALID_TOPOLOGY;
    }
  }
  postorder.assign(preorder.rbegin(), preorder.rend());

  // Flat per-problem offsets.
  nn_off.assign(N + 1, 0);
  n_off.assign(N + 1, 0);
  max_n = 0;
  for (int i = 0; i < N; ++i) {
    nn_off[i + 1] = nn_off[i] + n[i] * n[i];
    n_off[i + 1] = n_off[i] + n[i];
    max_n = std::max(max_n, n[i]);
  }
  nm_off.assign(E + 1, 0);
  mm_off.assign(E + 1, 0);
  m_off.assign(E + 1, 0);
  a_off.assign(E + 1, 0);
  b_off.assign(E + 1, 0);
  w_off.assign(E + 1, 0);
  k_off.assign(E + 1, 0);
  hxx_edge_off.assign(E + 1, 0);
  max_m = 0;
  for (int e = 0; e < E; ++e) {
    const int np = n[parents[e]], nc = n[children[e]], me = m[e];
    nm_off[e + 1] = nm_off[e] + np * me;
    mm_off[e + 1] = mm_off[e] + me * me;
    m_off[e + 1] = m_off[e] + me;
    a_off[e + 1] = a_off[e] + nc * np;
    b_off[e + 1] = b_off[e] + nc * me;
    w_off[e + 1] = w_off[e] + nc * nc;
    k_off[e + 1] = k_off[e] + me * np;
    hxx_edge_off[e + 1] = hxx_edge_off[e] + np * np;
    max_m = std::max(max_m, me);
  }

  // Wire-format offsets (types.cpp:24-64).  State i and control i interleave
  // BY INDEX, also on trees.
  x_state.assign(N, 0);
  x_control.assign(E, 0);
  int off = 0;
  for (int i = 0; i < N; ++i) {
    x_state[i] = off;
    if (i < E) {
      off += n[i];
      x_control[i] = off;
      off += m[i];
    }
  }
  sx_dim = n[E];
  for (int e = 0; e < E; ++e) sx_dim += n[e] + m[e];
  x_dim = sx_dim + theta_dim;  // theta closes the x vector (types.cpp:24-64)
  y_dyn.assign(N, 0);
  y_node_c.assign(N, 0);
  y_edge_c.assign(E, 0);
  off = 0;
  for (int i = 0; i < N; ++i) {
    y_dyn[i] = off;
    off += n[i];
    y_node_c[i] = off;
    off += node_c[i];
  }
  for (int e = 0; e < E; ++e) {
    y_edge_c[e] = off;
    off += edge_c[e];
  }
  y_dim = off;
  z_node.assign(N, 0);
  z_edge.assign(E, 0);
  off = 0;
  for (int i = 0; i < N; ++i) {
    z_node[i] = off;
    off += node_g[i];
  }
  for (int e = 0; e < E; ++e) {
    z_edge[e] = off;
    off += edge_g[e];
  }
  z_dim = off;
  kkt_dim = x_dim + y_dim + z_dim;

  jc_node_off.assign(N + 1, 0);
  jg_node_off.assign(N + 1, 0);
  for (int i = 0; i < N; ++i) {
    jc_node_off[i + 1] = jc_node_off[i] + node_c[i] * n[i];
    jg_node_off[i + 1] = jg_node_off[i] + node_g[i] * n[i];
  }
  jcx_off.assign(E + 1, 0);
  jcu_off.assign(E + 1, 0);
  jgx_off.assign(E + 1, 0);
  jgu_off.assign(E + 1, 0);
  for (int e = 0; e < E; ++e) {
    const int np = n[parents[e]];
    jcx_off[e + 1] = jcx_off[e] + edge_c[e] * np;
    jcu_off[e + 1] = jcu_off[e] + edge_c[e] * m[e];
    jgx_off[e + 1] = jgx_off[e] + edge_g[e] * np;
    jgu_off[e + 1] = jgu_off[e] + edge_g[e] * m[e];
  }
  pn_off.assign(E + 1, 0);
  cn_off.assign(E + 1, 0);
  for (int e = 0; e < E; ++e) {
    pn_off[e + 1] = pn_off[e] + n[parents[e]];
    cn_off[e + 1] = cn_off[e] + n[children[e]];
  }
  node_c_off = prefix(node_c);
  node_g_off = prefix(node_g);
  edge_c_off = prefix(edge_c);
  edge_g_off = prefix(edge_g);
  has_constraints = (node_c_off[N] + node_g_off[N] + edge_c_off[E] + edge_g_off[E]) > 0;

  is_chain = (root == 0);
  for (int e = 0; e < E && is_chain; ++e)
    is_chain = parents[e] == e && children[e] == e + 1;
  is_uniform = true;
  for (int i = 1; i < N; ++i) is_uniform = is_uniform && n[i] == n[0];
  for (int e = 1; e < E; ++e) is_uniform = is_uniform && m[e] == m[0];
  return SIPOC_OK;
}

std::vector<int> HostStructure::serialise(DevTables &t) const {
  std::vector<int> buf;
  auto put = [&buf](const std::vector<int> &v) -> const int * {
    const size_t at = buf.size();
    buf.insert(buf.end(), v.begin(), v.end());
    if (v.empty()) buf.push_back(0);  // keep offsets distinct and in range
    return reinterpret_cast<const int *>(at * sizeof(int));
  };
  t.E = E;
  t.N = N;
  t.root = root;
  t.max_n = max_n;
  t.max_m = max_m;
  t.parents = put(parents);
  t.children = put(children);
  t.n = put(n);
  t.m = put(m);
  t.child_offsets = put(child_offsets);
  t.child_edges = put(child_edges);
  t.preorder = put(preorder);
  t.postorder = put(postorder);
  t.in_edge = put(in_edge);
  t.nn_off = put(nn_off);
  t.n_off = put(n_off);
  t.nm_off = put(nm_off);
  t.mm_off = put(mm_off);
  t.m_off = put(m_off);
  t.a_off = put(a_off);
  t.b_off = put(b_off);
  t.w_off = put(w_off);
  t.k_off = put(k_off);
  t.hxx_edge_off = put(hxx_edge_off);
  t.node_c = put(node_c);
  t.node_g = put(node_g);
  t.edge_c = put(edge_c);
  t.edge_g = put(edge_g);
  t.x_state = put(x_state);
  t.y_dyn = put(y_dyn);
  t.y_node_c = put(y_node_c);
  t.z_node = put(z_node);
  t.x_control = put(x_control);
  t.y_edge_c = put(y_edge_c);
  t.z_edge = put(z_edge);
  t.jc_node_off = put(jc_node_off);
  t.jg_node_off = put(jg_node_off);
  t.jcx_off = put(jcx_off);
  t.jcu_off = put(jcu_off);
  t.jgx_off = put(jgx_off);
  t.jgu_off = put(jgu_off);
  t.node_c_off = put(node_c_off);
  t.node_g_off = put(node_g_off);
  t.edge_c_off = put(edge_c_off);
  t.edge_g_off = put(edge_g_off);
  t.pn_off = put(pn_off);
  t.cn_off = put(cn_off);
  t.theta_dim = theta_dim;
  t.sx_dim = sx_dim;
  t.x_dim = x_dim;
  t.y_dim = y_dim;
  t.z_dim = z_dim;
  t.kkt_dim = kkt_dim;
  return buf;
}

}  // namespace sipoc
