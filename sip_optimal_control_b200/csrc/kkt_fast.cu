// Newton-KKT -> LQR reduction for uniform chains (n, m compile-time; constraint
// dimensions per node / edge stay runtime): the prologue of the factorization
// (reference helpers.cpp:242-360).
//
// One CTA = 32 problems x (n + m) columns of the stage Hessian [Q M; M' R] of one
// node: thread (problem, column j) accumulates its column of the weighted Jacobian
// Gram products in REGISTERS and writes it once (the generic kernel read-modify-
// writes HBM per product).  The 1/r2 and 1/(w + r3) weights of the node and of its
// child edge are computed once per CTA into shared memory (and into the workspace
// the rhs build / dual recovery read later).  The column-threads of a problem sit in
// different warps of the same CTA: each copies its own column of every Jacobian row into
// shared memory ([row][column][problem], so a row's entries sit at compile-time offsets
// from one pointer) and the Gram loops read them from there; the loops are instantiated
// per column so that only the lower-triangle iterations exist.  All global accesses
// are 256 B coalesced across the 32 problems of a warp.  Exact-zero Jacobian entries are
// not skipped (adding the zero products changes nothing for finite data).
#include "generic_kernels.cuh"
#include "kkt_fast.cuh"

namespace sipoc {
namespace {

// Phase B of kkt_reduce_chain for one column, J = the column index (warp-uniform, so the
// dispatch below costs no divergence and the unrolled loops hold live iterations only).
// Jrow points at this lane's entry of Jacobian row 0, column 0: [row][column][32 lanes].
template <int N, int M, int J>
__device__ __forceinline__ void reduce_state_column(const DevTables &t, const KktModel &mdl,
                                                    const double *__restrict__ r1,
                                                    const KktWs &ws, const double *wrow,
                                                    const double *Jrow, int rows, int node,
                                                    bool has_edge, size_t L, int64_t b) {
  constexpr int NZ = N + M;
  constexpr int LIVE = N - J;  // rows J .. N - 1 of column J (lower triangle)
#define LD(ptr, off) __ldg((ptr) + static_cast<size_t>(off) * L + b)
  const int qo = t.nn_off[node], e = node;
  double accq[LIVE], accm[M];
#pragma unroll
  for (int i = 0; i < LIVE; ++i) accq[i] = LD(mdl.node_hxx, qo + (J + i) + J * N);
#pragma unroll
  for (int a = 0; a < M; ++a) accm[a] = 0.0;
  if (has_edge) {
#pragma unroll
    for (int i = 0; i < LIVE; ++i) accq[i] += LD(mdl.edge_hxx, t.hxx_edge_off[e] + (J + i) + J * N);
#pragma unroll
    for (int a = 0; a < M; ++a) accm[a] = LD(mdl.edge_hxu, t.nm_off[e] + J + a * N);
  }
  accq[0] += LD(r1, t.x_state[node] + J);
  // J' diag(weights) J over the node's and the child edge's constraint rows
  for (int q = 0; q < rows; ++q) {
    const double *row = Jrow + q * (NZ * 32);
    const double wj = wrow[q * 32] * row[J * 32];
#pragma unroll
    for (int i = 0; i < LIVE; ++i) accq[i] += wj * row[(J + i) * 32];
#pragma unroll
    for (int a = 0; a < M; ++a) accm[a] += wj * row[(N + a) * 32];  // zero on node rows
  }
#pragma unroll
  for (int i = 0; i < LIVE; ++i) {
    ws.Q_mod[static_cast<size_t>(qo + (J + i) + J * N) * L + b] = accq[i];
    if (i != 0) ws.Q_mod[static_cast<size_t>(qo + J + (J + i) * N) * L + b] = accq[i];  // mirror
  }
  if (has_edge) {
#pragma unroll
    for (int a = 0; a < M; ++a)
      ws.M_mod[static_cast<size_t>(t.nm_off[e] + J + a * N) * L + b] = accm[a];
  }
#undef LD
}

template <int N, int M, int A>
__device__ __forceinline__ void reduce_control_column(const DevTables &t, const KktModel &mdl,
                                                      const double *__restrict__ r1,
                                                      const KktWs &ws, const double *wrow,
                                                      const double *Jrow, int first_row, int rows,
                                                      int e, size_t L, int64_t b) {
  constexpr int NZ = N + M;
  constexpr int LIVE = M - A;
#define LD(ptr, off) __ldg((ptr) + static_cast<size_t>(off) * L + b)
  const int ro = t.mm_off[e];
  double accr[LIVE];
#pragma unroll
  for (int i = 0; i < LIVE; ++i) accr[i] = LD(mdl.edge_huu, ro + (A + i) + A * M);
  accr[0] += LD(r1, t.x_control[e] + A);
  for (int q = first_row; q < rows; ++q) {  // edge rows only: node rows have no control part
    const double *row = Jrow + q * (NZ * 32);
    const double wa = wrow[q * 32] * row[(N + A) * 32];
#pragma unroll
    for (int i = 0; i < LIVE; ++i) accr[i] += wa * row[(N + A + i) * 32];
  }
#pragma unroll
  for (int i = 0; i < LIVE; ++i) {
    ws.R_mod[static_cast<size_t>(ro + (A + i) + A * M) * L + b] = accr[i];
    if (i != 0) ws.R_mod[static_cast<size_t>(ro + A + (A + i) * M) * L + b] = accr[i];
  }
#undef LD
}

template <int N, int M, int J, typename... Args>
__device__ __forceinline__ void dispatch_state_column(int col, const Args &...args) {
  if (col == J) {
    reduce_state_column<N, M, J>(args...);
  } else if constexpr (J + 1 < N) {
    dispatch_state_column<N, M, J + 1>(col, args...);
  }
}
template <int N, int M, int A, typename... Args>
__device__ __forceinline__ void dispatch_control_column(int a, const Args &...args) {
  if (a == A) {
    reduce_control_column<N, M, A>(args...);
  } else if constexpr (A + 1 < M) {
    dispatch_control_column<N, M, A + 1>(a, args...);
  }
}

template <int N, int M>
__global__ void __launch_bounds__(32 * (N + M))
kkt_reduce_chain(DevTables t, KktModel mdl, const double *__restrict__ w,
                 const double *__restrict__ r1, const double *__restrict__ r2,
                 const double *__restrict__ r3, KktWs ws, int *ok, int64_t batch, int64_t ld,
                 int max_rows) {
  constexpr int NZ = N + M;
  extern __shared__ double wsm[];       // weights [row][32 problems], then
  double *jsm = wsm + max_rows * 32;    // Jacobian rows [row][column][32 problems]
  const int lane = threadIdx.x & 31, col = threadIdx.x >> 5;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * 32 + lane;  // < ld by construction
  const int node = blockIdx.y;
  const bool has_edge = node < t.E;
  const int e = node;  // chain: the child edge of node k is edge k
  const size_t L = static_cast<size_t>(ld);
#define LD(ptr, off) __ldg((ptr) + static_cast<size_t>(off) * L + b)

  const int c_n = t.node_c[node], g_n = t.node_g[node];
  const int c_e = has_edge ? t.edge_c[e] : 0, g_e = has_edge ? t.edge_g[e] : 0;
  const int o_gn = c_n, o_ce = c_n + g_n, o_ge = o_ce + c_e, rows = o_ge + g_e;

  // ---- phase A: weights and the dynamics regularization (helpers.cpp:251-295) ----
  bool good = true;
  for (int q = col; q < rows + N; q += NZ) {
    if (q < rows) {
      double reg;
      double *dst;
      if (q < o_gn) {
        reg = LD(r2, t.y_node_c[node] + q);
        dst = ws.node_c_r2_inv + static_cast<size_t>(t.node_c_off[node] + q) * L + b;
      } else if (q < o_ce) {
        const int o = t.z_node[node] + (q - o_gn);
        reg = LD(w, o) + LD(r3, o);
        dst = ws.node_mod_w_inv + static_cast<size_t>(t.node_g_off[node] + q - o_gn) * L + b;
      } else if (q < o_ge) {
        reg = LD(r2, t.y_edge_c[e] + (q - o_ce));
        dst = ws.edge_c_r2_inv + static_cast<size_t>(t.edge_c_off[e] + q - o_ce) * L + b;
      } else {
        const int o = t.z_edge[e] + (q - o_ge);
        reg = LD(w, o) + LD(r3, o);
        dst = ws.edge_mod_w_inv + static_cast<size_t>(t.edge_g_off[e] + q - o_ge) * L + b;
      }
      good = good && (reg > 0.0);
      const double inv = 1.0 / reg;
      *dst = inv;
      wsm[q * 32 + lane] = inv;
    } else {
      const int row = q - rows;
      const double reg = LD(r2, t.y_dyn[node] + row);
      good = good && (reg > 0.0);
      ws.dyn_r2[static_cast<size_t>(t.n_off[node] + row) * L + b] = reg;
    }
  }
  if (!good && b < batch) ok[b] = 0;  // benign race: every writer stores 0

  // ---- this thread's column of every Jacobian row -> shared memory (each entry is read
  // from HBM once; node rows have no control part) ----
  {
    double *dst = jsm + col * 32 + lane;
    const bool state = col < N;
    const int a = col - N;
    for (int k = 0; k < c_n; ++k)
      dst[k * (NZ * 32)] = state ? LD(mdl.node_jc, t.jc_node_off[node] + k + col * c_n) : 0.0;
    for (int k = 0; k < g_n; ++k)
      dst[(o_gn + k) * (NZ * 32)] =
          state ? LD(mdl.node_jg, t.jg_node_off[node] + k + col * g_n) : 0.0;
    for (int k = 0; k < c_e; ++k)
      dst[(o_ce + k) * (NZ * 32)] = state ? LD(mdl.edge_jcx, t.jcx_off[e] + k + col * c_e)
                                          : LD(mdl.edge_jcu, t.jcu_off[e] + k + a * c_e);
    for (int k = 0; k < g_e; ++k)
      dst[(o_ge + k) * (NZ * 32)] = state ? LD(mdl.edge_jgx, t.jgx_off[e] + k + col * g_e)
                                          : LD(mdl.edge_jgu, t.jgu_off[e] + k + a * g_e);
  }
  __syncthreads();

  // ---- phase B: one column of [Q M; M' R] per thread (helpers.cpp:297-360) ----
  const double *wrow = wsm + lane, *Jrow = jsm + lane;
  if (col < N) {
    dispatch_state_column<N, M, 0>(col, t, mdl, r1, ws, wrow, Jrow, rows, node, has_edge, L, b);
  } else if (has_edge) {
    dispatch_control_column<N, M, 0>(col - N, t, mdl, r1, ws, wrow, Jrow, o_ce, rows, e, L, b);
  }
#undef LD
}

template <int N, int M>
void launch(const DevTables &t, const KktModel &m, const double *w, const double *r1,
            const double *r2, const double *r3, const KktWs &ws, int *ok, int64_t batch,
            int64_t ld, int max_rows, cudaStream_t s) {
  launch_fill_int(ok, 1, ld, s);
  dim3 grid(static_cast<unsigned>(ld / 32), static_cast<unsigned>(t.N));
  const int rows = max_rows > 0 ? max_rows : 1;
  const size_t smem = kkt_reduce_smem_bytes(N, M, rows);
  auto kern = kkt_reduce_chain<N, M>;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  kern<<<grid, 32 * (N + M), smem, s>>>(t, m, w, r1, r2, r3, ws, ok, batch, ld, rows);
}

// y += K x for uniform chains (helpers.cpp:953-1368, theta == 0): thread (problem, node)
// owns the rows of the node's state, of its child edge's control, of the node / edge
// constraints and of the CHILD's dynamics, so that every model block of the node and of
// its child edge is read exactly once: each Jacobian / dynamics entry feeds both its C x
// (or G x) row and its C' y (or G' z) column while it is in a register, the cross Hessian
// feeds the state and the control rows.  x, u, the child's costate and the accumulators
// live in registers (the generic kernel re-reads the vectors per row and every
// off-diagonal block twice: 6.5 GB of DRAM reads per launch against 2.6 GB here at
// n = 12, m = 4, c = 6, g = 8, batch 8 192).  The sums run in a different order than the
// reference's row loops; the operator is a residual check, not part of the solve.
template <int N, int M>
__global__ void __launch_bounds__(128)
kkt_apply_chain(DevTables t, KktModel mdl, const double *__restrict__ w,
                const double *__restrict__ r1, const double *__restrict__ r2,
                const double *__restrict__ r3, const double *__restrict__ in,
                double *__restrict__ out, int64_t batch, int64_t ld) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int node = blockIdx.y;
  const bool has_edge = node < t.E;
  const int e = node, child = node + 1;  // chain: edge k joins node k to node k + 1
  const size_t L = static_cast<size_t>(ld);
  const size_t oy = static_cast<size_t>(t.x_dim), oz = oy + static_cast<size_t>(t.y_dim);
#define LD(ptr, off) __ldg((ptr) + static_cast<size_t>(off) * L + b)
#define ACC(off, v) out[static_cast<size_t>(off) * L + b] += (v)

  const int xs = t.x_state[node];
  double x[N], sx[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    x[i] = LD(in, xs + i);
    sx[i] = LD(r1, xs + i) * x[i] - LD(in, oy + t.y_dyn[node] + i);  // R1 x, -I' y_dyn
  }
  if (node == 0) {  // root row of C: -x_root - R2 y
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int o = t.y_dyn[0] + i;
      ACC(oy + o, -x[i] - LD(r2, o) * LD(in, oy + o));
    }
  }
  {  // node Hessian
    const int ho = t.nn_off[node];
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
      for (int i = 0; i < N; ++i) sx[i] += LD(mdl.node_hxx, ho + i + j * N) * x[j];
  }
  {  // node constraints: rows of C / G and columns of C' / G' from one read
    const int c = t.node_c[node], g = t.node_g[node];
    const int jo = t.jc_node_off[node], go = t.jg_node_off[node];
    for (int k = 0; k < c; ++k) {
      const int o = t.y_node_c[node] + k;
      const double yk = LD(in, oy + o);
      double acc = -LD(r2, o) * yk;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const double J = LD(mdl.node_jc, jo + k + j * c);
        sx[j] += J * yk;
        acc += J * x[j];
      }
      ACC(oy + o, acc);
    }
    for (int k = 0; k < g; ++k) {
      const int o = t.z_node[node] + k;
      const double zk = LD(in, oz + o);
      double acc = -(LD(w, o) + LD(r3, o)) * zk;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const double J = LD(mdl.node_jg, go + k + j * g);
        sx[j] += J * zk;
        acc += J * x[j];
      }
      ACC(oz + o, acc);
    }
  }
  if (has_edge) {
    const int xu = t.x_control[e];
    double u[M], su[M];
#pragma unroll
    for (int a = 0; a < M; ++a) {
      u[a] = LD(in, xu + a);
      su[a] = LD(r1, xu + a) * u[a];
    }
    {  // edge Hessian blocks
      const int ho = t.hxx_edge_off[e], co = t.nm_off[e], uo = t.mm_off[e];
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) sx[i] += LD(mdl.edge_hxx, ho + i + j * N) * x[j];
#pragma unroll
      for (int a = 0; a < M; ++a)
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const double h = LD(mdl.edge_hxu, co + i + a * N);
          sx[i] += h * u[a];
          su[a] += h * x[i];
        }
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int a = 0; a < M; ++a) su[a] += LD(mdl.edge_huu, uo + a + j * M) * u[j];
    }
    {  // dynamics of the child: A x + B u - x_child - R2 y_child, and A' y_child, B' y_child
      const int yo = t.y_dyn[child], xc = t.x_state[child];
      const int ao = t.a_off[e], bo = t.b_off[e];
      double yc[N], dyn[N];
#pragma unroll
      for (int p = 0; p < N; ++p) {
        yc[p] = LD(in, oy + yo + p);
        dyn[p] = -LD(in, xc + p) - LD(r2, yo + p) * yc[p];
      }
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int p = 0; p < N; ++p) {
          const double A = LD(mdl.edge_A, ao + p + i * N);
          sx[i] += A * yc[p];
          dyn[p] += A * x[i];
        }
#pragma unroll
      for (int a = 0; a < M; ++a)
#pragma unroll
        for (int p = 0; p < N; ++p) {
          const double B = LD(mdl.edge_B, bo + p + a * N);
          su[a] += B * yc[p];
          dyn[p] += B * u[a];
        }
#pragma unroll
      for (int p = 0; p < N; ++p) ACC(oy + yo + p, dyn[p]);
    }
    {  // edge constraints
      const int c = t.edge_c[e], g = t.edge_g[e];
      const int cx = t.jcx_off[e], cu = t.jcu_off[e], gx = t.jgx_off[e], gu = t.jgu_off[e];
      for (int k = 0; k < c; ++k) {
        const int o = t.y_edge_c[e] + k;
        const double yk = LD(in, oy + o);
        double acc = -LD(r2, o) * yk;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const double J = LD(mdl.edge_jcx, cx + k + j * c);
          sx[j] += J * yk;
          acc += J * x[j];
        }
#pragma unroll
        for (int a = 0; a < M; ++a) {
          const double J = LD(mdl.edge_jcu, cu + k + a * c);
          su[a] += J * yk;
          acc += J * u[a];
        }
        ACC(oy + o, acc);
      }
      for (int k = 0; k < g; ++k) {
        const int o = t.z_edge[e] + k;
        const double zk = LD(in, oz + o);
        double acc = -(LD(w, o) + LD(r3, o)) * zk;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const double J = LD(mdl.edge_jgx, gx + k + j * g);
          sx[j] += J * zk;
          acc += J * x[j];
        }
#pragma unroll
        for (int a = 0; a < M; ++a) {
          const double J = LD(mdl.edge_jgu, gu + k + a * g);
          su[a] += J * zk;
          acc += J * u[a];
        }
        ACC(oz + o, acc);
      }
    }
#pragma unroll
    for (int a = 0; a < M; ++a) ACC(xu + a, su[a]);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) ACC(xs + i, sx[i]);
#undef LD
#undef ACC
}

template <int N, int M>
void launch_apply(const DevTables &t, const KktModel &m, const double *w, const double *r1,
                  const double *r2, const double *r3, const double *x, double *y, int64_t batch,
                  int64_t ld, cudaStream_t s) {
  dim3 grid(static_cast<unsigned>((batch + 127) / 128), static_cast<unsigned>(t.N));
  kkt_apply_chain<N, M><<<grid, 128, 0, s>>>(t, m, w, r1, r2, r3, x, y, batch, ld);
}

}  // namespace

KktApplyFn select_kkt_apply(int n, int m) {
  if (n == 12 && m == 4) return &launch_apply<12, 4>;
  if (n == 4 && m == 1) return &launch_apply<4, 1>;
  if (n == 6 && m == 2) return &launch_apply<6, 2>;
  if (n == 8 && m == 3) return &launch_apply<8, 3>;
  return nullptr;
}

KktReduceFn select_kkt_reduce(int n, int m) {
  if (n == 12 && m == 4) return &launch<12, 4>;
  if (n == 4 && m == 1) return &launch<4, 1>;
  if (n == 6 && m == 2) return &launch<6, 2>;
  if (n == 8 && m == 3) return &launch<8, 3>;
  return nullptr;
}

}  // namespace sipoc
