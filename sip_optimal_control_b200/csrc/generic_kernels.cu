// See generic_kernels.cuh for the mapping and layout.
#include "generic_kernels.cuh"

namespace sipoc {

namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ int64_t problem_index() {
  return static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
}

// ---------------------------------------------------------------------------
// Dense helpers on batch-innermost storage (column-major inside a block).
// ---------------------------------------------------------------------------

// Unblocked left-looking lower Cholesky (Eigen LLT semantics: fail when a pivot
// is <= 0).  Returns false on failure; keeps going is the caller's choice.
__device__ bool chol_lower(GVec a, int off, int n) {
  for (int k = 0; k < n; ++k) {
    double x = a(off + k + k * n);
    for (int j = 0; j < k; ++j) {
      const double l = a(off + k + j * n);
      x -= l * l;
    }
    if (!(x > 0.0)) return false;
    x = sqrt(x);
    a(off + k + k * n) = x;
    for (int i = k + 1; i < n; ++i) {
      double s = a(off + i + k * n);
      for (int j = 0; j < k; ++j) s -= a(off + i + j * n) * a(off + k + j * n);
      a(off + i + k * n) = s / x;
    }
  }
  return true;
}

// L X = B, then L^T X = B, in place, column by column (lqr.cpp:517-519 etc.).
__device__ void chol_solve(GVec l, int loff, int n, GVec b, int boff, int nrhs) {
  for (int c = 0; c < nrhs; ++c) {
    const int xo = boff + c * n;
    for (int i = 0; i < n; ++i) {
      double s = b(xo + i);
      for (int j = 0; j < i; ++j) s -= l(loff + i + j * n) * b(xo + j);
      b(xo + i) = s / l(loff + i + i * n);
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = b(xo + i);
      for (int j = i + 1; j < n; ++j) s -= l(loff + j + i * n) * b(xo + j);
      b(xo + i) = s / l(loff + i + i * n);
    }
  }
}

// (I + D V)^-1 rhs = D^1/2 F^-1 D^-1/2 rhs  (lqr.cpp:531-549); result may alias
// nothing else.
__device__ void f_inv_mult(GVec Ff, int foff, GVec rhs, int roff, GVec res, int xoff,
                           GVec sd, GVec sdi, int doff, int n) {
  for (int i = 0; i < n; ++i) res(xoff + i) = sdi(doff + i) * rhs(roff + i);
  chol_solve(Ff, foff, n, res, xoff, 1);
  for (int i = 0; i < n; ++i) res(xoff + i) *= sd(doff + i);
}

// ---------------------------------------------------------------------------
// LQR factor: lqr.cpp:645-731.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
generic_lqr_factor_kernel(DevTables t, LqrIn in, LqrWs ws, int *status, int64_t batch,
                          int64_t ld) {
  const int64_t b = problem_index();
  if (b >= batch) return;
  const size_t L = static_cast<size_t>(ld);
  GCVec Q{in.Q + b, L}, Mi{in.M + b, L}, R{in.R + b, L}, A{in.A + b, L}, B{in.B + b, L},
      delta{in.delta + b, L};
  GVec W{ws.W + b, L}, K{ws.K + b, L}, V{ws.V + b, L}, Gf{ws.Gf + b, L}, Ff{ws.Ff + b, L},
      sd{ws.sd + b, L}, sdi{ws.sdi + b, L}, H{ws.H + b, L}, F{ws.F + b, L};

  int st = SIPOC_FACTOR_SUCCESS;
  for (int order = 0; order < t.N && st == SIPOC_FACTOR_SUCCESS; ++order) {
    const int node = t.postorder[order];
    const int n = t.n[node];
    const int vo = t.nn_off[node];
    for (int i = 0; i < n * n; ++i) V(vo + i) = Q(vo + i);  // :658

    for (int ci = t.child_offsets[node]; ci < t.child_offsets[node + 1]; ++ci) {
      const int e = t.child_edges[ci];
      const int child = t.children[e];
      const int nc = t.n[child], m = t.m[e];
      const int ao = t.a_off[e], bo = t.b_off[e], mo = t.nm_off[e], ro = t.mm_off[e];
      const int wo = t.w_off[e], ko = t.k_off[e];
      const int fco = t.nn_off[child], dco = t.n_off[child];

      // compute_regularized_W (lqr.cpp:511-529)
      for (int i = 0; i < nc * nc; ++i) W(wo + i) = 0.0;
      for (int i = 0; i < nc; ++i) W(wo + i + i * nc) = 1.0;
      chol_solve(Ff, fco, nc, W, wo, nc);
      for (int i = 0; i < nc * nc; ++i) W(wo + i) *= -1.0;
      for (int i = 0; i < nc; ++i) W(wo + i + i * nc) += 1.0;
      for (int col = 0; col < nc; ++col)
        for (int row = 0; row < nc; ++row)
          W(wo + row + col * nc) *= sdi(dco + row) * sdi(dco + col);

      // H_child = B^T W (m x nc)  (:692)
      for (int j = 0; j < nc; ++j)
        for (int a = 0; a < m; ++a) {
          double s = 0.0;
          for (int p = 0; p < nc; ++p) s += B(bo + p + a * nc) * W(wo + p + j * nc);
          H(a + j * m) = s;
        }
      // G = R + H_child B  (:693-694)
      for (int j = 0; j < m; ++j)
        for (int a = 0; a < m; ++a) {
          double s = R(ro + a + j * m);
          for (int p = 0; p < nc; ++p) s += H(a + p * m) * B(bo + p + j * nc);
          Gf(ro + a + j * m) = s;
        }
      if (!chol_lower(Gf, ro, m)) {  // :696-701
        st = SIPOC_FACTOR_G_FACTORIZATION_FAILURE;
        break;
      }
      // F = W A (nc x n)  (:703)
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < nc; ++i) {
          double s = 0.0;
          for (int p = 0; p < nc; ++p) s += W(wo + i + p * nc) * A(ao + p + j * nc);
          F(i + j * nc) = s;
        }
      // H_parent = M^T + B^T F (m x n)  (:704-705)
      for (int j = 0; j < n; ++j)
        for (int a = 0; a < m; ++a) {
          double s = Mi(mo + j + a * n);
          for (int p = 0; p < nc; ++p) s += B(bo + p + a * nc) * F(p + j * nc);
          H(a + j * m) = s;
        }
      // K = -G^-1 H_parent  (:707-713)
      for (int i = 0; i < m * n; ++i) K(ko + i) = H(i);
      chol_solve(Gf, ro, m, K, ko, n);
      for (int i = 0; i < m * n; ++i) K(ko + i) *= -1.0;
      // V += A^T F  (:715)
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
          double s = V(vo + i + j * n);
          for (int p = 0; p < nc; ++p) s += A(ao + p + i * nc) * F(p + j * nc);
          V(vo + i + j * n) = s;
        }
      // V += K^T H_parent  (:716-719)
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
          double s = 0.0;
          for (int a = 0; a < m; ++a) s += K(ko + a + i * m) * H(a + j * m);
          V(vo + i + j * n) += s;
        }
    }
    if (st != SIPOC_FACTOR_SUCCESS) break;

    // factor_F (lqr.cpp:487-509) with compute_delta_sqrt (:475-485)
    const int d0 = t.n_off[node];
    bool delta_ok = true;
    for (int i = 0; i < n; ++i) {
      const double d = delta(d0 + i);
      if (!(d > 0.0)) {
        delta_ok = false;
        break;
      }
      const double s = sqrt(d);
      sd(d0 + i) = s;
      sdi(d0 + i) = 1.0 / s;
    }
    if (!delta_ok) {
      st = SIPOC_FACTOR_INVALID_DELTA;
      break;
    }
    for (int col = 0; col < n; ++col) {
      for (int row = 0; row < n; ++row)
        Ff(vo + row + col * n) = sd(d0 + row) * V(vo + row + col * n) * sd(d0 + col);
      Ff(vo + col + col * n) += 1.0;
    }
    if (!chol_lower(Ff, vo, n)) st = SIPOC_FACTOR_F_FACTORIZATION_FAILURE;
  }
  if (status != nullptr) status[b] = st;
}

// ---------------------------------------------------------------------------
// LQR solve: lqr.cpp:735-871.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
generic_lqr_solve_kernel(DevTables t, LqrIn in, LqrWs ws, LqrOut out, int64_t batch,
                         int64_t ld) {
  const int64_t b = problem_index();
  if (b >= batch) return;
  const size_t L = static_cast<size_t>(ld);
  GCVec q{in.q + b, L}, r{in.r + b, L}, A{in.A + b, L}, B{in.B + b, L}, c{in.c + b, L},
      delta{in.delta + b, L};
  GVec W{ws.W + b, L}, K{ws.K + b, L}, V{ws.V + b, L}, Gf{ws.Gf + b, L}, Ff{ws.Ff + b, L},
      sd{ws.sd + b, L}, sdi{ws.sdi + b, L}, kk{ws.k + b, L}, v{ws.v + b, L},
      f{ws.f + b, L}, g{ws.g + b, L}, h{ws.h + b, L};
  GVec x{out.x + b, L}, u{out.u + b, L}, y{out.y + b, L};

  // Backward affine sweep (:738-796).
  for (int order = 0; order < t.N; ++order) {
    const int node = t.postorder[order];
    const int n = t.n[node];
    const int no = t.n_off[node];
    for (int i = 0; i < n; ++i) v(no + i) = q(no + i);

    for (int ci = t.child_offsets[node]; ci < t.child_offsets[node + 1]; ++ci) {
      const int e = t.child_edges[ci];
      const int child = t.children[e];
      const int nc = t.n[child], m = t.m[e];
      const int ao = t.a_off[e], bo = t.b_off[e], ro = t.mm_off[e], uo = t.m_off[e];
      const int wo = t.w_off[e], ko = t.k_off[e], co = t.n_off[child];

      for (int i = 0; i < nc; ++i) f(i) = delta(co + i) * v(co + i) - c(co + i);
      for (int i = 0; i < nc; ++i) {
        double s = 0.0;
        for (int j = 0; j < nc; ++j) s += W(wo + i + j * nc) * f(j);
        g(i) = v(co + i) - s;
      }
      for (int a = 0; a < m; ++a) {
        double s = 0.0;
        for (int i = 0; i < nc; ++i) s += B(bo + i + a * nc) * g(i);
        h(a) = r(uo + a) + s;
      }
      for (int a = 0; a < m; ++a) kk(uo + a) = h(a);
      chol_solve(Gf, ro, m, kk, uo, 1);
      for (int a = 0; a < m; ++a) kk(uo + a) *= -1.0;
      for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int i = 0; i < nc; ++i) s += A(ao + i + j * nc) * g(i);
        double w2 = 0.0;
        for (int a = 0; a < m; ++a) w2 += K(ko + a + j * m) * h(a);
        v(no + j) += s;
        v(no + j) += w2;
      }
    }
  }

  // Root (:798-819).
  {
    const int root = t.preorder[0];
    const int n = t.n[root];
    const int no = t.n_off[root], vo = t.nn_off[root];
    for (int i = 0; i < n; ++i) f(i) = delta(no + i) * v(no + i) - c(no + i);
    f_inv_mult(Ff, vo, f, 0, x, no, sd, sdi, no, n);
    for (int i = 0; i < n; ++i) x(no + i) *= -1.0;
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int j = 0; j < n; ++j) s += V(vo + i + j * n) * x(no + j);
      y(no + i) = v(no + i) + s;
    }
  }

  // Forward rollout (:821-870).
  for (int order = 0; order < t.N; ++order) {
    const int node = t.preorder[order];
    const int n = t.n[node];
    const int no = t.n_off[node];
    for (int ci = t.child_offsets[node]; ci < t.child_offsets[node + 1]; ++ci) {
      const int e = t.child_edges[ci];
      const int child = t.children[e];
      const int nc = t.n[child], m = t.m[e];
      const int ao = t.a_off[e], bo = t.b_off[e], uo = t.m_off[e], ko = t.k_off[e];
      const int co = t.n_off[child], vco = t.nn_off[child];

      for (int a = 0; a < m; ++a) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s += K(ko + a + j * m) * x(no + j);
        u(uo + a) = kk(uo + a) + s;
      }
      for (int i = 0; i < nc; ++i) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s += A(ao + i + j * nc) * x(no + j);
        double w2 = 0.0;
        for (int a = 0; a < m; ++a) w2 += B(bo + i + a * nc) * u(uo + a);
        double fi = c(co + i) - delta(co + i) * v(co + i);
        fi += s;
        fi += w2;
        f(i) = fi;
      }
      f_inv_mult(Ff, vco, f, 0, x, co, sd, sdi, co, nc);
      for (int i = 0; i < nc; ++i) {
        double s = 0.0;
        for (int j = 0; j < nc; ++j) s += V(vco + i + j * nc) * x(co + j);
        y(co + i) = v(co + i) + s;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Block-level reduction of {sum sq, max, failed, count} -> 4 atomics per block.
// ---------------------------------------------------------------------------
__device__ void accumulate_stats(double sq, double mx, double failed, double count,
                                 double *stats) {
  __shared__ double s_sq[kThreads / 32], s_mx[kThreads / 32], s_fail[kThreads / 32],
      s_cnt[kThreads / 32];
  for (int o = 16; o > 0; o >>= 1) {
    sq += __shfl_down_sync(0xffffffffu, sq, o);
    mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    failed += __shfl_down_sync(0xffffffffu, failed, o);
    count += __shfl_down_sync(0xffffffffu, count, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_sq[warp] = sq;
    s_mx[warp] = mx;
    s_fail[warp] = failed;
    s_cnt[warp] = count;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w) {
      sq += s_sq[w];
      mx = fmax(mx, s_mx[w]);
      failed += s_fail[w];
      count += s_cnt[w];
    }
    atomicAdd(stats + 0, sq);
    // Non-negative doubles order like their bit patterns.
    atomicMax(reinterpret_cast<unsigned long long *>(stats + 1),
              static_cast<unsigned long long>(__double_as_longlong(mx)));
    atomicAdd(stats + 2, failed);
    atomicAdd(stats + 3, count);
  }
}

// KKT residual of the LQR system (tests/lqr_test.cpp:152-186, :371-409); Q and
// R enter through their lower triangles.
__global__ void __launch_bounds__(kThreads)
lqr_residual_kernel(DevTables t, LqrIn in, LqrOut out, const int *status,
                    double *residual_norm, double *stats, int64_t batch, int64_t ld) {
  const int64_t b = problem_index();
  const bool active = b < batch;
  double sq = 0.0;
  bool failed = false;
  if (active) {
    failed = status != nullptr && status[b] != 0;
    const size_t L = static_cast<size_t>(ld);
    GCVec Q{in.Q + b, L}, Mi{in.M + b, L}, R{in.R + b, L}, q{in.q + b, L}, r{in.r + b, L},
        A{in.A + b, L}, B{in.B + b, L}, c{in.c + b, L}, delta{in.delta + b, L};
    GCVec x{out.x + b, L}, u{out.u + b, L}, y{out.y + b, L};
    if (!failed) {
      for (int node = 0; node < t.N; ++node) {
        const int n = t.n[node], no = t.n_off[node], qo = t.nn_off[node];
        for (int i = 0; i < n; ++i) {
          double s = q(no + i) - y(no + i);
          for (int j = 0; j < n; ++j)
            s += (i >= j ? Q(qo + i + j * n) : Q(qo + j + i * n)) * x(no + j);
          for (int ci = t.child_offsets[node]; ci < t.child_offsets[node + 1]; ++ci) {
            const int e = t.child_edges[ci];
            const int child = t.children[e];
            const int nc = t.n[child], m = t.m[e];
            for (int a = 0; a < m; ++a) s += Mi(t.nm_off[e] + i + a * n) * u(t.m_off[e] + a);
            for (int p = 0; p < nc; ++p)
              s += A(t.a_off[e] + p + i * nc) * y(t.n_off[child] + p);
          }
          sq += s * s;
        }
      }
      for (int e = 0; e < t.E; ++e) {
        const int parent = t.parents[e], child = t.children[e];
        const int n = t.n[parent], nc = t.n[child], m = t.m[e];
        const int po = t.n_off[parent], co = t.n_off[child], uo = t.m_off[e];
        for (int a = 0; a < m; ++a) {
          double s = r(uo + a);
          for (int j = 0; j < m; ++j)
            s += (a >= j ? R(t.mm_off[e] + a + j * m) : R(t.mm_off[e] + j + a * m)) *
                 u(uo + j);
          for (int i = 0; i < n; ++i) s += Mi(t.nm_off[e] + i + a * n) * x(po + i);
          for (int p = 0; p < nc; ++p) s += B(t.b_off[e] + p + a * nc) * y(co + p);
          sq += s * s;
        }
        for (int p = 0; p < nc; ++p) {
          double s = c(co + p) - x(co + p) - delta(co + p) * y(co + p);
          for (int i = 0; i < n; ++i) s += A(t.a_off[e] + p + i * nc) * x(po + i);
          for (int a = 0; a < m; ++a) s += B(t.b_off[e] + p + a * nc) * u(uo + a);
          sq += s * s;
        }
      }
      const int root = t.preorder[0];
      for (int i = 0; i < t.n[root]; ++i) {
        const int o = t.n_off[root] + i;
        const double s = -x(o) - delta(o) * y(o) + c(o);
        sq += s * s;
      }
    }
    if (residual_norm != nullptr) residual_norm[b] = failed ? -1.0 : sqrt(sq);
  }
  if (stats != nullptr) {
    const bool ok = active && !failed;
    accumulate_stats(ok ? sq : 0.0, ok ? sqrt(sq) : 0.0, (active && failed) ? 1.0 : 0.0,
                     active ? 1.0 : 0.0, stats);
  }
}

// ---------------------------------------------------------------------------
// Newton-KKT reduction.  Grid: x = problems, y = nodes.  Thread (b, node) owns
// everything whose accumulation target is indexed by `node` or by one of its
// child edges, in the reference's accumulation order (helpers.cpp:251-360).
// ---------------------------------------------------------------------------

// helpers.cpp:117-136 on strided storage.
__device__ void add_state_gram(GVec Q, int qo, int n, GCVec J, int jo, int rows, GVec wts,
                               int wo) {
  for (int k = 0; k < rows; ++k) {
    const double weight = wts(wo + k);
    for (int col = 0; col < n; ++col) {
      const double wj = weight * J(jo + k + col * rows);
      if (wj == 0.0) continue;
      for (int row = col; row < n; ++row) {
        const double j = J(jo + k + row * rows);
        if (j == 0.0) continue;
        Q(qo + row + col * n) += wj * j;
      }
    }
  }
}

// helpers.cpp:79-115
__device__ void add_control_grams(GVec M, int mo, GVec R, int ro, int n, int m, GCVec Jx,
                                  int jxo, GCVec Ju, int juo, int rows, GVec wts, int wo) {
  for (int k = 0; k < rows; ++k) {
    const double weight = wts(wo + k);
    for (int col = 0; col < m; ++col) {
      const double wju = weight * Ju(juo + k + col * rows);
      if (wju == 0.0) continue;
      for (int row = 0; row < n; ++row) {
        const double jx = Jx(jxo + k + row * rows);
        if (jx == 0.0) continue;
        M(mo + row + col * n) += jx * wju;
      }
    }
    for (int col = 0; col < m; ++col) {
      const double wju = weight * Ju(juo + k + col * rows);
      if (wju == 0.0) continue;
      for (int row = col; row < m; ++row) {
        const double ju = Ju(juo + k + row * rows);
        if (ju == 0.0) continue;
        R(ro + row + col * m) += wju * ju;
      }
    }
  }
}

// helpers.cpp:138-153
__device__ void sub_weighted_jt_rhs(GVec res, int ro, int cols, GCVec J, int jo, int rows,
                                    GVec wts, int wo, GCVec rhs, int rhso) {
  for (int k = 0; k < rows; ++k) {
    const double wr = wts(wo + k) * rhs(rhso + k);
    for (int col = 0; col < cols; ++col) {
      const double j = J(jo + k + col * rows);
      if (j == 0.0) continue;
      res(ro + col) -= j * wr;
    }
  }
}

__device__ void mirror_lower(GVec A, int ao, int n) {  // helpers.cpp:155-158
  for (int col = 0; col < n; ++col)
    for (int row = col + 1; row < n; ++row) A(ao + col + row * n) = A(ao + row + col * n);
}

__global__ void __launch_bounds__(kThreads)
kkt_reduce_kernel(DevTables t, KktModel mdl, const double *w, const double *r1,
                  const double *r2, const double *r3, KktWs ws, int *ok, int64_t batch,
                  int64_t ld) {
  const int64_t b = problem_index();
  if (b >= batch) return;
  const int node = blockIdx.y;
  const size_t L = static_cast<size_t>(ld);
  GCVec W{w + b, L}, R1{r1 + b, L}, R2{r2 + b, L}, R3{r3 + b, L};
  GCVec nh{mdl.node_hxx + b, L}, njc{mdl.node_jc + b, L}, njg{mdl.node_jg + b, L};
  GCVec eh{mdl.edge_hxx + b, L}, ehu{mdl.edge_hxu + b, L}, euu{mdl.edge_huu + b, L};
  GCVec jcx{mdl.edge_jcx + b, L}, jcu{mdl.edge_jcu + b, L}, jgx{mdl.edge_jgx + b, L},
      jgu{mdl.edge_jgu + b, L};
  GVec Q{ws.Q_mod + b, L}, M{ws.M_mod + b, L}, R{ws.R_mod + b, L}, dr2{ws.dyn_r2 + b, L};
  GVec ncw{ws.node_c_r2_inv + b, L}, ngw{ws.node_mod_w_inv + b, L},
      ecw{ws.edge_c_r2_inv + b, L}, egw{ws.edge_mod_w_inv + b, L};

  bool good = true;
  const int n = t.n[node], c = t.node_c[node], g = t.node_g[node];
  // helpers.cpp:251-277
  for (int row = 0; row < n; ++row) {
    const double reg = R2(t.y_dyn[node] + row);
    good = good && (reg > 0.0);
    dr2(t.n_off[node] + row) = reg;
  }
  for (int row = 0; row < c; ++row) {
    const double reg = R2(t.y_node_c[node] + row);
    good = good && (reg > 0.0);
    ncw(t.node_c_off[node] + row) = 1.0 / reg;
  }
  for (int row = 0; row < g; ++row) {
    const int o = t.z_node[node] + row;
    const double reg = W(o) + R3(o);
    good = good && (reg > 0.0);
    ngw(t.node_g_off[node] + row) = 1.0 / reg;
  }
  // helpers.cpp:297-316
  const int qo = t.nn_off[node];
  for (int col = 0; col < n; ++col)
    for (int row = 0; row < n; ++row)
      Q(qo + row + col * n) = row >= col ? nh(qo + row + col * n) : 0.0;
  for (int i = 0; i < n; ++i) Q(qo + i + i * n) += R1(t.x_state[node] + i);
  add_state_gram(Q, qo, n, njc, t.jc_node_off[node], c, ncw, t.node_c_off[node]);
  add_state_gram(Q, qo, n, njg, t.jg_node_off[node], g, ngw, t.node_g_off[node]);

  for (int ci = t.child_offsets[node]; ci < t.child_offsets[node + 1]; ++ci) {
    const int e = t.child_edges[ci];
    const int m = t.m[e], ec = t.edge_c[e], eg = t.edge_g[e];
    // helpers.cpp:279-295
    for (int row = 0; row < ec; ++row) {
      const double reg = R2(t.y_edge_c[e] + row);
      good = good && (reg > 0.0);
      ecw(t.edge_c_off[e] + row) = 1.0 / reg;
    }
    for (int row = 0; row < eg; ++row) {
      const int o = t.z_edge[e] + row;
      const double reg = W(o) + R3(o);
      good = good && (reg > 0.0);
      egw(t.edge_g_off[e] + row) = 1.0 / reg;
    }
    // helpers.cpp:318-354
    const int ho = t.hxx_edge_off[e];
    for (int col = 0; col < n; ++col)
      for (int row = col; row < n; ++row) Q(qo + row + col * n) += eh(ho + row + col * n);
    add_state_gram(Q, qo, n, jcx, t.jcx_off[e], ec, ecw, t.edge_c_off[e]);
    add_state_gram(Q, qo, n, jgx, t.jgx_off[e], eg, egw, t.edge_g_off[e]);

    const int mo = t.nm_off[e], ro = t.mm_off[e];
    for (int i = 0; i < n * m; ++i) M(mo + i) = ehu(mo + i);
    for (int col = 0; col < m; ++col)
      for (int row = 0; row < m; ++row)
        R(ro + row + col * m) = row >= col ? euu(ro + row + col * m) : 0.0;
    for (int i = 0; i < m; ++i) R(ro + i + i * m) += R1(t.x_control[e] + i);
    add_control_grams(M, mo, R, ro, n, m, jcx, t.jcx_off[e], jcu, t.jcu_off[e], ec, ecw,
                      t.edge_c_off[e]);
    add_control_grams(M, mo, R, ro, n, m, jgx, t.jgx_off[e], jgu, t.jgu_off[e], eg, egw,
                      t.edge_g_off[e]);
    mirror_lower(R, ro, m);
  }
  mirror_lower(Q, qo, n);  // helpers.cpp:356-360
  if (!good) ok[b] = 0;    // benign race: every writer stores 0
}

// ok[b] &= (lqr_status[b] == SUCCESS)   (helpers.cpp:368-370)
__global__ void kkt_finish_factor_kernel(const int *lqr_status, int *ok, int64_t batch) {
  const int64_t b = problem_index();
  if (b >= batch) return;
  if (lqr_status[b] != SIPOC_FACTOR_SUCCESS) ok[b] = 0;
}

// helpers.cpp:752-812: q_mod, c_mod per node; r_mod per edge.
__global__ void __launch_bounds__(kThreads)
kkt_build_rhs_kernel(DevTables t, KktModel mdl, KktWs ws, const double *bvec,
                     int64_t batch, int64_t ld) {
  const int64_t b = problem_index();
  if (b >= batch) return;
  const int node = blockIdx.y;
  const size_t L = static_cast<size_t>(ld);
  GCVec rhs{bvec + b, L};
  GCVec njc{mdl.node_jc + b, L}, njg{mdl.node_jg + b, L};
  GCVec jcx{mdl.edge_jcx + b, L}, jcu{mdl.edge_jcu + b, L}, jgx{mdl.edge_jgx + b, L},
      jgu{mdl.edge_jgu + b, L};
  GVec q{ws.q_mod + b, L}, r{ws.r_mod + b, L}, cm{ws.c_mod + b, L};
  GVec ncw{ws.node_c_r2_inv + b, L}, ngw{ws.node_mod_w_inv + b, L},
      ecw{ws.edge_c_r2_inv + b, L}, egw{ws.edge_mod_w_inv + b, L};
  const int xd = t.x_dim, yd = t.y_dim;
  const int n = t.n[node], no = t.n_off[node];

  for (int i = 0; i < n; ++i) q(no + i) = -rhs(t.x_state[node] + i);
  sub_weighted_jt_rhs(q, no, n, njc, t.jc_node_off[node], t.node_c[node], ncw,
                      t.node_c_off[node], rhs, xd + t.y_node_c[node]);
  sub_weighted_jt_rhs(q, no, n, njg, t.jg_node_off[node], t.node_g[node], ngw,
                      t.node_g_off[node], rhs, xd + yd + t.z_node[node]);
  for (int i = 0; i < n; ++i) cm(no + i) = -rhs(xd + t.y_dyn[node] + i);

  for (int ci = t.child_offsets[node]; ci < t.child_offsets[node + 1]; ++ci) {
    const int e = t.child_edges[ci];
    const int m = t.m[e], ec = t.edge_c[e], eg = t.edge_g[e], uo = t.m_off[e];
    const int byc = xd + t.y_edge_c[e], bz = xd + yd + t.z_edge[e];
    sub_weighted_jt_rhs(q, no, n, jcx, t.jcx_off[e], ec, ecw, t.edge_c_off[e], rhs, byc);
    sub_weighted_jt_rhs(q, no, n, jgx, t.jgx_off[e], eg, egw, t.edge_g_off[e], rhs, bz);
    for (int i = 0; i < m; ++i) r(uo + i) = -rhs(t.x_control[e] + i);
    sub_weighted_jt_rhs(r, uo, m, jcu, t.jcu_off[e], ec, ecw, t.edge_c_off[e], rhs, byc);
    sub_weighted_jt_rhs(r, uo, m, jgu, t.jgu_off[e], eg, egw, t.edge_g_off[e], rhs, bz);
  }
}

// helpers.cpp:817-893: scatter x, u, y_dyn into sol; recover y_c and z.
__global__ void __launch_bounds__(kThreads)
kkt_recover_kernel(DevTables t, KktModel mdl, KktWs ws, const double *bvec, double *solp,
                   int64_t batch, int64_t ld) {
  const int64_t b = problem_index();
  if (b >= batch) return;
  const int node = blockIdx.y;
  const size_t L = static_cast<size_t>(ld);
  GCVec rhs{bvec + b, L};
  GCVec njc{mdl.node_jc + b, L}, njg{mdl.node_jg + b, L};
  GCVec jcx{mdl.edge_jcx + b, L}, jcu{mdl.edge_jcu + b, L}, jgx{mdl.edge_jgx + b, L},
      jgu{mdl.edge_jgu + b, L};
  GVec x{ws.x + b, L}, u{ws.u + b, L}, y{ws.y + b, L}, sol{solp + b, L};
  GVec ncw{ws.node_c_r2_inv + b, L}, ngw{ws.node_mod_w_inv + b, L},
      ecw{ws.edge_c_r2_inv + b, L}, egw{ws.edge_mod_w_inv + b, L};
  const int xd = t.x_dim, yd = t.y_dim;
  const int n = t.n[node], no = t.n_off[node];

  for (int i = 0; i < n; ++i) {
    sol(t.x_state[node] + i) = x(no + i);
    sol(xd + t.y_dyn[node] + i) = y(no + i);
  }
  {  // :828-856
    const int c = t.node_c[node], g = t.node_g[node];
    for (int k = 0; k < c; ++k) {
      double s = 0.0;
      for (int j = 0; j < n; ++j) s += njc(t.jc_node_off[node] + k + j * c) * x(no + j);
      s -= rhs(xd + t.y_node_c[node] + k);
      sol(xd + t.y_node_c[node] + k) = ncw(t.node_c_off[node] + k) * s;
    }
    for (int k = 0; k < g; ++k) {
      double s = 0.0;
      for (int j = 0; j < n; ++j) s += njg(t.jg_node_off[node] + k + j * g) * x(no + j);
      s -= rhs(xd + yd + t.z_node[node] + k);
      sol(xd + yd + t.z_node[node] + k) = ngw(t.node_g_off[node] + k) * s;
    }
  }
  for (int ci = t.child_offsets[node]; ci < t.child_offsets[node + 1]; ++ci) {  // :858-893
    const int e = t.child_edges[ci];
    const int m = t.m[e], c = t.edge_c[e], g = t.edge_g[e], uo = t.m_off[e];
    for (int i = 0; i < m; ++i) sol(t.x_control[e] + i) = u(uo + i);
    for (int k = 0; k < c; ++k) {
      double s = 0.0;
      for (int j = 0; j < n; ++j) s += jcx(t.jcx_off[e] + k + j * c) * x(no + j);
      double s2 = 0.0;
      for (int j = 0; j < m; ++j) s2 += jcu(t.jcu_off[e] + k + j * c) * u(uo + j);
      s += s2;
      s -= rhs(xd + t.y_edge_c[e] + k);
      sol(xd + t.y_edge_c[e] + k) = ecw(t.edge_c_off[e] + k) * s;
    }
    for (int k = 0; k < g; ++k) {
      double s = 0.0;
      for (int j = 0; j < n; ++j) s += jgx(t.jgx_off[e] + k + j * g) * x(no + j);
      double s2 = 0.0;
      for (int j = 0; j < m; ++j) s2 += jgu(t.jgu_off[e] + k + j * g) * u(uo + j);
      s += s2;
      s -= rhs(xd + yd + t.z_edge[e] + k);
      sol(xd + yd + t.z_edge[e] + k) = egw(t.edge_g_off[e] + k) * s;
    }
  }
}

// y += K x (helpers.cpp:953-1368, theta == 0), output-driven: thread (b, node)
// produces every row of y whose block belongs to `node` or to one of its child
// edges, so no two threads write the same element.
template <bool ALL>  // ALL: the whole operator, the block mask folds away at compile time
__global__ void __launch_bounds__(kThreads)
kkt_apply_kernel(DevTables t, KktModel mdl, const double *w, const double *r1,
                 const double *r2, const double *r3, const double *in_x, const double *in_y,
                 const double *in_z, double *out_x, double *out_y, double *out_z, unsigned parts,
                 int64_t batch, int64_t ld) {
  // parts: which blocks of K = [H + R1, C', G'; C, -R2, 0; G, 0, -(W + R3)] are applied
  // (kKktH | kKktC | kKktCT | kKktG | kKktGT | kKktReg); pointers of unused vectors may be null.
  const int64_t b = problem_index();
  if (b >= batch) return;
  const int node = blockIdx.y;
  const size_t L = static_cast<size_t>(ld);
  const bool pH = ALL || (parts & kKktH), pC = ALL || (parts & kKktC), pCT = ALL || (parts & kKktCT),
             pG = ALL || (parts & kKktG), pGT = ALL || (parts & kKktGT), pR = ALL || (parts & kKktReg);
  GCVec W{w + b, L}, R1{r1 + b, L}, R2{r2 + b, L}, R3{r3 + b, L};
  GCVec X{in_x + b, L}, Y{in_y + b, L}, Z{in_z + b, L};
  GCVec nh{mdl.node_hxx + b, L}, njc{mdl.node_jc + b, L}, njg{mdl.node_jg + b, L};
  GCVec eh{mdl.edge_hxx + b, L}, ehu{mdl.edge_hxu + b, L}, euu{mdl.edge_huu + b, L},
      eA{mdl.edge_A + b, L}, eB{mdl.edge_B + b, L};
  GCVec jcx{mdl.edge_jcx + b, L}, jcu{mdl.edge_jcu + b, L}, jgx{mdl.edge_jgx + b, L},
      jgu{mdl.edge_jgu + b, L};
  GVec OX{out_x + b, L}, OY{out_y + b, L}, OZ{out_z + b, L};
  const int n = t.n[node], c = t.node_c[node], g = t.node_g[node];
  const int xs = t.x_state[node];
  const int ie = t.in_edge[node];

  // x rows of this node's state.
  if (pH || pCT || pGT || pR) {
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      if (pH)
        for (int j = 0; j < n; ++j) s += nh(t.nn_off[node] + i + j * n) * X(xs + j);
      if (pCT)
        for (int k = 0; k < c; ++k)
          s += njc(t.jc_node_off[node] + k + i * c) * Y(t.y_node_c[node] + k);
      if (pGT)
        for (int k = 0; k < g; ++k) s += njg(t.jg_node_off[node] + k + i * g) * Z(t.z_node[node] + k);
      if (pCT) s -= Y(t.y_dyn[node] + i);  // -I of the node's own dynamics / root row
      for (int ci = t.child_offsets[node]; ci < t.child_offsets[node + 1]; ++ci) {
        const int e = t.child_edges[ci];
        const int child = t.children[e];
        const int nc = t.n[child], m = t.m[e], ec = t.edge_c[e], eg = t.edge_g[e];
        if (pH) {
          for (int j = 0; j < n; ++j) s += eh(t.hxx_edge_off[e] + i + j * n) * X(xs + j);
          for (int a = 0; a < m; ++a) s += ehu(t.nm_off[e] + i + a * n) * X(t.x_control[e] + a);
        }
        if (pCT) {
          for (int p = 0; p < nc; ++p) s += eA(t.a_off[e] + p + i * nc) * Y(t.y_dyn[child] + p);
          for (int k = 0; k < ec; ++k) s += jcx(t.jcx_off[e] + k + i * ec) * Y(t.y_edge_c[e] + k);
        }
        if (pGT)
          for (int k = 0; k < eg; ++k) s += jgx(t.jgx_off[e] + k + i * eg) * Z(t.z_edge[e] + k);
      }
      if (pR) s += R1(xs + i) * X(xs + i);
      OX(xs + i) += s;
    }
  }
  // y rows: dynamics of this node (root: -x_root; else A x_p + B u - x_node).
  if (pC || pR) {
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      if (pC) {
        s = -X(xs + i);
        if (ie >= 0) {
          const int parent = t.parents[ie];
          const int np = t.n[parent], m = t.m[ie];
          for (int j = 0; j < np; ++j) s += eA(t.a_off[ie] + i + j * n) * X(t.x_state[parent] + j);
          for (int a = 0; a < m; ++a) s += eB(t.b_off[ie] + i + a * n) * X(t.x_control[ie] + a);
        }
      }
      const int o = t.y_dyn[node] + i;
      if (pR) s -= R2(o) * Y(o);
      OY(o) += s;
    }
    for (int k = 0; k < c; ++k) {
      double s = 0.0;
      if (pC)
        for (int j = 0; j < n; ++j) s += njc(t.jc_node_off[node] + k + j * c) * X(xs + j);
      const int o = t.y_node_c[node] + k;
      if (pR) s -= R2(o) * Y(o);
      OY(o) += s;
    }
  }
  if (pG || pR) {
    for (int k = 0; k < g; ++k) {
      double s = 0.0;
      if (pG)
        for (int j = 0; j < n; ++j) s += njg(t.jg_node_off[node] + k + j * g) * X(xs + j);
      const int o = t.z_node[node] + k;
      if (pR) s -= (W(o) + R3(o)) * Z(o);
      OZ(o) += s;
    }
  }
  // Rows owned by child edges: control stationarity, edge_c, edge_g.
  for (int ci = t.child_offsets[node]; ci < t.child_offsets[node + 1]; ++ci) {
    const int e = t.child_edges[ci];
    const int child = t.children[e];
    const int nc = t.n[child], m = t.m[e], ec = t.edge_c[e], eg = t.edge_g[e];
    const int xu = t.x_control[e];
    if (pH || pCT || pGT || pR) {
      for (int a = 0; a < m; ++a) {
        double s = 0.0;
        if (pH) {
          for (int i = 0; i < n; ++i) s += ehu(t.nm_off[e] + i + a * n) * X(xs + i);
          for (int j = 0; j < m; ++j) s += euu(t.mm_off[e] + a + j * m) * X(xu + j);
        }
        if (pCT) {
          for (int p = 0; p < nc; ++p) s += eB(t.b_off[e] + p + a * nc) * Y(t.y_dyn[child] + p);
          for (int k = 0; k < ec; ++k) s += jcu(t.jcu_off[e] + k + a * ec) * Y(t.y_edge_c[e] + k);
        }
        if (pGT)
          for (int k = 0; k < eg; ++k) s += jgu(t.jgu_off[e] + k + a * eg) * Z(t.z_edge[e] + k);
        if (pR) s += R1(xu + a) * X(xu + a);
        OX(xu + a) += s;
      }
    }
    if (pC || pR) {
      for (int k = 0; k < ec; ++k) {
        double s = 0.0;
        if (pC) {
          for (int j = 0; j < n; ++j) s += jcx(t.jcx_off[e] + k + j * ec) * X(xs + j);
          for (int j = 0; j < m; ++j) s += jcu(t.jcu_off[e] + k + j * ec) * X(xu + j);
        }
        const int o = t.y_edge_c[e] + k;
        if (pR) s -= R2(o) * Y(o);
        OY(o) += s;
      }
    }
    if (pG || pR) {
      for (int k = 0; k < eg; ++k) {
        double s = 0.0;
        if (pG) {
          for (int j = 0; j < n; ++j) s += jgx(t.jgx_off[e] + k + j * eg) * X(xs + j);
          for (int j = 0; j < m; ++j) s += jgu(t.jgu_off[e] + k + j * eg) * X(xu + j);
        }
        const int o = t.z_edge[e] + k;
        if (pR) s -= (W(o) + R3(o)) * Z(o);
        OZ(o) += s;
      }
    }
  }
}

// ||Ksol - b||_2 per problem (tests/variable_dimensions_test.cpp:167-180).
// One CTA = 32 problems x kResidualSlices row slices: slice s sums rows s, s + S, s + 2S, ...
// of its problem (every load 256 B coalesced over the 32 problems), the partial sums meet in
// shared memory and are added in slice order, so the result does not depend on scheduling.
// One thread per problem cannot hide the load latency by occupancy at Newton-KKT batch sizes.
constexpr int kResidualSlices = 8;

__global__ void __launch_bounds__(32 * kResidualSlices)
kkt_residual_kernel(DevTables t, const double *Ksol, const double *bvec, const int *ok,
                    double *residual_norm, double *stats, int64_t batch, int64_t ld) {
  __shared__ double part[kResidualSlices][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * 32 + lane;
  const bool active = b < batch;
  const bool failed = active && ok != nullptr && ok[b] == 0;
  double sq = 0.0;
  if (active && !failed) {
    const size_t L = static_cast<size_t>(ld);
    constexpr int S = kResidualSlices;
    int i = slice;
    for (; i + 3 * S < t.kkt_dim; i += 4 * S) {  // four rows (eight loads) in flight
      double d[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) d[u] = Ksol[(i + u * S) * L + b] - bvec[(i + u * S) * L + b];
#pragma unroll
      for (int u = 0; u < 4; ++u) sq += d[u] * d[u];
    }
    for (; i < t.kkt_dim; i += S) {
      const double d = Ksol[i * L + b] - bvec[i * L + b];
      sq += d * d;
    }
  }
  part[slice][lane] = sq;
  __syncthreads();
  if (slice != 0) return;
  sq = part[0][lane];
#pragma unroll
  for (int s = 1; s < kResidualSlices; ++s) sq += part[s][lane];
  if (active && residual_norm != nullptr) residual_norm[b] = failed ? -1.0 : sqrt(sq);
  if (stats == nullptr) return;
  const bool good = active && !failed;
  double mx = good ? sqrt(sq) : 0.0, nfail = failed ? 1.0 : 0.0, cnt = active ? 1.0 : 0.0;
  if (!good) sq = 0.0;
  for (int o = 16; o > 0; o >>= 1) {
    sq += __shfl_down_sync(0xffffffffu, sq, o);
    mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    nfail += __shfl_down_sync(0xffffffffu, nfail, o);
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) {
    atomicAdd(stats + 0, sq);
    // Non-negative doubles order like their bit patterns.
    atomicMax(reinterpret_cast<unsigned long long *>(stats + 1),
              static_cast<unsigned long long>(__double_as_longlong(mx)));
    atomicAdd(stats + 2, nfail);
    atomicAdd(stats + 3, cnt);
  }
}

// ---------------------------------------------------------------------------
// Layout conversion: 32 x 32 tiled transpose through shared memory so both the
// problem-major side (contiguous in `size`) and the engine side (contiguous in
// batch) are read / written with full 256-byte requests.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t size,
            int64_t batch, int64_t ld) {
  __shared__ double tile[32][33];
  const int64_t b0 = static_cast<int64_t>(blockIdx.y) * 32;
  const int64_t e0 = static_cast<int64_t>(blockIdx.x) * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int64_t bb = b0 + r, ee = e0 + threadIdx.x;
    tile[r][threadIdx.x] = (bb < batch && ee < size) ? src[bb * size + ee] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int64_t ee = e0 + r, bb = b0 + threadIdx.x;
    if (ee < size && bb < ld) dst[ee * ld + bb] = tile[threadIdx.x][r];
  }
}

__global__ void __launch_bounds__(256)
unpack_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t size,
              int64_t batch, int64_t ld) {
  __shared__ double tile[32][33];
  const int64_t b0 = static_cast<int64_t>(blockIdx.y) * 32;
  const int64_t e0 = static_cast<int64_t>(blockIdx.x) * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int64_t ee = e0 + r, bb = b0 + threadIdx.x;
    tile[r][threadIdx.x] = (ee < size && bb < ld) ? src[ee * ld + bb] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int64_t bb = b0 + r, ee = e0 + threadIdx.x;
    if (bb < batch && ee < size) dst[bb * size + ee] = tile[threadIdx.x][r];
  }
}

// ---------------------------------------------------------------------------
// Variable-dimension chain or tree <-> the same topology with uniform dims (NP, MP) >= every
// stage's dims.
// A stage is padded with states / controls that are decoupled from the real ones
// (identity on the diagonal of Q and R, delta = 1, zeros elsewhere: their terms enter
// every product as exact zeros), so the shape-specialised kernels serve chains whose
// dims change from stage to stage.  One thread per (stage, problem); consecutive threads
// are consecutive problems, so every access is a full coalesced request.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pad_chain_kernel(DevTables t, LqrIn src, LqrIn dst, int NP, int MP, unsigned mask, int64_t batch,
                 int64_t ld) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;  // node k, and edge k when k < E
  if (b >= batch) return;
  const size_t L = static_cast<size_t>(ld);
  const int nk = t.n[k];
  auto at = [&](const double *p, size_t flat) { return p[flat * L + b]; };
  // (dst holds the engine's own padded buffers; LqrIn's members are const for the readers)
  auto put = [&](const double *p, size_t flat, double v) { const_cast<double *>(p)[flat * L + b] = v; };
  if (mask & 1u)
    for (int j = 0; j < NP; ++j)
      for (int i = 0; i < NP; ++i)
        put(dst.Q, (static_cast<size_t>(k) * NP + j) * NP + i,
            (i < nk && j < nk) ? at(src.Q, t.nn_off[k] + j * nk + i) : (i == j ? 1.0 : 0.0));
  for (int i = 0; i < NP; ++i) {
    const bool in = i < nk;
    if (mask & 8u) put(dst.q, static_cast<size_t>(k) * NP + i, in ? at(src.q, t.n_off[k] + i) : 0.0);
    if (mask & 128u) put(dst.c, static_cast<size_t>(k) * NP + i, in ? at(src.c, t.n_off[k] + i) : 0.0);
    if (mask & 256u)
      put(dst.delta, static_cast<size_t>(k) * NP + i, in ? at(src.delta, t.n_off[k] + i) : 1.0);
  }
  if (k >= t.E) return;
  // edge k: parent state dims np, child state dims nc (on a chain: nodes k and k + 1)
  const int mk = t.m[k], np = t.n[t.parents[k]], nc = t.n[t.children[k]];
  for (int a = 0; a < MP; ++a) {
    if (mask & 2u)
      for (int x = 0; x < NP; ++x)
        put(dst.M, (static_cast<size_t>(k) * MP + a) * NP + x,
            (x < np && a < mk) ? at(src.M, t.nm_off[k] + a * np + x) : 0.0);
    if (mask & 4u)
      for (int a1 = 0; a1 < MP; ++a1)
        put(dst.R, (static_cast<size_t>(k) * MP + a) * MP + a1,
            (a1 < mk && a < mk) ? at(src.R, t.mm_off[k] + a * mk + a1) : (a1 == a ? 1.0 : 0.0));
    if (mask & 16u) put(dst.r, static_cast<size_t>(k) * MP + a, a < mk ? at(src.r, t.m_off[k] + a) : 0.0);
    if (mask & 64u)
      for (int i = 0; i < NP; ++i)
        put(dst.B, (static_cast<size_t>(k) * MP + a) * NP + i,
            (i < nc && a < mk) ? at(src.B, t.b_off[k] + a * nc + i) : 0.0);
  }
  if (mask & 32u)
    for (int j = 0; j < NP; ++j)
      for (int i = 0; i < NP; ++i)
        put(dst.A, (static_cast<size_t>(k) * NP + j) * NP + i,
            (i < nc && j < np) ? at(src.A, t.a_off[k] + j * nc + i) : 0.0);
}

__global__ void __launch_bounds__(128)
unpad_chain_kernel(DevTables t, LqrOut src, LqrOut dst, int NP, int MP, int64_t batch,
                   int64_t ld) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (b >= batch) return;
  const size_t L = static_cast<size_t>(ld);
  const int nk = t.n[k];
  for (int i = 0; i < nk; ++i) {
    dst.x[(static_cast<size_t>(t.n_off[k]) + i) * L + b] = src.x[(static_cast<size_t>(k) * NP + i) * L + b];
    dst.y[(static_cast<size_t>(t.n_off[k]) + i) * L + b] = src.y[(static_cast<size_t>(k) * NP + i) * L + b];
  }
  if (k >= t.E) return;
  const int mk = t.m[k];
  for (int a = 0; a < mk; ++a)
    dst.u[(static_cast<size_t>(t.m_off[k]) + a) * L + b] = src.u[(static_cast<size_t>(k) * MP + a) * L + b];
}

// stats = {0, 0, #problems with status != 0, #problems}: the per-iteration
// convergence / failure flags a multi-GPU driver all-reduces.
__global__ void __launch_bounds__(kThreads)
status_stats_kernel(const int *status, double *stats, int64_t batch) {
  double failed = 0.0, count = 0.0;
  for (int64_t b = problem_index(); b < batch;
       b += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    failed += status[b] != 0 ? 1.0 : 0.0;
    count += 1.0;
  }
  accumulate_stats(0.0, 0.0, failed, count, stats);
}

__global__ void fill_int_kernel(int *dst, int value, int64_t count) {
  const int64_t i = problem_index();
  if (i < count) dst[i] = value;
}

__global__ void zero_stats_kernel(double *stats) {
  if (threadIdx.x < 4) stats[threadIdx.x] = 0.0;
}

inline dim3 batch_grid(int64_t batch, int y = 1) {
  return dim3(static_cast<unsigned>((batch + kThreads - 1) / kThreads),
              static_cast<unsigned>(y));
}

}  // namespace

void launch_generic_lqr_factor(const DevTables &t, const LqrIn &in, const LqrWs &ws,
                               int *status, int64_t batch, int64_t ld, cudaStream_t s) {
  generic_lqr_factor_kernel<<<batch_grid(batch), kThreads, 0, s>>>(t, in, ws, status, batch,
                                                                   ld);
}

void launch_generic_lqr_solve(const DevTables &t, const LqrIn &in, const LqrWs &ws,
                              const LqrOut &out, int64_t batch, int64_t ld,
                              cudaStream_t s) {
  generic_lqr_solve_kernel<<<batch_grid(batch), kThreads, 0, s>>>(t, in, ws, out, batch, ld);
}

void launch_lqr_residual(const DevTables &t, const LqrIn &in, const LqrOut &out,
                         const int *status, double *residual_norm, double *stats,
                         int64_t batch, int64_t ld, cudaStream_t s) {
  if (stats != nullptr) zero_stats_kernel<<<1, 32, 0, s>>>(stats);
  lqr_residual_kernel<<<batch_grid(batch), kThreads, 0, s>>>(t, in, out, status,
                                                             residual_norm, stats, batch, ld);
}

void launch_kkt_reduce(const DevTables &t, const KktModel &m, const double *w,
                       const double *r1, const double *r2, const double *r3,
                       const KktWs &ws, int *ok, int64_t batch, int64_t ld,
                       cudaStream_t s) {
  fill_int_kernel<<<batch_grid(ld), kThreads, 0, s>>>(ok, 1, ld);
  kkt_reduce_kernel<<<batch_grid(batch, t.N), kThreads, 0, s>>>(t, m, w, r1, r2, r3, ws, ok,
                                                               batch, ld);
}

void launch_kkt_finish_factor(const int *lqr_status, int *ok, int64_t batch,
                              cudaStream_t s) {
  kkt_finish_factor_kernel<<<batch_grid(batch), kThreads, 0, s>>>(lqr_status, ok, batch);
}

void launch_kkt_build_rhs(const DevTables &t, const KktModel &m, const KktWs &ws,
                          const double *b, int64_t batch, int64_t ld, cudaStream_t s) {
  kkt_build_rhs_kernel<<<batch_grid(batch, t.N), kThreads, 0, s>>>(t, m, ws, b, batch, ld);
}

void launch_kkt_recover(const DevTables &t, const KktModel &m, const KktWs &ws,
                        const double *b, double *sol, int64_t batch, int64_t ld,
                        cudaStream_t s) {
  kkt_recover_kernel<<<batch_grid(batch, t.N), kThreads, 0, s>>>(t, m, ws, b, sol, batch, ld);
}

void launch_kkt_apply(const DevTables &t, const KktModel &m, const double *w,
                      const double *r1, const double *r2, const double *r3,
                      const double *x, double *y, int64_t batch, int64_t ld,
                      cudaStream_t s) {
  // the whole operator on [x | y | z] vectors
  const size_t oy = static_cast<size_t>(t.x_dim) * ld, oz = oy + static_cast<size_t>(t.y_dim) * ld;
  kkt_apply_kernel<true><<<batch_grid(batch, t.N), kThreads, 0, s>>>(
      t, m, w, r1, r2, r3, x, x + oy, x + oz, y, y + oy, y + oz, kKktAll, batch, ld);
}

void launch_kkt_apply_parts(const DevTables &t, const KktModel &m, unsigned parts,
                            const double *in_x, const double *in_y, const double *in_z,
                            double *out_x, double *out_y, double *out_z, int64_t batch,
                            int64_t ld, cudaStream_t s) {
  // component blocks only (no regularization): the weight pointers are never read
  kkt_apply_kernel<false><<<batch_grid(batch, t.N), kThreads, 0, s>>>(
      t, m, nullptr, nullptr, nullptr, nullptr, in_x, in_y, in_z, out_x, out_y, out_z,
      parts & ~kKktReg, batch, ld);
}

void launch_kkt_residual(const DevTables &t, const double *Ksol, const double *b,
                         const int *ok, double *residual_norm, double *stats,
                         int64_t batch, int64_t ld, cudaStream_t s) {
  if (stats != nullptr) zero_stats_kernel<<<1, 32, 0, s>>>(stats);
  const unsigned grid = static_cast<unsigned>((batch + 31) / 32);
  kkt_residual_kernel<<<grid, 32 * kResidualSlices, 0, s>>>(t, Ksol, b, ok, residual_norm, stats,
                                                            batch, ld);
}

void launch_pack(const double *src, double *dst, int64_t size, int64_t batch, int64_t ld,
                 cudaStream_t s) {
  if (size == 0) return;
  dim3 grid(static_cast<unsigned>((size + 31) / 32), static_cast<unsigned>((ld + 31) / 32));
  pack_kernel<<<grid, dim3(32, 8), 0, s>>>(src, dst, size, batch, ld);
}

void launch_unpack(const double *src, double *dst, int64_t size, int64_t batch, int64_t ld,
                   cudaStream_t s) {
  if (size == 0) return;
  dim3 grid(static_cast<unsigned>((size + 31) / 32), static_cast<unsigned>((ld + 31) / 32));
  unpack_kernel<<<grid, dim3(32, 8), 0, s>>>(src, dst, size, batch, ld);
}

void launch_pad_chain(const DevTables &t, const LqrIn &src, const LqrIn &dst, int np, int mp,
                      unsigned mask, int64_t batch, int64_t ld, cudaStream_t s) {
  dim3 grid(static_cast<unsigned>((batch + 127) / 128), static_cast<unsigned>(t.N));
  pad_chain_kernel<<<grid, 128, 0, s>>>(t, src, dst, np, mp, mask, batch, ld);
}

void launch_unpad_chain(const DevTables &t, const LqrOut &src, const LqrOut &dst, int np, int mp,
                        int64_t batch, int64_t ld, cudaStream_t s) {
  dim3 grid(static_cast<unsigned>((batch + 127) / 128), static_cast<unsigned>(t.N));
  unpad_chain_kernel<<<grid, 128, 0, s>>>(t, src, dst, np, mp, batch, ld);
}

void launch_status_stats(const int *status, double *stats, int64_t batch, cudaStream_t s) {
  zero_stats_kernel<<<1, 32, 0, s>>>(stats);
  const int64_t blocks = (batch + kThreads - 1) / kThreads;
  status_stats_kernel<<<static_cast<unsigned>(blocks < 592 ? blocks : 592), kThreads, 0, s>>>(
      status, stats, batch);
}

void launch_fill_int(int *dst, int value, int64_t count, cudaStream_t s) {
  fill_int_kernel<<<batch_grid(count), kThreads, 0, s>>>(dst, value, count);
}

}  // namespace sipoc
