// C++ host side of the B200 engine: the reference's lqr.hpp classes (same names,
// argument meaning and error behaviour) implemented over the C ABI of
// include/sipoc.h.  A translation unit written against the reference's
//   #include "sip_optimal_control/lqr.hpp"
// compiles against this header instead; LQR::factor / LQR::solve then run on the GPU.
//
// Reference interfaces mirrored (file:line in joaospinto/sip_optimal_control):
//   Topology            lqr.hpp:5-22,   lqr.cpp:12-60
//   Dimensions          lqr.hpp:24-64,  lqr.cpp:62-180
//   LQR::FactorStatus   lqr.hpp:68-74
//   LQR::Input / Output lqr.hpp:76-106
//   LQR::Workspace      lqr.hpp:109-187 (contents replaced by the device engine)
//   LQR                 lqr.hpp:189-199, lqr.cpp:635-871
// The classic classes solve ONE problem per call (batch of one through the
// host-buffer entry points, like a maintainer's drop-in shim would); BatchedLQR
// exposes the same operations on a batch of problems resident on the device.
#pragma once

#include <cstdint>
#include <vector>

struct sipoc_engine;

namespace sip::optimal_control {

struct Topology {
  int num_edges = 0;
  int root = 0;
  const int *edge_parents = nullptr;
  const int *edge_children = nullptr;

  int num_nodes() const;

  void reserve(int num_edges);
  void free();
  void set_chain();
  void set_tree(int root, const int *edge_parents, const int *edge_children);

 private:
  int *owned_parents_ = nullptr;
  int *owned_children_ = nullptr;
};

struct Dimensions {
  int theta_dim = 0;
  const int *state_dims = nullptr;
  const int *control_dims = nullptr;
  const int *node_c_dims = nullptr;
  const int *node_g_dims = nullptr;
  const int *edge_c_dims = nullptr;
  const int *edge_g_dims = nullptr;

  void reserve(int num_edges);
  void free();
  void set_uniform(int num_edges, int state_dim, int control_dim, int node_c_dim,
                   int node_g_dim, int edge_c_dim, int edge_g_dim, int theta_dim = 0);

  int get_schur_dim() const;
  int get_state_dim(int node) const;
  int get_control_dim(int edge) const;
  int get_node_c_dim(int node) const;
  int get_node_g_dim(int node) const;
  int get_edge_c_dim(int edge) const;
  int get_edge_g_dim(int edge) const;
  int get_stagewise_x_dim(int num_edges) const;
  int get_x_dim(int num_edges) const;
  int get_y_dim(int num_edges) const;
  int get_z_dim(int num_edges) const;
  int get_stagewise_kkt_dim(int num_edges) const;

 private:
  int *owned_[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

class LQR {
 public:
  enum class FactorStatus {
    SUCCESS = 0,
    INVALID_DELTA = 1,
    F_FACTORIZATION_FAILURE = 2,
    G_FACTORIZATION_FAILURE = 3,
    INVALID_TOPOLOGY = 4,
  };

  // Views of caller-owned memory, one pointer per node / edge, column-major blocks.
  struct Input {
    double **Q;
    double **M;
    double **R;
    double **q;
    double **r;
    double **A;
    double **B;
    double **c;
    double **delta;

    const Dimensions &dimensions;
    const Topology &topology;
  };

  struct Output {
    double **x;
    double **u;
    double **y;

    void reserve(int num_edges);  // the three pointer tables
    void free();
  };

  // The reference keeps W, K, V, the two Cholesky factors ... here; the GPU engine
  // keeps its factorization on the device, so the workspace owns the engine handle
  // and the flat host staging arrays.
  struct Workspace {
    void reserve(int state_dim, int control_dim, int num_edges);
    void reserve(const Dimensions &dimensions, const Topology &topology);
    void free(int num_edges = 0);

    sipoc_engine *engine = nullptr;
    std::vector<double> in[9];   // Q M R q r A B c delta, flat
    std::vector<double> out[3];  // x u y, flat
  };

  LQR(const Input &data, Workspace &workspace);

  auto compile_topology() -> FactorStatus;
  FactorStatus factor_with_status();
  bool factor();
  void solve(Output &output);

 private:
  const Input &input_;
  Workspace &workspace_;
  FactorStatus traversal_status_;
};

// The same operations on `batch` problems of one structure, device resident in the
// engine layout X[flat * batch_stride() + problem] (include/sipoc.h).
class BatchedLQR {
 public:
  BatchedLQR(const Dimensions &dimensions, const Topology &topology, int64_t batch,
             int device = -1);
  ~BatchedLQR();
  BatchedLQR(const BatchedLQR &) = delete;
  BatchedLQR &operator=(const BatchedLQR &) = delete;

  LQR::FactorStatus topology_status() const { return status_; }
  int64_t batch_stride() const;
  sipoc_engine *engine() const { return engine_; }

  // Device pointers; `status` is int[batch_stride()] with FactorStatus values.
  struct DeviceInput {
    const double *Q, *M, *R, *q, *r, *A, *B, *c, *delta;
  };
  struct DeviceOutput {
    double *x, *u, *y;
  };
  bool factor_with_status(const DeviceInput &in, int *status, void *stream = nullptr);
  bool solve(const DeviceInput &in, const DeviceOutput &out, void *stream = nullptr);
  bool factor_solve(const DeviceInput &in, const DeviceOutput &out, int *status,
                    void *stream = nullptr);

 private:
  sipoc_engine *engine_ = nullptr;
  LQR::FactorStatus status_ = LQR::FactorStatus::SUCCESS;
};

}  // namespace sip::optimal_control
