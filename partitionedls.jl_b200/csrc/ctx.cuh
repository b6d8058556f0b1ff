// The context object behind the C ABI and the internal helpers shared by api.cu and multi.cu.
#pragma once
#include <string>
#include <vector>

#include "common.cuh"

struct pls_ctx {
  int dev = 0, sm_count = 0;
  cudaStream_t stream = nullptr;
  pls::Problem pb;
  pls::SolveWs ws;
  std::vector<uint64_t> h_gmask;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  pls_stats stats;
  double *d_w = nullptr, *d_ssq = nullptr;
  double *h_pin = nullptr;   // pinned: winner record + ssq
  size_t z_bytes = 0;
  // staging of PAGEABLE host input (a plain Julia Array): two pinned bounce buffers + the events of their DMAs
  double *h_stage[2] = {nullptr, nullptr};
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};
  size_t h_stage_bytes = 0;
  int launches = 0, launch_mark = 0;
  // single-process multi-GPU (pls_create with n_dev > 1): one full context per device; this object only
  // orchestrates (multi.cu).  Empty for an ordinary one-GPU context.
  std::vector<pls_ctx *> subs;
  double *stage = nullptr;      // on subs[0]'s device: the other devices' raw Gram sums
  size_t stage_bytes = 0;
  std::vector<int64_t> row0;    // row shard boundaries (size subs.size() + 1)
  void *pool = nullptr;         // multi.cu: one persistent host worker per device
};

namespace pls {
int check_ctx(pls_ctx *c);
int host_d(const std::vector<uint64_t> &gm, int m, int64_t b);
void read_counters(pls_ctx *c, const unsigned long long *h);
int solve_range_dev(pls_ctx *c, int64_t b_begin, int64_t b_count, bool want_obj, bool want_alpha, bool pairs = false, int force_variant = 0);
int polish_if_drifted(pls_ctx *c, bool pairs, bool *did);
int fallback_if_stalled(pls_ctx *c, int64_t b_begin, int64_t b_count, bool want_obj, bool want_alpha, bool pairs, bool *did);
bool use_pairs(const pls_ctx *c, uint32_t flags, bool per_orthant_outputs);
double eta_term(const pls_ctx *c, const double *alpha_raw, int64_t b);
double eta_term_w(const pls_ctx *c, const double *w);
int residual_partial_w(pls_ctx *c, const double *w, double *ssq_out);
double now_ms();
// multi.cu
int multi_create(pls_ctx *c, const int *device_ids, int n_dev);
int multi_predict(pls_ctx *c, const double *w, double *yhat, int (*predict_dev)(pls_ctx *, const double *, double *));
void multi_destroy(pls_ctx *c);
int multi_load(pls_ctx *c, const double *X, int64_t N, int64_t ldx, int64_t M, const double *y, const int64_t *P,
               int64_t K, double eta);
int multi_opt_fit_resident(pls_ctx *c, uint32_t flags, double *alpha_raw, int64_t *b_best, double *obj_best,
                           double *all_obj, double *all_alpha, pls_stats *stats);
int multi_bnb_fit_resident(pls_ctx *c, uint32_t flags, double *alpha_signed, double *obj_out, int64_t *nopen, pls_stats *stats);
int multi_alt_fit_resident(pls_ctx *c, const double *beta0, int64_t R, double eps, int64_t T, uint32_t flags,
                           double *alpha, double *beta, double *obj_out, int64_t *best_restart, int64_t *iters,
                           double *all_obj, pls_stats *stats);
}  // namespace pls
