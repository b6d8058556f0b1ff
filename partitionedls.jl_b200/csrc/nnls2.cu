// K2 (fast path, M' <= 208): batched orthant NNLS with BLOCK pivoting on FP64 tensor cores.
//
// Same mathematics and the same reference lines as nnls.cu (src/PartitionedLSOpt.jl:85-96), but the
// inverse H = inv(G_FF) of the passive block is kept as a tile-packed SYMMETRIC matrix in shared
// memory (8x8 tiles, lower triangle of tiles only, XOR-swizzled so that every DMMA fragment access
// is bank-conflict free), and variables move in/out of the passive set up to 8 at a time:
//
//   remove R:  H <- H - H[:,R] inv(H[R,R]) H[R,:]                      (one rank-8 DMMA update)
//   add    A:  U = H G[F,A];  S = G[A,A] - G[A,F] U;  T = U inv(S)
//              H <- [H + T U', -T; -T', inv(S)]                         (one DMMA product + one update)
//
// A Gray-code step flips one group's sign, i.e. exchanges ~|group| variables: one removal block and
// one addition block instead of ~|group| rank-1 sweeps over H.  Slots freed by removals are reused by
// additions, so H is never compacted.  Gradient evaluations stream the passive columns of G from L2.
#include "common.cuh"

namespace pls {
namespace {

constexpr int T = 512;
constexpr int NW = T / 32;
constexpr int CAPMAX = 208;

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ bool lex_better(double oa, long long ba, double ob, long long bb) {
  if (bb < 0) return ba >= 0;
  if (ba < 0) return false;
  const bool na = oa != oa, nb = ob != ob;
  if (na != nb) return na;
  if (na) return ba < bb;
  return oa < ob || (oa == ob && ba < bb);
}

// ---- tile-packed symmetric storage --------------------------------------------------------------
__device__ __forceinline__ int tile_base(int ti, int tj) { return (((ti * (ti + 1)) >> 1) + tj) << 6; }
__device__ __forceinline__ int swz8(int r, int c) { return (r << 3) + (c ^ ((r & 2) << 1)); }
__device__ __forceinline__ double h_get(const double *H, int i, int j) {
  const int a = i > j ? i : j, b = i > j ? j : i;
  return H[tile_base(a >> 3, b >> 3) + swz8(a & 7, b & 7)];
}
__device__ __forceinline__ void h_set(double *H, int i, int j, double v) {
  const int a = i > j ? i : j, b = i > j ? j : i;
  const int base = tile_base(a >> 3, b >> 3);
  H[base + swz8(a & 7, b & 7)] = v;
  if ((a >> 3) == (b >> 3)) H[base + swz8(b & 7, a & 7)] = v;   // diagonal tiles hold both halves
}
// panels: [rows][8], swizzled like a tile column
__device__ __forceinline__ int pan(int row, int col) { return (row << 3) + (col ^ ((row & 2) << 1)); }

struct Sh {
  double *H, *Pa, *Pb, *w, *r, *wF, *rpart, *Sinv, *Gaa, *Spart, *rho, *theta, *cA, *red;
  int *F, *pos, *lst, *Rs, *As, *Av, *ctl;
  signed char *sg, *dd, *vflag;
};

// H(lower tiles) += Pa * Pb'   over the leading nt x nt tiles
__device__ __forceinline__ void rank_update(double *H, const double *Pa, const double *Pb, int nt) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int ntiles = (nt * (nt + 1)) >> 1;
  for (int q = wid; q < ntiles; q += NW) {
    int ti = (int)((sqrtf(8.f * (float)q + 1.f) - 1.f) * 0.5f);
    while (((ti + 1) * (ti + 2)) / 2 <= q) ++ti;
    while ((ti * (ti + 1)) / 2 > q) --ti;
    const int tj = q - (ti * (ti + 1)) / 2;
    const double a0 = Pa[pan(ti * 8 + fr, fk)], a1 = Pa[pan(ti * 8 + fr, 4 + fk)];
    const double b0 = Pb[pan(tj * 8 + fr, fk)], b1 = Pb[pan(tj * 8 + fr, 4 + fk)];
    double2 *cp = reinterpret_cast<double2 *>(H + (q << 6) + swz8(fr, fk * 2));
    double2 c = *cp;
    dmma(c.x, c.y, a0, b0);
    dmma(c.x, c.y, a1, b1);
    *cp = c;
  }
}

// Pout = H * Pin  (H symmetric, nt x nt tiles; panels nt*8 x 8)
__device__ __forceinline__ void hmul(const double *H, const double *Pin, double *Pout, int nt) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  for (int ti = wid; ti < nt; ti += NW) {
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
    for (int tj = 0; tj < nt; ++tj) {
      double a0, a1;
      if (tj <= ti) {
        const int base = tile_base(ti, tj);
        a0 = H[base + swz8(fr, fk)]; a1 = H[base + swz8(fr, 4 + fk)];
      } else {
        const int base = tile_base(tj, ti);
        a0 = H[base + swz8(fk, fr)]; a1 = H[base + swz8(4 + fk, fr)];
      }
      const double b0 = Pin[pan(tj * 8 + fk, fr)], b1 = Pin[pan(tj * 8 + 4 + fk, fr)];
      if (tj & 1) { dmma(e0, e1, a0, b0); dmma(e0, e1, a1, b1); }
      else { dmma(c0, c1, a0, b0); dmma(c0, c1, a1, b1); }
    }
    *reinterpret_cast<double2 *>(Pout + pan(ti * 8 + fr, fk * 2)) = make_double2(c0 + e0, c1 + e1);
  }
}

// In-place Gauss-Jordan inverse of an SPD 8x8 held by one warp: lane owns S[i][j0], S[i][j0+1]
// with i = lane >> 2, j0 = 2 * (lane & 3).  Only the leading n x n block is eliminated (the rest must
// be the identity).  Returns false if a pivot is not safely positive (relative to dref[k]).
__device__ __forceinline__ bool warp_inv8(double &e0, double &e1, int n, const double *dref, double tol) {
  const int lane = threadIdx.x & 31;
  const int i = lane >> 2, j0 = (lane & 3) << 1, j1 = j0 + 1;
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    const int srow = (k << 2) + (lane & 3);
    const double pk0 = __shfl_sync(0xffffffffu, e0, srow), pk1 = __shfl_sync(0xffffffffu, e1, srow);
    const int scol = (lane & ~3) + (k >> 1);
    const double ca = __shfl_sync(0xffffffffu, e0, scol), cb = __shfl_sync(0xffffffffu, e1, scol);
    const double cik = (k & 1) ? cb : ca;
    const int sp = (k << 2) + (k >> 1);
    const double pa = __shfl_sync(0xffffffffu, e0, sp), pb = __shfl_sync(0xffffffffu, e1, sp);
    const double pkk = (k & 1) ? pb : pa;
    if (!(pkk > tol * dref[k])) ok = false;
    const double d = 1.0 / pkk;
    if (i == k) {
      e0 = (j0 == k) ? d : pk0 * d;
      e1 = (j1 == k) ? d : pk1 * d;
    } else {
      const double f = cik * d;
      e0 = (j0 == k) ? -f : fma(-f, pk0, e0);
      e1 = (j1 == k) ? -f : fma(-f, pk1, e1);
    }
  }
  return ok;
}

// Pout[row][q] = sign * sum_j Pin[row][j] * Sinv[j][q];   w[F[row]] -= sum_j Pin[row][j] * coef[j]
__device__ __forceinline__ void panel_small(const Sh &s, const double *Pin, double *Pout, const double *coef,
                                            double sign, int nrows) {
  for (int item = threadIdx.x; item < nrows * 8; item += T) {
    const int row = item >> 3, q = item & 7;
    double acc = 0.0, z = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double x = Pin[pan(row, j)];
      acc = fma(x, s.Sinv[j * 8 + q], acc);
      z = fma(x, coef[j], z);
    }
    Pout[pan(row, q)] = sign * acc;
    if (q == 0) { const int var = s.F[row]; if (var >= 0) s.w[var] -= z; }
  }
}

struct State {
  int hwm, nt, p;
  unsigned long long n_piv, n_grad, n_sump, n_sump2, n_iter, n_rebuild, n_blocked, n_noconv;
};

// Remove the r <= 8 slots listed in s.Rs.
__device__ void block_remove(const Sh &s, State &st, int r) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nrows = st.nt * 8;
  for (int item = tid; item < nrows * 8; item += T) {
    const int q = item / nrows, row = item - q * nrows;
    s.Pb[pan(row, q)] = (q < r) ? h_get(s.H, row, s.Rs[q]) : 0.0;
  }
  __syncthreads();
  if (wid == 0) {
    const int i = lane >> 2, j0 = (lane & 3) << 1;
    double e0 = (i < r && j0 < r) ? s.Pb[pan(s.Rs[i], j0)] : (i == j0 ? 1.0 : 0.0);
    double e1 = (i < r && j0 + 1 < r) ? s.Pb[pan(s.Rs[i], j0 + 1)] : (i == j0 + 1 ? 1.0 : 0.0);
    if (lane < 8) s.rho[lane] = (lane < r) ? s.w[s.F[s.Rs[lane]]] : 0.0;    // w_R
    if (lane < 8) s.cA[lane] = 0.0;
    __syncwarp();
    warp_inv8(e0, e1, r, s.cA, -1.0);      // H[R,R] is SPD; no pivot test (tol < 0 with dref = 0)
    s.Sinv[i * 8 + j0] = e0; s.Sinv[i * 8 + j0 + 1] = e1;
    __syncwarp();
    if (lane < 8) {      // phi = inv(S) w_R
      double a = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) a = fma(s.Sinv[lane * 8 + j], s.rho[j], a);
      s.theta[lane] = a;
    }
  }
  __syncthreads();
  panel_small(s, s.Pb, s.Pa, s.theta, -1.0, nrows);
  __syncthreads();
  rank_update(s.H, s.Pa, s.Pb, st.nt);
  __syncthreads();
  for (int item = tid; item < nrows * 8; item += T) {
    const int q = item / nrows, row = item - q * nrows;
    if (q < r) h_set(s.H, s.Rs[q], row, 0.0);
  }
  if (tid < r) {
    const int sl = s.Rs[tid], var = s.F[sl];
    s.w[var] = 0.0; s.pos[var] = -1; s.F[sl] = -1;
  }
  __syncthreads();
  st.p -= r;
  st.n_piv += r; st.n_sump2 += (unsigned long long)r * st.p * st.p;
}

// Add the a <= 8 variables listed in s.Av into the free slots s.As.  Returns false (state
// untouched) if the Schur complement is not safely positive definite.
__device__ bool block_add(const Sh &s, State &st, const K2Args &A, int a) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int hw = st.hwm;
  for (int q = 0; q < a; ++q) hw = max(hw, s.As[q] + 1);
  const int nt = (hw + 7) >> 3, nrows = nt * 8;
  for (int item = tid; item < nrows * 8; item += T) {
    const int q = item / nrows, row = item - q * nrows;
    const int var = s.F[row];
    s.Pa[pan(row, q)] = (q < a && var >= 0) ? A.G[(size_t)A.ldg * s.Av[q] + var] : 0.0;
    if (q == 0) s.wF[row] = var >= 0 ? s.w[var] : 0.0;
  }
  if (tid < 64) {
    const int i = tid >> 3, j = tid & 7;
    s.Gaa[tid] = (i < a && j < a) ? A.G[(size_t)A.ldg * s.Av[j] + s.Av[i]] : (i == j ? 1.0 : 0.0);
  }
  if (tid >= 64 && tid < 72) s.cA[tid - 64] = (tid - 64 < a) ? A.c[s.Av[tid - 64]] : 0.0;
  __syncthreads();
  hmul(s.H, s.Pa, s.Pb, nt);                        // U = H V
  __syncthreads();
  if (wid < 4) {                                    // S partials = V' U over interleaved k-steps
    const int fr = lane >> 2, fk = lane & 3;
    double c0 = 0.0, c1 = 0.0;
    for (int ks = wid; ks < nt * 2; ks += 4) {
      const int row = ks * 4 + fk;
      dmma(c0, c1, s.Pa[pan(row, fr)], s.Pb[pan(row, fr)]);
    }
    s.Spart[wid * 64 + fr * 8 + fk * 2] = c0;
    s.Spart[wid * 64 + fr * 8 + fk * 2 + 1] = c1;
  } else if (wid < 12) {                            // rho_i = c_i - V[:,i]' w
    const int i = wid - 4;
    double acc = 0.0;
    for (int row = lane; row < nrows; row += 32) acc = fma(s.Pa[pan(row, i)], s.wF[row], acc);
    acc = warp_sum(acc);
    if (lane == 0) s.rho[i] = s.cA[i] - acc;
  }
  __syncthreads();
  if (wid == 0) {
    const int i = lane >> 2, j0 = (lane & 3) << 1;
    double e0 = s.Gaa[i * 8 + j0], e1 = s.Gaa[i * 8 + j0 + 1];
    if (i < a) {
#pragma unroll
      for (int q = 0; q < 4; ++q) { e0 -= s.Spart[q * 64 + i * 8 + j0]; e1 -= s.Spart[q * 64 + i * 8 + j0 + 1]; }
      if (j0 >= a) e0 = 0.0;
      if (j0 + 1 >= a) e1 = 0.0;
    }
    if (lane < 8) s.theta[lane] = s.Gaa[lane * 8 + lane];     // pivot reference: G_jj
    __syncwarp();
    const bool ok = warp_inv8(e0, e1, a, s.theta, 1e-13);
    const bool all_ok = __all_sync(0xffffffffu, ok);
    s.Sinv[i * 8 + j0] = e0; s.Sinv[i * 8 + j0 + 1] = e1;
    __syncwarp();
    double th = 0.0;
    if (lane < 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) th = fma(s.Sinv[lane * 8 + j], s.rho[j], th);
    }
    __syncwarp();
    if (lane < 8) s.theta[lane] = (lane < a) ? th : 0.0;
    if (lane == 0) s.ctl[4] = all_ok ? 1 : 0;
  }
  __syncthreads();
  if (!s.ctl[4]) { __syncthreads(); return false; }
  panel_small(s, s.Pb, s.Pa, s.theta, 1.0, nrows);  // T = U inv(S) -> Pa;  w_F -= U theta
  __syncthreads();
  rank_update(s.H, s.Pa, s.Pb, nt);                 // H += T U'
  __syncthreads();
  for (int item = tid; item < nrows * 8; item += T) {   // new rows / columns: -T
    const int q = item / nrows, row = item - q * nrows;
    if (q < a && s.F[row] >= 0) h_set(s.H, s.As[q], row, -s.Pa[pan(row, q)]);
  }
  __syncthreads();
  if (tid < 64) {
    const int i = tid >> 3, j = tid & 7;
    if (i < a && j <= i) h_set(s.H, s.As[i], s.As[j], s.Sinv[i * 8 + j]);
  }
  if (tid < a) {
    const int var = s.Av[tid], sl = s.As[tid];
    s.w[var] = s.theta[tid]; s.F[sl] = var; s.pos[var] = sl;
  }
  __syncthreads();
  st.n_piv += a; st.n_sump2 += (unsigned long long)a * st.p * st.p;
  st.hwm = hw; st.nt = nt; st.p += a;
  return true;
}

// r = c - G[:,F] w_F for all variables.  Threads 0..255 take the even half of the slots, 256..511
// the odd half.  Returns max |r_F| (normal-equation residual), same on all threads.
__device__ double grad_eval(const Sh &s, State &st, const K2Args &A) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int hw = st.hwm, Mp = A.Mp;
  for (int t = tid; t < hw; t += T) { const int var = s.F[t]; s.wF[t] = var >= 0 ? s.w[var] : 0.0; }
  __syncthreads();
  const int half = tid >> 8, m = tid & 255;
  const int mid = ((hw + 1) >> 1);
  const int t0 = half ? mid : 0, t1 = half ? hw : mid;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (m < Mp) {
    int t = t0;
    for (; t + 3 < t1; t += 4) {
      const int v0 = s.F[t], v1 = s.F[t + 1], v2 = s.F[t + 2], v3 = s.F[t + 3];
      const double g0 = v0 >= 0 ? A.G[(size_t)A.ldg * v0 + m] : 0.0;
      const double g1 = v1 >= 0 ? A.G[(size_t)A.ldg * v1 + m] : 0.0;
      const double g2 = v2 >= 0 ? A.G[(size_t)A.ldg * v2 + m] : 0.0;
      const double g3 = v3 >= 0 ? A.G[(size_t)A.ldg * v3 + m] : 0.0;
      a0 = fma(g0, s.wF[t], a0); a1 = fma(g1, s.wF[t + 1], a1);
      a2 = fma(g2, s.wF[t + 2], a2); a3 = fma(g3, s.wF[t + 3], a3);
    }
    for (; t < t1; ++t) { const int v0 = s.F[t]; if (v0 >= 0) a0 = fma(A.G[(size_t)A.ldg * v0 + m], s.wF[t], a0); }
  }
  const double part = (a0 + a1) + (a2 + a3);
  if (half) s.rpart[m] = part;
  __syncthreads();
  double mx = 0.0;
  if (!half && m < Mp) {
    const double rv = A.c[m] - (part + s.rpart[m]);
    s.r[m] = rv;
    if (s.pos[m] >= 0) mx = fabs(rv);
  }
  mx = warp_max(mx);
  if (lane == 0) s.red[wid] = mx;
  __syncthreads();
  double x = s.red[0];
#pragma unroll
  for (int i = 1; i < NW; ++i) x = fmax(x, s.red[i]);
  __syncthreads();
  st.n_grad++; st.n_sump += (unsigned long long)st.p;
  return x;
}

// w_F += H r_F   (one DMMA product with a single live column)
__device__ void refine(const Sh &s, State &st) {
  const int nrows = st.nt * 8;
  for (int item = threadIdx.x; item < nrows * 8; item += T) {
    const int q = item / nrows, row = item - q * nrows;
    const int var = s.F[row];
    s.Pa[pan(row, q)] = (q == 0 && var >= 0) ? s.r[var] : 0.0;
  }
  __syncthreads();
  hmul(s.H, s.Pa, s.Pb, st.nt);
  __syncthreads();
  for (int row = threadIdx.x; row < nrows; row += T) {
    const int var = s.F[row];
    if (var >= 0) s.w[var] += s.Pb[pan(row, 0)];
  }
  __syncthreads();
}

__device__ void clear_state(const Sh &s, State &st, const K2Args &A, int cap) {
  const int ntc = cap >> 3;
  const int words = ((ntc * (ntc + 1)) >> 1) << 6;
  for (int i = threadIdx.x; i < words; i += T) s.H[i] = 0.0;
  for (int t = threadIdx.x; t < cap; t += T) s.F[t] = -1;
  st.hwm = 0; st.nt = 0; st.p = 0;
}

__global__ void __launch_bounds__(T, 1) k2v2_orthant_chains(const K2Args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int Mp = A.Mp, cap = A.cap;            // cap % 8 == 0, cap >= Mp
  const int ntc = cap >> 3;
  Sh s;
  double *dp = reinterpret_cast<double *>(smem_raw);
  s.H = dp; dp += (((ntc * (ntc + 1)) >> 1) << 6);
  s.Pa = dp; dp += cap * 8;
  s.Pb = dp; dp += cap * 8;
  s.w = dp; dp += cap;
  s.r = dp; dp += cap;
  s.wF = dp; dp += cap;
  s.rpart = dp; dp += 256;
  s.Sinv = dp; dp += 64;
  s.Gaa = dp; dp += 64;
  s.Spart = dp; dp += 256;
  s.rho = dp; dp += 8;
  s.theta = dp; dp += 8;
  s.cA = dp; dp += 8;
  s.red = dp; dp += NW;
  int *ip = reinterpret_cast<int *>(dp);
  s.F = ip; ip += cap;
  s.pos = ip; ip += cap;
  s.lst = ip; ip += cap;
  s.Rs = ip; ip += 8;
  s.As = ip; ip += 8;
  s.Av = ip; ip += 8;
  s.ctl = ip; ip += 8;
  signed char *cp = reinterpret_cast<signed char *>(ip);
  s.sg = cp; cp += cap;
  s.dd = cp; cp += cap;
  s.vflag = cp; cp += cap;

  State st;
  st.n_piv = st.n_grad = st.n_sump = st.n_sump2 = st.n_iter = st.n_rebuild = st.n_blocked = st.n_noconv = 0;
  const double yy = A.scal[0], cmax = A.scal[1];
  const double told = 1e-12 * cmax;
  const long long L = 1ll << A.chain_log2;
  double best_obj = 0.0; long long best_b = -1;
  __shared__ unsigned long long s_chain;

  for (;;) {   // persistent CTA: fetch chains until the range is exhausted
    if (tid == 0) s_chain = atomicAdd(A.chain_counter, 1ull);
    __syncthreads();
    const unsigned long long chain = s_chain;
    __syncthreads();
    if (chain >= (unsigned long long)A.n_chains) break;
    const long long base = A.b_begin + (long long)chain * L;

    clear_state(s, st, A, cap);
    for (int m = tid; m < Mp; m += T) { s.w[m] = 0.0; s.r[m] = A.c[m]; s.pos[m] = -1; }
    __syncthreads();
    bool r_valid = true;

    for (long long i = 0; i < L; ++i) {
      const long long b = base + (i ^ (i >> 1));
      for (int m = tid; m < Mp; m += T) {   // d_m = sum_k Po[m,k] (2 bit_k(b) - 1)   (Opt.jl:28-29)
        const uint64_t gm = A.gmask[m];
        const int d = 2 * __popcll(gm & (uint64_t)b) - __popcll(gm);
        s.dd[m] = (signed char)d; s.sg[m] = (signed char)((d > 0) - (d < 0));
        s.vflag[m] = 0;
      }
      __syncthreads();

      int t_best = Mp + 1, pbar = 3, iters = 0;
      bool ok = true;
      for (;;) {
        if (!r_valid) {
          int rep = 0;
          for (;;) {
            const double rf = grad_eval(s, st, A);
            if (rf <= 1e-12 * cmax) break;          // carried solution is already exact to working accuracy
            refine(s, st);
            if (rf <= 1e-9 * cmax) break;
            if (++rep >= 4) {                       // inverse degraded: rebuild by re-adding the passive set
              int pn = 0;
              if (tid == 0) {
                int n = 0;
                for (int t = 0; t < st.hwm; ++t) if (s.F[t] >= 0) s.lst[n++] = s.F[t];
                s.ctl[5] = n;
              }
              __syncthreads();
              pn = s.ctl[5];
              clear_state(s, st, A, cap);
              for (int m = tid; m < Mp; m += T) { s.w[m] = 0.0; s.pos[m] = -1; }
              __syncthreads();
              for (int q0 = 0; q0 < pn; q0 += 8) {
                const int a = min(8, pn - q0);
                if (tid < a) { s.Av[tid] = s.lst[q0 + tid]; s.As[tid] = st.p + tid; }
                __syncthreads();
                block_add(s, st, A, a);
              }
              st.n_rebuild++;
              if (rep >= 6) { ok = false; break; }
            }
          }
          if (!ok) break;
        }
        r_valid = false;
        // ---- infeasibility sets
        for (int m = tid; m < Mp; m += T) {
          const int sl = s.pos[m], sg = s.sg[m];
          int f = 0;
          if (sl >= 0) { if (sg == 0 || (double)sg * s.w[m] < 0.0) f = 1; }
          else if (sg != 0 && s.vflag[m] != 3 && (double)sg * s.r[m] > told) f = 2;
          if (s.vflag[m] != 3) s.vflag[m] = (signed char)f;
        }
        __syncthreads();
        if (wid == 0) {   // deterministic compaction in index order: removals from the front, additions from the back
          int nr = 0, na = 0, mxi = -1;
          for (int m0 = 0; m0 < Mp; m0 += 32) {
            const int m = m0 + lane;
            const int f = (m < Mp) ? s.vflag[m] : 0;
            const unsigned br = __ballot_sync(0xffffffffu, f == 1);
            const unsigned ba = __ballot_sync(0xffffffffu, f == 2);
            if (f == 1) s.lst[nr + __popc(br & ((1u << lane) - 1))] = m;
            if (f == 2) s.lst[cap - 1 - (na + __popc(ba & ((1u << lane) - 1)))] = m;
            nr += __popc(br); na += __popc(ba);
            const unsigned any = br | ba;
            if (any) mxi = m0 + 31 - __clz(any);
          }
          if (lane == 0) { s.ctl[0] = nr; s.ctl[1] = na; s.ctl[2] = mxi; }
        }
        __syncthreads();
        int nr = s.ctl[0], na = s.ctl[1];
        const int nv = nr + na;
        if (nv == 0) { r_valid = true; break; }
        bool single = false;
        if (nv < t_best) { t_best = nv; pbar = 3; }
        else if (pbar >= 1) { --pbar; }
        else single = true;                         // Murty's rule: only the highest index moves
        if (single) {
          const int m = s.ctl[2];
          if (s.pos[m] >= 0) { nr = 1; na = 0; if (tid == 0) s.lst[0] = m; }
          else { nr = 0; na = 1; if (tid == 0) s.lst[cap - 1] = m; }
          __syncthreads();
        }
        // ---- removals, 8 at a time
        for (int q0 = 0; q0 < nr; q0 += 8) {
          const int r = min(8, nr - q0);
          if (tid < r) s.Rs[tid] = s.pos[s.lst[q0 + tid]];
          __syncthreads();
          block_remove(s, st, r);
        }
        // ---- additions, 8 at a time, into the lowest free slots
        for (int q0 = 0; q0 < na; q0 += 8) {
          const int a = min(8, na - q0);
          if (wid == 0) {
            if (lane < a) s.Av[lane] = s.lst[cap - 1 - (q0 + lane)];
            int found = 0;
            for (int s0 = 0; s0 < cap && found < a; s0 += 32) {
              const int sl = s0 + lane;
              const bool fr = sl < cap && s.F[sl] < 0;
              const unsigned bal = __ballot_sync(0xffffffffu, fr);
              const int rank = found + __popc(bal & ((1u << lane) - 1));
              if (fr && rank < a) s.As[rank] = sl;
              found += __popc(bal);
            }
          }
          __syncthreads();
          if (!block_add(s, st, A, a)) {
            // numerically dependent column in the block: retry one variable at a time
            for (int q = 0; q < a; ++q) {
              __syncthreads();
              if (wid == 0) {
                const int var = s.lst[cap - 1 - (q0 + q)];
                int found = 0;
                for (int s0 = 0; s0 < cap && found < 1; s0 += 32) {
                  const int sl = s0 + lane;
                  const bool fr = sl < cap && s.F[sl] < 0;
                  const unsigned bal = __ballot_sync(0xffffffffu, fr);
                  if (fr && __popc(bal & ((1u << lane) - 1)) == 0 && found == 0) s.As[0] = sl;
                  found += __popc(bal);
                }
                if (lane == 0) s.Av[0] = var;
              }
              __syncthreads();
              if (!block_add(s, st, A, 1)) {
                if (tid == 0) s.vflag[s.Av[0]] = 3;
                st.n_blocked++;
                __syncthreads();
              }
            }
          }
        }
        // high-water mark may have dropped
        if (wid == 0) {
          int hw = 0;
          for (int s0 = 0; s0 < cap; s0 += 32) {
            const int sl = s0 + lane;
            const unsigned bal = __ballot_sync(0xffffffffu, sl < cap && s.F[sl] >= 0);
            if (bal) hw = s0 + 32 - __clz(bal);
          }
          if (lane == 0) s.ctl[3] = hw;
        }
        __syncthreads();
        st.hwm = s.ctl[3]; st.nt = (st.hwm + 7) >> 3;
        st.n_iter++;
        if (++iters > 60 + 6 * Mp) { ok = false; break; }
      }
      if (!ok) st.n_noconv++;

      // ---- objective  sqrt(yy - c_F' w_F)   (= norm(Xa w - ya) at the KKT point, Opt.jl:90)
      double acc = 0.0;
      for (int m = tid; m < Mp; m += T) if (s.pos[m] >= 0) acc = fma(A.c[m], s.w[m], acc);
      acc = warp_sum(acc);
      if (lane == 0) s.red[wid] = acc;
      __syncthreads();
      double tot = 0.0;
#pragma unroll
      for (int q = 0; q < NW; ++q) tot += s.red[q];
      const double obj = ok ? sqrt(fmax(yy - tot, 0.0)) : __longlong_as_double(0x7ff8000000000000ll);
      const long long rel = b - A.b_begin;
      if (A.all_obj && tid == 0) A.all_obj[rel] = obj;
      if (A.all_alpha) {
        for (int m = tid; m < Mp; m += T) {
          const int d = s.dd[m];
          A.all_alpha[(size_t)rel * Mp + m] = (s.pos[m] >= 0 && d != 0) ? fmax(s.w[m] / (double)d, 0.0) : 0.0;
        }
      }
      if (lex_better(obj, b, best_obj, best_b)) {
        best_obj = obj; best_b = b;
        for (int m = tid; m < Mp; m += T) {
          const int d = s.dd[m];
          A.cta_w[(size_t)blockIdx.x * Mp + m] = (s.pos[m] >= 0 && d != 0) ? fmax(s.w[m] / (double)d, 0.0) : 0.0;
        }
      }
      __syncthreads();
    }
  }  // chains
  if (tid == 0) {
    A.cta_obj[blockIdx.x] = best_obj;
    A.cta_b[blockIdx.x] = best_b;
    atomicAdd(&A.counters[CNT_PIVOTS], st.n_piv);
    atomicAdd(&A.counters[CNT_GRAD], st.n_grad);
    atomicAdd(&A.counters[CNT_SUMP], st.n_sump);
    atomicAdd(&A.counters[CNT_SUMP2], st.n_sump2);
    atomicAdd(&A.counters[CNT_ITERS], st.n_iter);
    atomicAdd(&A.counters[CNT_REBUILDS], st.n_rebuild);
    atomicAdd(&A.counters[CNT_BLOCKED], st.n_blocked);
    atomicAdd(&A.counters[CNT_NOCONV], st.n_noconv);
  }
}

size_t v2_smem_bytes(int cap) {
  const int ntc = cap >> 3;
  size_t d = ((size_t)(ntc * (ntc + 1) / 2) << 6) + 2 * (size_t)cap * 8 + 3 * (size_t)cap + 256 + 64 + 64 + 256 + 24 + NW;
  size_t i = 3 * (size_t)cap + 32;
  size_t c = 3 * (size_t)cap;
  return d * sizeof(double) + i * sizeof(int) + c + 16;
}

}  // namespace

int k2v2_config(int Mp, int *cap, size_t *smem, int *occ) {
  if (Mp > CAPMAX) return PLS_EUNSUPPORTED;
  const int cp = (Mp + 7) & ~7;
  const size_t sm = v2_smem_bytes(cp);
  int dev = 0, max_smem = 0;
  PLS_CUDA_TRY(cudaGetDevice(&dev));
  PLS_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (sm > (size_t)max_smem) return PLS_EUNSUPPORTED;
  PLS_CUDA_TRY(cudaFuncSetAttribute(k2v2_orthant_chains, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  int oc = 1;
  PLS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc, k2v2_orthant_chains, T, sm));
  if (oc < 1) oc = 1;
  *cap = cp; *smem = sm; *occ = oc;
  return PLS_OK;
}

int k2v2_launch(const K2Args &A, int grid, cudaStream_t st) {
  size_t sm = v2_smem_bytes(A.cap);
  k2v2_orthant_chains<<<grid, T, sm, st>>>(A);
  PLS_CUDA_TRY(cudaGetLastError());
  return PLS_OK;
}

}  // namespace pls
