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
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
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
  double *H, *Pa, *Pb, *w, *r, *wF, *rpart, *Sinv, *Gaa, *Spart, *rho, *theta, *cA, *red, *cs;
  unsigned long long *gms;        // smem copies of c and the group masks
  long long *stat, *prof;
  int *F, *pos, *lst, *asl, *ctl, *pl;
  unsigned short *tmap;           // tile index -> (ti << 8) | tj
  signed char *sg, *dd, *vflag, *smark;
};

enum Phase { PH_START = 0, PH_PLAN, PH_REMOVE, PH_ADD, PH_GRAD, PH_REFINE, PH_OUT, PH_NREM, PH_NADD,
             PH_R_GATHER, PH_R_PANEL, PH_R_RANK, PH_R_ZERO, PH_A_GATHER, PH_A_HMUL, PH_A_SPART, PH_A_INV, PH_A_PANEL,
             PH_A_RANK, PH_A_ROWS, PH_A_INV1, PH_A_INV2, PH_A_INV3, PH_NUM };
#define SUBTICK(which) do { if (threadIdx.x == 0) { const long long now_ = clock64(); s.prof[which] += now_ - s.stat[ST_TSUB]; s.stat[ST_TSUB] = now_; } } while (0)

// statistics / phase timers live in shared memory and are touched by thread 0 only
enum Stat { ST_P = 0, ST_PIV, ST_GRAD, ST_SUMP, ST_SUMP2, ST_ITER, ST_REBUILD, ST_BLOCKED, ST_NOCONV, ST_TMARK, ST_TSUB, ST_NUM };
#define STAT_ADD(which, v) do { if (threadIdx.x == 0) s.stat[which] += (long long)(v); } while (0)

// Carves the dynamic shared memory; called (and fully inlined) in every device function so the
// pointers live in registers instead of a struct in local memory.
__device__ __forceinline__ Sh make_sh(int cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ntc = cap >> 3;
  const int ntiles_cap = (ntc * (ntc + 1)) >> 1;
  Sh s;
  double *dp = reinterpret_cast<double *>(smem_raw);
  s.H = dp; dp += (ntiles_cap << 6);
  s.Pa = dp; dp += cap * 8;
  s.Pb = dp; dp += cap * 8;
  s.w = dp; dp += cap;
  s.r = dp; dp += cap;
  s.wF = dp; dp += cap;
  s.Sinv = dp; dp += 64;
  s.Gaa = dp; dp += 64;
  s.Spart = dp; dp += 256;
  s.rho = dp; dp += 8;
  s.theta = dp; dp += 8;
  s.cA = dp; dp += 8;
  s.red = dp; dp += NW;
  s.cs = dp; dp += cap;
  s.gms = reinterpret_cast<unsigned long long *>(dp); dp += cap;
  s.stat = reinterpret_cast<long long *>(dp); dp += ST_NUM;
  s.prof = reinterpret_cast<long long *>(dp); dp += PH_NUM;
  s.rpart = nullptr;
  int *ip = reinterpret_cast<int *>(dp);
  s.F = ip; ip += cap;
  s.pos = ip; ip += cap;
  s.lst = ip; ip += cap;
  s.asl = ip; ip += cap;
  s.ctl = ip; ip += 8;
  s.pl = ip; ip += 32;
  s.tmap = reinterpret_cast<unsigned short *>(ip); ip += (ntiles_cap + 1) / 2;
  signed char *cp = reinterpret_cast<signed char *>(ip);
  s.sg = cp; cp += cap;
  s.dd = cp; cp += cap;
  s.vflag = cp; cp += cap;
  s.smark = cp; cp += cap;

  return s;
}

// H(lower tiles) += Pa * Pb'   over the leading nt x nt tiles
__device__ __noinline__ void rank_update(double *H, const double *Pa, const double *Pb, int nt,
                                         const unsigned short *tmap) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int ntiles = (nt * (nt + 1)) >> 1;
  const int coff = swz8(fr, fk * 2);
  // three tiles in flight per warp: loads of all three first, then their (independent) DMMA chains
  for (int q = wid; q < ntiles; q += 3 * NW) {
    const int q1 = q + NW, q2 = q + 2 * NW;
    const bool h1 = q1 < ntiles, h2 = q2 < ntiles;
    const int t0 = tmap[q], t1 = tmap[h1 ? q1 : q], t2 = tmap[h2 ? q2 : q];
    double2 *cp0 = reinterpret_cast<double2 *>(H + (q << 6) + coff);
    double2 *cp1 = reinterpret_cast<double2 *>(H + ((h1 ? q1 : q) << 6) + coff);
    double2 *cp2 = reinterpret_cast<double2 *>(H + ((h2 ? q2 : q) << 6) + coff);
    const int ra0 = ((t0 >> 8) << 3) + fr, rb0 = ((t0 & 255) << 3) + fr;
    const int ra1 = ((t1 >> 8) << 3) + fr, rb1 = ((t1 & 255) << 3) + fr;
    const int ra2 = ((t2 >> 8) << 3) + fr, rb2 = ((t2 & 255) << 3) + fr;
    const double a00 = Pa[pan(ra0, fk)], a01 = Pa[pan(ra0, 4 + fk)], b00 = Pb[pan(rb0, fk)], b01 = Pb[pan(rb0, 4 + fk)];
    const double a10 = Pa[pan(ra1, fk)], a11 = Pa[pan(ra1, 4 + fk)], b10 = Pb[pan(rb1, fk)], b11 = Pb[pan(rb1, 4 + fk)];
    const double a20 = Pa[pan(ra2, fk)], a21 = Pa[pan(ra2, 4 + fk)], b20 = Pb[pan(rb2, fk)], b21 = Pb[pan(rb2, 4 + fk)];
    double2 c0 = *cp0, c1 = *cp1, c2 = *cp2;
    dmma(c0.x, c0.y, a00, b00); dmma(c1.x, c1.y, a10, b10); dmma(c2.x, c2.y, a20, b20);
    dmma(c0.x, c0.y, a01, b01); dmma(c1.x, c1.y, a11, b11); dmma(c2.x, c2.y, a21, b21);
    *cp0 = c0;
    if (h1) *cp1 = c1;
    if (h2) *cp2 = c2;
  }
}

// Pout = H * Pin  (H symmetric, nt x nt tiles; panels nt*8 x 8)
__device__ __noinline__ void hmul(const double *H, const double *Pin, double *Pout, int nt) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int off_d0 = swz8(fr, fk), off_d1 = swz8(fr, 4 + fk);      // direct tile (tj <= ti)
  const int off_t0 = swz8(fk, fr), off_t1 = swz8(4 + fk, fr);      // transposed tile (tj > ti)
  for (int ti = wid; ti < nt; ti += NW) {
    // four independent accumulator pairs: (tile parity) x (k-step) -> short DMMA dependency chains
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0, e00 = 0.0, e01 = 0.0, e10 = 0.0, e11 = 0.0;
    const int rowbase = (ti * (ti + 1)) >> 1;
    int tj = 0;
    for (; tj + 1 < nt; tj += 2) {
      const int tk = tj + 1;
      const bool dj = tj <= ti, dk = tk <= ti;
      const int bj = dj ? ((rowbase + tj) << 6) : (tile_base(tj, ti));
      const int bk = dk ? ((rowbase + tk) << 6) : (tile_base(tk, ti));
      const double aj0 = H[bj + (dj ? off_d0 : off_t0)], aj1 = H[bj + (dj ? off_d1 : off_t1)];
      const double ak0 = H[bk + (dk ? off_d0 : off_t0)], ak1 = H[bk + (dk ? off_d1 : off_t1)];
      const double pj0 = Pin[pan(tj * 8 + fk, fr)], pj1 = Pin[pan(tj * 8 + 4 + fk, fr)];
      const double pk0 = Pin[pan(tk * 8 + fk, fr)], pk1 = Pin[pan(tk * 8 + 4 + fk, fr)];
      dmma(c00, c01, aj0, pj0); dmma(c10, c11, aj1, pj1);
      dmma(e00, e01, ak0, pk0); dmma(e10, e11, ak1, pk1);
    }
    if (tj < nt) {
      const bool dj = tj <= ti;
      const int bj = dj ? ((rowbase + tj) << 6) : (tile_base(tj, ti));
      const double aj0 = H[bj + (dj ? off_d0 : off_t0)], aj1 = H[bj + (dj ? off_d1 : off_t1)];
      dmma(c00, c01, aj0, Pin[pan(tj * 8 + fk, fr)]); dmma(c10, c11, aj1, Pin[pan(tj * 8 + 4 + fk, fr)]);
    }
    *reinterpret_cast<double2 *>(Pout + pan(ti * 8 + fr, fk * 2)) =
        make_double2((c00 + c10) + (e00 + e10), (c01 + c11) + (e01 + e11));
  }
}

// In-place Gauss-Jordan inverse of an SPD 8x8 held by one warp: lane owns S[i][j0], S[i][j0+1]
// with i = lane >> 2, j0 = 2 * (lane & 3).  Only the leading n x n block is eliminated (the rest must
// be the identity).  Returns false if a pivot is not safely positive (relative to dref[k]).
__device__ __forceinline__ bool warp_inv8(double &e0, double &e1, int n, const double *dref, double tol,
                                          double *Ssm) {
  // Gauss-Jordan through a 64-double shared scratch (one warp): every step publishes the current
  // matrix, then each lane reads pivot row / column / pivot.  Ssm doubles as the output (inverse).
  const int lane = threadIdx.x & 31;
  const int i = lane >> 2, j0 = (lane & 3) << 1, j1 = j0 + 1;
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    *reinterpret_cast<double2 *>(Ssm + i * 8 + j0) = make_double2(e0, e1);
    __syncwarp();
    const double2 pk = *reinterpret_cast<const double2 *>(Ssm + k * 8 + j0);   // S[k][j0], S[k][j1]
    const double cik = Ssm[i * 8 + k];                                            // S[i][k]
    const double pkk = Ssm[k * 8 + k];
    __syncwarp();
    ok = ok && (pkk > tol * dref[k]);
    const double d = 1.0 / pkk;
    const bool rowk = (i == k);
    const double f = cik * d;
    double n0 = rowk ? pk.x * d : fma(-f, pk.x, e0);
    double n1 = rowk ? pk.y * d : fma(-f, pk.y, e1);
    if (j0 == k) n0 = rowk ? d : -f;
    if (j1 == k) n1 = rowk ? d : -f;
    e0 = n0; e1 = n1;
  }
  return ok;
}

// Pout = sign * Pin * Sinv  (one 8x8x8 DMMA product per row tile);  w[F[row]] -= Pin[row,:] . coef
__device__ __forceinline__ void panel_small(const Sh &s, const double *Pin, double *Pout, const double *coef,
                                            double sign, int nrows) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  // B fragments of Sinv: B[k][n] = Sinv[kk*4 + k][n]
  const double b0 = sign * s.Sinv[fk * 8 + fr], b1 = sign * s.Sinv[(4 + fk) * 8 + fr];
  for (int ti = wid; ti < (nrows >> 3); ti += NW) {
    double c0 = 0.0, c1 = 0.0;
    dmma(c0, c1, Pin[pan(ti * 8 + fr, fk)], b0);
    dmma(c0, c1, Pin[pan(ti * 8 + fr, 4 + fk)], b1);
    *reinterpret_cast<double2 *>(Pout + pan(ti * 8 + fr, fk * 2)) = make_double2(c0, c1);
  }
  for (int row = threadIdx.x; row < nrows; row += T) {
    const int var = s.F[row];
    if (var >= 0) {
      double z = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) z = fma(Pin[pan(row, j)], coef[j], z);
      s.w[var] -= z;
    }
  }
}

// Remove the r <= 8 slots Rs[0..r).  nt covers every slot in use.
__device__ __noinline__ void block_remove(int cap, const int *Rs, int r, int nt) {
  const Sh s = make_sh(cap);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nrows = nt * 8;
  const int row = tid & 255;
  if (tid == 0) s.stat[ST_TSUB] = clock64();
  if (row < nrows)
    for (int q = tid >> 8; q < 8; q += 2) s.Pb[pan(row, q)] = (q < r) ? h_get(s.H, row, Rs[q]) : 0.0;
  if (wid == 15) {                              // S = H[R,R] straight from the tiles, then invert
    const int i = lane >> 2, j0 = (lane & 3) << 1;
    double e0 = (i < r && j0 < r) ? h_get(s.H, Rs[i], Rs[j0]) : (i == j0 ? 1.0 : 0.0);
    double e1 = (i < r && j0 + 1 < r) ? h_get(s.H, Rs[i], Rs[j0 + 1]) : (i == j0 + 1 ? 1.0 : 0.0);
    if (lane < 8) { s.rho[lane] = (lane < r) ? s.w[s.F[Rs[lane]]] : 0.0; s.cA[lane] = 0.0; }   // w_R
    __syncwarp();
    warp_inv8(e0, e1, r, s.cA, -1.0, s.Sinv);   // H[R,R] is SPD: no pivot test
    s.Sinv[i * 8 + j0] = e0; s.Sinv[i * 8 + j0 + 1] = e1;
    __syncwarp();
    if (lane < 8) {                             // phi = inv(S) w_R
      double a = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) a = fma(s.Sinv[lane * 8 + j], s.rho[j], a);
      s.theta[lane] = a;
    }
  }
  __syncthreads();
  SUBTICK(PH_R_GATHER);
  panel_small(s, s.Pb, s.Pa, s.theta, -1.0, nrows);      // Pa = -B inv(S);  w -= B phi
  __syncthreads();
  SUBTICK(PH_R_PANEL);
  rank_update(s.H, s.Pa, s.Pb, nt, s.tmap);              // H -= B inv(S) B'
  __syncthreads();
  SUBTICK(PH_R_RANK);
  if (row < nrows)
    for (int q = tid >> 8; q < r; q += 2) h_set(s.H, Rs[q], row, 0.0);
  if (tid < r) {
    const int sl = Rs[tid], var = s.F[sl];
    s.w[var] = 0.0; s.pos[var] = -1; s.F[sl] = -1;
  }
  __syncthreads();
  SUBTICK(PH_R_ZERO);
  if (tid == 0) { const long long p = s.stat[ST_P]; s.stat[ST_PIV] += r; s.stat[ST_SUMP2] += r * p * p; s.stat[ST_P] = p - r; }
}

// Add the a <= 8 variables Av[-k] (k = 0..a-1, stored downwards) into the free slots As[k].
// Returns false (state untouched) if the Schur complement is not safely positive definite.
__device__ __noinline__ bool block_add(int cap, const double *G, int ldg, const int *Av, const int *As, int a, int nt) {
  const Sh s = make_sh(cap);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nrows = nt * 8;
  const int row = tid & 255;
  if (tid == 0) s.stat[ST_TSUB] = clock64();
  if (row < nrows) {
    const int var = s.F[row];
    for (int q = tid >> 8; q < 8; q += 2)
      s.Pa[pan(row, q)] = (q < a && var >= 0) ? G[(size_t)ldg * Av[-q] + var] : 0.0;
    if (tid < 256) s.wF[row] = var >= 0 ? s.w[var] : 0.0;
  }
  if (tid >= 256 + 208 && tid < 256 + 208 + 8) {
    const int q = tid - (256 + 208);
    s.cA[q] = (q < a) ? s.cs[Av[-q]] : 0.0;
  }
  if (wid == 15) {                              // warp 15's rows are >= 224 > cap: free for G[A,A]
    for (int e = lane; e < 64; e += 32) {
      const int i = e >> 3, j = e & 7;
      s.Gaa[e] = (i < a && j < a) ? G[(size_t)ldg * Av[-j] + Av[-i]] : (i == j ? 1.0 : 0.0);
    }
  }
  __syncthreads();
  SUBTICK(PH_A_GATHER);
  hmul(s.H, s.Pa, s.Pb, nt);                    // U = H V
  __syncthreads();
  SUBTICK(PH_A_HMUL);
  if (wid < 4) {                                // S partials = V' U over interleaved k-steps
    const int fr = lane >> 2, fk = lane & 3;
    double c0 = 0.0, c1 = 0.0, g0 = 0.0, g1 = 0.0;
    int ks = wid;
    for (; ks + 4 < nt * 2; ks += 8) {
      const int rw = ks * 4 + fk, rx = (ks + 4) * 4 + fk;
      dmma(c0, c1, s.Pa[pan(rw, fr)], s.Pb[pan(rw, fr)]);
      dmma(g0, g1, s.Pa[pan(rx, fr)], s.Pb[pan(rx, fr)]);
    }
    if (ks < nt * 2) { const int rw = ks * 4 + fk; dmma(c0, c1, s.Pa[pan(rw, fr)], s.Pb[pan(rw, fr)]); }
    s.Spart[wid * 64 + fr * 8 + fk * 2] = c0 + g0;
    s.Spart[wid * 64 + fr * 8 + fk * 2 + 1] = c1 + g1;
  } else if (wid < 12) {                        // rho_i = c_i - V[:,i]' w
    const int i = wid - 4;
    double acc = 0.0;
    for (int rw = lane; rw < nrows; rw += 32) acc = fma(s.Pa[pan(rw, i)], s.wF[rw], acc);
    acc = warp_sum(acc);
    if (lane == 0) s.rho[i] = s.cA[i] - acc;
  }
  __syncthreads();
  SUBTICK(PH_A_SPART);
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
    SUBTICK(PH_A_INV1);
    const bool ok = warp_inv8(e0, e1, a, s.theta, 1e-13, s.Sinv);
    const bool all_ok = __all_sync(0xffffffffu, ok);
    SUBTICK(PH_A_INV2);
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
    SUBTICK(PH_A_INV3);
  }
  __syncthreads();
  SUBTICK(PH_A_INV);
  if (!s.ctl[4]) { __syncthreads(); return false; }
  panel_small(s, s.Pb, s.Pa, s.theta, 1.0, nrows);       // T = U inv(S) -> Pa;  w_F -= U theta
  __syncthreads();
  SUBTICK(PH_A_PANEL);
  rank_update(s.H, s.Pa, s.Pb, nt, s.tmap);              // H += T U'
  __syncthreads();
  SUBTICK(PH_A_RANK);
  if (row < nrows && s.F[row] >= 0)                      // new rows / columns: -T
    for (int q = tid >> 8; q < a; q += 2) h_set(s.H, As[q], row, -s.Pa[pan(row, q)]);
  __syncthreads();
  if (tid < 64) {
    const int i = tid >> 3, j = tid & 7;
    if (i < a && j <= i) h_set(s.H, As[i], As[j], s.Sinv[i * 8 + j]);
  }
  if (tid >= 64 && tid < 64 + a) {
    const int q = tid - 64;
    const int var = Av[-q], sl = As[q];
    s.w[var] = s.theta[q]; s.F[sl] = var; s.pos[var] = sl;
  }
  __syncthreads();
  SUBTICK(PH_A_ROWS);
  if (tid == 0) { const long long p = s.stat[ST_P]; s.stat[ST_PIV] += a; s.stat[ST_SUMP2] += a * p * p; s.stat[ST_P] = p + a; }
  return true;
}

// r = c - G[:,F] w_F for all variables; the passive columns of G stream from L2 as double2 row
// pairs, the slot range is split over nsl thread slices.  Returns max |r_F| (normal-equation
// residual), same on all threads.  Uses Pb as scratch.
__device__ __noinline__ double grad_eval(int cap, const double *G, int ldg, int Mp, int hw) {
  const Sh s = make_sh(cap);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int t = tid; t < hw; t += T) { const int var = s.F[t]; s.wF[t] = var >= 0 ? s.w[var] : 0.0; }
  __syncthreads();
  const int npairs = (Mp + 1) >> 1;
  int nsl = T / npairs;
  if (nsl > 8) nsl = 8;
  const int sl = tid / npairs, pr = tid - sl * npairs;
  double *part = s.Pb;                            // [nsl][2 * npairs]
  if (sl < nsl) {
    const int t0 = (hw * sl) / nsl, t1 = (hw * (sl + 1)) / nsl;
    const double *Gp = G + 2 * pr;
    double2 acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_double2(0.0, 0.0);
    for (int t = t0; t < t1; t += 8) {           // predicated batches of 8 loads: no serial tail
      double2 g[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int v = (t + i < t1) ? s.F[t + i] : -1;
        g[i] = v >= 0 ? *reinterpret_cast<const double2 *>(Gp + (size_t)ldg * v) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double ww = (t + i < t1) ? s.wF[t + i] : 0.0;
        acc[i & 3].x = fma(g[i].x, ww, acc[i & 3].x);
        acc[i & 3].y = fma(g[i].y, ww, acc[i & 3].y);
      }
    }
    *reinterpret_cast<double2 *>(part + sl * 2 * npairs + 2 * pr) =
        make_double2((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y));
  }
  __syncthreads();
  double mx = 0.0;
  if (tid < Mp) {
    double sum = 0.0;
    for (int q = 0; q < nsl; ++q) sum += part[q * 2 * npairs + tid];
    const double rv = s.cs[tid] - sum;
    s.r[tid] = rv;
    if (s.pos[tid] >= 0) mx = fabs(rv);
  }
  mx = warp_max(mx);
  if (lane == 0) s.red[wid] = mx;
  __syncthreads();
  double x = s.red[0];
#pragma unroll
  for (int i = 1; i < NW; ++i) x = fmax(x, s.red[i]);
  __syncthreads();
  if (tid == 0) { s.stat[ST_GRAD] += 1; s.stat[ST_SUMP] += s.stat[ST_P]; }
  return x;
}

// w_F += H r_F   (one DMMA product with a single live column)
__device__ __noinline__ void refine(int cap, int nt) {
  const Sh s = make_sh(cap);
  const int nrows = nt * 8;
  const int row = threadIdx.x & 255;
  if (row < nrows) {
    const int var = s.F[row];
    for (int q = threadIdx.x >> 8; q < 8; q += 2) s.Pa[pan(row, q)] = (q == 0 && var >= 0) ? s.r[var] : 0.0;
  }
  __syncthreads();
  hmul(s.H, s.Pa, s.Pb, nt);
  __syncthreads();
  for (int rw = threadIdx.x; rw < nrows; rw += T) {
    const int var = s.F[rw];
    if (var >= 0) s.w[var] += s.Pb[pan(rw, 0)];
  }
  __syncthreads();
}

__device__ __noinline__ void clear_state(int cap) {
  const Sh s = make_sh(cap);
  const int ntc = cap >> 3;
  const int words = ((ntc * (ntc + 1)) >> 1) << 6;
  for (int i = threadIdx.x; i < words; i += T) s.H[i] = 0.0;
  for (int t = threadIdx.x; t < cap; t += T) { s.F[t] = -1; s.smark[t] = 0; }
  if (threadIdx.x == 0) s.stat[ST_P] = 0;
}

__global__ void __launch_bounds__(T, 1) k2v2_orthant_chains(const K2Args A) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int Mp = A.Mp, cap = A.cap;            // cap % 8 == 0, cap >= Mp
  const int ntc = cap >> 3;
  const Sh s = make_sh(cap);
  for (int m = tid; m < cap; m += T) { s.cs[m] = m < Mp ? A.c[m] : 0.0; s.gms[m] = m < Mp ? A.gmask[m] : 0ull; }
  for (int ti = tid; ti < ntc; ti += T)
    for (int tj = 0; tj <= ti; ++tj) s.tmap[((ti * (ti + 1)) >> 1) + tj] = (unsigned short)((ti << 8) | tj);

  int hwm = 0, nt_cur = 0;
  if (tid == 0) {
    for (int i = 0; i < PH_NUM; ++i) s.prof[i] = 0;
    for (int i = 0; i < ST_NUM; ++i) s.stat[i] = 0;
    s.stat[ST_TMARK] = clock64();
  }
#define PH_TICK(which) do { if (tid == 0) { const long long now_ = clock64(); s.prof[which] += now_ - s.stat[ST_TMARK]; s.stat[ST_TMARK] = now_; } } while (0)
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

    clear_state(cap);
    hwm = 0; nt_cur = 0;
    for (int m = tid; m < Mp; m += T) { s.w[m] = 0.0; s.r[m] = s.cs[m]; s.pos[m] = -1; }
    __syncthreads();
    bool r_valid = true;
    PH_TICK(PH_START);

    for (long long i = 0; i < L; ++i) {
      const long long b = base + (i ^ (i >> 1));
      for (int m = tid; m < Mp; m += T) {   // d_m = sum_k Po[m,k] (2 bit_k(b) - 1)   (Opt.jl:28-29)
        const uint64_t gm = s.gms[m];
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
            PH_TICK(PH_OUT);
            const double rf = grad_eval(cap, A.G, A.ldg, Mp, hwm);
            PH_TICK(PH_GRAD);
            if (rf <= 1e-12 * cmax) break;          // carried solution already exact to working accuracy
            refine(cap, nt_cur);
            PH_TICK(PH_REFINE);
            if (rf <= 1e-9 * cmax) break;
            if (++rep >= 4) {                       // inverse degraded: rebuild by re-adding the passive set
              if (tid == 0) {
                int n = 0;
                for (int t = 0; t < hwm; ++t) if (s.F[t] >= 0) { s.lst[cap - 1 - n] = s.F[t]; s.asl[n] = n; ++n; }
                s.ctl[5] = n;
              }
              __syncthreads();
              const int pn = s.ctl[5];
              clear_state(cap);
              for (int m = tid; m < Mp; m += T) { s.w[m] = 0.0; s.pos[m] = -1; }
              __syncthreads();
              const int ntr = (pn + 7) >> 3;
              for (int q0 = 0; q0 < pn; q0 += 8)
                block_add(cap, A.G, A.ldg, s.lst + (cap - 1 - q0), s.asl + q0, min(8, pn - q0), ntr);
              hwm = pn; nt_cur = ntr;
              STAT_ADD(ST_REBUILD, 1);
              if (rep >= 6) { ok = false; break; }
            }
          }
          if (!ok) break;
        }
        r_valid = false;
        // ---- plan: infeasibility flags (passive variables with the wrong sign; active ones whose
        //      sign-adjusted gradient is positive), ordered lists in index order, pivoting rule, slots for
        //      the additions, new high-water mark.  One warp per 32 variables / 32 slots.  The first step
        //      also reduces c_F' w_F, so a converged orthant costs a single barrier here.
        const int nvw = (Mp + 31) >> 5, nsw = cap >> 5 ? (cap + 31) >> 5 : 1;
        int f_my = 0, rank_r = 0, rank_a = 0;
        if (wid < nvw) {
          const int m = (wid << 5) + lane;
          double cw = 0.0;
          if (m < Mp) {
            const int sl = s.pos[m], sg = s.sg[m];
            if (sl >= 0) { cw = s.cs[m] * s.w[m]; if (sg == 0 || (double)sg * s.w[m] < 0.0) f_my = 1; }
            else if (sg != 0 && s.vflag[m] != 3 && (double)sg * s.r[m] > told) f_my = 2;
          }
          const unsigned br = __ballot_sync(0xffffffffu, f_my == 1);
          const unsigned ba = __ballot_sync(0xffffffffu, f_my == 2);
          rank_r = __popc(br & ((1u << lane) - 1)); rank_a = __popc(ba & ((1u << lane) - 1));
          const unsigned any = br | ba;
          cw = warp_sum(cw);
          if (lane == 0) {
            s.pl[wid] = __popc(br) | (__popc(ba) << 16);
            s.pl[8 + wid] = any ? (wid << 5) + 31 - __clz(any) : -1;
            s.red[wid] = cw;
          }
        }
        __syncthreads();
        int nr = 0, na = 0, mxi = -1, pref_r = 0, pref_a = 0;
        for (int q = 0; q < nvw; ++q) {
          const int cnt = s.pl[q];
          if (q < wid) { pref_r += cnt & 0xffff; pref_a += cnt >> 16; }
          nr += cnt & 0xffff; na += cnt >> 16;
          mxi = max(mxi, s.pl[8 + q]);
        }
        const int nv = nr + na;
        if (nv == 0) { r_valid = true; PH_TICK(PH_PLAN); break; }   // KKT point; r stays valid for the next orthant
        const bool single = nv > 0 && !(nv < t_best) && pbar < 1;    // Murty's rule: only the highest index moves
        if (!single) {
          if (f_my == 1) { const int sl = s.pos[(wid << 5) + lane]; s.lst[pref_r + rank_r] = sl; s.smark[sl] = 1; }
          if (f_my == 2) s.lst[cap - 1 - (pref_a + rank_a)] = (wid << 5) + lane;
        } else {
          if (s.pos[mxi] >= 0) { nr = 1; na = 0; if (tid == 0) { const int sl = s.pos[mxi]; s.lst[0] = sl; s.smark[sl] = 1; } }
          else { nr = 0; na = 1; if (tid == 0) s.lst[cap - 1] = mxi; }
        }
        __syncthreads();
        bool fr_my = false, used_my = false;
        int rank_f = 0;
        if (wid < nsw) {
          const int sl = (wid << 5) + lane;
          used_my = sl < cap && s.F[sl] >= 0 && !s.smark[sl];
          fr_my = sl < cap && !used_my;
          const unsigned bal = __ballot_sync(0xffffffffu, fr_my);
          rank_f = __popc(bal & ((1u << lane) - 1));
          if (lane == 0) s.pl[16 + wid] = __popc(bal);
        }
        __syncthreads();
        if (wid < nsw) {
          const int sl = (wid << 5) + lane;
          int pref_f = 0;
          for (int q = 0; q < wid; ++q) pref_f += s.pl[16 + q];
          const bool take = fr_my && (pref_f + rank_f) < na;
          if (take) s.asl[pref_f + rank_f] = sl;
          if (sl < cap && s.smark[sl]) s.smark[sl] = 0;
          const unsigned after = __ballot_sync(0xffffffffu, used_my || take);
          if (lane == 0) s.pl[24 + wid] = after ? (wid << 5) + 32 - __clz(after) : 0;
        }
        __syncthreads();
        int hw_after = 0;
        for (int q = 0; q < nsw; ++q) hw_after = max(hw_after, s.pl[24 + q]);
        if (nv < t_best) { t_best = nv; pbar = 3; }
        else if (pbar >= 1) { --pbar; }
        const int nt_op = (max(hwm, hw_after) + 7) >> 3;
        PH_TICK(PH_PLAN);
        for (int q0 = 0; q0 < nr; q0 += 8) { block_remove(cap, s.lst + q0, min(8, nr - q0), nt_op); if (tid == 0) s.prof[PH_NREM]++; }
        PH_TICK(PH_REMOVE);
        for (int q0 = 0; q0 < na; q0 += 8) {
          if (tid == 0) s.prof[PH_NADD]++;
          const int a = min(8, na - q0);
          if (!block_add(cap, A.G, A.ldg, s.lst + (cap - 1 - q0), s.asl + q0, a, nt_op)) {
            // numerically dependent column in the block: retry one variable at a time
            for (int q = 0; q < a; ++q) {
              if (!block_add(cap, A.G, A.ldg, s.lst + (cap - 1 - q0 - q), s.asl + q0 + q, 1, nt_op)) {
                if (tid == 0) s.vflag[s.lst[cap - 1 - q0 - q]] = 3;
                STAT_ADD(ST_BLOCKED, 1);
                __syncthreads();
              }
            }
          }
        }
        PH_TICK(PH_ADD);
        hwm = hw_after; nt_cur = (hw_after + 7) >> 3;
        STAT_ADD(ST_ITER, 1);
        if (++iters > 60 + 6 * Mp) { ok = false; break; }
      }
      if (!ok) STAT_ADD(ST_NOCONV, 1);

      // ---- objective  sqrt(yy - c_F' w_F)   (= norm(Xa w - ya) at the KKT point, Opt.jl:90); the
      //      partial sums were left in s.red by the plan step that found no violation
      double tot = 0.0;
      for (int q = 0; q < ((Mp + 31) >> 5); ++q) tot += s.red[q];
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
    atomicAdd(&A.counters[CNT_PIVOTS], (unsigned long long)s.stat[ST_PIV]);
    atomicAdd(&A.counters[CNT_GRAD], (unsigned long long)s.stat[ST_GRAD]);
    atomicAdd(&A.counters[CNT_SUMP], (unsigned long long)s.stat[ST_SUMP]);
    atomicAdd(&A.counters[CNT_SUMP2], (unsigned long long)s.stat[ST_SUMP2]);
    atomicAdd(&A.counters[CNT_ITERS], (unsigned long long)s.stat[ST_ITER]);
    atomicAdd(&A.counters[CNT_REBUILDS], (unsigned long long)s.stat[ST_REBUILD]);
    atomicAdd(&A.counters[CNT_BLOCKED], (unsigned long long)s.stat[ST_BLOCKED]);
    atomicAdd(&A.counters[CNT_NOCONV], (unsigned long long)s.stat[ST_NOCONV]);
    PH_TICK(PH_OUT);
    for (int i = 0; i < PH_NUM; ++i) atomicAdd(&A.counters[CNT_NUM + 1 + i], (unsigned long long)s.prof[i]);
  }
}

size_t v2_smem_bytes(int cap) {
  const int ntc = cap >> 3;
  const size_t ntiles = (size_t)(ntc * (ntc + 1) / 2);
  size_t d = (ntiles << 6) + 2 * (size_t)cap * 8 + 5 * (size_t)cap + 64 + 64 + 256 + 24 + NW + ST_NUM + PH_NUM;
  size_t i = 4 * (size_t)cap + 8 + 32 + (ntiles + 1) / 2;
  size_t c = 4 * (size_t)cap;
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
