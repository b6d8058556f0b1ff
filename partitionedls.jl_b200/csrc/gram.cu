// K1: Gram build  S = Z'Z  for the augmented data Z = [X | 1 | y]  on FP64 tensor cores (DMMA),
// and the finalize step  G = S[0:M', 0:M'] + eta * Po*Po',  c = S[M'+0.., y],  yy.
//
// Replaces, once per fit instead of once per orthant:
//   homogeneousCoords  (src/PartitionedLS.jl:76-81)   -> the ones column of Z
//   regularizeProblem  (src/PartitionedLS.jl:108-123) -> + eta * Po*Po' (one sqrt(eta) row per group)
//   and every product with the data matrix that nonneg_lsq / norm(...) recompute per orthant
//   (src/PartitionedLSOpt.jl:88-90).
//
// Layout: Z is column-major with ldz % 16 == 0, rows [N, ldz) and columns [M+2, zcols_pad) are zero,
// so tiles never need bounds checks.  Only lower-triangle 64x64 tiles are computed; the row range is
// split into chunks (split-K) whose partial tiles are summed in a fixed order by the reduce kernel.
#include "common.cuh"

namespace pls {
namespace {

constexpr int BT = 64;      // output tile edge
constexpr int KB = 32;      // rows of Z per stage
constexpr int PITCH = 36;   // smem pitch (doubles) per column: 36 % 16 == 4 -> conflict-free frags
constexpr int T1 = 128;     // 4 warps, 2 x 2, each a 32 x 32 sub-tile

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// grid: (n_tiles_lower, n_chunks).  part[(chunk * n_tiles + tile) * BT*BT + col*BT + row]
__global__ void __launch_bounds__(T1) k1_gram_tiles(const double *__restrict__ Z, long long ldz,
                                                    long long rows_per_chunk, long long n_rows_pad,
                                                    int n_tiles, double *__restrict__ part) {
  __shared__ __align__(16) double As[BT * PITCH];
  __shared__ __align__(16) double Bs[BT * PITCH];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // tile index -> (ti >= tj)
  int tile = blockIdx.x, ti = 0;
  while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
  const int tj = tile - ti * (ti + 1) / 2;
  const bool diag = (ti == tj);
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > n_rows_pad) r1 = n_rows_pad;

  const double *Za = Z + (long long)ti * BT * ldz;
  const double *Zb = Z + (long long)tj * BT * ldz;
  // loader mapping: 16 lanes x double2 cover the 32 rows of one column; a warp covers 2 columns
  const int lrow = (lane & 15) * 2, lcol = wid * 2 + (lane >> 4);   // + 8 * q, q = 0..7
  double2 pa[8], pb[8];
  auto gload = [&](long long r) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int col = lcol + 8 * q;
      pa[q] = *reinterpret_cast<const double2 *>(Za + (long long)col * ldz + r + lrow);
      if (!diag) pb[q] = *reinterpret_cast<const double2 *>(Zb + (long long)col * ldz + r + lrow);
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int col = lcol + 8 * q;
      *reinterpret_cast<double2 *>(&As[col * PITCH + lrow]) = pa[q];
      if (!diag) *reinterpret_cast<double2 *>(&Bs[col * PITCH + lrow]) = pb[q];
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int wi = (wid >> 1) * 32, wj = (wid & 1) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  const double *Bsrc = diag ? As : Bs;

  if (r0 < r1) gload(r0);
  for (long long r = r0; r < r1; r += KB) {
    __syncthreads();
    sstore();
    __syncthreads();
    if (r + KB < r1) gload(r + KB);
#pragma unroll
    for (int kk = 0; kk < KB / 4; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[(wi + i * 8 + fr) * PITCH + kk * 4 + fk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bsrc[(wj + j * 8 + fr) * PITCH + kk * 4 + fk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  double *out = part + ((size_t)blockIdx.y * n_tiles + tile) * (BT * BT);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = wi + i * 8 + fr;            // index within tile ti (output row)
      const int col = wj + j * 8 + fk * 2;        // index within tile tj (output col)
      out[(size_t)col * BT + row] = acc[i][j][0];
      out[(size_t)(col + 1) * BT + row] = acc[i][j][1];
    }
}

// S[i + j*zc] (i >= j) = sum over chunks of the partial tiles, fixed order.
__global__ void k1_reduce(const double *__restrict__ part, int n_tiles, int n_chunks, int zc,
                          double *__restrict__ S) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)zc * zc) return;
  const int i = (int)(idx % zc), j = (int)(idx / zc);
  if (i < j) { S[idx] = 0.0; return; }
  const int ti = i / BT, tj = j / BT;
  const int tile = ti * (ti + 1) / 2 + tj;
  const size_t off = (size_t)(j % BT) * BT + (i % BT);
  double s = 0.0;
  for (int ch = 0; ch < n_chunks; ++ch) s += part[((size_t)ch * n_tiles + tile) * (BT * BT) + off];
  S[idx] = s;
}

// G = S (mirrored) + eta * |groups(i) & groups(j)|;  c = S[y row];  scal = {yy, max|c|, max diag}
__global__ void k1_finalize(const double *__restrict__ S, int zc, int Mp, double eta,
                            const uint64_t *__restrict__ gmask, double *__restrict__ G, int ldg,
                            double *__restrict__ c) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)ldg * Mp) return;
  const int i = (int)(idx % ldg), j = (int)(idx / ldg);
  if (i >= Mp) { G[idx] = 0.0; return; }
  const int hi = i > j ? i : j, lo = i > j ? j : i;
  double g = S[(size_t)lo * zc + hi];
  if (eta != 0.0) g += eta * (double)__popcll(gmask[i] & gmask[j]);
  G[idx] = g;
  if (j == 0) c[i] = S[(size_t)i * zc + (Mp)];   // row index of y in Z'Z is Mp = M+1
}

__global__ void k1_scalars(const double *__restrict__ S, int zc, int Mp, const double *__restrict__ c,
                           const double *__restrict__ G, int ldg, double *__restrict__ scal) {
  __shared__ double sm[256], sd[256];
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  double mx = 0.0, md = 0.0;
  int nonfinite = 0;
  for (int i = threadIdx.x; i < Mp; i += 256) {
    mx = fmax(mx, fabs(c[i]));
    md = fmax(md, fabs(G[(size_t)i * ldg + i]));
    if (!isfinite(c[i])) nonfinite = 1;
  }
  for (long long idx = threadIdx.x; idx < (long long)Mp * Mp; idx += 256)
    if (!isfinite(G[(size_t)(idx / Mp) * ldg + idx % Mp])) nonfinite = 1;
  if (nonfinite) bad = 1;
  sm[threadIdx.x] = mx; sd[threadIdx.x] = md;
  __syncthreads();
  for (int st = 128; st; st >>= 1) {
    if (threadIdx.x < st) {
      sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + st]);
      sd[threadIdx.x] = fmax(sd[threadIdx.x], sd[threadIdx.x + st]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    scal[0] = S[(size_t)Mp * zc + Mp];
    scal[1] = sm[0];
    scal[2] = sd[0];
    scal[3] = (bad || !isfinite(scal[0])) ? 1.0 : 0.0;   // non-finite input flag
  }
}

}  // namespace

int k1_gram_build(Problem &pb, cudaStream_t st, int *launches) {
  const int nt = pb.zcols_pad / BT;
  const int n_tiles = nt * (nt + 1) / 2;
  const long long n_rows_pad = round_up(pb.N, KB);      // <= ldz (ldz % 32 == 0)
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // ~4 waves of CTAs (several CTAs fit per SM), chunks of whole stages
  long long want_chunks = (4ll * 4 * sms + n_tiles - 1) / n_tiles;
  long long max_chunks = n_rows_pad / (KB * 8);
  if (max_chunks < 1) max_chunks = 1;
  if (want_chunks > max_chunks) want_chunks = max_chunks;
  if (want_chunks < 1) want_chunks = 1;
  long long rows_per_chunk = round_up((n_rows_pad + want_chunks - 1) / want_chunks, KB);
  const int n_chunks = (int)((n_rows_pad + rows_per_chunk - 1) / rows_per_chunk);
  const size_t need = (size_t)n_chunks * n_tiles * BT * BT * sizeof(double);
  if (need > pb.part_bytes) {
    if (pb.part) cudaFree(pb.part);
    pb.part = nullptr; pb.part_bytes = 0;
    PLS_CUDA_TRY(cudaMalloc(&pb.part, need));
    pb.part_bytes = need;
  }
  dim3 grid(n_tiles, n_chunks);
  k1_gram_tiles<<<grid, T1, 0, st>>>(pb.Z, pb.ldz, rows_per_chunk, n_rows_pad, n_tiles, pb.part);
  PLS_CUDA_TRY(cudaGetLastError());
  ++*launches;
  const int zc = pb.zcols;
  const long long tot = (long long)zc * zc;
  k1_reduce<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(pb.part, n_tiles, n_chunks, zc, pb.S);
  PLS_CUDA_TRY(cudaGetLastError());
  ++*launches;
  return PLS_OK;
}

int k1_gram_finalize(Problem &pb, cudaStream_t st, int *launches) {
  const long long tot = (long long)pb.ldg * pb.Mp;
  k1_finalize<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(pb.S, pb.zcols, pb.Mp, pb.eta, pb.gmask,
                                                            pb.G, pb.ldg, pb.c);
  PLS_CUDA_TRY(cudaGetLastError());
  ++*launches;
  k1_scalars<<<1, 256, 0, st>>>(pb.S, pb.zcols, pb.Mp, pb.c, pb.G, pb.ldg, pb.scal);
  PLS_CUDA_TRY(cudaGetLastError());
  ++*launches;
  pb.gram_ready = true;
  return PLS_OK;
}

}  // namespace pls
