// K4: data-space objective of one orthant's solution,  sum_n (Xo[n,:] . w - y[n])^2  over the
// loaded rows -- the reference's  norm(Xo * (Po .* alpha) * beta - yo)  (src/PartitionedLSOpt.jl:90)
// without the eta rows (added on the host side of the ABI, they are K' scalars).
// One streaming pass over the passive columns of Z: HBM-bound, 8*N*|F| bytes.
#include "common.cuh"

namespace pls {
namespace {

constexpr int T4 = 256;
constexpr int MAXNZ = 2048;

// w_sparse: list of (column, weight) for the nonzero entries of w = d .* alpha; ycol = M+1.
__global__ void __launch_bounds__(T4) k4_residual_rows(const double *__restrict__ Z, long long ldz,
                                                       long long n_rows, int ycol,
                                                       const double *__restrict__ w, int Mp,
                                                       double *__restrict__ block_part) {
  __shared__ int s_col[MAXNZ];
  __shared__ double s_w[MAXNZ];
  __shared__ int s_n;
  __shared__ double red[T4 / 32];
  if (threadIdx.x == 0) {
    int n = 0;
    for (int m = 0; m < Mp && n < MAXNZ; ++m) {
      const double x = w[m];
      if (x != 0.0) { s_col[n] = m; s_w[n] = x; ++n; }
    }
    s_n = n;
  }
  __syncthreads();
  const int nz = s_n;
  double ssq = 0.0;
  const long long n_pairs = (n_rows + 1) / 2;     // rows [n_rows, ldz) are zero in every column
  for (long long pr = (long long)blockIdx.x * T4 + threadIdx.x; pr < n_pairs;
       pr += (long long)gridDim.x * T4) {
    const long long row = pr * 2;
    const double2 yv = *reinterpret_cast<const double2 *>(Z + (long long)ycol * ldz + row);
    double a0 = -yv.x, a1 = -yv.y;
    int t = 0;
    for (; t + 3 < nz; t += 4) {
      const double2 z0 = *reinterpret_cast<const double2 *>(Z + (long long)s_col[t] * ldz + row);
      const double2 z1 = *reinterpret_cast<const double2 *>(Z + (long long)s_col[t + 1] * ldz + row);
      const double2 z2 = *reinterpret_cast<const double2 *>(Z + (long long)s_col[t + 2] * ldz + row);
      const double2 z3 = *reinterpret_cast<const double2 *>(Z + (long long)s_col[t + 3] * ldz + row);
      a0 = fma(z0.x, s_w[t], a0); a1 = fma(z0.y, s_w[t], a1);
      a0 = fma(z1.x, s_w[t + 1], a0); a1 = fma(z1.y, s_w[t + 1], a1);
      a0 = fma(z2.x, s_w[t + 2], a0); a1 = fma(z2.y, s_w[t + 2], a1);
      a0 = fma(z3.x, s_w[t + 3], a0); a1 = fma(z3.y, s_w[t + 3], a1);
    }
    for (; t < nz; ++t) {
      const double2 z0 = *reinterpret_cast<const double2 *>(Z + (long long)s_col[t] * ldz + row);
      a0 = fma(z0.x, s_w[t], a0); a1 = fma(z0.y, s_w[t], a1);
    }
    ssq = fma(a0, a0, ssq);
    ssq = fma(a1, a1, ssq);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ssq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < T4 / 32; ++i) s += red[i];
    block_part[blockIdx.x] = s;
  }
}

__global__ void k4_sum_parts(const double *__restrict__ part, int n, double *__restrict__ out) {
  __shared__ double sm[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += part[i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int st = 128; st; st >>= 1) {
    if (threadIdx.x < st) sm[threadIdx.x] += sm[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sm[0];
}

}  // namespace

// d_w: device vector of signed weights w (length Mp).  The pad row (if N is odd) holds zeros in
// every column including the ones column and y, so it adds nothing.
int k4_residual(const Problem &pb, SolveWs &ws, const double *d_w, double *d_ssq, int sm_count,
                cudaStream_t st, int *launches) {
  if (pb.Mp > MAXNZ) { set_error("k4: M' = %d exceeds %d", pb.Mp, MAXNZ); return PLS_EUNSUPPORTED; }
  long long n_pairs = (pb.N + 1) / 2;
  long long blocks = (n_pairs + T4 - 1) / T4;
  const long long cap = (long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if ((int)blocks > ws.resid_blocks) {
    if (ws.resid_part) cudaFree(ws.resid_part);
    ws.resid_part = nullptr; ws.resid_blocks = 0;
    PLS_CUDA_TRY(cudaMalloc(&ws.resid_part, sizeof(double) * (size_t)cap));
    ws.resid_blocks = (int)cap;
  }
  k4_residual_rows<<<(unsigned)blocks, T4, 0, st>>>(pb.Z, pb.ldz, pb.N, pb.M + 1, d_w, pb.Mp,
                                                    ws.resid_part);
  PLS_CUDA_TRY(cudaGetLastError());
  ++*launches;
  k4_sum_parts<<<1, 256, 0, st>>>(ws.resid_part, (int)blocks, d_ssq);
  PLS_CUDA_TRY(cudaGetLastError());
  ++*launches;
  return PLS_OK;
}

}  // namespace pls
