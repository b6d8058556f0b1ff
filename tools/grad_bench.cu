// Microbenchmark for K2's gradient evaluation r = c - G[:,F] w_F: 148 CTAs x 512 threads, each
// CTA repeatedly streams ~100 random columns of a 201 x 201 FP64 matrix from L2 (smem carve-out
// set like the real kernel so L1 is tiny).  Prints cycles per evaluation for several variants.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int T = 512;
template <int VAR>
__global__ void __launch_bounds__(T, 1) grad_kernel(const double *G, int ldg, int Mp, const int *Fall, int p,
                                                    int reps, double *out, long long *cyc) {
  extern __shared__ __align__(16) unsigned char smem[];
  double *part = reinterpret_cast<double *>(smem);
  double *wF = part + 16 * 208;
  int *F = reinterpret_cast<int *>(wF + 208);
  const int tid = threadIdx.x;
  for (int t = tid; t < p; t += T) { F[t] = Fall[(blockIdx.x % 16) * 256 + t]; wF[t] = 1.0 + t * 1e-3; }
  __syncthreads();
  const int npairs = (Mp + 1) >> 1;
  double total = 0.0;
  const long long t0c = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    if (VAR == 0) {            // row pairs x 5 slices, batches of 8 (as in the kernel)
      int nsl = T / npairs; if (nsl > 8) nsl = 8;
      const int sl = tid / npairs, pr = tid - sl * npairs;
      if (sl < nsl) {
        const int s0 = (p * sl) / nsl, s1 = (p * (sl + 1)) / nsl;
        const double *Gp = G + 2 * pr;
        double2 acc[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        for (int t = s0; t < s1; t += 8) {
          double2 g[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int v = (t + i < s1) ? F[t + i] : -1;
            g[i] = v >= 0 ? *reinterpret_cast<const double2 *>(Gp + (size_t)ldg * v) : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const double ww = (t + i < s1) ? wF[t + i] : 0.0;
            acc[i & 3].x = fma(g[i].x, ww, acc[i & 3].x); acc[i & 3].y = fma(g[i].y, ww, acc[i & 3].y);
          }
        }
        *reinterpret_cast<double2 *>(part + sl * 2 * npairs + 2 * pr) =
            make_double2(acc[0].x + acc[1].x + acc[2].x + acc[3].x, acc[0].y + acc[1].y + acc[2].y + acc[3].y);
      }
      __syncthreads();
      if (tid < Mp) { double s = 0; for (int q = 0; q < nsl; ++q) s += part[q * 2 * npairs + tid]; total += s; }
      __syncthreads();
    } else if (VAR == 1) {     // warp per column: lanes over rows (8 B), columns strided by warp; all loads first
      const int lane = tid & 31, wid = tid >> 5;
      double acc[7] = {0, 0, 0, 0, 0, 0, 0};
      for (int t = wid; t < p; t += 32) {
        const int v0 = F[t], v1 = (t + 16 < p) ? F[t + 16] : -1;
        double g0[7], g1[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          const int m = lane + 32 * i;
          g0[i] = (m < Mp) ? G[(size_t)ldg * v0 + m] : 0.0;
          g1[i] = (m < Mp && v1 >= 0) ? G[(size_t)ldg * v1 + m] : 0.0;
        }
        const double w0 = wF[t], w1 = (t + 16 < p) ? wF[t + 16] : 0.0;
#pragma unroll
        for (int i = 0; i < 7; ++i) acc[i] = fma(g0[i], w0, fma(g1[i], w1, acc[i]));
      }
#pragma unroll
      for (int i = 0; i < 7; ++i) { const int m = lane + 32 * i; if (m < 208) part[wid * 208 + m] = acc[i]; }
      __syncthreads();
      if (tid < Mp) { double s = 0; for (int q = 0; q < 16; ++q) s += part[q * 208 + tid]; total += s; }
      __syncthreads();
    } else if (VAR == 2) {     // like 0 but 2 batches of 10-12: everything in flight at once
      int nsl = T / npairs; if (nsl > 8) nsl = 8;
      const int sl = tid / npairs, pr = tid - sl * npairs;
      if (sl < nsl) {
        const int s0 = (p * sl) / nsl, s1 = (p * (sl + 1)) / nsl;
        const double *Gp = G + 2 * pr;
        double2 acc[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        for (int t = s0; t < s1; t += 12) {
          double2 g[12];
#pragma unroll
          for (int i = 0; i < 12; ++i) {
            const int v = (t + i < s1) ? F[t + i] : -1;
            g[i] = v >= 0 ? *reinterpret_cast<const double2 *>(Gp + (size_t)ldg * v) : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int i = 0; i < 12; ++i) {
            const double ww = (t + i < s1) ? wF[t + i] : 0.0;
            acc[i & 3].x = fma(g[i].x, ww, acc[i & 3].x); acc[i & 3].y = fma(g[i].y, ww, acc[i & 3].y);
          }
        }
        *reinterpret_cast<double2 *>(part + sl * 2 * npairs + 2 * pr) =
            make_double2(acc[0].x + acc[1].x + acc[2].x + acc[3].x, acc[0].y + acc[1].y + acc[2].y + acc[3].y);
      }
      __syncthreads();
      if (tid < Mp) { double s = 0; for (int q = 0; q < nsl; ++q) s += part[q * 2 * npairs + tid]; total += s; }
      __syncthreads();
    }
  }
  const long long t1c = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1c - t0c;
  out[blockIdx.x * T + tid] = total;
}

template <int VAR> void run(const char *name, const double *G, int ldg, int Mp, const int *F, int p, size_t smem, int grid) {
  double *out; long long *cyc; cudaMalloc(&out, sizeof(double) * grid * T); cudaMalloc(&cyc, sizeof(long long) * grid);
  cudaFuncSetAttribute(grad_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int reps = 200;
  grad_kernel<VAR><<<grid, T, smem>>>(G, ldg, Mp, F, p, reps, out, cyc);
  grad_kernel<VAR><<<grid, T, smem>>>(G, ldg, Mp, F, p, reps, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(grid); cudaMemcpy(h.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0, sum = 0; for (auto v : h) { mx = v > mx ? v : mx; sum += v; }
  printf("%-28s grid %3d smem %6zu: %8.0f cycles/eval (mean) %8.0f (max)  [%s]\n", name, grid, smem, (double)sum / grid / reps,
         (double)mx / reps, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  const int Mp = 201, ldg = 204, p = 100;
  std::vector<double> G((size_t)208 * Mp, 0.5);
  std::vector<int> F(16 * 256);
  srand(1);
  for (int c = 0; c < 16; ++c) {
    std::vector<int> perm(Mp); for (int i = 0; i < Mp; ++i) perm[i] = i;
    for (int i = Mp - 1; i > 0; --i) { int j = rand() % (i + 1); std::swap(perm[i], perm[j]); }
    for (int t = 0; t < 256; ++t) F[c * 256 + t] = perm[t % Mp];
  }
  double *dG; int *dF; cudaMalloc(&dG, G.size() * 8); cudaMalloc(&dF, F.size() * 4);
  cudaMemcpy(dG, G.data(), G.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dF, F.data(), F.size() * 4, cudaMemcpyHostToDevice);
  for (int ld : {204, 208}) {
    printf("ldg = %d\n", ld);
    for (size_t smem : {(size_t)220000, (size_t)100000, (size_t)32768}) {
      for (int grid : {148, 74, 1}) {
        run<0>("pairs x slices, batch 8", dG, ld, Mp, dF, p, smem, grid);
        run<1>("warp per column", dG, ld, Mp, dF, p, smem, grid);
      }
    }
  }
  return 0;
}
