// Measures the FP64 peaks MEASURED_PEAKS.json does not hold: DFMA (vector pipe), DMMA (mma.sync
// m8n8k4 and m16n8k16 f64), and L2 read bandwidth on an 8 MB resident buffer.  Build:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double *out, int iters) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double x = 1.0000001, y = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma884_kernel(double *out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma16816_kernel(double *out, int iters) {
  double c[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-4 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  double s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void l2_read_kernel(const double2 *buf, size_t n, int reps, double *out) {
  double s = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      double2 v = buf[i]; s += v.x + v.y;
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int l2 = 0; cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"smem_optin\": %zu", p.name, p.multiProcessorCount, l2, p.sharedMemPerBlockOptin);
  const int blocks = p.multiProcessorCount * 8, threads = 256;
  double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  const int iters = 1 << 14;
  float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters); });
  printf(", \"dfma_tflops\": %.2f", 2.0 * 8 * iters * (double)blocks * threads / ms / 1e9);
  ms = time_ms([&] { dmma884_kernel<<<blocks, threads>>>(out, iters); });
  printf(", \"dmma_m8n8k4_tflops\": %.2f", 512.0 * 8 * iters * (double)blocks * (threads / 32) / ms / 1e9);
  ms = time_ms([&] { dmma16816_kernel<<<blocks, threads>>>(out, iters); });
  printf(", \"dmma_m16n8k16_tflops\": %.2f", 2.0 * 16 * 8 * 16 * 4 * iters * (double)blocks * (threads / 32) / ms / 1e9);
  const size_t nbytes = 8u << 20; double2 *buf; cudaMalloc(&buf, nbytes); cudaMemset(buf, 0, nbytes);
  ms = time_ms([&] { l2_read_kernel<<<blocks, threads>>>(buf, nbytes / 16, 64, out); });
  printf(", \"l2_read_gbs_8MB\": %.0f", 64.0 * nbytes / ms / 1e6);
  const size_t nb2 = 323200; 
  ms = time_ms([&] { l2_read_kernel<<<blocks, threads>>>(buf, nb2 / 16, 2048, out); });
  printf(", \"l2_read_gbs_323KB\": %.0f}\n", 2048.0 * nb2 / ms / 1e6);
  return 0;
}
