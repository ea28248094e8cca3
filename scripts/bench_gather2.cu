// Microbenchmark (dev tool, round 2): gather rate of ROWB-byte rows (32 / 64 / 128 B) out of a
// [rows][ROWB] array, for uniformly random sources and for sources that ascend inside 512-edge
// tasks (the order k_sweep_long sees), at two state sizes.  Decides whether a topic split
// (narrower rows per GPU) costs per-GPU byte throughput.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bench_gather2 scripts/bench_gather2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
// mode 0: uniform random; mode 1: ascending inside every 512-edge task (stride = rows / 512, jittered)
__global__ void k_fill_idx(uint32_t* idx, uint64_t n, uint32_t rows, int mode) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mode == 0) { idx[i] = (uint32_t)(mix64(i) % rows); return; }
  const uint64_t task = i / 512, j = i % 512;
  const uint64_t stride = rows / 512;
  idx[i] = (uint32_t)(j * stride + mix64(task * 1315423911ull + j) % stride);
}
__global__ void k_fill_y(double* y, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = 1.0;
}

// LPR lanes x 16 B per row; a warp handles 32 indices per step, U = LPR loads in flight per lane
template <int LPR>
__global__ void __launch_bounds__(256, 4) k_gather(const double* __restrict__ y, const uint32_t* __restrict__ idx,
                                                  uint64_t n, double* out) {
  constexpr int GPW = 32 / LPR, TP = LPR * 2;
  const int lane = threadIdx.x & 31, l = lane % LPR, g = lane / LPR;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  double a0 = 0, a1 = 0;
  uint32_t nxt = (warp * 32 + lane < n) ? idx[warp * 32 + lane] : 0;
  for (uint64_t base = warp * 32; base < n; base += nw * 32) {
    const uint32_t my = nxt;
    const uint64_t nb = base + nw * 32;
    nxt = (nb + lane < n) ? idx[nb + lane] : 0;
    double2 r[LPR];
#pragma unroll
    for (int j = 0; j < LPR; ++j) {
      uint32_t u = __shfl_sync(0xFFFFFFFFu, my, j * GPW + g);
      r[j] = __ldg(reinterpret_cast<const double2*>(y + (uint64_t)u * TP) + l);
    }
#pragma unroll
    for (int j = 0; j < LPR; ++j) { a0 += r[j].x; a1 += r[j].y; }
  }
  if (a0 + a1 == 12345.678) out[0] = a0;
}

template <class K>
void run(const char* name, K kern, int rowb, const double* y, const uint32_t* idx, uint64_t n, double* out) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int grid = 148 * 4;
  kern<<<grid, 256>>>(y, idx, n, out);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < 3; ++i) kern<<<grid, 256>>>(y, idx, n, out);
  cudaEventRecord(b); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
  printf("  %-10s row %3d B  %.3f ms  %6.0f GB/s  %5.1f G rows/s\n", name, rowb, ms, n * (double)rowb / ms / 1e6, n / ms / 1e6);
}

int main() {
  const uint64_t n = 150000000;
  double* y; uint32_t* idx; double* out;
  CK(cudaMalloc(&y, (size_t)80000000 * 128)); CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&out, 8));
  k_fill_y<<<(unsigned)(((uint64_t)80000000 * 16 + 255) / 256), 256>>>(y, (uint64_t)80000000 * 16);
  for (uint32_t rows : {10000000u, 80000000u}) {
    for (int mode = 0; mode < 2; ++mode) {
      k_fill_idx<<<(unsigned)((n + 255) / 256), 256>>>(idx, n, rows, mode);
      CK(cudaDeviceSynchronize());
      printf("== %u rows, %s sources\n", rows, mode ? "ascending-in-task" : "uniform random");
      run("lpr8", k_gather<8>, 128, y, idx, n, out);
      run("lpr4", k_gather<4>, 64, y, idx, n, out);
      run("lpr2", k_gather<2>, 32, y, idx, n, out);
    }
  }
  return 0;
}
