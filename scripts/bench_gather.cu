// Microbenchmark: how fast can a B200 gather random 128-byte rows?  (dev tool)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bench_gather scripts/bench_gather.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
__global__ void k_fill_idx(uint32_t* idx, uint64_t n, uint32_t rows, uint32_t hot_rows, int hot_pct) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t h = mix64(i);
  bool hot = (int)(h % 100) < hot_pct;
  uint64_t h2 = mix64(h);
  idx[i] = hot ? (uint32_t)(mix64((h2 % hot_rows) * 7919) % rows) : (uint32_t)(h2 % rows);
}
__global__ void k_fill_y(double* y, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = 1.0;
}

// 8 lanes x 16 B per row, U rows in flight per lane-group, warp handles 32 indices per outer step
template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB) k_gather16(const double* __restrict__ y, const uint32_t* __restrict__ idx,
                                                       uint64_t n, double* out) {
  const int lane = threadIdx.x & 31, l8 = lane & 7, g = lane >> 3;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  double a0 = 0, a1 = 0;
  for (uint64_t base = warp * (4 * U); base < n; base += nw * (4 * U)) {
    uint32_t my = (base + lane < n && lane < 4 * U) ? idx[base + lane] : 0;
    double2 r[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      uint32_t u = __shfl_sync(0xFFFFFFFFu, my, j * 4 + g);
      r[j] = __ldg(reinterpret_cast<const double2*>(y + (uint64_t)u * 16) + l8);
    }
#pragma unroll
    for (int j = 0; j < U; ++j) { a0 += r[j].x; a1 += r[j].y; }
  }
  if (a0 + a1 == 12345.678) out[0] = a0;
}
// 4 lanes x 32 B per row
struct __align__(32) d4 { double x, y, z, w; };
template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB) k_gather32(const double* __restrict__ y, const uint32_t* __restrict__ idx,
                                                       uint64_t n, double* out) {
  const int lane = threadIdx.x & 31, l4 = lane & 3, g = lane >> 2;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  double a0 = 0, a1 = 0;
  for (uint64_t base = warp * (8 * U); base < n; base += nw * (8 * U)) {
    uint32_t my[(8 * U + 31) / 32];
#pragma unroll
    for (int k = 0; k < (8 * U + 31) / 32; ++k) my[k] = (base + k * 32 + lane < n) ? idx[base + k * 32 + lane] : 0;
    d4 r[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int e = j * 8 + g;
      uint32_t u = __shfl_sync(0xFFFFFFFFu, my[e / 32], e % 32);
      const double* p = y + (uint64_t)u * 16 + l4 * 4;
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r[j].x), "=d"(r[j].y), "=d"(r[j].z), "=d"(r[j].w) : "l"(p));
    }
#pragma unroll
    for (int j = 0; j < U; ++j) { a0 += r[j].x + r[j].z; a1 += r[j].y + r[j].w; }
  }
  if (a0 + a1 == 12345.678) out[0] = a0;
}

template <class K>
void run(const char* name, K kern, int ctas_per_sm, const double* y, const uint32_t* idx, uint64_t n, double* out) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int grid = 148 * ctas_per_sm;
  kern<<<grid, 256>>>(y, idx, n, out);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < 3; ++i) kern<<<grid, 256>>>(y, idx, n, out);
  cudaEventRecord(b); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
  printf("%-28s ctas/sm %d  %.3f ms  %.0f GB/s rows  %.1f Grows/s\n", name, ctas_per_sm, ms, n * 128.0 / ms / 1e6, n / ms / 1e6);
}

int main(int argc, char** argv) {
  const uint32_t rows = 10000000; const uint64_t n = 150000000;
  double* y; uint32_t* idx; double* out;
  CK(cudaMalloc(&y, (size_t)rows * 128)); CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&out, 8));
  k_fill_y<<<(rows * 16 + 255) / 256, 256>>>(y, (uint64_t)rows * 16);
  for (uint32_t hot_rows : {131072u, 262144u, 524288u, 1048576u}) {
    for (int hot_pct : {67, 100}) {
      k_fill_idx<<<(unsigned)((n + 255) / 256), 256>>>(idx, n, rows, hot_rows, hot_pct);
      CK(cudaDeviceSynchronize());
      printf("== %d%% of gathers go to a %.0f MB hot set, rest uniform over 1.28 GB\n", hot_pct, hot_rows * 128.0 / 1048576.0);
      run("16B x U=8  minb4", k_gather16<8, 4>, 4, y, idx, n, out);
      run("32B x U=4  minb4", k_gather32<4, 4>, 4, y, idx, n, out);
    }
  }
  return 0;
}
