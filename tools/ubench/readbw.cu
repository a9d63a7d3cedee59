// Micro-benchmark: read bandwidth of the Lloyd pass's access pattern (persistent grid, tiles of
// THREADS x 4 float4, one-tile register prefetch) with no arithmetic worth mentioning.
#include <cstdio>
#include <cuda_runtime.h>
template <int THREADS, int P>
__global__ void __launch_bounds__(THREADS) k(const float4* __restrict__ w, unsigned long long n, float* out) {
  constexpr unsigned long long TILE = (unsigned long long)THREADS * P;
  const unsigned long long tiles = n / TILE;
  float acc = 0.f;
  float4 cur[P], nxt[P];
  unsigned long long t = blockIdx.x;
  if (t < tiles) for (int i = 0; i < P; ++i) cur[i] = __ldcs(w + t * TILE + i * THREADS + threadIdx.x);
  for (; t < tiles; t += gridDim.x) {
    unsigned long long nx = t + gridDim.x;
    if (nx < tiles) for (int i = 0; i < P; ++i) nxt[i] = __ldcs(w + nx * TILE + i * THREADS + threadIdx.x);
    for (int i = 0; i < P; ++i) acc += cur[i].x + cur[i].y + cur[i].z + cur[i].w;
    for (int i = 0; i < P; ++i) cur[i] = nxt[i];
  }
  if (acc == 123.456f) out[0] = acc;
}
int main() {
  const unsigned long long n = 8192ull * 8192ull;
  float4* w; float* out;
  cudaMalloc(&w, n * 16); cudaMalloc(&out, 4); cudaMemset(w, 0, n * 16);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int bps : {1, 2, 3, 4, 8}) {
    int grid = 148 * bps;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(a);
      for (int i = 0; i < 20; ++i) k<256, 4><<<grid, 256>>>(w, n, out);
      cudaEventRecord(b); cudaEventSynchronize(b);
    }
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 20;
    printf("256 thr x 4 px, %d blocks/SM: %.4f ms  %.0f GB/s\n", bps, ms, n * 16 / ms / 1e6);
  }
  return 0;
}
