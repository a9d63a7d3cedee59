// Micro-benchmark: issue cost of packed FFMA2 / FFMA with register, uniform-register and immediate
// scalar operands (register-file bank model of B200).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi){f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c){f32x2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;}
__constant__ float ctab[64];
constexpr int CH = 8;      // independent chains
constexpr int IT = 2048;   // loop trips
// MODE 0: FFMA2, scalar b in a (per-thread) vector register
// MODE 1: FFMA2, scalar b from the constant bank (uniform register)
// MODE 2: FFMA 3-register
// MODE 3: FFMA with constant-bank b
// MODE 4: FFMA2 with the same `a` for all chains (operand reuse possible), register b
// MODE 5: FFMA2, register b different from one instruction to the next (no operand reuse)
// MODE 6: FFMA 3-register, b different from one instruction to the next
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float* in, long long* cyc) {
  float b[5];
  for (int i = 0; i < 5; ++i) b[i] = MODE == 1 || MODE == 3 ? ctab[i] : in[i + (threadIdx.x & 1)];
  f32x2 a[CH], acc[CH];
  float fa[CH], facc[CH];
  for (int i = 0; i < CH; ++i) {
    a[i] = pack2(in[8 + i] + threadIdx.x, in[16 + i]);
    acc[i] = pack2(in[24 + i], in[32 + i]);
    fa[i] = in[8 + i] + threadIdx.x;
    facc[i] = in[24 + i];
  }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < IT; ++it) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (MODE == 0 || MODE == 1) acc[i] = fma2(a[i], pack2(b[j], b[j]), acc[i]);
        if (MODE == 4) acc[i] = fma2(a[0], pack2(b[j], b[j]), acc[i]);
        if (MODE == 5) acc[i] = fma2(a[i], pack2(b[(i + j) % 5], b[(i + j) % 5]), acc[i]);
        if (MODE == 6) facc[i] = fmaf(fa[i], b[(i + j) % 5], facc[i]);
        if (MODE == 2 || MODE == 3) facc[i] = fmaf(fa[i], b[j], facc[i]);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < CH; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
    s += lo + hi + facc[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, float* out, float* in, long long* cyc, int warps_per_smsp) {
  int threads = 128 * warps_per_smsp;  // one block per SM, warps spread over the 4 SMSPs
  k<MODE><<<148, threads>>>(out, in, cyc);
  cudaDeviceSynchronize();
  k<MODE><<<148, threads>>>(out, in, cyc);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / ((double)IT * 5 * CH * warps_per_smsp);
  printf("%-44s warps/SMSP %d: %.2f cycles per instruction per SMSP\n", name, warps_per_smsp, per);
}
int main() {
  float *out, *in; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 256); cudaMalloc(&cyc, 8);
  float h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0f + i * 1e-3f;
  cudaMemcpy(in, h, 256, cudaMemcpyHostToDevice); cudaMemcpyToSymbol(ctab, h, 256);
  for (int w : {1, 2, 4}) {
    run<0>("FFMA2 reg scalar", out, in, cyc, w);
    run<1>("FFMA2 uniform/const scalar", out, in, cyc, w);
    run<4>("FFMA2 reg scalar, shared a operand", out, in, cyc, w);
    run<5>("FFMA2 reg scalar, no reuse", out, in, cyc, w);
    run<2>("FFMA 3-reg", out, in, cyc, w);
    run<6>("FFMA 3-reg, no reuse", out, in, cyc, w);
    run<3>("FFMA const b", out, in, cyc, w);
  }
  return 0;
}
