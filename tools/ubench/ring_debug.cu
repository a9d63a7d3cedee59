// Debug harness for k_lloyd_ring: one launch on a synthetic plane, dumps what lane 0..3 of block 0 /
// warp 0 saw in the first iteration (built with -DRING_DEBUG).
#include <cstdio>
#include <vector>
#include "../../kmeans-gpu_b200/csrc/kmg_lloyd_ring.cuh"
using namespace kmg;
int main() {
  const unsigned long long n = 1 << 20;
  uint32_t* rgba; float4* work; float* lut;
  cudaMalloc(&rgba, n * 4); cudaMalloc(&work, n * 16); cudaMalloc(&lut, 1024);
  k_build_srgb_table<<<1, 256>>>(lut);
  k_synth<<<148, 256>>>(rgba, 0, n, 0, 2, 16);
  k_convert<<<148 * 4, 256>>>(rgba, n, 0, lut, work);
  const unsigned k = 8;
  unsigned char* blob; cudaMalloc(&blob, 1 << 20); cudaMemset(blob, 0, 1 << 20);
  JobPtrs J;
  J.st = (JobState*)blob; J.cent = (float4*)(blob + 256); J.tab = (CentRec*)(blob + 1024);
  J.acc = (long long*)(blob + 8192); J.last = (long long*)(blob + 65536); J.keys = (unsigned long long*)(blob + 70000 / 8 * 8);
  J.pal = (uint32_t*)(blob + 80000); J.acc_copies = 8;
  JobState st = {}; st.k = k; st.max_iter = 1000; st.check_every = 0; st.conv_threshold = 1.0f;
  cudaMemcpy(J.st, &st, sizeof(st), cudaMemcpyHostToDevice);
  std::vector<float4> hw(n); cudaMemcpy(hw.data(), work, n * 16, cudaMemcpyDeviceToHost);
  std::vector<float4> cent(k);
  for (unsigned i = 0; i < k; ++i) { cent[i] = hw[i * 1000 + 17]; cent[i].w = 1.0f; }
  cudaMemcpy(J.cent, cent.data(), k * 16, cudaMemcpyHostToDevice);
  k_prepare<<<1, 256>>>(J, 0, 0);
  void* ctab; cudaGetSymbolAddress(&ctab, c_tab);
  cudaMemcpy(ctab, J.tab, 8 * sizeof(CentRec), cudaMemcpyDeviceToDevice);
  using L = RingLayout<8, 8, 2, 8>;
  auto fn = k_lloyd_ring<8, 8, 2, 8, 2>;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES);
  PeerXchg X = {};
  fn<<<296, L::THREADS, L::BYTES>>>(J, work, n, 0, 0, X, 0, k);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  JobState out; cudaMemcpy(&out, J.st, sizeof(out), cudaMemcpyDeviceToHost);
  printf("passes %u slow %llu of %llu\n", out.passes, out.slow_pixels, n);
  std::vector<float> dbg(64 * 32); cudaMemcpyFromSymbol(dbg.data(), g_dbg, dbg.size() * 4);
  for (int l = 0; l < 4; ++l) {
    printf("lane %d:", l);
    for (int i = 0; i < 24; ++i) printf(" %g", dbg[l * 64 + i]);
    printf("\n   ua=%08x\n", *(unsigned*)&dbg[l * 64 + 24]);
  }
  printf("px0 true: %g %g %g %g\n", hw[0].x, hw[0].y, hw[0].z, hw[0].w);
  return 0;
}
