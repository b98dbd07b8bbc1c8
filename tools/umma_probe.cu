// umma_probe.cu -- stand-alone probe of tcgen05.mma shapes on sm_100a (tools only, not part of libdevo_b200):
//   (1) functional: D = A * B^T for one K-block of 64 halfs with cta_group::{1,2}; the raw TMEM image of every CTA
//       is dumped ([rank][lane 0..127][column]) and decoded on the host against the layouts the kernels assume
//       (cta_group::1 M=128: lane = row, column = n;  cta_group::2 M=128: lane = row % 64 + 64 * (n / (N/2)),
//       column = n % (N/2), CTA r holds rows 64r..64r+63;  cta_group::2 M=256: lane = row % 128, column = n);
//   (2) rate: cycles per K=16 MMA instruction for a stream of back-to-back MMAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/umma_probe tools/umma_probe.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../devo_b200/csrc/tc05.cuh"
using namespace tc05;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

template <int CG>
__device__ __forceinline__ void mma_issue(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 1) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void commit_all(uint32_t bar) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
  }
}

constexpr int kRot = 4;     // distinct operand buffers the rate loop rotates through (defeats any operand reuse inside the MMA unit)

template <int CG, int M, int N>
__global__ void __launch_bounds__(640, 1) probe(const __half* A, const __half* B, float* Draw, int iters, long long* cyc, int rot) {
  constexpr int RA = M / CG, RB = N / CG;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* As = base;
  unsigned char* Bs = base + kRot * RA * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(Bs + kRot * RB * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  volatile int* done = reinterpret_cast<volatile int*>(bar + 6);
  unsigned char* scratch = reinterpret_cast<unsigned char*>(bar + 8);     // 32 KB written by the 'store' warps
  const int rank = (CG == 2) ? (int)cluster_ctarank() : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // K-major SWIZZLE_128B: 16-byte chunk c of row r at r*128 + ((c ^ (r & 7)) << 4)
  for (int q = threadIdx.x; q < RA * 8; q += blockDim.x) {
    const int r = q >> 3, c = q & 7;
    for (int b = 0; b < kRot; b++)
      *reinterpret_cast<uint4*>(As + b * RA * 128 + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + (size_t)(rank * RA + r) * 64 + c * 8);
  }
  for (int q = threadIdx.x; q < RB * 8; q += blockDim.x) {
    const int r = q >> 3, c = q & 7;
    for (int b = 0; b < kRot; b++)
      *reinterpret_cast<uint4*>(Bs + b * RB * 128 + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + (size_t)(rank * RB + r) * 64 + c * 8);
  }
  if (threadIdx.x == 0) {
    *done = 0;
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    mbar_init(bar + 2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // f32 accum, f16 x f16, K-major
  const uint64_t ad = umma_desc_sw128(smem_u32(As)), bd = umma_desc_sw128(smem_u32(Bs));
  uint32_t phase = 0;
  if (rank == 0 && threadIdx.x == 0) {
    for (int k4 = 0; k4 < 4; k4++) mma_issue<CG>(tmem, ad + 2 * k4, bd + 2 * k4, idesc, k4 > 0);
    commit_all<CG>(smem_u32(bar));
  }
  mbar_wait(bar, phase); phase ^= 1;
  tc_fence_after();
  // dump this warp's 32 lanes x 512 columns... only the first NC columns are meaningful
  constexpr int NC = (CG == 2 && M == 128) ? N / 2 : N;
  for (int c0 = 0; c0 < NC && warp < 4; c0 += 8) {
    uint32_t v[8];
    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_wait_ld();
    for (int k = 0; k < 8; k++) Draw[((size_t)rank * 128 + warp * 32 + lane) * 512 + c0 + k] = __uint_as_float(v[k]);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  // ---- rate: iters x 4 MMAs back to back
  if (rank == 0 && threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      const int b = (rot & 1) ? (it % kRot) : 0;
      const uint64_t a2 = ad + (uint64_t)(b * ((RA * 128) >> 4)), b2 = bd + (uint64_t)(b * ((RB * 128) >> 4));
      for (int k4 = 0; k4 < 4; k4++) mma_issue<CG>(tmem, a2 + 2 * k4, b2 + 2 * k4, idesc, 1u);
      if (rot & 2) commit_all<CG>(smem_u32(bar + 1));      // a stage-release commit after every 4 MMAs, like a GEMM main loop
    }
    commit_all<CG>(smem_u32(bar));
    mbar_wait(bar, phase);
    const long long t1 = clock64();
    cyc[blockIdx.x / CG] = t1 - t0;
    *done = 1;
  } else if (threadIdx.x == 0) {
    mbar_wait(bar, phase);      // the peer CTA: the multicast commit lands here too; release this CTA's background warps
    *done = 1;
  } else if (warp >= 4 && warp < 8 && (rot & 8)) {
    // background shared-memory writes (what a TMA weight stream does to the banks): 4 warps x 512 B per instruction
    uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    int o = 0;
    while (!*done) {
      *reinterpret_cast<uint4*>(scratch + (warp - 4) * 8192 + ((o * 512 + lane * 16) & 8191)) = v;
      o++;
    }
  } else if (warp >= 8 && (rot & 4)) {
    // pollers: warps spinning on an mbarrier phase that never completes (what idle epilogue warps do)
    uint32_t ok = 0;
    while (!*done && !ok) {
      asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar + 2)), "r"(0u) : "memory");
    }
  } else {
    mbar_wait(bar, phase);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

template <int CG, int M, int N>
static int run(const char* name, int nclusters, int rot) {
  std::vector<__half> hA((size_t)M * 64), hB((size_t)N * 64);
  for (int r = 0; r < M; r++) for (int k = 0; k < 64; k++) hA[(size_t)r * 64 + k] = __float2half((float)((r * 7 + k * 3) % 5 - 2));
  for (int n = 0; n < N; n++) for (int k = 0; k < 64; k++) hB[(size_t)n * 64 + k] = __float2half((float)((n * 5 + k) % 7 - 3));
  __half *dA, *dB; float* dD; long long* dC;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, (size_t)CG * 128 * 512 * 4)); CK(cudaMalloc(&dC, 256 * 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, (size_t)CG * 128 * 512 * 4));
  const size_t smem = 1024 + (size_t)kRot * (M / CG + N / CG) * 128 + 128 + 32768;
  CK(cudaFuncSetAttribute(probe<CG, M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 2000;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * CG); cfg.blockDim = dim3(640); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, probe<CG, M, N>, (const __half*)dA, (const __half*)dB, dD, iters, dC, rot));
  CK(cudaDeviceSynchronize());
  std::vector<float> hD((size_t)CG * 128 * 512);
  std::vector<long long> hC(256);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hC.data(), dC, 256 * 8, cudaMemcpyDeviceToHost));
  // decode with the assumed layout (the LAST cluster wrote Draw; all clusters write identical data)
  int bad = 0;
  for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) {
    float ref = 0.f;
    for (int k = 0; k < 64; k++) ref += __half2float(hA[(size_t)m * 64 + k]) * __half2float(hB[(size_t)n * 64 + k]);
    int rank, lane, col;
    if (CG == 1) { rank = 0; lane = m; col = n; }
    else if (M == 128) { rank = m / 64; lane = (m % 64) + 64 * (n / (N / 2)); col = n % (N / 2); }
    else { rank = m / 128; lane = m % 128; col = n; }
    const float got = hD[((size_t)rank * 128 + lane) * 512 + col];
    if (got != ref) { if (bad < 5) printf("  %s mismatch m=%d n=%d got %g ref %g\n", name, m, n, got, ref); bad++; }
  }
  long long cmin = hC[0], cmax = hC[0];
  for (int i = 0; i < nclusters; i++) { if (hC[i] < cmin) cmin = hC[i]; if (hC[i] > cmax) cmax = hC[i]; }
  const double per = (double)cmin / (iters * 4.0);
  printf("%-14s rot %d clusters %3d  layout %s (%d bad)  cycles/MMA(K=16) min %.1f max %.1f  => %.0f MAC/cyc/SM\n", name, rot, nclusters,
         bad ? "MISMATCH" : "ok", bad, per, (double)cmax / (iters * 4.0), (double)M * N * 16 / per / CG);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return bad;
}

int main() {
  int bad = 0;
  // rot bits: 1 rotate operands, 2 commit per 4 MMAs, 4 sixteen polling warps, 8 four warps streaming st.shared
  for (int rot : {3, 7, 11, 15})
    for (int nc : {1}) {
      bad += run<1, 128, 192>("cg1 M128 N192", nc, rot);
      bad += run<2, 128, 192>("cg2 M128 N192", nc, rot);
      bad += run<2, 128, 256>("cg2 M128 N256", nc, rot);
      bad += run<2, 256, 192>("cg2 M256 N192", nc, rot);
      bad += run<2, 256, 256>("cg2 M256 N256", nc, rot);
    }
  printf(bad ? "PROBE: LAYOUT MISMATCH\n" : "PROBE: all layouts as assumed\n");
  return bad ? 1 : 0;
}
