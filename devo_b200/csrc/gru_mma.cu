// gru_mma.cu -- the recurrent update operator ("ConvGRU", devo/enet.py:32-99, devo/blocks.py:15-48) as a
// handful of fused tensor-core kernels.
//
// Reference: ~17 cuBLAS GEMMs of [E,384]x[384,384] plus ~40 ATen element-wise launches per iteration; every
// intermediate [E,384] tensor makes a round trip through HBM/L2.  Here one CTA owns a tile of 128 edges and
// walks it through a whole CHAIN of layers without leaving the SM:
//
//   A operand  : the tile's activations, [128 rows x 384] f16 in shared memory (6 K-blocks of 128 x 64, the
//                UMMA K-major SWIZZLE_128B canonical layout), written by the epilogue of the previous layer
//   B operand  : the layer's weights W[384 out, K in] (K-major as stored by nn.Linear), streamed from L2 by TMA in
//                [192 x 64] boxes through a 5-stage mbarrier ring (the producer warp runs ahead across layers)
//   accumulator: [128 x 384] f32 in TMEM (two N=192 halves), tcgen05.mma issued by one elected lane of warp 1
//   epilogue   : 4 warps, one thread per row: tcgen05.ld 32 columns at a time, bias, then the layer's element-wise
//                tail (ReLU / LayerNorm / residual / gate / heads) in registers, and the result goes straight back
//                into the A tile for the next layer.  A LayerNorm sees a whole row inside one thread: no shuffles.
//
// Kernel boundaries remain only where rows of different tiles meet: the neighbour gathers (net[ix], net[jx]) and
// the two SoftAgg segment reductions.  10 launches per update instead of ~60.
//
// Rounding points follow torch.autocast exactly as devo_b200/update.py::forward_fused does (Linear outputs are
// rounded to half, LayerNorm in float32, element-wise ops round to their promoted type); accumulation is fp32 in
// both, only the summation order inside a dot product differs from cuBLAS.
//
// fp32 per-row state (net32 / n32) and the gate scratch use a tile-friendly layout [tile][col/4][128 rows][4]
// (resp. [tile][col/8][128][8] halfs) so that "one thread per row" accesses are fully coalesced.
#include <string.h>
#include <type_traits>
#include "common.cuh"
#include "tc05.cuh"

extern "C" int devo_segment_softmax_sum(const void* g, const void* f, const int32_t* perm, const int32_t* gstart,
                                         const int32_t* ngroups, int max_groups, void* y, int dtype, int n_rows,
                                         int dim, void* stream);

namespace {
using namespace tc05;
using devo::ElemTraits;

constexpr int kRows = 128;                 // edges per tile = MMA M
constexpr int kD = 384;                    // hidden width: N of every layer, K of all but the first
constexpr int kNH = 192;                   // N per MMA / per weight stage
constexpr int kABlk = kRows * 128;         // one K-block of A: 128 rows x 64 halfs = 16 KB
constexpr int kASlots = kD / 64;           // 6
constexpr int kWStage = kNH * 128;         // 24 KB
constexpr int kWStages = 4;
constexpr int kEpiPer = 2;                 // epilogue warps per TMEM lane quarter (each owns a share of the columns)
constexpr int kEpiThreads = 128 * kEpiPer;
constexpr int kThreads = 128 + kEpiThreads;
constexpr int kCB = kD / 32;               // 32-column blocks per row (12)
constexpr int kMaxLayers = 6;
constexpr int kChunks = kD / 8;            // 16-byte chunks per row (48)
constexpr int kCol4 = kD / 4;              // float4 groups per row (96)

enum { PRO_NONE = 0, PRO_GATHER = 1, PRO_CAST = 2, PRO_RESID = 3, PRO_RESID_LN = 4 };
enum { EPI_RELU_A = 0, EPI_LNRELU_A = 1, EPI_ADD3_LN = 2, EPI_RESID = 3, EPI_STORE_A = 4, EPI_STORE_B = 5,
       EPI_GATE = 6, EPI_GATED_LN = 7, EPI_GATED_HEADS = 8 };

template <typename T>
struct GruProg {
  int rows, src_rows;                      // valid rows of this launch; rows of x16_in (gather source)
  int n_layers, pro;
  int kblocks0, stream_a0, use_w0;         // layer 0: K-blocks; A streamed by TMA (tm_a); weights from tm_w0
  int w_row[kMaxLayers];                   // row offset of the layer in the stacked weight matrix (tm_w)
  int epi[kMaxLayers];
  const T* bias[kMaxLayers];
  const float* ln_g[2];
  const float* ln_b[2];
  float eps;
  const T* x16_in;                         // row-major [src_rows,384]: gather source / hidden state in (ADD3)
  const int64_t* idx64;                    // PRO_GATHER: source row per row (-1 => zero row); null => identity
  const int32_t* gid;                      // PRO_RESID*: group of each row
  const T* y16;                            // PRO_RESID*: [groups,384] values added through gid
  const T* inp16;                          // ADD3: imap [n_patches,384]
  const int64_t* kk;                       // ADD3: patch of each row
  float* net32;                            // tile layout
  float* n32;                              // tile layout
  T* gate16;                               // tile layout
  T* out16_a;                              // row-major outputs
  T* out16_b;
  const T* headW;                          // [4,384]
  const T* headB;                          // [4]
  T* delta;                                // [rows,2]
  T* weight;                               // [rows,2]
};

template <typename T> __device__ __forceinline__ float rnd(float v) { return ElemTraits<T>::to_float(ElemTraits<T>::from_float(v)); }

template <typename T> struct Pack8 {
  union { uint4 u; T h[8]; };
};
template <typename T> __device__ __forceinline__ uint4 pack8(const float* v) {
  Pack8<T> p;
#pragma unroll
  for (int k = 0; k < 8; k++) p.h[k] = ElemTraits<T>::from_float(v[k]);
  return p.u;
}
template <typename T> __device__ __forceinline__ void unpack8(uint4 u, float* v) {
  Pack8<T> p;
  p.u = u;
#pragma unroll
  for (int k = 0; k < 8; k++) v[k] = ElemTraits<T>::to_float(p.h[k]);
}

// byte offset of 16-byte chunk c (0..47) of row r inside the A tile (K-major SWIZZLE_128B, 6 K-blocks)
__device__ __forceinline__ uint32_t a_off(int r, int c) {
  return (uint32_t)((c >> 3) * kABlk + r * 128 + (((c & 7) ^ (r & 7)) << 4));
}
// tile layouts: float4 group q (0..95) / half8 chunk c (0..47) of row r of tile t
__device__ __forceinline__ size_t t32(int tile, int q, int r) { return (((size_t)tile * kCol4 + q) * kRows + r) * 4; }
__device__ __forceinline__ size_t t16(int tile, int c, int r) { return (((size_t)tile * kChunks + c) * kRows + r) * 8; }

__device__ __forceinline__ float4 as_f4(uint4 u) {
  return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// the 128 epilogue threads only
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) gru_mma_kernel(const __grid_constant__ CUtensorMap tm_w,
                                                              const __grid_constant__ CUtensorMap tm_w0,
                                                              const __grid_constant__ CUtensorMap tm_a,
                                                              const __grid_constant__ GruProg<T> P) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* As = base;                                   // kASlots x 16 KB
  unsigned char* Ws = base + kASlots * kABlk;                 // kWStages x 24 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ws + kWStages * kWStage);
  uint64_t* w_full = bars;                    // [kWStages]
  uint64_t* w_empty = w_full + kWStages;      // [kWStages]
  uint64_t* a_full = w_empty + kWStages;      // [kASlots]
  uint64_t* a_empty = a_full + kASlots;       // [kASlots]
  uint64_t* acc_full = a_empty + kASlots;     // [1]
  uint64_t* a_ready = acc_full + 1;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);    // bars: 24 x 8 B = 192 B; slot + pad = 16 B
  int* s_idx = reinterpret_cast<int*>(tmem_slot + 4);   // [128] gather sources of the tile
  float* s_stat = reinterpret_cast<float*>(s_idx + kRows);            // [kEpiPer][128][2] LayerNorm partials
  float* s_hacc = s_stat + kEpiPer * kRows * 2;                       // [kEpiPer][128][4] head partials
  float* s_ln = s_hacc + kEpiPer * kRows * 4;                         // [2][2][384]: gamma, beta of the program's LayerNorms
  T* s_bias = reinterpret_cast<T*>(s_ln + 4 * kD);                    // [kMaxLayers][384]
  T* s_head = s_bias + kMaxLayers * kD;                               // [4][384] + [4] (+4 pad)

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int row0 = tile * kRows;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWStages; s++) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < kASlots; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(a_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else if (warp >= 2) {
    // stage the program's small parameters (biases, LayerNorm affine, heads) in shared memory once: the epilogues
    // read them as broadcasts instead of dependent global loads
    const int t = threadIdx.x - 64, nt = kThreads - 64;
    for (int l = 0; l < P.n_layers; l++)
      for (int q = t; q < kD / 8; q += nt)
        reinterpret_cast<uint4*>(s_bias + l * kD)[q] = __ldg(reinterpret_cast<const uint4*>(P.bias[l]) + q);
    for (int k = 0; k < 2; k++) {
      if (P.ln_g[k] == nullptr) continue;
      for (int q = t; q < kD / 4; q += nt) {
        reinterpret_cast<float4*>(s_ln + (2 * k) * kD)[q] = __ldg(reinterpret_cast<const float4*>(P.ln_g[k]) + q);
        reinterpret_cast<float4*>(s_ln + (2 * k + 1) * kD)[q] = __ldg(reinterpret_cast<const float4*>(P.ln_b[k]) + q);
      }
    }
    if (P.headW != nullptr) {
      for (int q = t; q < 4 * kD / 8; q += nt) reinterpret_cast<uint4*>(s_head)[q] = __ldg(reinterpret_cast<const uint4*>(P.headW) + q);
      if (t < 4) s_head[4 * kD + t] = P.headB[t];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool has_pro = (P.pro != PRO_NONE);

  if (warp == 0) {
    // =========================== TMA producer: weights of every layer (+ the streamed A of layer 0) ============
    if (lane == 0) { prefetch_tensormap(&tm_w); if (P.use_w0) prefetch_tensormap(&tm_w0); if (P.stream_a0) prefetch_tensormap(&tm_a); }
    uint32_t stage = 0, phase = 0;
    for (int l = 0; l < P.n_layers; l++) {
      const int nkb = (l == 0) ? P.kblocks0 : kASlots;
      const CUtensorMap* wm = (l == 0 && P.use_w0) ? &tm_w0 : &tm_w;
      const int wrow = P.w_row[l];
      for (int kb = 0; kb < nkb; kb++) {
        if (l == 0 && P.stream_a0) {
          const int slot = kb % kASlots, use = kb / kASlots;
          mbar_wait(&a_empty[slot], (uint32_t)(use & 1) ^ 1u);
          mbar_arrive_expect_tx_elect(smem_u32(&a_full[slot]), kABlk);
          tma_load_2d_elect(smem_u32(As) + slot * kABlk, &tm_a, smem_u32(&a_full[slot]), kb * 64, row0);
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
          mbar_wait(&w_empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx_elect(smem_u32(&w_full[stage]), kWStage);
          tma_load_2d_elect(smem_u32(Ws) + stage * kWStage, wm, smem_u32(&w_full[stage]), kb * 64, wrow + h * kNH);
          if (++stage == kWStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ====================================================================
    const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
    const uint32_t idesc = umma_idesc_f16(fmt, kNH);
    const uint64_t ad0 = umma_desc_sw128(smem_u32(As));
    const uint64_t bd0 = umma_desc_sw128(smem_u32(Ws));
    uint32_t stage = 0, phase = 0, ready_uses = 0;
    for (int l = 0; l < P.n_layers; l++) {
      const int nkb = (l == 0) ? P.kblocks0 : kASlots;
      const bool streamed = (l == 0 && P.stream_a0);
      if (l > 0 || has_pro) {          // the A tile is written by the prologue / the previous layer's epilogue
        mbar_wait(a_ready, ready_uses & 1u);
        ready_uses++;
        tc_fence_after();
      }
      for (int kb = 0; kb < nkb; kb++) {
        const int slot = kb % kASlots;
        if (streamed) { mbar_wait(&a_full[slot], (uint32_t)((kb / kASlots) & 1)); tc_fence_after(); }
        const uint64_t ad = ad0 + (uint64_t)(slot * (kABlk >> 4));
#pragma unroll
        for (int h = 0; h < 2; h++) {
          mbar_wait(&w_full[stage], phase);
          tc_fence_after();
          const uint64_t bd = bd0 + (uint64_t)(stage * (kWStage >> 4));
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++)
            tc_mma_f16_elect(tmem_base + h * kNH, ad + 2 * k4, bd + 2 * k4, idesc, (kb > 0 || k4 > 0) ? 1u : 0u);
          tc_commit_elect(smem_u32(&w_empty[stage]));
          if (++stage == kWStages) { stage = 0; phase ^= 1u; }
        }
        if (streamed) tc_commit_elect(smem_u32(&a_empty[slot]));
      }
      tc_commit_elect(smem_u32(acc_full));
    }
    __syncwarp();
  } else if (warp >= 4) {
    // =========================== prologue + epilogues ==========================================================
    // kEpiPer warps per TMEM lane quarter; a thread owns one row and the column blocks [cb0, cb1) of 32 columns.
    const int quarter = warp & 3;
    const int part = (warp - 4) >> 2;               // 0 .. kEpiPer-1
    const int et = threadIdx.x - 128;               // 0 .. kEpiThreads-1
    const int r = quarter * 32 + lane;              // tile row = TMEM lane
    const int grow = row0 + r;                      // global row
    const bool live = grow < P.rows;
    const int cb0 = part * (kCB / kEpiPer), cb1 = cb0 + kCB / kEpiPer;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    unsigned char* arow = As;                       // + a_off(r, c)

    // row statistics of a LayerNorm: combine the partial (sum, sum of squares) of the kEpiPer threads of a row
    auto ln_stats = [&](float s1, float s2, float& mean, float& rstd) {
      if (kEpiPer > 1) {
        s_stat[(part * kRows + r) * 2 + 0] = s1;
        s_stat[(part * kRows + r) * 2 + 1] = s2;
        epi_bar();
        s1 = 0.f; s2 = 0.f;
#pragma unroll
        for (int p = 0; p < kEpiPer; p++) { s1 += s_stat[(p * kRows + r) * 2 + 0]; s2 += s_stat[(p * kRows + r) * 2 + 1]; }
        epi_bar();                                  // s_stat may be reused right away
      }
      mean = s1 * (1.0f / kD);
      rstd = rsqrtf(fmaxf(s2 * (1.0f / kD) - mean * mean, 0.f) + P.eps);
    };

    // ---------------- prologue ----------------
    if (P.pro == PRO_GATHER) {
      if (et < kRows) {
        const int gr = row0 + et;
        int src = -1;
        if (gr < P.rows) {
          const long long j = P.idx64 ? (long long)P.idx64[gr] : (long long)gr;
          src = (j >= 0 && j < (long long)P.src_rows) ? (int)j : -1;
        }
        s_idx[et] = src;
      }
      epi_bar();
      // cooperative: 48 consecutive threads copy one 768-byte source row; 12 independent 16-byte loads in flight
      constexpr int kPer = kRows * kChunks / kEpiThreads;     // chunks per thread
      constexpr int kBatch = 12;
      static_assert(kPer % kBatch == 0, "gather batches");
#pragma unroll 1
      for (int b = 0; b < kPer / kBatch; b++) {
        uint4 v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; u++) {
          const int q = et + (b * kBatch + u) * kEpiThreads;
          const int rr = q / kChunks, c = q - rr * kChunks;
          const int s = s_idx[rr];
          v[u] = make_uint4(0u, 0u, 0u, 0u);
          if (s >= 0) v[u] = *reinterpret_cast<const uint4*>(P.x16_in + (size_t)s * kD + c * 8);
        }
#pragma unroll
        for (int u = 0; u < kBatch; u++) {
          const int q = et + (b * kBatch + u) * kEpiThreads;
          const int rr = q / kChunks, c = q - rr * kChunks;
          *reinterpret_cast<uint4*>(As + a_off(rr, c)) = v[u];
        }
      }
    } else if (P.pro == PRO_CAST || P.pro == PRO_RESID || P.pro == PRO_RESID_LN) {
      const bool resid = (P.pro != PRO_CAST);
      const bool with_ln = (P.pro == PRO_RESID_LN);
      const int g = (resid && live) ? P.gid[grow] : 0;
      const T* yrow = resid ? P.y16 + (size_t)g * kD : nullptr;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int cb = cb0; cb < cb1; cb++) {
        float4 f[8];
        uint4 yu[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int c = cb * 4 + j;
          f[2 * j] = *reinterpret_cast<const float4*>(P.net32 + t32(tile, 2 * c, r));
          f[2 * j + 1] = *reinterpret_cast<const float4*>(P.net32 + t32(tile, 2 * c + 1, r));
          yu[j] = make_uint4(0u, 0u, 0u, 0u);
          if (resid && live) yu[j] = *reinterpret_cast<const uint4*>(yrow + c * 8);
        }
        uint32_t st[32];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int c = cb * 4 + j;
          float v[8] = {f[2 * j].x, f[2 * j].y, f[2 * j].z, f[2 * j].w, f[2 * j + 1].x, f[2 * j + 1].y, f[2 * j + 1].z, f[2 * j + 1].w};
          if (resid) {
            float y[8];
            unpack8<T>(yu[j], y);
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] += y[k];
          }
          if (with_ln) {
#pragma unroll
            for (int k = 0; k < 8; k++) { s1 += v[k]; s2 += v[k] * v[k]; st[j * 8 + k] = __float_as_uint(v[k]); }
          } else {
            if (resid) {
              *reinterpret_cast<float4*>(P.net32 + t32(tile, 2 * c, r)) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(P.net32 + t32(tile, 2 * c + 1, r)) = make_float4(v[4], v[5], v[6], v[7]);
            }
            *reinterpret_cast<uint4*>(arow + a_off(r, c)) = pack8<T>(v);
          }
        }
        if (with_ln) tmem_st32(trow + cb * 32, st);      // park the fp32 row in TMEM (free until the first MMA)
      }
      if (with_ln) {     // n = LayerNorm(net) -> n32 (float, needed by the gated residual) and the A tile (half)
        tmem_wait_st();
        float mean, rstd;
        ln_stats(s1, s2, mean, rstd);
        const float* gm = s_ln;
        const float* bt = s_ln + kD;
#pragma unroll 1
        for (int cb = cb0; cb < cb1; cb++) {
          uint32_t raw[32];
          tmem_ld32(trow + cb * 32, raw);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int c = cb * 4 + j;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = (__uint_as_float(raw[j * 8 + k]) - mean) * rstd * gm[c * 8 + k] + bt[c * 8 + k];
            *reinterpret_cast<float4*>(P.n32 + t32(tile, 2 * c, r)) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(P.n32 + t32(tile, 2 * c + 1, r)) = make_float4(v[4], v[5], v[6], v[7]);
            *reinterpret_cast<uint4*>(arow + a_off(r, c)) = pack8<T>(v);
          }
        }
        tc_fence_before();
      }
    }
    if (has_pro) {
      fence_proxy_async();
      epi_bar();
      if (et == 0) mbar_arrive(a_ready);
    }

    // ---------------- per-layer epilogues ----------------
    int ln_used = (P.pro == PRO_RESID_LN) ? 1 : 0;
    for (int l = 0; l < P.n_layers; l++) {
      const int epi = P.epi[l];
      const T* bias = s_bias + l * kD;
      float s1 = 0.f, s2 = 0.f;
      float hacc[4] = {0.f, 0.f, 0.f, 0.f};
      const T* netrow = nullptr;
      const T* inprow = nullptr;
      if (epi == EPI_ADD3_LN && live) {
        netrow = P.x16_in + (size_t)grow * kD;
        inprow = P.inp16 + (size_t)P.kk[grow] * kD;
      }
      T* orow = nullptr;
      if (live) {
        if (epi == EPI_STORE_A || epi == EPI_RESID || epi == EPI_GATED_HEADS || epi == EPI_ADD3_LN) orow = P.out16_a ? P.out16_a + (size_t)grow * kD : nullptr;
        if (epi == EPI_STORE_B) orow = P.out16_b ? P.out16_b + (size_t)grow * kD : nullptr;
      }
      // per-row operands of the element-wise tail, fetched one column block ahead of their use
      struct Aux { uint4 q[12]; };     // RESID: q[0..7] net32 ; GATED: q[0..7] n32, q[8..11] gate ; ADD3: q[0..3] net16, q[4..7] inp16
      auto load_aux = [&](int cb, Aux& a) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int c = cb * 4 + j;
          if (epi == EPI_RESID) {
            a.q[2 * j] = *reinterpret_cast<const uint4*>(P.net32 + t32(tile, 2 * c, r));
            a.q[2 * j + 1] = *reinterpret_cast<const uint4*>(P.net32 + t32(tile, 2 * c + 1, r));
          } else if (epi == EPI_GATED_LN || epi == EPI_GATED_HEADS) {
            a.q[2 * j] = *reinterpret_cast<const uint4*>(P.n32 + t32(tile, 2 * c, r));
            a.q[2 * j + 1] = *reinterpret_cast<const uint4*>(P.n32 + t32(tile, 2 * c + 1, r));
            a.q[8 + j] = *reinterpret_cast<const uint4*>(P.gate16 + t16(tile, c, r));
          } else if (epi == EPI_ADD3_LN) {
            a.q[j] = make_uint4(0u, 0u, 0u, 0u);
            a.q[4 + j] = a.q[j];
            if (live) { a.q[j] = *reinterpret_cast<const uint4*>(netrow + c * 8); a.q[4 + j] = __ldg(reinterpret_cast<const uint4*>(inprow + c * 8)); }
          }
        }
      };
      Aux cur, nxt;
      load_aux(cb0, cur);                           // issued before the accumulator is ready: overlaps the MMAs
      mbar_wait(acc_full, (uint32_t)(l & 1));
      tc_fence_after();
#pragma unroll 1
      for (int cb = cb0; cb < cb1; cb++) {
        uint32_t raw[32];
        tmem_ld32(trow + cb * 32, raw);
        if (cb + 1 < cb1) load_aux(cb + 1, nxt);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int c = cb * 4 + j;                 // 16-byte chunk index (8 columns)
          float o[8];
          {
            float bv[8];
            unpack8<T>(*reinterpret_cast<const uint4*>(bias + c * 8), bv);
#pragma unroll
            for (int k = 0; k < 8; k++) o[k] = rnd<T>(__uint_as_float(raw[j * 8 + k]) + bv[k]);   // Linear output (half)
          }
          if (epi == EPI_RELU_A) {
#pragma unroll
            for (int k = 0; k < 8; k++) o[k] = fmaxf(o[k], 0.f);
            *reinterpret_cast<uint4*>(arow + a_off(r, c)) = pack8<T>(o);
          } else if (epi == EPI_LNRELU_A) {
#pragma unroll
            for (int k = 0; k < 8; k++) { s1 += o[k]; s2 += o[k] * o[k]; }
            *reinterpret_cast<uint4*>(arow + a_off(r, c)) = pack8<T>(o);
          } else if (epi == EPI_ADD3_LN) {
            float a[8], b[8];
            unpack8<T>(cur.q[j], a);
            unpack8<T>(cur.q[4 + j], b);
#pragma unroll
            for (int k = 0; k < 8; k++) { o[k] = rnd<T>(rnd<T>(a[k] + b[k]) + o[k]); s1 += o[k]; s2 += o[k] * o[k]; }
            *reinterpret_cast<uint4*>(arow + a_off(r, c)) = pack8<T>(o);
          } else if (epi == EPI_RESID) {
            float4 a = as_f4(cur.q[2 * j]), b = as_f4(cur.q[2 * j + 1]);
            a.x += o[0]; a.y += o[1]; a.z += o[2]; a.w += o[3]; b.x += o[4]; b.y += o[5]; b.z += o[6]; b.w += o[7];
            *reinterpret_cast<float4*>(P.net32 + t32(tile, 2 * c, r)) = a;
            *reinterpret_cast<float4*>(P.net32 + t32(tile, 2 * c + 1, r)) = b;
            if (orow) {
              const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
              *reinterpret_cast<uint4*>(orow + c * 8) = pack8<T>(v);
            }
          } else if (epi == EPI_STORE_A || epi == EPI_STORE_B) {
            if (orow) *reinterpret_cast<uint4*>(orow + c * 8) = pack8<T>(o);
          } else if (epi == EPI_GATE) {
            *reinterpret_cast<uint4*>(P.gate16 + t16(tile, c, r)) = pack8<T>(o);
          } else {   // EPI_GATED_LN / EPI_GATED_HEADS:  x = n + half(half(sigmoid(gate)) * res)
            float g[8];
            unpack8<T>(cur.q[8 + j], g);
            const float4 a = as_f4(cur.q[2 * j]), b = as_f4(cur.q[2 * j + 1]);
            float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] += rnd<T>(rnd<T>(sigmoidf_(g[k])) * o[k]);
            if (epi == EPI_GATED_LN) {
#pragma unroll
              for (int k = 0; k < 8; k++) { s1 += x[k]; s2 += x[k] * x[k]; raw[j * 8 + k] = __float_as_uint(x[k]); }
            } else {
              if (orow) *reinterpret_cast<uint4*>(orow + c * 8) = pack8<T>(x);       // new hidden state (half)
              float hw[8];
#pragma unroll
              for (int k = 0; k < 8; k++) x[k] = rnd<T>(fmaxf(x[k], 0.f));
#pragma unroll
              for (int o4 = 0; o4 < 4; o4++) {
                unpack8<T>(*reinterpret_cast<const uint4*>(s_head + o4 * kD + c * 8), hw);
#pragma unroll
                for (int k = 0; k < 8; k++) hacc[o4] += x[k] * hw[k];
              }
            }
          }
        }
        if (epi == EPI_GATED_LN) tmem_st32(trow + cb * 32, raw);  // fp32 row parked in its own accumulator columns
        if (cb + 1 < cb1) cur = nxt;
      }
      // ---------------- row-wise tails ----------------
      if (epi == EPI_LNRELU_A || epi == EPI_ADD3_LN) {
        // the row (half values) sits in this thread's slice of the A tile: one more pass over shared memory
        float mean, rstd;
        ln_stats(s1, s2, mean, rstd);
        const float* gm = s_ln + ln_used * 2 * kD;
        const float* bt = gm + kD;
#pragma unroll 2
        for (int c = cb0 * 4; c < cb1 * 4; c++) {
          float v[8];
          unpack8<T>(*reinterpret_cast<const uint4*>(arow + a_off(r, c)), v);
#pragma unroll
          for (int k = 0; k < 8; k++) v[k] = (v[k] - mean) * rstd * gm[c * 8 + k] + bt[c * 8 + k];
          if (epi == EPI_LNRELU_A) {
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = fmaxf(v[k], 0.f);
            *reinterpret_cast<uint4*>(arow + a_off(r, c)) = pack8<T>(v);
          } else {
            *reinterpret_cast<float4*>(P.net32 + t32(tile, 2 * c, r)) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(P.net32 + t32(tile, 2 * c + 1, r)) = make_float4(v[4], v[5], v[6], v[7]);
            if (orow) *reinterpret_cast<uint4*>(orow + c * 8) = pack8<T>(v);
          }
        }
        ln_used++;
      } else if (epi == EPI_GATED_LN) {
        tmem_wait_st();
        float mean, rstd;
        ln_stats(s1, s2, mean, rstd);
        const float* gm = s_ln + ln_used * 2 * kD;
        const float* bt = gm + kD;
#pragma unroll 1
        for (int cb = cb0; cb < cb1; cb++) {
          uint32_t raw[32];
          tmem_ld32(trow + cb * 32, raw);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int c = cb * 4 + j;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = (__uint_as_float(raw[j * 8 + k]) - mean) * rstd * gm[c * 8 + k] + bt[c * 8 + k];
            *reinterpret_cast<float4*>(P.n32 + t32(tile, 2 * c, r)) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(P.n32 + t32(tile, 2 * c + 1, r)) = make_float4(v[4], v[5], v[6], v[7]);
            *reinterpret_cast<uint4*>(arow + a_off(r, c)) = pack8<T>(v);
          }
        }
        ln_used++;
      } else if (epi == EPI_GATED_HEADS) {
        if (kEpiPer > 1) {
#pragma unroll
          for (int o4 = 0; o4 < 4; o4++) s_hacc[(part * kRows + r) * 4 + o4] = hacc[o4];
          epi_bar();
#pragma unroll
          for (int o4 = 0; o4 < 4; o4++) {
            hacc[o4] = 0.f;
#pragma unroll
            for (int p = 0; p < kEpiPer; p++) hacc[o4] += s_hacc[(p * kRows + r) * 4 + o4];
          }
        }
        if (live && part == 0) {
          const float d0 = rnd<T>(hacc[0] + ElemTraits<T>::to_float(s_head[4 * kD + 0]));
          const float d1 = rnd<T>(hacc[1] + ElemTraits<T>::to_float(s_head[4 * kD + 1]));
          const float w0 = rnd<T>(hacc[2] + ElemTraits<T>::to_float(s_head[4 * kD + 2]));
          const float w1 = rnd<T>(hacc[3] + ElemTraits<T>::to_float(s_head[4 * kD + 3]));
          P.delta[(size_t)grow * 2 + 0] = ElemTraits<T>::from_float(d0);
          P.delta[(size_t)grow * 2 + 1] = ElemTraits<T>::from_float(d1);
          P.weight[(size_t)grow * 2 + 0] = ElemTraits<T>::from_float(sigmoidf_(w0));
          P.weight[(size_t)grow * 2 + 1] = ElemTraits<T>::from_float(sigmoidf_(w1));
        }
      }
      tc_fence_before();
      if (l + 1 < P.n_layers) {       // hand the A tile / the TMEM accumulator back to the MMA warp
        fence_proxy_async();
        epi_bar();
        if (et == 0) mbar_arrive(a_ready);
      }
    }
  }
  // teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// row-major [rows, cols] 16-bit matrix, box = [box_rows x 64 cols], SWIZZLE_128B
static int make_map_2d(CUtensorMap* m, int dtype, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  DEVO_REQUIRE(enc != nullptr, DEVO_EUNSUPPORTED, "gru_update: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, dtype == DEVO_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DEVO_REQUIRE(r == CUDA_SUCCESS, DEVO_EINVAL, "gru_update: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return DEVO_OK;
}

constexpr size_t kSmemBytes = 1024 + (size_t)kASlots * kABlk + (size_t)kWStages * kWStage + 24 * sizeof(uint64_t) + 16 + kRows * sizeof(int) +
                              (size_t)kEpiPer * kRows * 6 * sizeof(float) + 4 * kD * sizeof(float) + (kMaxLayers * kD + 4 * kD + 8) * 2 + 64;

template <typename T>
static int launch_prog(const CUtensorMap& tw, const CUtensorMap& tw0, const CUtensorMap& ta, const GruProg<T>& P, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    DEVO_CUDA(cudaFuncSetAttribute(gru_mma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    configured = true;
  }
  const int tiles = (P.rows + kRows - 1) / kRows;
  if (tiles <= 0) return DEVO_OK;
  gru_mma_kernel<T><<<tiles, kThreads, kSmemBytes, s>>>(tw, tw0, ta, P);
  DEVO_LAUNCH_CHECK("gru_mma");
  return DEVO_OK;
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
struct GruWs {
  size_t net32, n32, gate16, x16a, x16b, g16, f16, y16, hy16, total;
};
static GruWs gru_ws(int E, int max_groups) {
  GruWs w;
  const size_t tiles = (size_t)(E + kRows - 1) / kRows;
  const size_t G = (size_t)(max_groups > 0 ? max_groups : 1);
  size_t off = 0;
  w.net32 = off; off += al256(tiles * kRows * kD * 4);
  w.n32 = off;   off += al256(tiles * kRows * kD * 4);
  w.gate16 = off; off += al256(tiles * kRows * kD * 2);
  w.x16a = off;  off += al256((size_t)E * kD * 2);
  w.x16b = off;  off += al256((size_t)E * kD * 2);
  w.g16 = off;   off += al256((size_t)E * kD * 2);
  w.f16 = off;   off += al256((size_t)E * kD * 2);
  w.y16 = off;   off += al256(G * kD * 2);
  w.hy16 = off;  off += al256(G * kD * 2);
  w.total = off;
  return w;
}

template <typename T>
static int gru_update_impl(const devo_gru_weights_t* Wt, const devo_gru_io_t* io, int dtype, void* workspace, cudaStream_t s) {
  const int E = io->E;
  const int maxG = io->max_groups_kk > io->max_groups_ij ? io->max_groups_kk : io->max_groups_ij;
  const GruWs L = gru_ws(E, maxG);
  char* w = (char*)workspace;
  float* net32 = (float*)(w + L.net32);
  float* n32 = (float*)(w + L.n32);
  T* gate16 = (T*)(w + L.gate16);
  T* x16a = (T*)(w + L.x16a);
  T* x16b = (T*)(w + L.x16b);
  T* g16 = (T*)(w + L.g16);
  T* f16 = (T*)(w + L.f16);
  T* y16 = (T*)(w + L.y16);
  T* hy16 = (T*)(w + L.hy16);
  const T* bias = (const T*)Wt->bias;            // [19,384]: row 0 = corr[0], row 1+i = stacked layer i
  auto B = [&](int layer) { return bias + (size_t)(1 + layer) * kD; };
  CUtensorMap tw, tw0, ta;
  int rc = make_map_2d(&tw, dtype, Wt->W, 18 * kD, kD, kD, kNH);
  if (rc != DEVO_OK) return rc;
  rc = make_map_2d(&tw0, dtype, Wt->W0, kD, (uint64_t)io->corr_ld, (uint64_t)io->corr_ld, kNH);
  if (rc != DEVO_OK) return rc;
  rc = make_map_2d(&ta, dtype, io->corr16, (uint64_t)E, (uint64_t)io->corr_ld, (uint64_t)io->corr_ld, kRows);
  if (rc != DEVO_OK) return rc;

  GruProg<T> base;
  memset(&base, 0, sizeof(base));
  base.rows = E; base.src_rows = E; base.eps = Wt->ln_eps;
  base.kblocks0 = kASlots;
  base.net32 = net32; base.n32 = n32; base.gate16 = gate16;

  // (1) corr MLP + norm(net + inp + corr)  (enet.py:59-66,82-83)
  {
    GruProg<T> P = base;
    P.n_layers = 3; P.pro = PRO_NONE; P.kblocks0 = io->corr_ld / 64; P.stream_a0 = 1; P.use_w0 = 1;
    P.w_row[0] = 0; P.epi[0] = EPI_RELU_A; P.bias[0] = bias;
    P.w_row[1] = 0 * kD; P.epi[1] = EPI_LNRELU_A; P.bias[1] = B(0);
    P.w_row[2] = 1 * kD; P.epi[2] = EPI_ADD3_LN; P.bias[2] = B(1);
    P.ln_g[0] = Wt->ln_gamma; P.ln_b[0] = Wt->ln_beta;
    P.ln_g[1] = Wt->ln_gamma + kD; P.ln_b[1] = Wt->ln_beta + kD;
    P.x16_in = (const T*)io->net16; P.inp16 = (const T*)io->imap16; P.kk = io->kk;
    P.out16_a = x16a;
    rc = launch_prog<T>(tw, tw0, ta, P, s);
    if (rc != DEVO_OK) return rc;
  }
  // (2,3) net += c1(mask * net[ix]) ; net += c2(mask * net[jx])   (enet.py:86-91)
  for (int k = 0; k < 2; k++) {
    GruProg<T> P = base;
    P.n_layers = 2; P.pro = PRO_GATHER;
    P.w_row[0] = (2 + 2 * k) * kD; P.epi[0] = EPI_RELU_A; P.bias[0] = B(2 + 2 * k);
    P.w_row[1] = (3 + 2 * k) * kD; P.epi[1] = EPI_RESID;  P.bias[1] = B(3 + 2 * k);
    P.x16_in = k == 0 ? x16a : x16b; P.idx64 = k == 0 ? io->ix : io->jx;
    P.out16_a = k == 0 ? x16b : nullptr;
    rc = launch_prog<T>(tw, tw0, ta, P, s);
    if (rc != DEVO_OK) return rc;
  }
  // (4,5) net += SoftAgg(net) over patches, then over frame pairs  (enet.py:93-94, blocks.py:40-48)
  for (int k = 0; k < 2; k++) {
    const int32_t* perm = k == 0 ? io->perm_kk : io->perm_ij;
    const int32_t* gstart = k == 0 ? io->gstart_kk : io->gstart_ij;
    const int32_t* ngroups = k == 0 ? io->ngroups_kk : io->ngroups_ij;
    const int mg = k == 0 ? io->max_groups_kk : io->max_groups_ij;
    {
      GruProg<T> P = base;
      P.n_layers = 2; P.pro = k == 0 ? PRO_CAST : PRO_RESID;
      P.gid = io->gid_kk; P.y16 = hy16;                       // k == 1: the patch-wise aggregate is added first
      P.w_row[0] = (6 + 3 * k) * kD; P.epi[0] = EPI_STORE_A; P.bias[0] = B(6 + 3 * k);     // g
      P.w_row[1] = (7 + 3 * k) * kD; P.epi[1] = EPI_STORE_B; P.bias[1] = B(7 + 3 * k);     // f
      P.out16_a = g16; P.out16_b = f16;
      rc = launch_prog<T>(tw, tw0, ta, P, s);
      if (rc != DEVO_OK) return rc;
    }
    rc = devo_segment_softmax_sum(g16, f16, perm, gstart, ngroups, mg, y16, dtype, E, kD, (void*)s);
    if (rc != DEVO_OK) return rc;
    {
      GruProg<T> P = base;
      P.rows = mg; P.src_rows = mg;
      P.n_layers = 1; P.pro = PRO_GATHER; P.x16_in = y16; P.idx64 = nullptr;
      P.w_row[0] = (8 + 3 * k) * kD; P.epi[0] = EPI_STORE_A; P.bias[0] = B(8 + 3 * k);     // h
      P.out16_a = hy16;
      rc = launch_prog<T>(tw, tw0, ta, P, s);
      if (rc != DEVO_OK) return rc;
    }
  }
  // (6) gru: LN, GatedResidual, LN, GatedResidual; heads  (enet.py:68-77,96-99)
  {
    GruProg<T> P = base;
    P.n_layers = 6; P.pro = PRO_RESID_LN;
    P.gid = io->gid_ij; P.y16 = hy16;
    P.ln_g[0] = Wt->ln_gamma + 2 * kD; P.ln_b[0] = Wt->ln_beta + 2 * kD;
    P.ln_g[1] = Wt->ln_gamma + 3 * kD; P.ln_b[1] = Wt->ln_beta + 3 * kD;
    const int epis[6] = {EPI_GATE, EPI_RELU_A, EPI_GATED_LN, EPI_GATE, EPI_RELU_A, EPI_GATED_HEADS};
    for (int l = 0; l < 6; l++) { P.w_row[l] = (12 + l) * kD; P.epi[l] = epis[l]; P.bias[l] = B(12 + l); }
    P.out16_a = (T*)io->net16_out;
    P.headW = (const T*)Wt->head_W; P.headB = (const T*)Wt->head_b;
    P.delta = (T*)io->delta; P.weight = (T*)io->weight;
    rc = launch_prog<T>(tw, tw0, ta, P, s);
    if (rc != DEVO_OK) return rc;
  }
  return DEVO_OK;
}

}  // namespace

extern "C" {

size_t devo_gru_workspace(int E, int max_groups) { return gru_ws(E, max_groups).total; }

int devo_gru_update(const devo_gru_weights_t* weights, const devo_gru_io_t* io, int dtype, void* workspace,
                    size_t workspace_bytes, void* stream) {
  DEVO_REQUIRE(weights && io, DEVO_EINVAL, "gru_update: NULL argument");
  DEVO_REQUIRE(dtype == DEVO_F16 || dtype == DEVO_BF16, DEVO_EUNSUPPORTED, "gru_update: dtype must be f16 or bf16");
  DEVO_REQUIRE(io->dim == kD, DEVO_EUNSUPPORTED, "gru_update: hidden width must be %d (got %d)", kD, io->dim);
  DEVO_REQUIRE(io->corr_ld > 0 && io->corr_ld % 64 == 0, DEVO_EINVAL,
               "gru_update: correlation rows must be zero-padded to a multiple of 64 (ld = %d)", io->corr_ld);
  DEVO_REQUIRE(io->max_groups_kk > 0 && io->max_groups_ij > 0, DEVO_EINVAL, "gru_update: max_groups must be > 0");
  if (io->E <= 0) return DEVO_OK;
  const int maxG = io->max_groups_kk > io->max_groups_ij ? io->max_groups_kk : io->max_groups_ij;
  DEVO_REQUIRE(workspace && workspace_bytes >= gru_ws(io->E, maxG).total, DEVO_EWORKSPACE, "gru_update: workspace too small");
  DEVO_REQUIRE((((uintptr_t)weights->W | (uintptr_t)weights->W0 | (uintptr_t)io->corr16 | (uintptr_t)io->net16 |
                 (uintptr_t)io->imap16 | (uintptr_t)io->net16_out | (uintptr_t)weights->bias | (uintptr_t)weights->head_W |
                 (uintptr_t)workspace) & 15) == 0, DEVO_EINVAL, "gru_update: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == DEVO_F16) return gru_update_impl<__half>(weights, io, dtype, workspace, s);
  return gru_update_impl<__nv_bfloat16>(weights, io, dtype, workspace, s);
}

}  // extern "C"
