// gru_mma.cu -- the recurrent update operator ("ConvGRU", devo/enet.py:32-99, devo/blocks.py:15-48) as a
// handful of fused tensor-core kernels.
//
// Reference: ~17 cuBLAS GEMMs of [E,384]x[384,384] plus ~40 ATen element-wise launches per iteration; every
// intermediate [E,384] tensor makes a round trip through HBM/L2.  Here a CLUSTER OF 2 CTAs owns a tile of 128 edges and
// walks it through a whole CHAIN of layers without leaving the two SMs; each CTA computes 192 of the 384 output columns:
//
//   A operand  : the tile's activations, [128 rows x 384] f16 in shared memory of BOTH CTAs (6 K-blocks of 128 x 64,
//                the UMMA K-major SWIZZLE_128B canonical layout); the epilogue of layer l writes the next A operand
//                in place -- its own column slice locally, the peer's copy through distributed shared memory
//   B operand  : this CTA's slice of the layer's weights W[384 out, K in] (K-major as stored by nn.Linear), streamed
//                from L2 by TMA in [192 x 64] boxes through a 4-stage mbarrier ring (the producer warp runs ahead
//                across layers: weights do not depend on activations)
//   accumulator: two [128 x 192] f32 tiles in TMEM (layer l uses l & 1); tcgen05.mma issued by one elected lane of
//                warp 1; tcgen05.commit multicast to both CTAs says "this layer's MMAs are done everywhere", which is
//                what allows a CTA to overwrite its peer's A tile.  When an epilogue leaves the A tile alone (gate,
//                g/f stores) the next layer's MMAs run while it still reads the other accumulator.
//   epilogue   : 12 warps, one thread per row and column share: tcgen05.ld 8 columns at a time, bias, then the
//                layer's element-wise tail (ReLU / LayerNorm / residual / gate / heads) in registers.  LayerNorm row
//                statistics are exchanged between the two CTAs through DSMEM + an mbarrier; the fp32 row waits in
//                the thread's own TMEM columns between the two passes.
//   launches   : programmatic dependent launch -- barrier / TMEM / parameter set-up and the weight prefetch of a
//                launch overlap the tail of its predecessor (griddepcontrol).
//
// Kernel boundaries remain only where rows of different tiles meet: the neighbour gathers (net[ix], net[jx]) and
// the two SoftAgg segment reductions (their `h` layers are applied per edge inside the consuming kernel).  8 launches per
// update (6 of this kernel + 2 segment reductions) instead of ~60.  History, measurements and rejected variants:
// DESIGN.md 2.6, profiles/r01_notes.md, tools/gru_timing.py.
//
// Rounding points follow torch.autocast exactly as devo_b200/update.py::forward_fused does (Linear outputs are
// rounded to half, LayerNorm in float32, element-wise ops round to their promoted type); accumulation is fp32 in
// both, only the summation order inside a dot product differs from cuBLAS.
//
// fp32 per-row state (net32 / n32) and the gate scratch use a tile-friendly layout [tile][col/4][128 rows][4]
// (resp. [tile][col/8][128][8] halfs) so that "one thread per row" accesses are fully coalesced.
#include <stdlib.h>
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
#ifndef DEVO_GRU_CHUNKS_PER_ITER
#define DEVO_GRU_CHUNKS_PER_ITER 1
#endif
#ifndef DEVO_GRU_SPLIT
#define DEVO_GRU_SPLIT 2
#endif
constexpr int kSplit = DEVO_GRU_SPLIT;     // CTAs per cluster: each owns kD/kSplit output columns of the same 128 rows
constexpr int kNC = kD / kSplit;           // N per CTA (128 or 192)
constexpr int kABlk = kRows * 128;         // one K-block of A: 128 rows x 64 halfs = 16 KB
constexpr int kASlots = kD / 64;           // 6
constexpr int kWStage = kNC * 128;         // one K-block of this CTA's weight slice (16 / 24 KB)
constexpr int kWStages = (96 * 1024) / kWStage;   // 6 / 4
constexpr int kEpiPer = 3;                 // epilogue warps per TMEM lane quarter (each owns a share of the columns)
constexpr int kEpiThreads = 128 * kEpiPer;
constexpr int kFirstEpiWarp = 2;            // warps 0 (TMA) and 1 (MMA) + the epilogue warps; any 4 consecutive warps cover the 4 TMEM lane quarters
constexpr int kThreads = 32 * kFirstEpiWarp + kEpiThreads;
constexpr int kMaxLayers = 7;
constexpr int kChunks = kD / 8;            // 16-byte chunks per row (48)
constexpr int kCol4 = kD / 4;              // float4 groups per row (96)
constexpr int kCB = kD / 32;               // 32-column blocks per row (12)
constexpr int kCBc = kNC / 32;             // ... per CTA (4 / 6)
constexpr int kCBp = kCBc / kEpiPer;       // ... per epilogue thread (2 / 3)
constexpr int kStageStride = kNC / 8 + 1;  // output staging tile: 16-byte chunks per row (+1: bank-conflict padding)
constexpr int kStageBytes = kRows * kStageStride * 16;
static_assert(kD % kSplit == 0 && kNC % 32 == 0 && kCBc % kEpiPer == 0 && kNC % 16 == 0 && kNC <= 256, "bad split");

enum { PRO_NONE = 0, PRO_GATHER = 1, PRO_CAST = 2 };
enum { EPI_RELU_A = 0, EPI_LNRELU_A = 1, EPI_ADD3_LN = 2, EPI_RESID = 3, EPI_STORE_A = 4, EPI_STORE_B = 5,
       EPI_GATE = 6, EPI_GATED_LN = 7, EPI_GATED_HEADS = 8, EPI_RESID_A = 9, EPI_RESID_LN_A = 10 };

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
  const int32_t* idx32;                    // PRO_GATHER: ... or a 32-bit index (the group of each row) when idx64 is null
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
  const float* coords;                     // optional [rows,2,3,3]: target32 = centre coordinate + delta
  float* target32;                         // optional [rows,2]
  float* weight32;                         // optional [rows,2]
  long long* dbg;                          // optional: %globaltimer stamps of CTA 0 (tools/gru_timing.py)
};

// float <-> half conversions always two at a time: cvt.rn.f16x2.f32 (F2FP.PACK_AB) runs at full rate, the scalar
// cvt.rn.f16.f32 (F2F) only at 16 lanes/clk/SM -- with four scalar conversions per element the epilogue was bound by
// that one pipe (measured: 2.1 us per layer for 64 elements per thread; tools/gru_timing.py).
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t u);
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u)); }

template <typename T> __device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack2<T>(v[0], v[1]), pack2<T>(v[2], v[3]), pack2<T>(v[4], v[5]), pack2<T>(v[6], v[7]));
}
template <typename T> __device__ __forceinline__ void unpack8(uint4 u, float* v) {
  const float2 a = unpack2<T>(u.x), b = unpack2<T>(u.y), c = unpack2<T>(u.z), d = unpack2<T>(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
template <typename T> __device__ __forceinline__ uint32_t relu2(uint32_t u);
template <> __device__ __forceinline__ uint32_t relu2<__half>(uint32_t u) {
  const __half2 h = __hmax2(*reinterpret_cast<const __half2*>(&u), __float2half2_rn(0.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t relu2<__nv_bfloat16>(uint32_t u) {
  const __nv_bfloat162 h = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&u), __float2bfloat162_rn(0.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <typename T> __device__ __forceinline__ uint4 relu8(uint4 u) {
  return make_uint4(relu2<T>(u.x), relu2<T>(u.y), relu2<T>(u.z), relu2<T>(u.w));
}
// round eight floats to T and back (the autocast rounding point of a half-typed intermediate)
template <typename T> __device__ __forceinline__ void rnd8(float* v) { unpack8<T>(pack8<T>(v), v); }

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
__device__ __forceinline__ void stamp(long long* dbg, int slot) {
  if (dbg != nullptr && blockIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[slot] = t;
  }
}
// sigmoid in 4 instructions (FMUL, MUFU.EX2, FADD, MUFU.RCP): the .ftz forms skip the denormal range fix-ups of
// __expf / __fdividef; the result is rounded to half right away, far coarser than the approximation error
__device__ __forceinline__ float sigmoidf_(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

// the epilogue threads only
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// ---- the epilogue role: one thread per (row, column share).  Every element-wise tail is a separate template
// instantiation (constexpr EPI / PRO), selected once per layer by a warp-uniform switch: tight straight-line loops
// instead of a branch chain per 8 columns (the first version spent ~150 instructions per 8 columns; tools/gru_timing.py).
template <typename T>
struct Epi {
  const GruProg<T>& P;
  unsigned char* As;
  unsigned char* Ws;           // the weight ring: idle during the LAST layer's epilogue, reused as an output staging tile
  float* s_stat; float* s_hacc; const float* s_ln; const float* s_bias; const T* s_head; int* s_idx;
  uint64_t* acc_full; uint64_t* a_ready; uint64_t* a_local; uint64_t* pro_ready; uint64_t* stat_bar;
  int rank, tile, row0, quarter, part, et, r, grow;
  bool live;
  int lcb0, gcb0;
  uint32_t trow;
  uint32_t peer_as[kSplit];
  uint32_t stat_uses;
  int ln_used;
  float* net_r;        // tile-layout bases of this thread's row: float4 group q at net_r + q * (kRows * 4)
  float* n32_r;
  T* gate_r;           // half8 chunk c at gate_r + c * (kRows * 8)

  __device__ __forceinline__ Epi(const GruProg<T>& P_) : P(P_) {}

  __device__ __forceinline__ static const uint4* f4(const float* base, int q) { return reinterpret_cast<const uint4*>(base + (size_t)q * (kRows * 4)); }
  __device__ __forceinline__ static float4* f4w(float* base, int q) { return reinterpret_cast<float4*>(base + (size_t)q * (kRows * 4)); }

  // a 16-byte chunk of the next layer's A operand: written into the LOCAL tile only; the finished column slice is
  // pushed to the peer CTAs by bulk DSMEM copies at the end of the layer (deliver_slice)
  __device__ __forceinline__ void a_store_all(int c, uint4 v) {
    *reinterpret_cast<uint4*>(As + a_off(r, c)) = v;
  }
  // end of an epilogue that rewrote this CTA's slice of the A tile (K-blocks rank*kKBc .. +kKBc): one thread arms the
  // local a_ready barrier for the bytes the peers will send and pushes the own slice into every peer's tile with
  // cp.async.bulk (shared::cta -> shared::cluster), which completes the bytes on the PEER's a_ready barrier.
  // The MMA warp of a CTA therefore starts the next layer when its own epilogue has arrived and every slice has landed.
  __device__ __forceinline__ void deliver_slice(bool wrote_a, int l) {
    constexpr int kKBc = kNC / 64;                 // K-blocks per CTA slice (3 at kSplit = 2)
    constexpr uint32_t kSliceBytes = kKBc * kABlk;
    if (wrote_a) fence_proxy_async();              // generic-proxy writes of the tile -> async proxy (bulk copy, UMMA)
    epi_bar();
    if (et == 0) {
      uint64_t* a_ready_l = a_ready + (l & 1);
      mbar_arrive(a_local + (l & 1));              // the MMA warp may start on the K-blocks of the local slice right away
      if (wrote_a && kSplit > 1) {
        mbar_arrive_expect_tx(a_ready_l, (kSplit - 1) * kSliceBytes);
        const uint32_t src = smem_u32(As) + (uint32_t)rank * kSliceBytes;
#pragma unroll
        for (int p = 0; p < kSplit; p++) {
          if (p == rank) continue;
          const uint32_t bar = mapa(smem_u32(a_ready_l), (uint32_t)p);
#pragma unroll
          for (int j = 0; j < kKBc; j++)
            bulk_copy_to_peer(peer_as[p] + (uint32_t)rank * kSliceBytes + j * kABlk, src + j * kABlk, kABlk, bar);
        }
      } else {
        mbar_arrive(a_ready_l);
      }
    }
  }
  // exchange per-row partials (N floats at s_buf[rank][part][r][*]) between all epilogue threads of the cluster
  template <int N>
  __device__ __forceinline__ void exchange(float* s_buf, const float* vals) {
    float* mine = s_buf + ((rank * kEpiPer + part) * kRows + r) * N;
#pragma unroll
    for (int k = 0; k < N; k++) mine[k] = vals[k];
    if (kSplit > 1) {
      // this CTA's block of partials ([kEpiPer][128][N] floats, contiguous) goes to every peer as ONE bulk DSMEM copy
      // that completes on the peer's stat_bar; the local arrival carries the bytes expected from the peers
      constexpr uint32_t kBlockBytes = kEpiPer * kRows * N * 4;
      fence_proxy_async();
      epi_bar();
      if (et == 0) {
        mbar_arrive_expect_tx(stat_bar, (kSplit - 1) * kBlockBytes);
        const uint32_t src = smem_u32(s_buf) + (uint32_t)rank * kBlockBytes;
#pragma unroll
        for (int p = 0; p < kSplit; p++) {
          if (p == rank) continue;
          bulk_copy_to_peer(mapa(src, (uint32_t)p), src, kBlockBytes, mapa(smem_u32(stat_bar), (uint32_t)p));
        }
      }
      mbar_wait(stat_bar, stat_uses & 1u);
      stat_uses++;
    } else {
      epi_bar();
    }
  }
  __device__ __forceinline__ void ln_stats(float s1, float s2, float& mean, float& rstd) {
    const float v[2] = {s1, s2};
    exchange<2>(s_stat, v);
    s1 = 0.f; s2 = 0.f;
#pragma unroll
    for (int p = 0; p < kSplit * kEpiPer; p++) { s1 += s_stat[(p * kRows + r) * 2 + 0]; s2 += s_stat[(p * kRows + r) * 2 + 1]; }
    mean = s1 * (1.0f / kD);
    rstd = rsqrtf(fmaxf(s2 * (1.0f / kD) - mean * mean, 0.f) + P.eps);
  }

  // ---------------- prologue: every CTA builds the full [128 x 384] A tile itself (no exchange) ----------------
  template <int PRO>
  __device__ __forceinline__ void prologue() {
    if constexpr (PRO == PRO_GATHER) {
      if (et < kRows) {
        const int gr = row0 + et;
        int src = -1;
        if (gr < P.rows) {
          const long long j = P.idx64 ? (long long)P.idx64[gr] : (P.idx32 ? (long long)P.idx32[gr] : (long long)gr);
          src = (j >= 0 && j < (long long)P.src_rows) ? (int)j : -1;
        }
        s_idx[et] = src;
      }
      epi_bar();
      // cooperative: 48 consecutive threads copy one 768-byte source row; kBatch independent 16-byte loads in flight
      constexpr int kPer = kRows * kChunks / kEpiThreads;     // chunks per thread
      constexpr int kBatch = (kPer % 12 == 0) ? 12 : 8;
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
    } else {
      // PRO_CAST: A = half(net32), full rows, redundantly in every CTA of the cluster (nothing is written back: a peer
      // reads the same net32 columns concurrently)
      constexpr int kPcb = kCB / kEpiPer;
      const int pcb0 = part * kPcb;
#pragma unroll 1
      for (int i = 0; i < kPcb; i++) {
        const int cb = pcb0 + i;
        uint4 f[8];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int c = cb * 4 + j;
          f[2 * j] = *f4(net_r, 2 * c);
          f[2 * j + 1] = *f4(net_r, 2 * c + 1);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int c = cb * 4 + j;
          const float4 a = as_f4(f[2 * j]), b = as_f4(f[2 * j + 1]);
          const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
          *reinterpret_cast<uint4*>(As + a_off(r, c)) = pack8<T>(v);
        }
      }
    }
    fence_proxy_async();
    epi_bar();
    if (et == 0) mbar_arrive(pro_ready);
  }

  // ---------------- one layer's epilogue over this thread's columns ----------------
  // Compact loops on purpose: one 16-byte chunk (8 columns) per iteration, not unrolled.  The fully unrolled version
  // was ~2500 straight-line instructions per gated layer and stalled on instruction fetch (ncu: 35 % "no_instructions").
  template <int EPI>
  __device__ __forceinline__ void layer(int l) {
    constexpr bool kGated = (EPI == EPI_GATED_LN || EPI == EPI_GATED_HEADS);
    constexpr bool kResid = (EPI == EPI_RESID || EPI == EPI_RESID_A || EPI == EPI_RESID_LN_A);
    constexpr int kAux = (kResid || EPI == EPI_ADD3_LN) ? 2 : (kGated ? 3 : 1);
    constexpr int kIter = kCBp * 4;                 // 16-byte chunks per thread
    const float* bias = s_bias + l * kD;
    float s1 = 0.f, s2 = 0.f;
    float hacc[4] = {0.f, 0.f, 0.f, 0.f};
    const T* netrow = nullptr;
    const T* inprow = nullptr;
    if (EPI == EPI_ADD3_LN && live) {
      netrow = P.x16_in + (size_t)grow * kD;
      inprow = P.inp16 + (size_t)P.kk[grow] * kD;
    }
    T* orow = nullptr;
    T* obase = nullptr;
    if (EPI == EPI_STORE_A || EPI == EPI_RESID || EPI == EPI_GATED_HEADS || EPI == EPI_ADD3_LN) obase = P.out16_a;
    if (EPI == EPI_STORE_B) obase = P.out16_b;
    if (live && obase) orow = obase + (size_t)grow * kD;
    // Row-major [rows,384] outputs: one thread per row means a warp's 16-byte stores hit 32 different rows (32 sectors
    // per instruction, ~1.5 us per layer).  In the LAST layer of a launch the weight ring is idle (every stage has been
    // consumed), so the slice is staged there ([128 rows][kNC/8 + 1 chunks], padded against bank conflicts) and copied
    // out with fully coalesced stores after the loop.
    const bool staged = (obase != nullptr) && (l + 1 == P.n_layers) && (kStageBytes <= kWStages * kWStage);
    uint4* stage = reinterpret_cast<uint4*>(Ws);
    auto out_store = [&](int c, uint4 v) {          // c: global 16-byte chunk index
      if (staged) stage[r * kStageStride + (c - rank * (kNC / 8))] = v;
      else if (orow) *reinterpret_cast<uint4*>(orow + c * 8) = v;
    };
    const int c0 = gcb0 * 4;                        // first global chunk of this thread
    const uint32_t tcol = trow + (uint32_t)(l & 1) * kNC + lcb0 * 32;   // its first column of this layer's accumulator
    // per-row operands of the element-wise tail, fetched one chunk ahead of their use
    // RESID: q[0..1] net32 ; GATED: q[0..1] n32, q[2] gate ; ADD3: q[0] net16, q[1] inp16
    auto load_aux = [&](int c, uint4* q) {
      if constexpr (kResid) {
        q[0] = *f4(net_r, 2 * c);
        q[1] = *f4(net_r, 2 * c + 1);
      } else if constexpr (kGated) {
        q[0] = *f4(n32_r, 2 * c);
        q[1] = *f4(n32_r, 2 * c + 1);
        q[2] = *reinterpret_cast<const uint4*>(gate_r + (size_t)c * (kRows * 8));
      } else if constexpr (EPI == EPI_ADD3_LN) {
        q[0] = make_uint4(0u, 0u, 0u, 0u);
        q[1] = q[0];
        if (live) { q[0] = *reinterpret_cast<const uint4*>(netrow + c * 8); q[1] = __ldg(reinterpret_cast<const uint4*>(inprow + c * 8)); }
      }
    };
    // kU chunks per loop iteration.  kU = 2 (two independent instruction streams per thread) was measured SLOWER:
    // 168 registers with spills, gated layers 12 us instead of 8.8 us -- so one chunk per iteration it is.
    constexpr int kU = DEVO_GRU_CHUNKS_PER_ITER;
    static_assert(kIter % kU == 0, "chunks per iteration");
    uint4 cur[kU][kAux], nxt[kU][kAux];
    if constexpr (kAux > 1) {                            // issued before the accumulator is ready: overlaps the MMAs
#pragma unroll
      for (int u = 0; u < kU; u++) load_aux(c0 + u, cur[u]);
    }
    mbar_wait(&acc_full[l & 1], (uint32_t)((l >> 1) & 1));   // every CTA of the cluster is done reading its A tile
    tc_fence_after();
    if (et == 0) stamp(P.dbg, 6 + 4 * l);
    // (prefetching the next accumulator chunk as well was measured slower: 168 instead of 126 registers, +7 %)
#pragma unroll 1
    for (int i0 = 0; i0 < kIter; i0 += kU) {
      uint32_t rawu[kU][8];
#pragma unroll
      for (int u = 0; u < kU; u++) tmem_ld8(tcol + (i0 + u) * 8, rawu[u]);
      if constexpr (kAux > 1) {
        if (i0 + kU < kIter) {
#pragma unroll
          for (int u = 0; u < kU; u++) load_aux(c0 + i0 + kU + u, nxt[u]);
        }
      }
      tmem_wait_ld();
#pragma unroll
      for (int u = 0; u < kU; u++) {
      const int i = i0 + u;
      const int c = c0 + i;                         // global 16-byte chunk index (8 columns)
      uint32_t (&raw)[8] = rawu[u];
      uint4 (&cur_)[kAux] = cur[u];
      float o[8];
      {
        const float4 b0 = *reinterpret_cast<const float4*>(bias + c * 8), b1 = *reinterpret_cast<const float4*>(bias + c * 8 + 4);
        o[0] = __uint_as_float(raw[0]) + b0.x; o[1] = __uint_as_float(raw[1]) + b0.y;
        o[2] = __uint_as_float(raw[2]) + b0.z; o[3] = __uint_as_float(raw[3]) + b0.w;
        o[4] = __uint_as_float(raw[4]) + b1.x; o[5] = __uint_as_float(raw[5]) + b1.y;
        o[6] = __uint_as_float(raw[6]) + b1.z; o[7] = __uint_as_float(raw[7]) + b1.w;
      }
      const uint4 oh = pack8<T>(o);                 // the Linear output, rounded to half (autocast)
      if constexpr (EPI == EPI_RELU_A) {
        a_store_all(c, relu8<T>(oh));               // max(.,0) commutes with the rounding: done on packed halves
      } else if constexpr (EPI == EPI_STORE_A || EPI == EPI_STORE_B) {
        out_store(c, oh);
      } else if constexpr (EPI == EPI_GATE) {
        *reinterpret_cast<uint4*>(gate_r + (size_t)c * (kRows * 8)) = oh;
      } else if constexpr (EPI == EPI_LNRELU_A) {
        unpack8<T>(oh, o);
#pragma unroll
        for (int k = 0; k < 8; k++) { s1 += o[k]; s2 += o[k] * o[k]; }
        *reinterpret_cast<uint4*>(As + a_off(r, c)) = oh;                     // parked in the local tile until normalised
      } else if constexpr (EPI == EPI_ADD3_LN) {
        float a[8], b[8];
        unpack8<T>(oh, o);
        unpack8<T>(cur_[0], a);
        unpack8<T>(cur_[1], b);
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] += b[k];
        rnd8<T>(a);
#pragma unroll
        for (int k = 0; k < 8; k++) o[k] += a[k];
        const uint4 vh = pack8<T>(o);
        unpack8<T>(vh, o);
#pragma unroll
        for (int k = 0; k < 8; k++) { s1 += o[k]; s2 += o[k] * o[k]; }
        *reinterpret_cast<uint4*>(As + a_off(r, c)) = vh;
      } else if constexpr (EPI == EPI_RESID) {
        unpack8<T>(oh, o);
        float4 a = as_f4(cur_[0]), b = as_f4(cur_[1]);
        a.x += o[0]; a.y += o[1]; a.z += o[2]; a.w += o[3]; b.x += o[4]; b.y += o[5]; b.z += o[6]; b.w += o[7];
        *f4w(net_r, 2 * c) = a;
        *f4w(net_r, 2 * c + 1) = b;
        if (obase) {
          const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
          out_store(c, pack8<T>(v));
        }
      } else if constexpr (EPI == EPI_RESID_A || EPI == EPI_RESID_LN_A) {
        // net += half Linear output (the SoftAgg `h` layer applied per edge: h(y)[:, gid] == h(y[:, gid]), row by row
        // bit-identical to the per-group product), then the sum becomes the next A operand (optionally after LayerNorm)
        unpack8<T>(oh, o);
        const float4 a = as_f4(cur_[0]), b = as_f4(cur_[1]);
        float x[8] = {a.x + o[0], a.y + o[1], a.z + o[2], a.w + o[3], b.x + o[4], b.y + o[5], b.z + o[6], b.w + o[7]};
        if constexpr (EPI == EPI_RESID_A) {
          *f4w(net_r, 2 * c) = make_float4(x[0], x[1], x[2], x[3]);      // own column slice only: nobody else reads it here
          *f4w(net_r, 2 * c + 1) = make_float4(x[4], x[5], x[6], x[7]);
          a_store_all(c, pack8<T>(x));
        } else {
          uint32_t xs[8];
#pragma unroll
          for (int k = 0; k < 8; k++) { s1 += x[k]; s2 += x[k] * x[k]; xs[k] = __float_as_uint(x[k]); }
          tmem_st8(tcol + i * 8, xs);
        }
      } else {   // EPI_GATED_LN / EPI_GATED_HEADS:  x = n + half(half(sigmoid(gate)) * res)
        float g[8];
        unpack8<T>(oh, o);
        unpack8<T>(cur_[2], g);
        const float4 a = as_f4(cur_[0]), b = as_f4(cur_[1]);
        float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; k++) g[k] = sigmoidf_(g[k]);
        rnd8<T>(g);
#pragma unroll
        for (int k = 0; k < 8; k++) g[k] *= o[k];
        rnd8<T>(g);
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] += g[k];
        if constexpr (EPI == EPI_GATED_LN) {
#pragma unroll
          uint32_t xs[8];
#pragma unroll
          for (int k = 0; k < 8; k++) { s1 += x[k]; s2 += x[k] * x[k]; xs[k] = __float_as_uint(x[k]); }
          tmem_st8(tcol + i * 8, xs);               // fp32 row parked in its own accumulator columns
        } else {
          out_store(c, pack8<T>(x));                  // new hidden state (half)
          float hw[8];
#pragma unroll
          for (int k = 0; k < 8; k++) x[k] = fmaxf(x[k], 0.f);
          rnd8<T>(x);
#pragma unroll
          for (int o4 = 0; o4 < 4; o4++) {
            unpack8<T>(*reinterpret_cast<const uint4*>(s_head + o4 * kD + c * 8), hw);
#pragma unroll
            for (int k = 0; k < 8; k++) hacc[o4] += x[k] * hw[k];
          }
        }
      }
      }
      if constexpr (kAux > 1) {
#pragma unroll
        for (int u = 0; u < kU; u++)
#pragma unroll
          for (int k = 0; k < kAux; k++) cur[u][k] = nxt[u][k];
      }
    }
    if (et == 0 && l == 0) stamp(P.dbg, 40);
    // ---------------- row-wise tails ----------------
    if constexpr (EPI == EPI_LNRELU_A || EPI == EPI_ADD3_LN) {
      // this thread's slice of the row (half values) sits in the local A tile: one more pass over shared memory
      float mean, rstd;
      ln_stats(s1, s2, mean, rstd);
      const float* gm = s_ln + ln_used * 2 * kD;
      const float* bt = gm + kD;
#pragma unroll 1
      for (int i = 0; i < kIter; i++) {
        const int c = c0 + i;
        float v[8];
        unpack8<T>(*reinterpret_cast<const uint4*>(As + a_off(r, c)), v);
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (v[k] - mean) * rstd * gm[c * 8 + k] + bt[c * 8 + k];
        if constexpr (EPI == EPI_LNRELU_A) {
#pragma unroll
          for (int k = 0; k < 8; k++) v[k] = fmaxf(v[k], 0.f);
          a_store_all(c, pack8<T>(v));
        } else {
          *f4w(net_r, 2 * c) = make_float4(v[0], v[1], v[2], v[3]);
          *f4w(net_r, 2 * c + 1) = make_float4(v[4], v[5], v[6], v[7]);
          out_store(c, pack8<T>(v));
        }
      }
      ln_used++;
    } else if constexpr (EPI == EPI_GATED_LN || EPI == EPI_RESID_LN_A) {
      tmem_wait_st();
      float mean, rstd;
      ln_stats(s1, s2, mean, rstd);
      const float* gm = s_ln + ln_used * 2 * kD;
      const float* bt = gm + kD;
#pragma unroll 1
      for (int i = 0; i < kIter; i++) {
        const int c = c0 + i;
        uint32_t raw[8];
        tmem_ld8(tcol + i * 8, raw);
        tmem_wait_ld();
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (__uint_as_float(raw[k]) - mean) * rstd * gm[c * 8 + k] + bt[c * 8 + k];
        *f4w(n32_r, 2 * c) = make_float4(v[0], v[1], v[2], v[3]);
        *f4w(n32_r, 2 * c + 1) = make_float4(v[4], v[5], v[6], v[7]);
        a_store_all(c, pack8<T>(v));
      }
      ln_used++;
    } else if constexpr (EPI == EPI_GATED_HEADS) {
      exchange<4>(s_hacc, hacc);
      if (live && part == 0 && rank == 0) {
#pragma unroll
        for (int o4 = 0; o4 < 4; o4++) {
          hacc[o4] = 0.f;
#pragma unroll
          for (int p = 0; p < kSplit * kEpiPer; p++) hacc[o4] += s_hacc[(p * kRows + r) * 4 + o4];
        }
        const uint32_t d = pack2<T>(hacc[0] + ElemTraits<T>::to_float(s_head[4 * kD + 0]), hacc[1] + ElemTraits<T>::to_float(s_head[4 * kD + 1]));
        const float2 w = unpack2<T>(pack2<T>(hacc[2] + ElemTraits<T>::to_float(s_head[4 * kD + 2]), hacc[3] + ElemTraits<T>::to_float(s_head[4 * kD + 3])));
        *reinterpret_cast<uint32_t*>(P.delta + (size_t)grow * 2) = d;
        const uint32_t wh = pack2<T>(1.0f / (1.0f + expf(-w.x)), 1.0f / (1.0f + expf(-w.y)));
        *reinterpret_cast<uint32_t*>(P.weight + (size_t)grow * 2) = wh;
        if (P.target32 != nullptr) {            // BA inputs, fused: target = reprojected centre + delta (devo.py:326-331)
          const float2 df = unpack2<T>(d), wf = unpack2<T>(wh);
          const float* cr = P.coords + (size_t)grow * 18;
          *reinterpret_cast<float2*>(P.target32 + (size_t)grow * 2) = make_float2(cr[4] + df.x, cr[13] + df.y);
          *reinterpret_cast<float2*>(P.weight32 + (size_t)grow * 2) = wf;
        }
      }
    }
    if (staged) {                                   // coalesced copy-out of the staged slice (kNC/8 chunks per row)
      epi_bar();
      constexpr int kCpr = kNC / 8;
      T* dst = obase + (size_t)row0 * kD + rank * kNC;
      for (int q = et; q < kRows * kCpr; q += kEpiThreads) {
        const int rr = q / kCpr, ch = q - rr * kCpr;
        if (row0 + rr < P.rows) *reinterpret_cast<uint4*>(dst + (size_t)rr * kD + ch * 8) = stage[rr * kStageStride + ch];
      }
    }
    tc_fence_before();
    if (et == 0 && l == 0) stamp(P.dbg, 41);
    if (l + 1 < P.n_layers) {       // hand the A tiles / the TMEM accumulator back to the MMA warps of the cluster
      constexpr bool kWritesA = (EPI == EPI_RELU_A || EPI == EPI_LNRELU_A || EPI == EPI_GATED_LN || EPI == EPI_RESID_A ||
                                 EPI == EPI_RESID_LN_A);
      deliver_slice(kWritesA, l);
    }
    if (et == 0) stamp(P.dbg, 7 + 4 * l);
  }
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) gru_mma_kernel(const __grid_constant__ CUtensorMap tm_w,
                                                              const __grid_constant__ CUtensorMap tm_w0,
                                                              const __grid_constant__ CUtensorMap tm_a,
                                                              const __grid_constant__ GruProg<T> P) {
  extern __shared__ unsigned char smem_dyn[];
  // identical carve-up in every CTA of the cluster (mapa addresses the same offset in a peer)
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* As = base;                                   // kASlots x 16 KB
  unsigned char* Ws = base + kASlots * kABlk;                 // kWStages x kWStage
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ws + kWStages * kWStage);
  uint64_t* w_full = bars;                    // [kWStages]
  uint64_t* w_empty = w_full + kWStages;      // [kWStages]
  uint64_t* a_full = w_empty + kWStages;      // [kASlots]
  uint64_t* a_empty = a_full + kASlots;       // [kASlots]
  uint64_t* acc_full = a_empty + kASlots;     // [2] (one per TMEM accumulator) count kSplit: the MMAs of a layer are done in EVERY CTA of the cluster
  // a_ready / a_local exist twice: the epilogue of layer q signals barrier q & 1, so each barrier only sees every other
  // hand-off.  The MMA warp may legitimately run two hand-offs behind (it skips waiting for an epilogue that leaves the
  // A tile alone); on a single phase-parity barrier a two-phase lag is indistinguishable from "not yet" and would hang.
  uint64_t* a_ready = acc_full + 2;           // [2] own epilogue done + the peers' slices of the next A have landed (tx bytes)
  uint64_t* pro_ready = a_ready + 2;          // [1] local prologue finished
  uint64_t* a_local = pro_ready + 2;          // [2] own epilogue done: this CTA's slice of the next A is in place, accumulator free
  uint64_t* stat_bar = pro_ready + 1;         // [1] count kSplit: LayerNorm / head partials of all CTAs have arrived
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);      // 32 barrier slots = 256 B; slot + pad = 16 B
  int* s_idx = reinterpret_cast<int*>(tmem_slot + 4);                 // [128] gather sources of the tile
  float* s_stat = reinterpret_cast<float*>(s_idx + kRows);            // [kSplit][kEpiPer][128][2] LayerNorm partials
  float* s_hacc = s_stat;                                             // [kSplit][kEpiPer][128][4] head partials (last layer only: same storage)
  float* s_ln = s_stat + kSplit * kEpiPer * kRows * 4;                // [2][2][384]: gamma, beta of the program's LayerNorms
  float* s_bias = s_ln + 4 * kD;                                      // [kMaxLayers][384] biases as fp32
  T* s_head = reinterpret_cast<T*>(s_bias + kMaxLayers * kD);         // [4][384] + [4] (+4 pad)

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();    // which column slice of the layer outputs this CTA computes
  const int tile = blockIdx.x / kSplit;
  const int row0 = tile * kRows;
  const int col0 = rank * kNC;

  if (threadIdx.x == 0) stamp(P.dbg, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWStages; s++) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < kASlots; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(&acc_full[0], kSplit);
    mbar_init(&acc_full[1], kSplit);
    mbar_init(&a_ready[0], 1);                     // local epilogue arrival (+ the bytes of the peers' slices)
    mbar_init(&a_ready[1], 1);
    mbar_init(pro_ready, 1);
    mbar_init(&a_local[0], 1);
    mbar_init(&a_local[1], 1);
    mbar_init(stat_bar, 1);                        // local arrival (+ the bytes of the peers' partials)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else if (warp >= 2) {
    // stage the program's small parameters (biases, LayerNorm affine, heads) in shared memory once: the epilogues
    // read them as broadcasts instead of dependent global loads
    const int t = threadIdx.x - 32 * kFirstEpiWarp, nt = kThreads - 32 * kFirstEpiWarp;
    for (int l = 0; l < P.n_layers; l++)
      for (int q = t; q < kD / 8; q += nt) {
        float bv[8];
        unpack8<T>(__ldg(reinterpret_cast<const uint4*>(P.bias[l]) + q), bv);
        reinterpret_cast<float4*>(s_bias + l * kD)[2 * q] = make_float4(bv[0], bv[1], bv[2], bv[3]);
        reinterpret_cast<float4*>(s_bias + l * kD)[2 * q + 1] = make_float4(bv[4], bv[5], bv[6], bv[7]);
      }
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
  if (kSplit > 1) cluster_sync_all();         // peers' barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool has_pro = (P.pro != PRO_NONE);
  if (threadIdx.x == 0) stamp(P.dbg, 1);
  // PDL: everything above (barriers, TMEM, parameters) and the weight stream below only touch constants, so this grid
  // may start while its predecessor still runs; the next grid of the stream may start now as well.
  if (threadIdx.x == 0) pdl_launch_dependents();

  if (warp == 0) {
    // =========================== TMA producer: this CTA's weight slice of every layer (+ the streamed A of layer 0)
    if (lane == 0) { prefetch_tensormap(&tm_w); if (P.use_w0) prefetch_tensormap(&tm_w0); if (P.stream_a0) prefetch_tensormap(&tm_a); }
    if (P.stream_a0) pdl_wait();                    // the streamed A operand is the previous kernel's output
    uint32_t stage = 0, phase = 0;
    for (int l = 0; l < P.n_layers; l++) {
      const int nkb = (l == 0) ? P.kblocks0 : kASlots;
      const CUtensorMap* wm = (l == 0 && P.use_w0) ? &tm_w0 : &tm_w;
      const int wrow = P.w_row[l] + col0;
      for (int kb = 0; kb < nkb; kb++) {
        if (l == 0 && P.stream_a0) {
          const int slot = kb % kASlots, use = kb / kASlots;
          mbar_wait(&a_empty[slot], (uint32_t)(use & 1) ^ 1u);
          mbar_arrive_expect_tx_elect(smem_u32(&a_full[slot]), kABlk);
          tma_load_2d_elect(smem_u32(As) + slot * kABlk, &tm_a, smem_u32(&a_full[slot]), kb * 64, row0);
        }
        mbar_wait(&w_empty[stage], phase ^ 1u);
        mbar_arrive_expect_tx_elect(smem_u32(&w_full[stage]), kWStage);
        // layers > 0 consume the K-blocks of the LOCAL slice first (they are ready before the peer's slice has landed)
        const int kbw = (l == 0) ? kb : (rank * (kNC / 64) + kb) % kASlots;
        tma_load_2d_elect(smem_u32(Ws) + stage * kWStage, wm, smem_u32(&w_full[stage]), kbw * 64, wrow);
        if (++stage == kWStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ====================================================================
    const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
    const uint32_t idesc = umma_idesc_f16(fmt, kNC);
    const uint64_t ad0 = umma_desc_sw128(smem_u32(As));
    const uint64_t bd0 = umma_desc_sw128(smem_u32(Ws));
    const uint16_t all = (uint16_t)((1u << kSplit) - 1u);
    uint32_t stage = 0, phase = 0, waited = 0, waited_r = 0;
    for (int l = 0; l < P.n_layers; l++) {
      const int nkb = (l == 0) ? P.kblocks0 : kASlots;
      const bool streamed = (l == 0 && P.stream_a0);
      // Two TMEM accumulators (layer l uses l & 1): when the previous epilogue does not rewrite the A tile (gate, g/f
      // stores) this layer's MMAs run WHILE that epilogue still reads the other accumulator.  `waited` counts the
      // epilogues (a_ready phases) consumed so far; phases are never skipped, only deferred.
      if (l == 0) {
        if (has_pro) { mbar_wait(pro_ready, 0u); fence_proxy_async(); tc_fence_after(); }
      } else {
        const int pe = P.epi[l - 1];
        const bool writes_a = (pe == EPI_RELU_A || pe == EPI_LNRELU_A || pe == EPI_GATED_LN || pe == EPI_RESID_A || pe == EPI_RESID_LN_A);
        const int need = writes_a ? l - 1 : l - 2;     // last epilogue that must be complete: A operand / accumulator reuse
        while ((int)waited <= need) {
          mbar_wait(&a_local[waited & 1u], (waited >> 1) & 1u);   // own epilogue arrived: local slice in place, accumulator free
          waited++;
        }
        fence_proxy_async();
        tc_fence_after();
      }
      const uint32_t tacc = tmem_base + (uint32_t)(l & 1) * kNC;
      if (lane == 0) stamp(P.dbg, 4 + 4 * l);
      for (int kb = 0; kb < nkb; kb++) {
        if (l > 0 && kb == kNC / 64) {                 // the remaining K-blocks belong to the peers' slices
          const int pe = P.epi[l - 1];
          const bool writes_a = (pe == EPI_RELU_A || pe == EPI_LNRELU_A || pe == EPI_GATED_LN || pe == EPI_RESID_A || pe == EPI_RESID_LN_A);
          const int need = writes_a ? l - 1 : l - 2;
          while ((int)waited_r <= need) {
            mbar_wait(&a_ready[waited_r & 1u], (waited_r >> 1) & 1u);   // the peers' slices (bulk copies) have landed
            waited_r++;
          }
          tc_fence_after();
        }
        const int slot = (l == 0) ? kb % kASlots : (rank * (kNC / 64) + kb) % kASlots;
        if (streamed) { mbar_wait(&a_full[slot], (uint32_t)((kb / kASlots) & 1)); tc_fence_after(); }
        const uint64_t ad = ad0 + (uint64_t)(slot * (kABlk >> 4));
        mbar_wait(&w_full[stage], phase);
        tc_fence_after();
        const uint64_t bd = bd0 + (uint64_t)(stage * (kWStage >> 4));
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++)
          tc_mma_f16_elect(tacc, ad + 2 * k4, bd + 2 * k4, idesc, (kb > 0 || k4 > 0) ? 1u : 0u);
        tc_commit_elect(smem_u32(&w_empty[stage]));
        if (++stage == kWStages) { stage = 0; phase ^= 1u; }
        if (streamed) tc_commit_elect(smem_u32(&a_empty[slot]));
      }
      if (kSplit > 1) tc_commit_mc_elect(smem_u32(&acc_full[l & 1]), all);
      else tc_commit_elect(smem_u32(&acc_full[l & 1]));
      if (lane == 0) stamp(P.dbg, 5 + 4 * l);
    }
    __syncwarp();
  } else if (warp >= kFirstEpiWarp) {
    // =========================== prologue + epilogues (struct Epi) =============================================
    Epi<T> e(P);
    e.As = As; e.Ws = Ws; e.s_stat = s_stat; e.s_hacc = s_hacc; e.s_ln = s_ln; e.s_bias = s_bias; e.s_head = s_head; e.s_idx = s_idx;
    e.acc_full = acc_full; e.a_ready = a_ready; e.a_local = a_local; e.pro_ready = pro_ready; e.stat_bar = stat_bar;
    e.rank = rank; e.tile = tile; e.row0 = row0;
    e.quarter = warp & 3;
    e.part = (warp - kFirstEpiWarp) >> 2;           // 0 .. kEpiPer-1
    e.et = threadIdx.x - 32 * kFirstEpiWarp;        // 0 .. kEpiThreads-1
    e.r = e.quarter * 32 + lane;                    // tile row = TMEM lane
    e.grow = row0 + e.r;
    e.live = e.grow < P.rows;
    e.lcb0 = e.part * kCBp;                         // first local column block of this thread
    e.gcb0 = rank * kCBc + e.lcb0;                  // ... as a global column block
    e.trow = tmem_base + ((uint32_t)(e.quarter * 32) << 16);
#pragma unroll
    for (int p = 0; p < kSplit; p++) e.peer_as[p] = mapa(smem_u32(As), (uint32_t)p);
    e.stat_uses = 0;
    e.ln_used = 0;
    e.net_r = P.net32 + t32(tile, 0, e.r);
    e.n32_r = P.n32 + t32(tile, 0, e.r);
    e.gate_r = P.gate16 + t16(tile, 0, e.r);
    pdl_wait();                                     // first use of the previous kernels' results (and first global writes)
    switch (P.pro) {                                // warp-uniform
      case PRO_GATHER: e.template prologue<PRO_GATHER>(); break;
      case PRO_CAST: e.template prologue<PRO_CAST>(); break;
      default: break;
    }
    if (e.et == 0) stamp(P.dbg, 2);
    for (int l = 0; l < P.n_layers; l++) {
      switch (P.epi[l]) {
        case EPI_RELU_A: e.template layer<EPI_RELU_A>(l); break;
        case EPI_LNRELU_A: e.template layer<EPI_LNRELU_A>(l); break;
        case EPI_ADD3_LN: e.template layer<EPI_ADD3_LN>(l); break;
        case EPI_RESID: e.template layer<EPI_RESID>(l); break;
        case EPI_STORE_A: e.template layer<EPI_STORE_A>(l); break;
        case EPI_STORE_B: e.template layer<EPI_STORE_B>(l); break;
        case EPI_GATE: e.template layer<EPI_GATE>(l); break;
        case EPI_GATED_LN: e.template layer<EPI_GATED_LN>(l); break;
        case EPI_RESID_A: e.template layer<EPI_RESID_A>(l); break;
        case EPI_RESID_LN_A: e.template layer<EPI_RESID_LN_A>(l); break;
        default: e.template layer<EPI_GATED_HEADS>(l); break;
      }
    }
  }
  // teardown: no CTA may exit while a peer can still write into its shared memory
  tc_fence_before();
  __syncthreads();
  if (kSplit > 1) cluster_sync_all();
  if (threadIdx.x == 0) stamp(P.dbg, 3);
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

constexpr size_t kSmemBytes = 1024 + (size_t)kASlots * kABlk + (size_t)kWStages * kWStage + 32 * sizeof(uint64_t) + 16 + kRows * sizeof(int) +
                              (size_t)kSplit * kEpiPer * kRows * 4 * sizeof(float) + 4 * kD * sizeof(float) + kMaxLayers * kD * sizeof(float) + (4 * kD + 8) * 2 + 64;
static_assert(kSmemBytes <= 232448, "shared memory budget");

static long long* g_dbg = nullptr;     // 16 launches x 48 stamps, allocated when DEVO_GRU_TIMING is set
static int g_dbg_launch = 0;

template <typename T>
static int launch_prog(const CUtensorMap& tw, const CUtensorMap& tw0, const CUtensorMap& ta, const GruProg<T>& P, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    DEVO_CUDA(cudaFuncSetAttribute(gru_mma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    configured = true;
  }
  const int tiles = (P.rows + kRows - 1) / kRows;
  if (tiles <= 0) return DEVO_OK;
  GruProg<T> Pd = P;
  Pd.dbg = g_dbg ? g_dbg + 48 * (g_dbg_launch++ % 16) : nullptr;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(tiles * kSplit), 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;      // kSplit CTAs share one 128-row tile
  attr[0].val.clusterDim.x = kSplit; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: overlap the set-up with the previous kernel's tail
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static int use_pdl = -1;
  if (use_pdl < 0) { const char* e = getenv("DEVO_GRU_PDL"); use_pdl = (e && e[0] == '0') ? 0 : 1; }
  cfg.numAttrs = use_pdl ? 2 : 1;
  DEVO_CUDA(cudaLaunchKernelEx(&cfg, gru_mma_kernel<T>, tw, tw0, ta, Pd));
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
  int rc = make_map_2d(&tw, dtype, Wt->W, 18 * kD, kD, kD, kNC);
  if (rc != DEVO_OK) return rc;
  rc = make_map_2d(&tw0, dtype, Wt->W0, kD, (uint64_t)io->corr_ld, (uint64_t)io->corr_ld, kNC);
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
  // (4,5,6) net += SoftAgg(net) over patches, then over frame pairs; gru + heads  (enet.py:93-99, blocks.py:40-48).
  // The `h` layer of a SoftAgg is applied per EDGE inside the kernel that consumes it (h(y)[:, gid] == h(y[:, gid])):
  // its A operand is the gathered group row y[gid[e]], its epilogue adds the result to the fp32 state.
  {
    GruProg<T> P = base;                                      // g, f of the patch-wise aggregation
    P.n_layers = 2; P.pro = PRO_CAST;
    P.w_row[0] = 6 * kD; P.epi[0] = EPI_STORE_A; P.bias[0] = B(6);
    P.w_row[1] = 7 * kD; P.epi[1] = EPI_STORE_B; P.bias[1] = B(7);
    P.out16_a = g16; P.out16_b = f16;
    rc = launch_prog<T>(tw, tw0, ta, P, s);
    if (rc != DEVO_OK) return rc;
  }
  rc = devo_segment_softmax_sum(g16, f16, io->perm_kk, io->gstart_kk, io->ngroups_kk, io->max_groups_kk, y16, dtype, E, kD, (void*)s);
  if (rc != DEVO_OK) return rc;
  {
    GruProg<T> P = base;                                      // net += h_kk(y_kk)[gid_kk]; g, f of the pair-wise aggregation
    P.n_layers = 3; P.pro = PRO_GATHER; P.x16_in = y16; P.idx32 = io->gid_kk; P.src_rows = io->max_groups_kk;
    P.w_row[0] = 8 * kD;  P.epi[0] = EPI_RESID_A; P.bias[0] = B(8);
    P.w_row[1] = 9 * kD;  P.epi[1] = EPI_STORE_A; P.bias[1] = B(9);
    P.w_row[2] = 10 * kD; P.epi[2] = EPI_STORE_B; P.bias[2] = B(10);
    P.out16_a = g16; P.out16_b = f16;
    rc = launch_prog<T>(tw, tw0, ta, P, s);
    if (rc != DEVO_OK) return rc;
  }
  rc = devo_segment_softmax_sum(g16, f16, io->perm_ij, io->gstart_ij, io->ngroups_ij, io->max_groups_ij, hy16, dtype, E, kD, (void*)s);
  if (rc != DEVO_OK) return rc;
  {
    GruProg<T> P = base;                                      // net += h_ij(y_ij)[gid_ij]; LN, GatedResidual x2; heads
    P.n_layers = 7; P.pro = PRO_GATHER; P.x16_in = hy16; P.idx32 = io->gid_ij; P.src_rows = io->max_groups_ij;
    P.ln_g[0] = Wt->ln_gamma + 2 * kD; P.ln_b[0] = Wt->ln_beta + 2 * kD;
    P.ln_g[1] = Wt->ln_gamma + 3 * kD; P.ln_b[1] = Wt->ln_beta + 3 * kD;
    P.w_row[0] = 11 * kD; P.epi[0] = EPI_RESID_LN_A; P.bias[0] = B(11);
    const int epis[6] = {EPI_GATE, EPI_RELU_A, EPI_GATED_LN, EPI_GATE, EPI_RELU_A, EPI_GATED_HEADS};
    for (int l = 0; l < 6; l++) { P.w_row[1 + l] = (12 + l) * kD; P.epi[1 + l] = epis[l]; P.bias[1 + l] = B(12 + l); }
    P.out16_a = (T*)io->net16_out;
    P.headW = (const T*)Wt->head_W; P.headB = (const T*)Wt->head_b;
    P.delta = (T*)io->delta; P.weight = (T*)io->weight;
    if (io->coords && io->target32 && io->weight32) { P.coords = io->coords; P.target32 = io->target32; P.weight32 = io->weight32; }
    rc = launch_prog<T>(tw, tw0, ta, P, s);
    if (rc != DEVO_OK) return rc;
  }
  return DEVO_OK;
}

}  // namespace

extern "C" {

size_t devo_gru_workspace(int E, int max_groups) { return gru_ws(E, max_groups).total; }

// debug (tools/gru_timing.py): enable / read back the %globaltimer stamps of CTA 0 of the last 16 launches
int devo_gru_debug_timing(long long* host_out) {
  if (!g_dbg) {
    if (cudaMalloc(&g_dbg, 16 * 48 * sizeof(long long)) != cudaSuccess) return -1;
    cudaMemset(g_dbg, 0, 16 * 48 * sizeof(long long));
  }
  g_dbg_launch = 0;
  {
    int nclusters = -1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(48 * kSplit, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kSplit; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaFuncSetAttribute(gru_mma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, gru_mma_kernel<__half>, &cfg);
    fprintf(stderr, "gru_mma: split %d, max co-resident clusters %d (%s)\n", kSplit, nclusters, cudaGetErrorString(e));
  }
  if (host_out) return (int)cudaMemcpy(host_out, g_dbg, 16 * 48 * sizeof(long long), cudaMemcpyDeviceToHost);
  return 0;
}

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
