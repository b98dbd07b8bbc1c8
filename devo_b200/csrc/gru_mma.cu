// gru_mma.cu -- the recurrent update operator ("ConvGRU", devo/enet.py:32-99, devo/blocks.py:15-48) as a
// handful of fused tensor-core kernels.
//
// Reference: ~17 cuBLAS GEMMs of [E,384]x[384,384] plus ~40 ATen element-wise launches per iteration; every
// intermediate [E,384] tensor makes a round trip through HBM/L2.  Here a PAIR of CTAs (one cluster, two SMs of a TPC)
// owns a tile of 128 edges -- 64 WHOLE ROWS per CTA -- and walks it through a chain of layers without leaving the SMs:
//
//   MMA        : tcgen05.mma.cta_group::2.kind::f16, M = 128 (64 rows per CTA), N = 192 (two N-tiles per layer), K = 16,
//                issued by elected lanes of the LEADER CTA for the pair -- ONE ISSUING WARP PER N-TILE, each with its own
//                weight ring: a 57-cycle MMA is shorter than the ~75-130 cycles one thread needs to wait for a stage,
//                build descriptors and issue it (measured: a single issuer ran the tensor pipe at 40 %).  Measured (tools/umma_probe.cu,
//                profiles/r02_umma_probe.txt): 48 cycles per instruction = 4092 MAC/cycle/SM, i.e. full rate, while
//                each CTA stages only HALF of every weight block (the tensor cores read the B operand from both CTAs'
//                shared memory).  Every output row is local to one CTA: no activation exchange between the CTAs,
//                LayerNorm / head reductions never leave the CTA.  (Round 1 split the N dimension over the two CTAs with
//                cta_group::1 MMAs and had to ship a 48 KB activation slice + LayerNorm partials through DSMEM per layer.)
//   A operand  : the CTA's activations [64 rows x 384] f16, K-major SWIZZLE_128B, DOUBLE BUFFERED in shared memory: the
//                epilogue of layer l writes the A operand of layer l+1 into the other buffer while the MMAs of layer l
//                may still read theirs.  For the first Linear of the corr MLP (K = 896 padded) the K-blocks are streamed by
//                TMA through the same 12 slots.
//   B operand  : weights W[384 out, K in] (K-major as stored by nn.Linear), streamed from L2 by TMA in [96 x 64] boxes
//                per CTA through an 8-stage mbarrier ring; the producer warps run ahead across layers.
//   accumulator: TMEM, 2 sets (layer parity) x 2 N-tiles x 96 columns.  cta_group::2 M = 128 layout: lanes 0-63 hold
//                rows 0-63 x columns n in [0,96) of the N-tile, lanes 64-127 the same rows x n in [96,192).
//   epilogue   : 16 warps; thread = (row, 24-column slice of each N-tile): tcgen05.ld, bias, the layer's element-wise tail
//                (ReLU / LayerNorm / residual / gate / heads) in registers, then the next A operand (or a staging image
//                that a TMA store writes to global memory row-major).  Each warp signals "my part of N-tile h is in
//                place" on an mbarrier of the leader; the MMA warp issues layer l+1 in WAVEFRONT order -- the K-blocks fed
//                by N-tile 0 for both output tiles first -- so the tensor pipe runs under the epilogue of the other half.
//   launches   : programmatic dependent launch -- barrier / TMEM / parameter set-up and the weight prefetch of a
//                launch overlap the tail of its predecessor (griddepcontrol).
//
// Kernel boundaries remain only where rows of different tiles meet: the neighbour gathers (net[ix], net[jx]) and
// the two SoftAgg segment reductions (their `h` layers are applied per edge inside the consuming kernel).  7 launches per
// update (5 of this kernel + 2 segment reductions) instead of ~60.
//
// Rounding points follow torch.autocast exactly as the reference's module forward does (Linear outputs are rounded to
// half, LayerNorm in float32, element-wise ops round to their promoted type); accumulation is fp32, only the summation
// order inside a dot product differs from cuBLAS.  The recurrent hidden state is FLOAT32 (GatedResidual returns float32
// under autocast and devo.py:232-233 concatenates half zeros onto it, so every update after the first sees a float32
// state); a half state (the very first update) is accepted as well and follows the half dtype flow.
//
// fp32 per-row state (state32 / n32) and the gate scratch use a tile layout [tile64][col/4][64 rows][4]
// (resp. [tile64][col/8][64][8] halfs) so that "one thread per row" accesses are fully coalesced.
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

constexpr int kRows = 64;                  // rows (edges) per CTA; the pair's MMA has M = 128
constexpr int kD = 384;                    // hidden width: N of every layer, K of all but the first
constexpr int kNT = 192;                   // N per MMA instruction; 2 N-tiles per layer
constexpr int kNTc = kNT / 2;              // TMEM columns per N-tile (cta_group::2, M = 128: n / 96 selects the lane half)
constexpr int kABlk = kRows * 128;         // one K-block of A: 64 rows x 64 halfs = 8 KB
constexpr int kKB = kD / 64;               // 6 K-blocks
constexpr int kABuf = kKB * kABlk;         // 48 KB
constexpr int kASlots = 2 * kKB;           // both buffers as a ring for the streamed first layer
constexpr int kWStage = kNTc * 128;        // this CTA's half of a [192 x 64] weight block: 96 rows x 64 halfs = 12 KB
constexpr int kRing = 4;                   // weight stages per N-tile ring (each N-tile has its own MMA-issuing warp)
constexpr int kWStages = 2 * kRing;
constexpr int kParts = 4;                  // epilogue warps per TMEM lane quarter: each takes 24 of the 96 columns
constexpr int kEpiWarps = 4 * kParts;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kFirstEpiWarp = 2;           // warp 0: TMA producer, warp 1: MMA issuer of N-tile 0 (leader CTA) + TMEM allocation
constexpr int kMma1Warp = kFirstEpiWarp + kEpiWarps;      // the last warp: MMA issuer of N-tile 1 (leader CTA)
constexpr int kThreads = 32 * (kMma1Warp + 1);
constexpr int kMaxLayers = 12;
constexpr int kChunks = kD / 8;            // 16-byte chunks per row (48)
constexpr int kCol4 = kD / 4;              // float4 groups per row (96)
constexpr int kCPT = kNTc / kParts / 8;    // 16-byte chunks per thread per N-tile (3)
constexpr int kSlots = 2 * kParts;         // threads sharing one row (partials of row reductions)
static_assert(kCPT * 8 * kParts == kNTc, "column split");

// shared-memory carve-up (byte offsets from the 1024-aligned base; identical in both CTAs of the pair)
constexpr int kOffW = 2 * kABuf;                               // weight ring
constexpr int kOffBars = kOffW + kWStages * kWStage;           // 56 mbarrier slots
constexpr int kOffTmem = kOffBars + 56 * 8;                    // TMEM base address (+ pad)
constexpr int kOffIdx = kOffTmem + 16;                         // [64] gather sources of the tile; merged program: [2][64] scatter rows
constexpr int kOffStat = kOffIdx + 3 * kRows * 4 + 16;         // [kSlots][64][4] row-reduction partials
constexpr int kOffLn = kOffStat + kSlots * kRows * 4 * 4;      // [2][2][384] f32: gamma, beta of the program's LayerNorms
constexpr int kOffBias = kOffLn + 4 * kD * 4;                  // [kMaxLayers][384] biases (T: 12 layers of floats would not fit)
constexpr int kOffHead = kOffBias + kMaxLayers * kD * 2;       // [4][384] + [4] (+4 pad) head weights
constexpr int kOffEnd = kOffHead + (4 * kD + 8) * 2;
// barrier slots
constexpr int kBarWFull = 0, kBarWEmpty = kWStages, kBarAFull = 2 * kWStages, kBarAEmpty = 2 * kWStages + kASlots,
              kBarAccFull = 2 * kWStages + 2 * kASlots, kBarAReady = kBarAccFull + 4, kBarEpiDone = kBarAReady + 4,
              kBarProReady = kBarEpiDone + 2;
static_assert(kBarProReady < 56, "barrier slots");

enum { PRO_NONE = 0, PRO_GATHER = 1, PRO_CAST = 2 };
enum { EPI_RELU_A = 0, EPI_LNRELU_A = 1, EPI_ADD3_LN = 2, EPI_RESID = 3, EPI_STORE_A = 4, EPI_STORE_B = 5,
       EPI_GATE = 6, EPI_GATED_LN = 7, EPI_GATED_HEADS = 8, EPI_RESID_A = 9, EPI_RESID_LN_A = 10,
       // the "tile-local" program (neighbour links of every edge stay inside the edge's 64-row tile): like ADD3_LN / RESID,
       // but the half result becomes the next A operand with its rows PERMUTED through the neighbour links -- row r is
       // written where the next layer's gather net[ix] / net[jx] would have fetched it from, rows without a neighbour
       // are zero -- so the exchange that used to need a kernel boundary is a scatter inside the CTA's own A tile
       EPI_ADD3_LN_SC = 11, EPI_RESID_SC = 12,
       // ... and the patch-wise SoftAgg inside the tile: STORE_G parks half(g) in the free A buffer, STORE_F_AGG parks
       // half(f) in the buffer its own MMAs have finished reading, then the epilogue warps walk each patch's edge chain
       // (the neighbour links), take the softmax-weighted sum per channel and write the group row to every member's row
       // of the next A operand: segment_softmax_sum + the gather y[gid] without leaving shared memory
       EPI_STORE_G = 13, EPI_STORE_F_AGG = 14 };
__host__ __device__ constexpr bool epi_writes_a(int e) {
  return e == EPI_RELU_A || e == EPI_LNRELU_A || e == EPI_GATED_LN || e == EPI_RESID_A || e == EPI_RESID_LN_A ||
         e == EPI_ADD3_LN_SC || e == EPI_RESID_SC || e == EPI_STORE_F_AGG;
}

template <typename T>
struct GruProg {
  int rows, src_rows;                      // valid rows of this launch; rows of x16_in (gather source)
  int n_layers, pro;
  int kblocks0, stream_a0, use_w0;         // layer 0: K-blocks; A streamed by TMA (tm_a); weights from tm_w0
  int a0_hint;                             // the streamed A rows (read once) are loaded with an L2 evict-first hint
  int state_half;                          // ADD3: the hidden state comes in as half row-major (x16_in) instead of state32
  int w_row[kMaxLayers];                   // row offset of the layer in the stacked weight matrix (tm_w)
  int epi[kMaxLayers];
  const T* bias[kMaxLayers];
  const float* ln_g[2];
  const float* ln_b[2];
  float eps;
  const T* x16_in;                         // row-major [src_rows,384]: gather source / half hidden state in (ADD3)
  const int64_t* idx64;                    // PRO_GATHER: source row per row (-1 => zero row); null => identity
  const int32_t* idx32;                    // PRO_GATHER: ... or a 32-bit index (the group of each row) when idx64 is null
  const int64_t* sc_ix;                    // tile-local program: previous / next edge of the same patch (graph plan), both in
  const int64_t* sc_jx;                    //   the edge's own 64-row tile or -1
  const T* inp16;                          // ADD3: imap [n_patches,384]
  const int64_t* kk;                       // ADD3: patch of each row
  float* net32;                            // tile layout: the fp32 hidden state / running `net`
  float* n32;                              // tile layout
  T* gate16;                               // tile layout
  int out_a, out_b;                        // the launch has a row-major half output through tm_oa / tm_ob
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
// cvt.rn.f16.f32 (F2F) only at 16 lanes/clk/SM.
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
// eight products of packed halves, each rounded to T (HMUL2 / HMUL2.BF16)
template <typename T> __device__ __forceinline__ uint32_t mul2(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t mul2<__half>(uint32_t a, uint32_t b) {
  const __half2 h = __hmul2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t mul2<__nv_bfloat16>(uint32_t a, uint32_t b) {
  const __nv_bfloat162 h = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <typename T> __device__ __forceinline__ uint4 mul8(uint4 a, uint4 b) {
  return make_uint4(mul2<T>(a.x, b.x), mul2<T>(a.y, b.y), mul2<T>(a.z, b.z), mul2<T>(a.w, b.w));
}
// sigmoid of eight packed 16-bit values, rounded to T:  0.5 * tanh(0.5 x) + 0.5  with ONE MUFU per PAIR (tanh.approx.f16x2 /
// .bf16x2) instead of ex2 + rcp per element.  The gated layers were MUFU-bound (16 MUFU per 8 columns at 4 lanes/clk/SMSP =
// 512 cycles per chunk round with 4 warps per scheduler: tools/gru_timing.py).  Absolute error <= 2^-11 (tanh.approx) / 2,
// i.e. below half an ulp of T for sigmoid values >= 0.25 and <= 2.5e-4 absolute everywhere.
template <typename T> __device__ __forceinline__ uint32_t sigmoid2(uint32_t x);
template <> __device__ __forceinline__ uint32_t sigmoid2<__half>(uint32_t x) {
  const __half2 half = __float2half2_rn(0.5f);
  __half2 h = __hmul2(*reinterpret_cast<const __half2*>(&x), half);
  uint32_t t, hu = *reinterpret_cast<const uint32_t*>(&h);
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(hu));
  const __half2 r = __hfma2(*reinterpret_cast<const __half2*>(&t), half, half);
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <> __device__ __forceinline__ uint32_t sigmoid2<__nv_bfloat16>(uint32_t x) {
  const __nv_bfloat162 half = __float2bfloat162_rn(0.5f);
  __nv_bfloat162 h = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&x), half);
  uint32_t t, hu = *reinterpret_cast<const uint32_t*>(&h);
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(t) : "r"(hu));
  const __nv_bfloat162 r = __hfma2(*reinterpret_cast<const __nv_bfloat162*>(&t), half, half);
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <typename T> __device__ __forceinline__ uint4 sigmoid8(uint4 u) {
  return make_uint4(sigmoid2<T>(u.x), sigmoid2<T>(u.y), sigmoid2<T>(u.z), sigmoid2<T>(u.w));
}
// round eight floats to T and back (the autocast rounding point of a half-typed intermediate)
template <typename T> __device__ __forceinline__ void rnd8(float* v) { unpack8<T>(pack8<T>(v), v); }

// byte offset of 16-byte chunk c (0..47) of row r (0..63) inside an A buffer (K-major SWIZZLE_128B, 6 K-blocks of 8 KB)
__device__ __forceinline__ uint32_t a_off(int r, int c) {
  return (uint32_t)((c >> 3) * kABlk + r * 128 + (((c & 7) ^ (r & 7)) << 4));
}
// tile layouts: float4 group q (0..95) / half8 chunk c (0..47) of row r of 64-row tile t
__device__ __forceinline__ size_t t32(int tile, int q, int r) { return (((size_t)tile * kCol4 + q) * kRows + r) * 4; }
__device__ __forceinline__ size_t t16(int tile, int c, int r) { return (((size_t)tile * kChunks + c) * kRows + r) * 8; }

__device__ __forceinline__ float4 as_f4(uint4 u) {
  return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
}
constexpr int kDbgSlots = 80;             // stamps per launch: 4 + 6 per layer, up to kMaxLayers = 12
__device__ __forceinline__ void stamp(long long* dbg, int slot) {
  if (dbg != nullptr && blockIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[slot] = t;
    dbg[16 * kDbgSlots + slot] = clock64(); // SM cycles next to wall time: separates clock ramps from pipeline stalls
  }
}
// sigmoid in 4 instructions (FMUL, MUFU.EX2, FADD, MUFU.RCP); the result is rounded to half right away
__device__ __forceinline__ float sigmoidf_(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

// ---- pair-scope synchronisation helpers ---------------------------------------------------------------------------
// arrive on an mbarrier of the LEADER CTA (address already mapped with mapa).  Default (.release.cta) semantics, like
// CUTLASS's ClusterBarrier::arrive(cta_id): what the leader's MMA warp consumes is shared memory read by the tensor cores
// through the async proxy (ordered by the writer's fence.proxy.async) and TMEM (ordered by tcgen05.fence).  A
// .release.cluster arrive compiles to MEMBAR.ALL.GPU and waits for every outstanding GLOBAL store of the thread (measured:
// the top stall of the epilogue warps), a cluster-scope acquire on the waiting side to a CCTL.IVALL per wait.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA load whose completion bytes are credited to an mbarrier that may live in the PEER CTA of the pair (the leader's)
__device__ __forceinline__ void tma_load_2d_pair_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t"
      "}" ::"r"(dst), "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
// the same load with an L2 evict-first hint (rows that are read exactly once)
__device__ __forceinline__ void tma_load_2d_pair_elect_evict_first(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      ".reg .b64 pol;\n\t"
      "createpolicy.fractional.L2::evict_first.b64 pol, 1.0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], pol;\n\t"
      "}" ::"r"(dst), "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_mma2_f16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// commit the MMAs issued so far; the arrive lands on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit2_elect(uint32_t bar_addr) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
      "}" ::"r"(bar_addr), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void epi_bar_all() { asm volatile("bar.sync 3, %0;" ::"n"(kEpiThreads) : "memory"); }
// the 8 warps that share the rows of one row half (LayerNorm statistics, head partials)
__device__ __forceinline__ void epi_bar_half(int rowhalf) { asm volatile("bar.sync %0, %1;" ::"r"(1 + rowhalf), "n"(kEpiThreads / 2) : "memory"); }

// ---- the epilogue role ---------------------------------------------------------------------------------------------
template <typename T>
struct Epi {
  const GruProg<T>& P;
  const CUtensorMap* tm_oa; const CUtensorMap* tm_ob;
  unsigned char* As;           // shared-memory base: both A buffers, then everything at the kOff* offsets
  uint32_t bars_l;             // the LEADER's barrier array as a shared::cluster address
  int row0, lane, rowhalf, cbase, slot, et, r, grow;
  bool live;
  uint32_t trow;               // TMEM address of this thread's lane, column 0
  int cur;                     // A buffer the current layer's MMAs read; the epilogue writes the other one
  int na;                      // A-writing epilogues so far
  bool store_pending;          // a TMA store may still be reading the buffer the next epilogue writes
  int ln_used;
  float* net_r;                // tile-layout bases of this thread's row: float4 group q at net_r + q * (kRows * 4)
  float* n32_r;
  T* gate_r;                   // half8 chunk c at gate_r + c * (kRows * 8)

  __device__ __forceinline__ Epi(const GruProg<T>& P_) : P(P_) {}

  __device__ __forceinline__ static const uint4* f4(const float* base, int q) { return reinterpret_cast<const uint4*>(base + (size_t)q * (kRows * 4)); }
  __device__ __forceinline__ static float4* f4w(float* base, int q) { return reinterpret_cast<float4*>(base + (size_t)q * (kRows * 4)); }
  // global 16-byte chunk index (0..47) of this thread's j-th chunk of N-tile h
  __device__ __forceinline__ int chunk_of(int h, int j) const { return h * (kNT / 8) + cbase + j; }
  __device__ __forceinline__ float* s_stat() const { return reinterpret_cast<float*>(As + kOffStat); }
  __device__ __forceinline__ const float* s_ln() const { return reinterpret_cast<const float*>(As + kOffLn); }
  __device__ __forceinline__ const T* s_bias() const { return reinterpret_cast<const T*>(As + kOffBias); }
  __device__ __forceinline__ const T* s_head() const { return reinterpret_cast<const T*>(As + kOffHead); }
  __device__ __forceinline__ int* s_idx() const { return reinterpret_cast<int*>(As + kOffIdx); }
  // tile-local program: s_sc(0)[r] = where the ADD3_LN_SC epilogue writes row r (as the c1 gather net[ix] would read it:
  // row jx[r]), s_sc(1)[r] = the same for RESID_SC / the c2 gather net[jx] (row ix[r]).  Low 16 bits: tile row + 1
  // (0: nobody reads this row); bit 16: row r itself has no source and must be zero in that operand.
  __device__ __forceinline__ int* s_sc(int k) const { return reinterpret_cast<int*>(As + kOffIdx) + k * kRows; }
  __device__ __forceinline__ void scatter_index() {
    if (et < kRows) {
      const int gr = row0 + et;
      long long ixv = -1, jxv = -1;
      if (gr < P.rows) { ixv = P.sc_ix[gr]; jxv = P.sc_jx[gr]; }
      const int dj = jxv >= 0 ? (int)(jxv - row0) : -1, di = ixv >= 0 ? (int)(ixv - row0) : -1;
      if ((jxv >= 0 && (dj < 0 || dj >= kRows)) || (ixv >= 0 && (di < 0 || di >= kRows))) __trap();   // the caller's promise is broken
      s_sc(0)[et] = (dj + 1) | ((ixv < 0) ? (1 << 16) : 0);
      s_sc(1)[et] = (di + 1) | ((jxv < 0) ? (1 << 16) : 0);
    }
    epi_bar_all();
    if (et == 0) {        // the first edge of every patch chain, compacted: the work list of the in-tile aggregation
      int* heads = s_sc(2);
      int n = 0;
      for (int q = 0; q < kRows; q++)
        if (s_sc(0)[q] >> 16) heads[n++] = q;
      heads[kRows] = n;
    }
    epi_bar_all();
  }
  __device__ __forceinline__ uint64_t* acc_full() const { return reinterpret_cast<uint64_t*>(As + kOffBars) + kBarAccFull; }
  __device__ __forceinline__ unsigned char* next_a() const { return As + (cur ^ 1) * kABuf; }

  // this warp's part of N-tile h of the next A operand is in place (generic-proxy writes -> async proxy, then one arrive
  // per warp on the leader's barrier)
  __device__ __forceinline__ void signal_a(int h) {
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(bars_l + (uint32_t)((kBarAReady + (na & 1) * 2 + h) * 8));
  }
  // Row reductions over the 8 threads that share a row (they sit in 8 different warps of the same row half).
  // The scratch is NOT reused without an intervening data dependency: the launch's LayerNorm #k uses the [8][64][2] block
  // k (k = 0, 1), the head partials ([8][64][4]) use both blocks -- and between two uses every epilogue warp of the CTA has
  // passed an accumulator barrier that itself depends on all warps having finished the earlier use (A-writing layers
  // signal the MMA warp only after their second pass).  So ONE named barrier per reduction suffices.
  template <int N>
  __device__ __forceinline__ void row_reduce(float* buf, float* vals) {
    float* mine = buf + (slot * kRows + r) * N;
#pragma unroll
    for (int k = 0; k < N; k++) mine[k] = vals[k];
    epi_bar_half(rowhalf);
#pragma unroll
    for (int k = 0; k < N; k++) vals[k] = 0.f;
#pragma unroll
    for (int p = 0; p < kSlots; p++)
#pragma unroll
      for (int k = 0; k < N; k++) vals[k] += buf[(p * kRows + r) * N + k];
  }
  __device__ __forceinline__ void ln_stats(float s1, float s2, float& mean, float& rstd) {
    float v[2] = {s1, s2};
    row_reduce<2>(s_stat() + (ln_used & 1) * (kSlots * kRows * 2), v);
    mean = v[0] * (1.0f / kD);
    rstd = rsqrtf(fmaxf(v[1] * (1.0f / kD) - mean * mean, 0.f) + P.eps);
  }
  // a TMA store of the previous layer may still read the buffer this layer is about to write
  __device__ __forceinline__ void settle_store() {
    if (store_pending) {
      if (et == 0) tma_store_wait_read();
      epi_bar_all();
      store_pending = false;
    }
  }
  // write the staged [64 x 384] half image (the buffer the epilogue just filled) to global memory, row-major
  __device__ __forceinline__ void store_image(const CUtensorMap* map) {
    fence_proxy_async();
    epi_bar_all();
    if (et == 0) {
      const uint32_t src = smem_u32(next_a());
#pragma unroll
      for (int kb = 0; kb < kKB; kb++) tma_store_2d(map, src + kb * kABlk, kb * 64, row0);
      tma_store_commit();
    }
    store_pending = true;
  }

  // ---------------- prologue: build the full [64 x 384] A tile of layer 0 in buffer 0 ----------------
  // the source row of every tile row of a gather prologue.  The index arrays (neighbour links, group ids) belong to the
  // graph plan -- inputs of the whole update, complete before its first program started -- so this runs BEFORE the wait
  // for the previous program and its L2 round trip is off the boundary between two programs.
  __device__ __forceinline__ void prologue_index() {
    if (et < kRows) {
      const int gr = row0 + et;
      int src = -1;
      if (gr < P.rows) {
        const long long j = P.idx64 ? (long long)P.idx64[gr] : (P.idx32 ? (long long)P.idx32[gr] : (long long)gr);
        src = (j >= 0 && j < (long long)P.src_rows) ? (int)j : -1;
      }
      s_idx()[et] = src;
    }
    epi_bar_all();
  }

  template <int PRO>
  __device__ __forceinline__ void prologue() {
    unsigned char* A0 = As;
    if constexpr (PRO == PRO_GATHER) {
      // cooperative: 48 consecutive threads copy one 768-byte source row; all 6 loads of a thread are in flight together
      constexpr int kPer = kRows * kChunks / kEpiThreads;     // 6
      uint4 v[kPer];
#pragma unroll
      for (int u = 0; u < kPer; u++) {
        const int q = et + u * kEpiThreads;
        const int rr = q / kChunks, c = q - rr * kChunks;
        const int s = s_idx()[rr];
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (s >= 0) v[u] = *reinterpret_cast<const uint4*>(P.x16_in + (size_t)s * kD + c * 8);
      }
#pragma unroll
      for (int u = 0; u < kPer; u++) {
        const int q = et + u * kEpiThreads;
        const int rr = q / kChunks, c = q - rr * kChunks;
        *reinterpret_cast<uint4*>(A0 + a_off(rr, c)) = v[u];
      }
    } else {
      // PRO_CAST: A = half(net32), every thread its own chunks of its own row
#pragma unroll 1
      for (int h = 0; h < 2; h++) {           // one N-tile's worth of loads in flight at a time (register budget)
        uint4 f[2 * kCPT];
#pragma unroll
        for (int j = 0; j < kCPT; j++) {
          const int c = chunk_of(h, j);
          f[2 * j] = *f4(net_r, 2 * c);
          f[2 * j + 1] = *f4(net_r, 2 * c + 1);
        }
#pragma unroll
        for (int j = 0; j < kCPT; j++) {
          const int c = chunk_of(h, j);
          const float4 a = as_f4(f[2 * j]), b = as_f4(f[2 * j + 1]);
          const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
          *reinterpret_cast<uint4*>(A0 + a_off(r, c)) = pack8<T>(v);
        }
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(bars_l + (uint32_t)(kBarProReady * 8));
  }

  // ---------------- one layer's epilogue over this thread's 6 chunks (2 N-tiles x 3) ----------------
  template <int EPI>
  __device__ __forceinline__ void layer(int l) {
    constexpr bool kGated = (EPI == EPI_GATED_LN || EPI == EPI_GATED_HEADS);
    constexpr bool kAdd3 = (EPI == EPI_ADD3_LN || EPI == EPI_ADD3_LN_SC);
    constexpr bool kScatter = (EPI == EPI_ADD3_LN_SC || EPI == EPI_RESID_SC);
    constexpr bool kResid = (EPI == EPI_RESID || EPI == EPI_RESID_SC || EPI == EPI_RESID_A || EPI == EPI_RESID_LN_A);
    constexpr bool kTwoPass = (EPI == EPI_LNRELU_A || kAdd3 || EPI == EPI_GATED_LN || EPI == EPI_RESID_LN_A);
    constexpr bool kWritesA = epi_writes_a(EPI);
    constexpr int kAux = (kResid || kAdd3) ? 2 : (kGated ? 3 : 1);
    int drow = -1;                       // kScatter: tile row of the next A operand this thread's row goes to
    bool zrow = false;                   // kScatter: this thread's own row of the next A operand is zero
    if constexpr (kScatter) {
      const int sc = s_sc(EPI == EPI_ADD3_LN_SC ? 0 : 1)[r];
      drow = (sc & 0xffff) - 1;
      zrow = (sc >> 16) != 0;
    }
    constexpr int kIter = 2 * kCPT;
    const T* bias = s_bias() + l * kD;
    const int set = l & 1;
    const uint32_t accph = (uint32_t)((l >> 1) & 1);
    float s1 = 0.f, s2 = 0.f;
    float hacc[4] = {0.f, 0.f, 0.f, 0.f};
    const T* netrow = nullptr;
    const T* inprow = nullptr;
    if (kAdd3 && live) {
      netrow = P.x16_in + (size_t)grow * kD;
      inprow = P.inp16 + (size_t)P.kk[grow] * kD;
    }
    const bool out_img = (EPI == EPI_STORE_A || EPI == EPI_ADD3_LN) ? (P.out_a != 0)
                       : (EPI == EPI_STORE_B) ? (P.out_b != 0)
                       : (EPI == EPI_RESID || EPI == EPI_GATED_HEADS) ? (P.out_a != 0 && l + 1 == P.n_layers) : false;
    const bool writes_smem = kWritesA || out_img || EPI == EPI_STORE_G;
    // (STORE_F_AGG parks its tile in the buffer its OWN MMAs read -- free once both N-tiles are complete, see below --
    //  because the other buffer holds g)
    unsigned char* An = (EPI == EPI_STORE_F_AGG) ? (As + cur * kABuf) : next_a();
    const uint32_t tcol = trow + (uint32_t)(set * kNT);      // + h * kNTc + j * 8 (trow already points at this thread's 24 columns)
    // per-row operands of the element-wise tail, fetched one chunk ahead of their use
    // RESID: q[0..1] net32 ; GATED: q[0..1] n32, q[2] gate ; ADD3: q[0..1] state32 (or q[0] half state), q[... ] inp
    auto load_aux = [&](int c, uint4* q) {
      if constexpr (kResid) {
        q[0] = *f4(net_r, 2 * c);
        q[1] = *f4(net_r, 2 * c + 1);
      } else if constexpr (kGated) {
        q[0] = *f4(n32_r, 2 * c);
        q[1] = *f4(n32_r, 2 * c + 1);
        q[2] = *reinterpret_cast<const uint4*>(gate_r + (size_t)c * (kRows * 8));
      } else if constexpr (kAdd3) {
        q[0] = make_uint4(0u, 0u, 0u, 0u);
        q[1] = q[0];
        if (P.state_half) { if (live) q[0] = *reinterpret_cast<const uint4*>(netrow + c * 8); }
        else { q[0] = *f4(net_r, 2 * c); q[1] = *f4(net_r, 2 * c + 1); }
      }
    };
    if (writes_smem) settle_store();
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
      // every per-row operand of this N-tile's three chunks is requested BEFORE the accumulator wait: the L2 round trips
      // (state, gate, context rows) overlap the MMAs instead of serialising chunk by chunk
      uint4 aux[kCPT][kAux];
      uint4 inp[kCPT];
#pragma unroll
      for (int j = 0; j < kCPT; j++) {
        const int c = chunk_of(h, j);
        if constexpr (kAux > 1) load_aux(c, aux[j]);
        if constexpr (kAdd3) {
          inp[j] = make_uint4(0u, 0u, 0u, 0u);
          if (live) inp[j] = __ldg(reinterpret_cast<const uint4*>(inprow + c * 8));
        }
      }
      // a streamed layer's epilogue overwrites A ring slots: BOTH tiles' MMAs must be done before the first store
      if (((l == 0 && P.stream_a0) || EPI == EPI_STORE_F_AGG) && h == 0) mbar_wait(acc_full() + set * 2 + 1, accph);
      mbar_wait(acc_full() + set * 2 + h, accph);           // N-tile h of this layer is complete in TMEM (both CTAs)
      tc_fence_after();
      if (et == 0) stamp(P.dbg, 6 + 6 * l + h);
#pragma unroll
      for (int j = 0; j < kCPT; j++) {
        const int c = chunk_of(h, j);                         // global 16-byte chunk index (8 columns)
        uint4 (&cur_)[kAux] = aux[j];
        uint32_t raw[8];
        tmem_ld8(tcol + h * kNTc + j * 8, raw);
        tmem_wait_ld();
        float o[8];
        {
          float bb[8];
          unpack8<T>(*reinterpret_cast<const uint4*>(bias + c * 8), bb);
#pragma unroll
          for (int k = 0; k < 8; k++) o[k] = __uint_as_float(raw[k]) + bb[k];
        }
        const uint4 oh = pack8<T>(o);                 // the Linear output, rounded to half (autocast)
        if constexpr (EPI == EPI_RELU_A) {
          *reinterpret_cast<uint4*>(An + a_off(r, c)) = relu8<T>(oh);   // max(.,0) commutes with the rounding
        } else if constexpr (EPI == EPI_STORE_A || EPI == EPI_STORE_B || EPI == EPI_STORE_G || EPI == EPI_STORE_F_AGG) {
          *reinterpret_cast<uint4*>(An + a_off(r, c)) = oh;             // staging image (written out by TMA below) / g, f tiles
        } else if constexpr (EPI == EPI_GATE) {
          *reinterpret_cast<uint4*>(gate_r + (size_t)c * (kRows * 8)) = oh;
        } else if constexpr (EPI == EPI_LNRELU_A) {
          unpack8<T>(oh, o);
#pragma unroll
          for (int k = 0; k < 8; k++) { s1 += o[k]; s2 += o[k] * o[k]; }
          *reinterpret_cast<uint4*>(An + a_off(r, c)) = oh;             // parked (half) until normalised
        } else if constexpr (kAdd3) {
          float x[8], b[8];
          unpack8<T>(oh, o);
          unpack8<T>(inp[j], b);
          if (P.state_half) {        // half state: T(T(net + inp) + corr)   (the very first update of a sequence)
            unpack8<T>(cur_[0], x);
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] += b[k];
            rnd8<T>(x);
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] += o[k];
            rnd8<T>(x);
          } else {                   // float32 state: (net + float(inp)) + float(corr), no rounding
            const float4 a0 = as_f4(cur_[0]), a1 = as_f4(cur_[1]);
            const float n8[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = (n8[k] + b[k]) + o[k];
          }
          uint32_t xs[8];
#pragma unroll
          for (int k = 0; k < 8; k++) { s1 += x[k]; s2 += x[k] * x[k]; xs[k] = __float_as_uint(x[k]); }
          tmem_st8(tcol + h * kNTc + j * 8, xs);      // fp32 row parked in its own accumulator columns
        } else if constexpr (EPI == EPI_RESID || EPI == EPI_RESID_SC) {
          unpack8<T>(oh, o);
          float4 a = as_f4(cur_[0]), b = as_f4(cur_[1]);
          a.x += o[0]; a.y += o[1]; a.z += o[2]; a.w += o[3]; b.x += o[4]; b.y += o[5]; b.z += o[6]; b.w += o[7];
          *f4w(net_r, 2 * c) = a;
          *f4w(net_r, 2 * c + 1) = b;
          if constexpr (EPI == EPI_RESID_SC) {
            const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            if (drow >= 0) *reinterpret_cast<uint4*>(An + a_off(drow, c)) = pack8<T>(v);
            if (zrow) *reinterpret_cast<uint4*>(An + a_off(r, c)) = make_uint4(0u, 0u, 0u, 0u);
          } else if (out_img) {
            const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            *reinterpret_cast<uint4*>(An + a_off(r, c)) = pack8<T>(v);
          }
        } else if constexpr (EPI == EPI_RESID_A || EPI == EPI_RESID_LN_A) {
          // net += half Linear output (the SoftAgg `h` layer applied per edge: h(y)[:, gid] == h(y[:, gid]), row by row
          // bit-identical to the per-group product), then the sum becomes the next A operand (optionally after LayerNorm)
          unpack8<T>(oh, o);
          const float4 a = as_f4(cur_[0]), b = as_f4(cur_[1]);
          float x[8] = {a.x + o[0], a.y + o[1], a.z + o[2], a.w + o[3], b.x + o[4], b.y + o[5], b.z + o[6], b.w + o[7]};
          if constexpr (EPI == EPI_RESID_A) {
            *f4w(net_r, 2 * c) = make_float4(x[0], x[1], x[2], x[3]);
            *f4w(net_r, 2 * c + 1) = make_float4(x[4], x[5], x[6], x[7]);
            *reinterpret_cast<uint4*>(An + a_off(r, c)) = pack8<T>(x);
          } else {
            uint32_t xs[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { s1 += x[k]; s2 += x[k] * x[k]; xs[k] = __float_as_uint(x[k]); }
            tmem_st8(tcol + h * kNTc + j * 8, xs);
          }
        } else {   // EPI_GATED_LN / EPI_GATED_HEADS:  x = n + half(half(sigmoid(gate)) * res)
          float g[8];
          // half(sigmoid(gate)) * half(res) rounded to half: packed 16-bit arithmetic is exactly that dtype flow
          unpack8<T>(mul8<T>(sigmoid8<T>(cur_[2]), oh), g);
          const float4 a = as_f4(cur_[0]), b = as_f4(cur_[1]);
          float x[8] = {a.x + g[0], a.y + g[1], a.z + g[2], a.w + g[3], b.x + g[4], b.y + g[5], b.z + g[6], b.w + g[7]};
          if constexpr (EPI == EPI_GATED_LN) {
            uint32_t xs[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { s1 += x[k]; s2 += x[k] * x[k]; xs[k] = __float_as_uint(x[k]); }
            tmem_st8(tcol + h * kNTc + j * 8, xs);
          } else {
            *f4w(net_r, 2 * c) = make_float4(x[0], x[1], x[2], x[3]);          // the new hidden state (float32)
            *f4w(net_r, 2 * c + 1) = make_float4(x[4], x[5], x[6], x[7]);
            if (out_img) *reinterpret_cast<uint4*>(An + a_off(r, c)) = pack8<T>(x);
            float hw[8];
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = fmaxf(x[k], 0.f);
            rnd8<T>(x);
#pragma unroll
            for (int o4 = 0; o4 < 4; o4++) {
              unpack8<T>(*reinterpret_cast<const uint4*>(s_head() + o4 * kD + c * 8), hw);
#pragma unroll
              for (int k = 0; k < 8; k++) hacc[o4] += x[k] * hw[k];
            }
          }
        }
      }
      if constexpr (kWritesA && !kTwoPass && EPI != EPI_STORE_F_AGG) {
        if (l + 1 < P.n_layers) signal_a(h);     // the MMA warps may start on the K-blocks fed by this N-tile
      }
    }
    if (et == 0) stamp(P.dbg, 8 + 6 * l);
    if constexpr (EPI == EPI_STORE_F_AGG) {
      // ---------------- the patch-wise SoftAgg of this tile (blocks.py:40-43 on rows that all live here) ----------------
      // g sits in the other A buffer, f in this layer's own; one (patch, 8-channel chunk) item per thread and round: walk
      // the patch's edge chain (kb order, as the segment kernel's sort) for the softmax-weighted sum, then write the group row
      // to every member's row of the next A operand -- in place over g: an item owns its (rows, chunk) cells
      epi_bar_all();                                         // every warp's g (previous layer) and f rows are in place
      unsigned char* G = next_a();
      const unsigned char* F = As + cur * kABuf;
      const int* heads = s_sc(2);
      const int* link = s_sc(0);
      const int nitems = heads[kRows] * kChunks;
#pragma unroll 1
      for (int item = et; item < nitems; item += kEpiThreads) {
        const int c = item % kChunks;
        const int head = heads[item / kChunks];
        // two passes over the chain (the rows sit in shared memory): the maximum first, then ONE exponential per element
        // (the online form of the segment kernel costs two and a data-dependent branch: 5.9 us for this phase at S8)
        float m[8], den[8], num[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { m[k] = -INFINITY; den[k] = 0.f; num[k] = 0.f; }
        for (int rr = head; rr >= 0; rr = (link[rr] & 0xffff) - 1) {
          float gv[8];
          unpack8<T>(*reinterpret_cast<const uint4*>(G + a_off(rr, c)), gv);
#pragma unroll
          for (int k = 0; k < 8; k++) m[k] = fmaxf(m[k], gv[k]);
        }
        for (int rr = head; rr >= 0; rr = (link[rr] & 0xffff) - 1) {
          float gv[8], fv[8];
          unpack8<T>(*reinterpret_cast<const uint4*>(G + a_off(rr, c)), gv);
          unpack8<T>(*reinterpret_cast<const uint4*>(F + a_off(rr, c)), fv);
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const float e = __expf(gv[k] - m[k]);
            den[k] += e;
            num[k] += e * fv[k];
          }
        }
        float y[8];
#pragma unroll
        for (int k = 0; k < 8; k++) y[k] = den[k] > 0.f ? num[k] / den[k] : 0.f;
        const uint4 yh = pack8<T>(y);
        for (int rr = head; rr >= 0; rr = (link[rr] & 0xffff) - 1) *reinterpret_cast<uint4*>(G + a_off(rr, c)) = yh;
      }
      if (l + 1 < P.n_layers) { signal_a(0); signal_a(1); }
    }
    // ---------------- row-wise tails ----------------
    if constexpr (EPI == EPI_LNRELU_A) {
      float mean, rstd;
      ln_stats(s1, s2, mean, rstd);
      const float* gm = s_ln() + ln_used * 2 * kD;
      const float* bt = gm + kD;
#pragma unroll 1
      for (int i = 0; i < kIter; i++) {
        const int c = chunk_of(i / kCPT, i % kCPT);
        float v[8], gk[8], bk[8];
        unpack8<T>(*reinterpret_cast<const uint4*>(An + a_off(r, c)), v);
        *reinterpret_cast<float4*>(gk) = *reinterpret_cast<const float4*>(gm + c * 8); *reinterpret_cast<float4*>(gk + 4) = *reinterpret_cast<const float4*>(gm + c * 8 + 4);
        *reinterpret_cast<float4*>(bk) = *reinterpret_cast<const float4*>(bt + c * 8); *reinterpret_cast<float4*>(bk + 4) = *reinterpret_cast<const float4*>(bt + c * 8 + 4);
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = fmaxf((v[k] - mean) * rstd * gk[k] + bk[k], 0.f);
        *reinterpret_cast<uint4*>(An + a_off(r, c)) = pack8<T>(v);
        // the K-blocks fed by N-tile 0 are complete after its kCPT chunks: the MMA warps start on them while this
        // thread still normalises N-tile 1
        if (kWritesA && i == kCPT - 1 && l + 1 < P.n_layers) signal_a(0);
      }
      if (kWritesA && l + 1 < P.n_layers) signal_a(1);
      ln_used++;
    } else if constexpr (kAdd3 || EPI == EPI_GATED_LN || EPI == EPI_RESID_LN_A) {
      tmem_wait_st();
      float mean, rstd;
      ln_stats(s1, s2, mean, rstd);
      const float* gm = s_ln() + ln_used * 2 * kD;
      const float* bt = gm + kD;
      float* dst = kAdd3 ? net_r : n32_r;
#pragma unroll 1
      for (int h = 0; h < 2; h++) {
        uint32_t raw3[kCPT][8];
#pragma unroll
        for (int j = 0; j < kCPT; j++) tmem_ld8(tcol + h * kNTc + j * 8, raw3[j]);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < kCPT; j++) {
        const int c = chunk_of(h, j);
        uint32_t (&raw)[8] = raw3[j];
        float v[8], gk[8], bk[8];
        *reinterpret_cast<float4*>(gk) = *reinterpret_cast<const float4*>(gm + c * 8); *reinterpret_cast<float4*>(gk + 4) = *reinterpret_cast<const float4*>(gm + c * 8 + 4);
        *reinterpret_cast<float4*>(bk) = *reinterpret_cast<const float4*>(bt + c * 8); *reinterpret_cast<float4*>(bk + 4) = *reinterpret_cast<const float4*>(bt + c * 8 + 4);
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (__uint_as_float(raw[k]) - mean) * rstd * gk[k] + bk[k];
        *f4w(dst, 2 * c) = make_float4(v[0], v[1], v[2], v[3]);
        *f4w(dst, 2 * c + 1) = make_float4(v[4], v[5], v[6], v[7]);
        if constexpr (EPI == EPI_ADD3_LN_SC) {
          if (drow >= 0) *reinterpret_cast<uint4*>(An + a_off(drow, c)) = pack8<T>(v);
          if (zrow) *reinterpret_cast<uint4*>(An + a_off(r, c)) = make_uint4(0u, 0u, 0u, 0u);
        } else if (EPI != EPI_ADD3_LN || out_img) {
          *reinterpret_cast<uint4*>(An + a_off(r, c)) = pack8<T>(v);
        }
        }
        if (kWritesA && l + 1 < P.n_layers) signal_a(h);      // (as above: N-tile h of the next A operand is complete)
      }
      ln_used++;
    } else if constexpr (EPI == EPI_GATED_HEADS) {
      row_reduce<4>(s_stat(), hacc);
      if (live && slot == 0) {
        const uint32_t d = pack2<T>(hacc[0] + ElemTraits<T>::to_float(s_head()[4 * kD + 0]), hacc[1] + ElemTraits<T>::to_float(s_head()[4 * kD + 1]));
        const float2 w = unpack2<T>(pack2<T>(hacc[2] + ElemTraits<T>::to_float(s_head()[4 * kD + 2]), hacc[3] + ElemTraits<T>::to_float(s_head()[4 * kD + 3])));
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
    if (out_img) store_image((EPI == EPI_STORE_B) ? tm_ob : tm_oa);
    // this warp is done with the layer's accumulators (TMEM set l & 1 may be overwritten by layer l + 2)
    tc_fence_before();
    __syncwarp();
    if (lane == 0 && l + 2 < P.n_layers) mbar_arrive_remote(bars_l + (uint32_t)((kBarEpiDone + set) * 8));
    if (kWritesA) { cur ^= 1; na++; }
    if (et == 0) stamp(P.dbg, 9 + 6 * l);
  }
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) gru_mma_kernel(const __grid_constant__ CUtensorMap tm_w,
                                                              const __grid_constant__ CUtensorMap tm_w0,
                                                              const __grid_constant__ CUtensorMap tm_a,
                                                              const __grid_constant__ CUtensorMap tm_oa,
                                                              const __grid_constant__ CUtensorMap tm_ob,
                                                              const __grid_constant__ GruProg<T> P) {
  extern __shared__ unsigned char smem_dyn[];
  // identical carve-up in both CTAs of the pair (the MMA descriptors and mapa address the same offset in the peer)
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* As = base;                                   // 2 x 48 KB
  unsigned char* Ws = base + kOffW;                           // kWStages x 12 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kOffBars);
  uint64_t* w_full = bars + kBarWFull;        // [kWStages] leader: both CTAs' halves of the weight block have landed
  uint64_t* w_empty = bars + kBarWEmpty;      // [kWStages] both: the MMAs that read the stage are done (multicast commit)
  uint64_t* a_full = bars + kBarAFull;        // [kASlots]  leader: streamed A K-block landed in both CTAs
  uint64_t* a_empty = bars + kBarAEmpty;      // [kASlots]  both
  uint64_t* acc_full = bars + kBarAccFull;    // [2][2]     both: N-tile h of TMEM set s is complete
  uint64_t* a_ready = bars + kBarAReady;      // [2][2]     leader: N-tile h of the next A operand is written (count: all epilogue warps of the pair)
  uint64_t* epi_done = bars + kBarEpiDone;    // [2]        leader: every epilogue warp of the pair is done with TMEM set s
  uint64_t* pro_ready = bars + kBarProReady;  // [1]        leader: prologue finished in both CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + kOffTmem);
  float* s_ln = reinterpret_cast<float*>(base + kOffLn);
  T* s_bias = reinterpret_cast<T*>(base + kOffBias);
  T* s_head = reinterpret_cast<T*>(base + kOffHead);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();    // 0 = leader (issues the pair's MMAs)
  const int row0 = blockIdx.x * kRows;        // this CTA's 64 rows

  if (threadIdx.x == 0) stamp(P.dbg, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWStages; s++) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < kASlots; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 2); }     // a_empty: both MMA warps
    for (int s = 0; s < 4; s++) { mbar_init(&acc_full[s], 1); mbar_init(&a_ready[s], 2 * kEpiWarps); }
    mbar_init(&epi_done[0], 2 * kEpiWarps);
    mbar_init(&epi_done[1], 2 * kEpiWarps);
    mbar_init(pro_ready, 2 * kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else if (warp >= kFirstEpiWarp) {
    // stage the program's small parameters (biases, LayerNorm affine, heads) in shared memory once
    const int t = threadIdx.x - 32 * kFirstEpiWarp, nt = kEpiThreads;
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
  cluster_sync_all();                         // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool has_pro = (P.pro != PRO_NONE);
  if (threadIdx.x == 0) stamp(P.dbg, 1);
  // PDL: everything above (barriers, TMEM, parameters) and the weight stream below only touch constants, so this grid
  // may start while its predecessor still runs; the next grid of the stream may start now as well.
  if (threadIdx.x == 0) pdl_launch_dependents();

  if (warp == 0) {
    // =========================== TMA producer: this CTA's half of every weight block (+ the streamed A of layer 0).
    // Completion bytes of BOTH CTAs are credited to the leader's barrier; the leader arms it for the pair.
    if (lane == 0) { prefetch_tensormap(&tm_w); if (P.use_w0) prefetch_tensormap(&tm_w0); if (P.stream_a0) prefetch_tensormap(&tm_a); }
    if (P.stream_a0) pdl_wait();                    // the streamed A operand is the previous kernel's output
    // one ring of kRing stages per N-tile (stage index h * kRing + st[h]); K-blocks in ascending order for both tiles
    uint32_t st0 = 0, st1 = 0, ph0 = 0, ph1 = 0;
    for (int l = 0; l < P.n_layers; l++) {
      const int nkb = (l == 0) ? P.kblocks0 : kKB;
      const CUtensorMap* wm = (l == 0 && P.use_w0) ? &tm_w0 : &tm_w;
      const bool streamed = (l == 0 && P.stream_a0);
      for (int kb = 0; kb < nkb; kb++) {
        if (streamed) {
          const int slot = kb % kASlots, use = kb / kASlots;
          mbar_wait(&a_empty[slot], (uint32_t)(use & 1) ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx_elect(smem_u32(&a_full[slot]), 2 * kABlk);
          if (P.a0_hint) tma_load_2d_pair_elect_evict_first(smem_u32(As) + slot * kABlk, &tm_a, mapa(smem_u32(&a_full[slot]), 0u), kb * 64, row0);
          else tma_load_2d_pair_elect(smem_u32(As) + slot * kABlk, &tm_a, mapa(smem_u32(&a_full[slot]), 0u), kb * 64, row0);
        }
        {
          mbar_wait(&w_empty[st0], ph0 ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx_elect(smem_u32(&w_full[st0]), 2 * kWStage);
          tma_load_2d_pair_elect(smem_u32(Ws) + st0 * kWStage, wm, mapa(smem_u32(&w_full[st0]), 0u), kb * 64,
                                 P.w_row[l] + rank * kNTc);
          if (++st0 == kRing) { st0 = 0; ph0 ^= 1u; }
        }
        {
          mbar_wait(&w_empty[kRing + st1], ph1 ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx_elect(smem_u32(&w_full[kRing + st1]), 2 * kWStage);
          tma_load_2d_pair_elect(smem_u32(Ws) + (kRing + st1) * kWStage, wm, mapa(smem_u32(&w_full[kRing + st1]), 0u), kb * 64,
                                 P.w_row[l] + kNT + rank * kNTc);
          if (++st1 == kRing) { st1 = 0; ph1 ^= 1u; }
        }
      }
    }
  } else if (warp == 1 || warp == kMma1Warp) {
    if (rank == 0) {
      // =========================== MMA issuers (leader CTA): warp 1 owns N-tile 0, the last warp N-tile 1; one elected
      // lane issues for the pair.  The two instruction streams touch different accumulators and different weight rings;
      // they share the A operand and the barriers that guard it.
      const int h = (warp == 1) ? 0 : 1;
      const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kNT >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t ad0 = umma_desc_sw128(smem_u32(As));
      const uint64_t bd0 = umma_desc_sw128(smem_u32(Ws) + h * kRing * kWStage);
      uint64_t* wf = w_full + h * kRing;
      uint64_t* we = w_empty + h * kRing;
      uint32_t stage = 0, phase = 0;
      int cur = P.stream_a0 ? 1 : 0;            // A buffer read by the current layer (a streamed layer 0 uses both as a ring)
      int na = 0;                               // A-writing epilogues before this layer
      for (int l = 0; l < P.n_layers; l++) {
        const int nkb = (l == 0) ? P.kblocks0 : kKB;
        const bool streamed = (l == 0 && P.stream_a0);
        const bool wave = (l > 0) && epi_writes_a(P.epi[l - 1]);
        const int set = l & 1;
        if (l == 0) {
          if (has_pro) { mbar_wait_cluster(pro_ready, 0u); tc_fence_after(); }
        } else if (l >= 2) {
          mbar_wait_cluster(&epi_done[set], (uint32_t)(((l - 2) >> 1) & 1));    // TMEM set free again
          tc_fence_after();
        }
        if (lane == 0 && h == 0) stamp(P.dbg, 4 + 6 * l);
        const uint32_t abase = streamed ? 0u : (uint32_t)(cur * (kABuf >> 4));
        const uint32_t tacc = tmem_base + (uint32_t)(set * kNT + h * kNTc);
        const int pa = na - 1;
        for (int kb = 0; kb < nkb; kb++) {
          if (wave && (kb == 0 || kb == kKB / 2)) {
            // K-blocks 0-2 are written by N-tile 0 of the previous epilogue, 3-5 by N-tile 1: the first half of this layer
            // runs under the second half of that epilogue
            mbar_wait_cluster(&a_ready[(pa & 1) * 2 + (kb ? 1 : 0)], (uint32_t)((pa >> 1) & 1));
            tc_fence_after();
          }
          const int slot = streamed ? kb % kASlots : kb;
          // w_full / a_full are completed by TMA transaction bytes
          if (streamed) { mbar_wait(&a_full[slot], (uint32_t)((kb / kASlots) & 1)); tc_fence_after(); }
          mbar_wait(&wf[stage], phase);
          tc_fence_after();
          const uint64_t ad = ad0 + (uint64_t)(abase + (uint32_t)slot * (kABlk >> 4));
          const uint64_t bd = bd0 + (uint64_t)(stage * (kWStage >> 4));
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++)
            tc_mma2_f16_elect(tacc, ad + 2 * k4, bd + 2 * k4, idesc, (kb == 0 && k4 == 0) ? 0u : 1u);
          tc_commit2_elect(smem_u32(&we[stage]));
          if (++stage == kRing) { stage = 0; phase ^= 1u; }
          if (streamed) tc_commit2_elect(smem_u32(&a_empty[slot]));
        }
        tc_commit2_elect(smem_u32(&acc_full[set * 2 + h]));
        if (epi_writes_a(P.epi[l])) { cur ^= 1; na++; }
        if (lane == 0 && h == 0) stamp(P.dbg, 5 + 6 * l);
      }
      __syncwarp();
    }
  } else if (warp >= kFirstEpiWarp) {
    // =========================== prologue + epilogues (struct Epi) =============================================
    Epi<T> e(P);
    e.tm_oa = &tm_oa; e.tm_ob = &tm_ob;
    e.As = As;
    e.bars_l = mapa(smem_u32(bars), 0u);
    e.row0 = row0; e.lane = lane;
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    e.rowhalf = q & 1;
    const int ch = q >> 1;                          // lane half = column half of every N-tile
    const int part = (warp - kFirstEpiWarp) >> 2;   // 0 .. kParts-1: which 24 of the 96 columns
    e.cbase = ch * (kNTc / 8) + part * kCPT;
    e.slot = ch * kParts + part;
    e.et = threadIdx.x - 32 * kFirstEpiWarp;        // 0 .. kEpiThreads-1
    e.r = e.rowhalf * 32 + lane;                    // tile row
    e.grow = row0 + e.r;
    e.live = e.grow < P.rows;
    e.trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * (kNTc / kParts));
    e.cur = P.stream_a0 ? 1 : 0;
    e.na = 0;
    e.store_pending = false;
    e.ln_used = 0;
    e.net_r = P.net32 + t32((int)blockIdx.x, 0, e.r);
    e.n32_r = P.n32 + t32((int)blockIdx.x, 0, e.r);
    e.gate_r = P.gate16 + t16((int)blockIdx.x, 0, e.r);
    if (lane == 0 && P.out_a) prefetch_tensormap(&tm_oa);
    if (P.pro == PRO_GATHER) e.prologue_index();
    if (P.sc_ix != nullptr) e.scatter_index();      // (graph-plan data: read ahead of the wait, like the gather indices)
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
        case EPI_ADD3_LN_SC: e.template layer<EPI_ADD3_LN_SC>(l); break;
        case EPI_RESID_SC: e.template layer<EPI_RESID_SC>(l); break;
        case EPI_STORE_G: e.template layer<EPI_STORE_G>(l); break;
        case EPI_STORE_F_AGG: e.template layer<EPI_STORE_F_AGG>(l); break;
        default: e.template layer<EPI_GATED_HEADS>(l); break;
      }
    }
    if (e.et == 0 && e.store_pending) tma_store_wait_read();  // the staging buffer must outlive the bulk stores' reads (CUTLASS's tma_store_wait<0>)
  }
  // teardown: no CTA of the pair may exit (or free TMEM) while the other can still be using its shared memory / TMEM
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0) stamp(P.dbg, 3);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- state layout conversion: row-major [E,384] <-> tile layout, with an optional row gather -----------------------
// dst row e = (idx ? (idx[e] >= 0 ? src[idx[e]] : 0) : src[e]); layouts: 0 row-major, 1 tile ([tile64][96][64][4])
__global__ void state_gather_kernel(const float* __restrict__ src, int src_layout, int src_rows, const int64_t* __restrict__ idx,
                                    float* __restrict__ dst, int dst_layout, int dst_rows) {
  const int e = blockIdx.x * 4 + (threadIdx.x / kCol4);        // 4 rows per 384-thread block
  const int q = threadIdx.x % kCol4;
  if (e >= dst_rows) return;
  long long s = idx ? (long long)idx[e] : (long long)e;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s >= 0 && s < src_rows) {
    const size_t so = src_layout ? t32((int)(s / kRows), q, (int)(s % kRows)) : ((size_t)s * kD + q * 4);
    v = *reinterpret_cast<const float4*>(src + so);
  }
  const size_t d_o = dst_layout ? t32(e / kRows, q, e % kRows) : ((size_t)e * kD + q * 4);
  *reinterpret_cast<float4*>(dst + d_o) = v;
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

constexpr size_t kSmemBytes = 1024 + (size_t)kOffEnd + 64;
static_assert(kSmemBytes <= 232448, "shared memory budget");

static long long* g_dbg = nullptr;     // 16 launches x kDbgSlots stamps (ns), then the same in SM cycles
static int g_dbg_launch = 0;

template <typename T>
static int launch_prog(const CUtensorMap& tw, const CUtensorMap& tw0, const CUtensorMap& ta, const CUtensorMap& toa,
                       const CUtensorMap& tob, const GruProg<T>& P, cudaStream_t s) {
  // the attribute is per device: one flag per device ordinal (a process may drive several GPUs)
  static bool configured[64] = {};
  int dev = 0;
  DEVO_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    DEVO_CUDA(cudaFuncSetAttribute(gru_mma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const int tiles = (P.rows + kRows - 1) / kRows;
  if (tiles <= 0) return DEVO_OK;
  GruProg<T> Pd = P;
  Pd.dbg = g_dbg ? g_dbg + kDbgSlots * (g_dbg_launch++ % 16) : nullptr;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)((tiles + 1) / 2 * 2), 1, 1);      // whole pairs; a CTA without live rows still plays its part
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;      // the CTA pair of a cta_group::2 MMA
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: overlap the set-up with the previous kernel's tail
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static int use_pdl = -1;
  if (use_pdl < 0) { const char* e = getenv("DEVO_GRU_PDL"); use_pdl = (e && e[0] == '0') ? 0 : 1; }
  cfg.numAttrs = use_pdl ? 2 : 1;
  DEVO_CUDA(cudaLaunchKernelEx(&cfg, gru_mma_kernel<T>, tw, tw0, ta, toa, tob, Pd));
  DEVO_LAUNCH_CHECK("gru_mma");
  return DEVO_OK;
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
static size_t state_floats(int E) { return (size_t)((E + 2 * kRows - 1) / (2 * kRows)) * 2 * kRows * kD; }
struct GruWs {
  size_t n32, gate16, x16a, x16b, g16, f16, y16, hy16, total;
};
static GruWs gru_ws(int E, int max_groups) {
  GruWs w;
  const size_t G = (size_t)(max_groups > 0 ? max_groups : 1);
  size_t off = 0;
  w.n32 = off;   off += al256(state_floats(E) * 4);
  w.gate16 = off; off += al256(state_floats(E) * 2);
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
  float* net32 = io->state32;
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
  CUtensorMap tw, tw0, ta, t_x16a, t_x16b, t_g16, t_f16, t_out;
  int rc = make_map_2d(&tw, dtype, Wt->W, 18 * kD, kD, kD, kNTc);
  if (rc != DEVO_OK) return rc;
  rc = make_map_2d(&tw0, dtype, Wt->W0, kD, (uint64_t)io->corr_ld, (uint64_t)io->corr_ld, kNTc);
  if (rc != DEVO_OK) return rc;
  rc = make_map_2d(&ta, dtype, io->corr16, (uint64_t)E, (uint64_t)io->corr_ld, (uint64_t)io->corr_ld, kRows);
  if (rc != DEVO_OK) return rc;
  if ((rc = make_map_2d(&t_x16a, dtype, x16a, (uint64_t)E, kD, kD, kRows)) != DEVO_OK) return rc;
  if ((rc = make_map_2d(&t_x16b, dtype, x16b, (uint64_t)E, kD, kD, kRows)) != DEVO_OK) return rc;
  if ((rc = make_map_2d(&t_g16, dtype, g16, (uint64_t)E, kD, kD, kRows)) != DEVO_OK) return rc;
  if ((rc = make_map_2d(&t_f16, dtype, f16, (uint64_t)E, kD, kD, kRows)) != DEVO_OK) return rc;
  t_out = t_x16a;
  if (io->net16_out && (rc = make_map_2d(&t_out, dtype, io->net16_out, (uint64_t)E, kD, kD, kRows)) != DEVO_OK) return rc;

  GruProg<T> base;
  memset(&base, 0, sizeof(base));
  base.rows = E; base.src_rows = E; base.eps = Wt->ln_eps;
  base.kblocks0 = kKB;
  base.net32 = net32; base.n32 = n32; base.gate16 = gate16;
  // the correlation rows are read exactly once (layer 0 of the first program): loaded evict-first they do not push the
  // operator's own weights / state out of L2 (DEVO_GRU_L2_HINT=0 turns it off; back-to-back steps +1.2 %, e2e +0.6 %)
  static const int a0_hint_env = [] { const char* e = getenv("DEVO_GRU_L2_HINT"); return e ? atoi(e) : 1; }();
  base.a0_hint = a0_hint_env;

  if (io->tile_local) {
    // (1-3) as ONE program: corr MLP + norm | c1 | c2 | g, f of the patch-wise aggregation.  Every neighbour link stays
    // inside the CTA's 64 rows (the caller's promise), so `mask * net[ix]` / `mask * net[jx]` are row permutations of the
    // tile the CTA has just produced: the ADD3_LN / RESID epilogues write their half rows straight to where the next
    // layer's A operand wants them.  Two kernel boundaries (~5 us each: bulk-store completion, skew, dependent gather)
    // and the two [E,384] images that carried the rows across them are gone.
    // The patch-wise SoftAgg follows in the same launch: all edges of a patch are rows of this CTA, so the segment
    // softmax + sum and the gather of the group row back to the edges (the A operand of `h`) are a walk along the
    // patch's chain in shared memory (EPI_STORE_G / EPI_STORE_F_AGG) -- the segment-reduction kernel and its two
    // boundaries are gone as well; `h`, then g / f of the frame-pair aggregation, close the program.
    GruProg<T> P = base;
    P.n_layers = 12; P.pro = PRO_NONE; P.kblocks0 = io->corr_ld / 64; P.stream_a0 = 1; P.use_w0 = 1;
    P.w_row[0] = 0;      P.epi[0] = EPI_RELU_A;     P.bias[0] = bias;
    P.w_row[1] = 0 * kD; P.epi[1] = EPI_LNRELU_A;   P.bias[1] = B(0);
    P.w_row[2] = 1 * kD; P.epi[2] = EPI_ADD3_LN_SC; P.bias[2] = B(1);
    P.w_row[3] = 2 * kD; P.epi[3] = EPI_RELU_A;     P.bias[3] = B(2);
    P.w_row[4] = 3 * kD; P.epi[4] = EPI_RESID_SC;   P.bias[4] = B(3);
    P.w_row[5] = 4 * kD; P.epi[5] = EPI_RELU_A;     P.bias[5] = B(4);
    P.w_row[6] = 5 * kD; P.epi[6] = EPI_RESID_A;    P.bias[6] = B(5);
    P.w_row[7] = 6 * kD; P.epi[7] = EPI_STORE_G;     P.bias[7] = B(6);
    P.w_row[8] = 7 * kD; P.epi[8] = EPI_STORE_F_AGG; P.bias[8] = B(7);
    P.w_row[9] = 8 * kD;   P.epi[9] = EPI_RESID_A;  P.bias[9] = B(8);
    P.w_row[10] = 9 * kD;  P.epi[10] = EPI_STORE_A; P.bias[10] = B(9);
    P.w_row[11] = 10 * kD; P.epi[11] = EPI_STORE_B; P.bias[11] = B(10);
    P.ln_g[0] = Wt->ln_gamma; P.ln_b[0] = Wt->ln_beta;
    P.ln_g[1] = Wt->ln_gamma + kD; P.ln_b[1] = Wt->ln_beta + kD;
    P.state_half = io->net16 != nullptr;
    P.x16_in = (const T*)io->net16; P.inp16 = (const T*)io->imap16; P.kk = io->kk;
    P.sc_ix = io->ix; P.sc_jx = io->jx;
    P.out_a = 1; P.out_b = 1;
    rc = launch_prog<T>(tw, tw0, ta, t_g16, t_f16, P, s);
    if (rc != DEVO_OK) return rc;
  } else {
  // (1) corr MLP + norm(net + inp + corr)  (enet.py:59-66,82-83)
  {
    GruProg<T> P = base;
    P.n_layers = 3; P.pro = PRO_NONE; P.kblocks0 = io->corr_ld / 64; P.stream_a0 = 1; P.use_w0 = 1;
    P.w_row[0] = 0; P.epi[0] = EPI_RELU_A; P.bias[0] = bias;
    P.w_row[1] = 0 * kD; P.epi[1] = EPI_LNRELU_A; P.bias[1] = B(0);
    P.w_row[2] = 1 * kD; P.epi[2] = EPI_ADD3_LN; P.bias[2] = B(1);
    P.ln_g[0] = Wt->ln_gamma; P.ln_b[0] = Wt->ln_beta;
    P.ln_g[1] = Wt->ln_gamma + kD; P.ln_b[1] = Wt->ln_beta + kD;
    P.state_half = io->net16 != nullptr;
    P.x16_in = (const T*)io->net16; P.inp16 = (const T*)io->imap16; P.kk = io->kk;
    P.out_a = 1;
    rc = launch_prog<T>(tw, tw0, ta, t_x16a, t_x16a, P, s);
    if (rc != DEVO_OK) return rc;
  }
  // (2) net += c1(mask * net[ix])   (enet.py:86-88)
  {
    GruProg<T> P = base;
    P.n_layers = 2; P.pro = PRO_GATHER;
    P.w_row[0] = 2 * kD; P.epi[0] = EPI_RELU_A; P.bias[0] = B(2);
    P.w_row[1] = 3 * kD; P.epi[1] = EPI_RESID;  P.bias[1] = B(3);
    P.x16_in = x16a; P.idx64 = io->ix;
    P.out_a = 1;
    rc = launch_prog<T>(tw, tw0, ta, t_x16b, t_x16b, P, s);
    if (rc != DEVO_OK) return rc;
  }
  // (3) net += c2(mask * net[jx]); then g, f of the patch-wise aggregation on the same rows -- no other tile's rows are
  // needed between the two, so they share one launch: the residual epilogue leaves half(net) as the next A operand
  // (enet.py:89-93, blocks.py:40-43).  The `h` layer of a SoftAgg is applied per EDGE inside the kernel that consumes it
  // (h(y)[:, gid] == h(y[:, gid])): its A operand is the gathered group row y[gid[e]], its epilogue adds to the fp32 state.
  {
    GruProg<T> P = base;
    P.n_layers = 4; P.pro = PRO_GATHER;
    P.w_row[0] = 4 * kD; P.epi[0] = EPI_RELU_A;  P.bias[0] = B(4);
    P.w_row[1] = 5 * kD; P.epi[1] = EPI_RESID_A; P.bias[1] = B(5);
    P.w_row[2] = 6 * kD; P.epi[2] = EPI_STORE_A; P.bias[2] = B(6);
    P.w_row[3] = 7 * kD; P.epi[3] = EPI_STORE_B; P.bias[3] = B(7);
    P.x16_in = x16b; P.idx64 = io->jx;
    P.out_a = 1; P.out_b = 1;
    rc = launch_prog<T>(tw, tw0, ta, t_g16, t_f16, P, s);
    if (rc != DEVO_OK) return rc;
  }
  }   // !tile_local
  if (!io->tile_local) {
  rc = devo::segment_softmax_sum(g16, f16, io->perm_kk, io->gstart_kk, io->ngroups_kk, io->max_groups_kk, y16, dtype, E, kD, (void*)s, 1);
  if (rc != DEVO_OK) return rc;
  {
    GruProg<T> P = base;                                      // net += h_kk(y_kk)[gid_kk]; g, f of the pair-wise aggregation
    P.n_layers = 3; P.pro = PRO_GATHER; P.x16_in = y16; P.idx32 = io->gid_kk; P.src_rows = io->max_groups_kk;
    P.w_row[0] = 8 * kD;  P.epi[0] = EPI_RESID_A; P.bias[0] = B(8);
    P.w_row[1] = 9 * kD;  P.epi[1] = EPI_STORE_A; P.bias[1] = B(9);
    P.w_row[2] = 10 * kD; P.epi[2] = EPI_STORE_B; P.bias[2] = B(10);
    P.out_a = 1; P.out_b = 1;
    rc = launch_prog<T>(tw, tw0, ta, t_g16, t_f16, P, s);
    if (rc != DEVO_OK) return rc;
  }
  }   // !tile_local
  rc = devo::segment_softmax_sum(g16, f16, io->perm_ij, io->gstart_ij, io->ngroups_ij, io->max_groups_ij, hy16, dtype, E, kD, (void*)s, 1);
  if (rc != DEVO_OK) return rc;
  {
    GruProg<T> P = base;                                      // net += h_ij(y_ij)[gid_ij]; LN, GatedResidual x2; heads
    P.n_layers = 7; P.pro = PRO_GATHER; P.x16_in = hy16; P.idx32 = io->gid_ij; P.src_rows = io->max_groups_ij;
    P.ln_g[0] = Wt->ln_gamma + 2 * kD; P.ln_b[0] = Wt->ln_beta + 2 * kD;
    P.ln_g[1] = Wt->ln_gamma + 3 * kD; P.ln_b[1] = Wt->ln_beta + 3 * kD;
    P.w_row[0] = 11 * kD; P.epi[0] = EPI_RESID_LN_A; P.bias[0] = B(11);
    const int epis[6] = {EPI_GATE, EPI_RELU_A, EPI_GATED_LN, EPI_GATE, EPI_RELU_A, EPI_GATED_HEADS};
    for (int l = 0; l < 6; l++) { P.w_row[1 + l] = (12 + l) * kD; P.epi[1 + l] = epis[l]; P.bias[1 + l] = B(12 + l); }
    P.out_a = io->net16_out != nullptr;
    P.headW = (const T*)Wt->head_W; P.headB = (const T*)Wt->head_b;
    P.delta = (T*)io->delta; P.weight = (T*)io->weight;
    if (io->coords && io->target32 && io->weight32) { P.coords = io->coords; P.target32 = io->target32; P.weight32 = io->weight32; }
    rc = launch_prog<T>(tw, tw0, ta, t_out, t_out, P, s);
    if (rc != DEVO_OK) return rc;
  }
  return DEVO_OK;
}

}  // namespace

extern "C" {

size_t devo_gru_workspace(int E, int max_groups) { return gru_ws(E, max_groups).total; }
size_t devo_gru_state_floats(int E) { return state_floats(E); }

int devo_gru_state_gather(const float* src, int src_layout, int src_rows, const int64_t* idx, float* dst, int dst_layout,
                          int dst_rows, void* stream) {
  DEVO_REQUIRE(src && dst && src_rows >= 0 && dst_rows >= 0, DEVO_EINVAL, "gru_state_gather: bad arguments");
  DEVO_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, DEVO_EINVAL, "gru_state_gather: pointers must be 16-byte aligned");
  if (dst_rows == 0) return DEVO_OK;
  state_gather_kernel<<<(dst_rows + 3) / 4, 4 * kCol4, 0, (cudaStream_t)stream>>>(src, src_layout, src_rows, idx, dst, dst_layout, dst_rows);
  DEVO_LAUNCH_CHECK("gru_state_gather");
  return DEVO_OK;
}

// debug (tools/gru_timing.py): enable / read back the %globaltimer stamps of CTA 0 of the last 16 launches
int devo_gru_debug_timing(long long* host_out) {
  if (!g_dbg) {
    if (cudaMalloc(&g_dbg, 2 * 16 * kDbgSlots * sizeof(long long)) != cudaSuccess) return -1;
    cudaMemset(g_dbg, 0, 2 * 16 * kDbgSlots * sizeof(long long));
  }
  g_dbg_launch = 0;
  if (host_out) return (int)cudaMemcpy(host_out, g_dbg, 2 * 16 * kDbgSlots * sizeof(long long), cudaMemcpyDeviceToHost);
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
  DEVO_REQUIRE(io->state32 != nullptr, DEVO_EINVAL, "gru_update: state32 (tile-layout float32 hidden state) is required");
  if (io->E <= 0) return DEVO_OK;
  const int maxG = io->max_groups_kk > io->max_groups_ij ? io->max_groups_kk : io->max_groups_ij;
  DEVO_REQUIRE(workspace && workspace_bytes >= gru_ws(io->E, maxG).total, DEVO_EWORKSPACE, "gru_update: workspace too small");
  DEVO_REQUIRE((((uintptr_t)weights->W | (uintptr_t)weights->W0 | (uintptr_t)io->corr16 | (uintptr_t)io->net16 |
                 (uintptr_t)io->imap16 | (uintptr_t)io->net16_out | (uintptr_t)weights->bias | (uintptr_t)weights->head_W |
                 (uintptr_t)io->state32 | (uintptr_t)workspace) & 15) == 0, DEVO_EINVAL, "gru_update: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == DEVO_F16) return gru_update_impl<__half>(weights, io, dtype, workspace, s);
  return gru_update_impl<__nv_bfloat16>(weights, io, dtype, workspace, s);
}

}  // extern "C"
