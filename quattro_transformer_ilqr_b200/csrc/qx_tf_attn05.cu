// Causal multi-head attention of the Quattro transformer on tcgen05 (UMMA + TMEM + TMA), sm_100a.
//
//   = the self-attention inside nn.TransformerEncoderLayer (transformer_model.py:112-113, :134-136 of the reference:
//     bool causal mask triu(ones(L, L), 1), 4 heads x 32, softmax(Q K^T / sqrt(32)) V), for every (trajectory, head).
//
// A (trajectory, head) item is a 101x32 . 32x101 . 101x32 problem (61 for cart-pole): ONE M = 128 UMMA tile in the
// query dimension.  Per item, one 128-thread CTA (thread = query row) runs
//   1. TMA: the Q, K, V head slices of the packed qkv rows as three [128 tokens x 32 dims] boxes of a
//      [trajectory][token][feature] tensor map (tokens past L are zero-filled), 128-B swizzle;
//   2. split: every tile into TF32 (hi, lo) halves in shared memory (3xTF32: hi*hi + lo*hi + hi*lo are
//      fp32-accurate products - the 1e-5 parity bar rules plain TF32 out);
//   3. S = Q K^T: 12 tcgen05.mma kind::tf32 (M128, N = keys padded to 16, K8), both operands K-major from shared
//      memory, accumulator S in TMEM columns [0, NK);
//   4. softmax, thread = row: two passes over the row's TMEM columns (tcgen05.ld) - exact max-subtracted softmax
//      like the reference, base 2 - causal chunks only (warp q reads 32 (q+1) keys); P is split and written back
//      to TMEM (tcgen05.st): P_hi over the S columns, P_lo next to them;
//   5. O = P V: A = P from TMEM, B = V from shared memory as an MN-major operand: the TMA tile [key][dim], loaded with
//      the 32-B-atom flavour of the 128-B swizzle (the only MN-major layout kind::tf32 takes), is the canonical atom
//      sequence - 4 keys of 128 B per atom, two atoms per K = 8 step - so V is never transposed; N = 32;
//   6. epilogue: O / rowsum -> swizzled staging chunk -> TMA store (rows past L are clipped by the map).
// The steps of one item are serial inside a CTA (two __syncthreads and four mbarriers), so TWO CTAs share an SM
// (112 KB of shared memory and 256 TMEM columns each): one does its softmax on the CUDA cores while the other's
// MMAs run.  The next item's Q/K tiles are requested as soon as S is done, its V tile as soon as O is done.
#include <stdio.h>
#include <stdlib.h>

#include "qx_tc05.cuh"

namespace qx {

namespace {

constexpr int kHd = 32;                    // head dimension = one 128-B row
constexpr int kRows = 128;                 // UMMA M: query rows (and key rows of a tile)
constexpr int kMaxKeys = 112;              // TMEM budget: S/P_hi [0,112) | P_lo [112,224) | O [224,256)
constexpr int kTile = kRows * kHd;         // floats per tile (16 KB)
constexpr uint32_t kColPlo = kMaxKeys, kColO = 2 * kMaxKeys;
constexpr uint32_t kColQ = kColPlo;            // Q_hi [112,144) | Q_lo [144,176) while S is computed: P_lo is written after it

// One trajectory per tile: K and V are boxes of kMaxKeys = 112 token rows, which buys a landing tile of its own for V:
// the next item's V is requested as soon as this item's has been split, a whole item ahead of its use.  (With V landing
// in v_hi its request had to wait for the item's P V product, and the TMA latency sat on every item's critical path.)
// The epilogue's staging chunks (quarter w: 32 rows x 128 B in the TMA swizzle layout) then no longer fit next to the
// tiles: they ARE rows 32 w .. 32 w + 31 of v_hi, free between the item's P V product and the next item's V split -
// which the same two warps perform on exactly these rows, after the lane that issued the store has seen it read (v_hi
// keeps 128 rows for that: rows 112-127 are staging only).
// PAIR: 128-row tiles, V lands in v_hi and the staging chunks are their own.
template <bool PAIR>
struct AttnSmemT {
  static constexpr int kKeyRows = PAIR ? kRows : kMaxKeys;
  float q_hi[kTile], q_lo[PAIR ? kTile : 256];   // single trajectory: q_hi is the landing tile only, the halves of Q live in TMEM
  float k_hi[kKeyRows * kHd], k_lo[kKeyRows * kHd], v_hi[kTile], v_lo[kKeyRows * kHd];
  float4 stage[PAIR ? 4 : 1][32][8];       // PAIR: per quarter 32 rows x 32 dims in the TMA 128-B swizzle layout
  float v_raw[PAIR ? 4 : kKeyRows * kHd];
  float pmax[PAIR ? 4 : 256], psum[PAIR ? 4 : 256];   // the two threads of a row exchange their partial maximum / sum
  uint64_t bar_qk, bar_v, bar_s, bar_o, bar_vfree;
  uint32_t tmem_base;
};
static_assert(2 * (sizeof(AttnSmemT<false>) + 1024) <= 228 * 1024 && 2 * (sizeof(AttnSmemT<true>) + 1024) <= 228 * 1024,
              "two attention CTAs must fit the shared memory of an SM");

// instruction descriptors: D = F32, A = B = TF32, M = 128; N and the B major differ
constexpr uint32_t idesc_tf32(int n, bool b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (b_mn_major ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
}

// QK16: Q and K arrive as fp16 (hi, lo) pairs - D = F32, A = B = F16 (format bits 0), both K-major
constexpr uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24); }

// The operand split of this kernel: hi = x rounded to the nearest TF32, lo = x - hi, exact and NOT rounded: the tensor
// core reads the upper 19 bits of a kind::tf32 operand and ignores the rest, so lo enters the products truncated to its
// top 11 bits - an error of 2^-10 |lo| <= 2^-21 |x| where the rounded lo of tf32_split() has 2^-22 |x|, for three
// instructions per element instead of five (the splits were 45 % of the kernel's instructions and it is bound by them).
__device__ __forceinline__ void attn_split(float v, float& hi, float& lo) {
  hi = tf32_rn(v);
#ifdef QX_ATTN_SPLIT_RN
  lo = tf32_rn(v - hi);
#else
  lo = v - hi;
#endif
}

template <int W>
struct TmemRow;      // tcgen05.ld / st of W consecutive columns of this thread's lane (32x32b shape)
template <>
struct TmemRow<32> {
  static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
  }
  static __device__ __forceinline__ void st(uint32_t taddr, const uint32_t (&r)[32]) { tmem_st_32x32b_x32(taddr, r); }
};
template <>
struct TmemRow<16> {
  static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
  }
  static __device__ __forceinline__ void st(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
  }
};
template <>
struct TmemRow<8> {
  static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
  }
  static __device__ __forceinline__ void st(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
  }
};
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Two threads per tile row (half = 0 / 1 takes the 16-B pieces 4 half .. 4 half + 3 of the row's 128 B).
// hi in place, lo into the second tile; pieces visited in swizzle order (bank-conflict free)
__device__ __forceinline__ void split_tile_row(float* hi_tile, float* lo_tile, int row, int half) {
  float4* hi4 = reinterpret_cast<float4*>(hi_tile) + row * 8;
  float4* lo4 = reinterpret_cast<float4*>(lo_tile) + row * 8;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int c = (4 * half + jj) ^ (row & 7);
    const float4 v = hi4[c];
    float4 h, l;
    attn_split(v.x, h.x, l.x);
    attn_split(v.y, h.y, l.y);
    attn_split(v.z, h.z, l.z);
    attn_split(v.w, h.w, l.w);
    hi4[c] = h;
    lo4[c] = l;
  }
}

// The same from a landing tile of its own: raw -> (hi, lo)
__device__ __forceinline__ void split_raw_row(const float* raw_tile, float* hi_tile, float* lo_tile, int row, int half) {
  const float4* raw4 = reinterpret_cast<const float4*>(raw_tile) + row * 8;
  float4* hi4 = reinterpret_cast<float4*>(hi_tile) + row * 8;
  float4* lo4 = reinterpret_cast<float4*>(lo_tile) + row * 8;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int c = (4 * half + jj) ^ (row & 7);
    const float4 v = raw4[c];
    float4 h, l;
    attn_split(v.x, h.x, l.x);
    attn_split(v.y, h.y, l.y);
    attn_split(v.z, h.z, l.z);
    attn_split(v.w, h.w, l.w);
    hi4[c] = h;
    lo4[c] = l;
  }
}

// The Q tile is scaled by log2(e) / sqrt(32) on the way (the scores then are base-2 logits)
__device__ __forceinline__ void split_q_row(float* hi_tile, float* lo_tile, int row, int half, float scale) {
  float4* hi4 = reinterpret_cast<float4*>(hi_tile) + row * 8;
  float4* lo4 = reinterpret_cast<float4*>(lo_tile) + row * 8;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int c = (4 * half + jj) ^ (row & 7);
    float4 v = hi4[c];
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    float4 h, l;
    attn_split(v.x, h.x, l.x);
    attn_split(v.y, h.y, l.y);
    attn_split(v.z, h.z, l.z);
    attn_split(v.w, h.w, l.w);
    hi4[c] = h;
    lo4[c] = l;
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// the two threads of a row (warps w and w + 4) meet on named barrier 1 + (w % 4)
__device__ __forceinline__ void pair_sync(int wq) {
  switch (wq) {      // immediate barrier numbers: the kernel then reserves 5 of the 16 barriers, not all of them
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
  __syncwarp();
}

// Softmax of one query row whose scores sit in this thread's TMEM lane; the row's two threads take the lower / upper
// half of the columns of every chunk: Q full chunks of 32 keys (16 columns each), then the chunk that holds the
// diagonal (WD = 32 or 16 columns wide, HW = WD / 2 per thread; key 32 Q + t is unmasked for t <= lrev).  One sweep:
// all loads first (their latencies overlap), partial maxima exchanged through shared memory, P = 2^(s - m), partial
// row sum, (hi, lo) halves back into TMEM (P_hi over the scores).  Returns this thread's partial row sum.
template <int Q, int WD>
__device__ __forceinline__ float softmax_half_row(uint32_t tmem_row, int lrev, int half, int wq, float* pmax, int tid) {
  constexpr int HW = WD / 2;
  uint32_t r[Q + 1][16];
#pragma unroll
  for (int c = 0; c < Q; ++c) TmemRow<16>::ld(tmem_row + 32 * c + 16 * half, r[c]);
  TmemRow<HW>::ld(tmem_row + 32 * Q + HW * half, r[Q]);
  tmem_wait_ld();
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};      // four chains: the reductions are latency-bound
#pragma unroll
  for (int c = 0; c < Q; ++c) {
#pragma unroll
    for (int t = 0; t < 16; ++t) m4[t & 3] = fmaxf(m4[t & 3], __uint_as_float(r[c][t]));
  }
#pragma unroll
  for (int t = 0; t < HW; ++t) {
    if (HW * half + t <= lrev) m4[t & 3] = fmaxf(m4[t & 3], __uint_as_float(r[Q][t]));
  }
  float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
  pmax[tid] = m;
  pair_sync(wq);
  m = fmaxf(m, pmax[tid ^ 128]);           // key 0 is never masked, so the lower half's maximum is always finite
  float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c <= Q; ++c) {
    uint32_t lo[16];
#pragma unroll
    for (int t = 0; t < ((c < Q) ? 16 : HW); ++t) {
      float p = ex2_approx(__uint_as_float(r[c][t]) - m);
      if (c == Q && HW * half + t > lrev) p = 0.f;
      l4[t & 3] += p;
      float h, lw;
      attn_split(p, h, lw);
      r[c][t] = __float_as_uint(h);
      lo[t] = __float_as_uint(lw);
    }
    if (c < Q) {
      TmemRow<16>::st(tmem_row + 32 * c + 16 * half, r[c]);
      TmemRow<16>::st(tmem_row + kColPlo + 32 * c + 16 * half, lo);
    } else {
      TmemRow<HW>::st(tmem_row + 32 * c + HW * half, r[c]);
      TmemRow<HW>::st(tmem_row + kColPlo + 32 * c + HW * half, lo);
    }
  }
  return (l4[0] + l4[1]) + (l4[2] + l4[3]);
}
// ---- PAIR mode (L <= 64, the cart-pole's 61 tokens): TWO trajectories share the M = 128 tile ---------------------------
// rows 0-63 = trajectory 2p, rows 64-127 = trajectory 2p + 1 (a TMA box of 64 tokens x 2 trajectories lands exactly so).
// S = Q [K_a ; K_b]^T is one N = 128 product; a row of trajectory t reads its scores in columns [64 t, 64 t + 64) and
// leaves P_hi in columns [0, 64), P_lo in [64, 128) whatever t is (all loads of a row precede its stores); the second
// product runs once per trajectory (B = V_a into O_a, B = V_b into O_b: a row reads the one that is its own).
// TMEM: S / P [0,128) | O_a [128,160) | O_b [160,192) | pair exchange [192,196) - the K_lo tile has no spare rows here.
constexpr uint32_t kPairPlo = 64, kPairO = 128, kPairX = 192;
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)) : "memory");
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {      // completes the load
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return __uint_as_float(v);
}
// softmax_half_row for a PAIR tile: Q (0 / 1) full chunks, then the 32-wide chunk with the diagonal; the row's two
// threads swap their partial maxima through two TMEM columns of their (shared) lane
template <int Q>
__device__ __forceinline__ float softmax_pair_row(uint32_t tmem_row, uint32_t s_col, int lrev, int half, int wq) {
  uint32_t r[Q + 1][16];
#pragma unroll
  for (int c = 0; c <= Q; ++c) TmemRow<16>::ld(tmem_row + s_col + 32 * c + 16 * half, r[c]);
  tmem_wait_ld();
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int c = 0; c < Q; ++c) {
#pragma unroll
    for (int t = 0; t < 16; ++t) m4[t & 3] = fmaxf(m4[t & 3], __uint_as_float(r[c][t]));
  }
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    if (16 * half + t <= lrev) m4[t & 3] = fmaxf(m4[t & 3], __uint_as_float(r[Q][t]));
  }
  float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
  tmem_st1(tmem_row + kPairX + half, m);
  tmem_wait_st();
  tcgen05_fence_before();
  pair_sync(wq);
  tcgen05_fence_after();
  m = fmaxf(m, tmem_ld1(tmem_row + kPairX + (half ^ 1)));      // -inf from an upper half that is fully masked
  float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c <= Q; ++c) {
    uint32_t lo[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      float p = ex2_approx(__uint_as_float(r[c][t]) - m);
      if (c == Q && 16 * half + t > lrev) p = 0.f;
      l4[t & 3] += p;
      float h, lw;
      attn_split(p, h, lw);
      r[c][t] = __float_as_uint(h);
      lo[t] = __float_as_uint(lw);
    }
    TmemRow<16>::st(tmem_row + 32 * c + 16 * half, r[c]);
    TmemRow<16>::st(tmem_row + kPairPlo + 32 * c + 16 * half, lo);
  }
  return (l4[0] + l4[1]) + (l4[2] + l4[3]);
}

template <int HW>
__device__ __forceinline__ void chunk_zero(uint32_t tmem_row, int col0) {
  uint32_t z[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) z[t] = 0u;
  TmemRow<HW>::st(tmem_row + col0, z);
  TmemRow<HW>::st(tmem_row + kColPlo + col0, z);
}

// -DQX_ATTN_TRACE (experiments only; QX_EXTRA_NVCC_FLAGS of build.sh): CTA 0 sums the cycles each phase of an item takes
// as seen by warp 0 (MMA / TMA issuer, quarter 0) and warp 7 (quarter 3, the longest causal rows); the launcher prints
// them to stderr.
#ifdef QX_ATTN_TRACE
__device__ long long g_attn_trace[2][16];
#define QX_TRACE_DECL long long tr_t = clock64(), tr_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define QX_TRACE(i) do { const long long tr_n = clock64(); tr_acc[i] += tr_n - tr_t; tr_t = tr_n; } while (0)
#define QX_TRACE_DUMP do { if (blockIdx.x == 0 && (tid == 0 || tid == 224)) { for (int i = 0; i < 16; ++i) g_attn_trace[tid ? 1 : 0][i] = tr_acc[i]; } } while (0)
#else
#define QX_TRACE_DECL
#define QX_TRACE(i)
#define QX_TRACE_DUMP
#endif

}  // namespace

// grid: min(B * nhead, 2 * SMs) persistent CTAs of 256 threads; item = trajectory * nhead + head.
// Two threads per query row (warps w and w + 4 share TMEM lane quarter w % 4 and split every chunk's columns), so an
// SM holds 16 warps: the softmax and the operand splits are latency-bound CUDA-core work and need the occupancy.
// (Tried and dropped, twice: token quarters in reverse lane order in every second CTA / in the second CTA to arrive on
// an SM, so that the long causal rows of two co-resident CTAs do not share a lane quarter - 0.63 ms against 0.57 ms.
// The hardware already spreads the warps of the two CTAs evenly: smsp__inst_executed differs by 3 % across the
// schedulers (profiles/r2_ncu_full_attention_summary.txt).)
// QK16: the Q and K columns of qkv hold fp16 (hi, lo) pairs written by the QKV projection (tf_gemm_ta_kernel,
// kGemmPackQk; Q already scaled): per head slice and token 64 B of hi halves, then 64 B of lo halves.  The same TMA box
// (32 "floats" x tokens, 128-B swizzle) lands them as ONE K-major fp16 tile of 64 k per row - k 0-31 the hi halves,
// k 32-63 the lo halves - so the three products hi hi + hi lo + lo hi are six kind::f16 MMAs (K = 16 each) whose
// descriptors start 0 / 32 / 64 / 96 B into the swizzle atom.  Neither tile is touched by a thread: the Q / K splits
// (a quarter of the kernel's CUDA-core instructions) and the Q detour through tensor memory are gone.
template <bool PAIR, bool QK16>
__global__ void __launch_bounds__(256, 2)
tf_attention_umma_kernel(const __grid_constant__ CUtensorMap tmQkv, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmOut, float* __restrict__ out_rows, int n_items, int n_traj, int nhead, int L,
                         int d, int row_lo, int dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  using Smem = AttnSmemT<PAIR>;
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wq = warp & 3, half = warp >> 2;        // TMEM lane quarter; which half of the row's columns
  const int row = wq * 32 + lane;                   // tile row / TMEM lane of this thread
  const int NK = PAIR ? kRows : (L + 15) & ~15;     // keys padded to the UMMA N granularity (PAIR: 2 x 64)
  const int nks = (L + 7) >> 3;                     // K = 8 steps of the second product
  const uint32_t idesc_s = QK16 ? idesc_f16(NK) : idesc_tf32(NK, false), idesc_o = idesc_tf32(kHd, true);
  const int th = PAIR ? (wq >> 1) : 0;              // PAIR: which trajectory of the tile this row belongs to
  const int qrev = PAIR ? (wq & 1) : wq, lrev = lane;   // this thread's token is 32 qrev + lrev

  if (tid == 0) {
    if (smem_u32(smem_raw) & 1023u) __trap();       // the swizzled tiles need 1 KB alignment
    mbar_init(&s.bar_qk, 1);
    mbar_init(&s.bar_v, 1);
    mbar_init(&s.bar_s, 1);
    mbar_init(&s.bar_o, 1);
    mbar_init(&s.bar_vfree, 8);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "n"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s.tmem_base;
  const uint32_t tmem_row = tmem + ((uint32_t)(wq * 32) << 16);       // this warp's lane quarter
  const float scale_log2 = rsqrtf((float)kHd) * 1.4426950408889634f;  // softmax in base 2: exp(x) = 2^(x log2 e)

  QX_TRACE_DECL;
  auto load_qk = [&](int item) {
    const int b = (PAIR ? 2 : 1) * (item / nhead), h = item % nhead;
    mbar_arrive_expect_tx(&s.bar_qk, (kTile + Smem::kKeyRows * kHd) * 4);
    tma_load_3d(&tmQkv, &s.bar_qk, s.q_hi, h * kHd, 0, b);
    tma_load_3d(&tmK, &s.bar_qk, s.k_hi, d + h * kHd, 0, b);
  };
  auto load_v = [&](int item) {
    const int b = (PAIR ? 2 : 1) * (item / nhead), h = item % nhead;
    mbar_arrive_expect_tx(&s.bar_v, Smem::kKeyRows * kHd * 4);
    tma_load_3d(&tmV, &s.bar_v, PAIR ? s.v_hi : s.v_raw, 2 * d + h * kHd, 0, b);
  };
  float* const pmax = s.pmax;
  float* const psum = s.psum;
  auto split_qk = [&](uint32_t ph) {
    mbar_wait(&s.bar_qk, ph);
    QX_TRACE(11);    // wait Q, K
    if (!(dbg & 8)) {
      if (PAIR || row < kMaxKeys) split_tile_row(s.k_hi, s.k_lo, row, half);
      if (PAIR) split_q_row(s.q_hi, s.q_lo, row, half, scale_log2);
    }
    fence_proxy_async();
  };
  // Single trajectory: Q is the A operand of the first product FROM TENSOR MEMORY (lane = query row, column = dim): an
  // M128 N112 K8 product with both operands in shared memory reads 7.5 KB in its 56 cycles - more than the 128 B/clk an
  // SM has, while the V split of this CTA and whatever the other CTA does want the same banks.  The halves sit in the
  // columns of P_lo, which is dead from the end of one item's P V product to the end of the next item's S.
  auto split_q_tmem = [&]() {
    const float4* src = reinterpret_cast<const float4*>(s.q_hi) + row * 8;
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float4 v = src[(4 * half + jj) ^ (row & 7)];
      float h, l;
      attn_split(v.x * scale_log2, h, l); hi[4 * jj + 0] = __float_as_uint(h); lo[4 * jj + 0] = __float_as_uint(l);
      attn_split(v.y * scale_log2, h, l); hi[4 * jj + 1] = __float_as_uint(h); lo[4 * jj + 1] = __float_as_uint(l);
      attn_split(v.z * scale_log2, h, l); hi[4 * jj + 2] = __float_as_uint(h); lo[4 * jj + 2] = __float_as_uint(l);
      attn_split(v.w * scale_log2, h, l); hi[4 * jj + 3] = __float_as_uint(h); lo[4 * jj + 3] = __float_as_uint(l);
    }
    __syncwarp();        // reconverge after the per-lane spins: tcgen05.st / wait::st are .sync.aligned
    TmemRow<16>::st(tmem_row + kColQ + 16 * half, hi);
    TmemRow<16>::st(tmem_row + kColQ + kHd + 16 * half, lo);
    tmem_wait_st();
  };
  if ((int)blockIdx.x < n_items) {
    if (tid == 0) {
      load_qk(blockIdx.x);
      load_v(blockIdx.x);
    }
    if (!QK16) {
      split_qk(0);
      if (!PAIR) split_q_tmem();
    }
  }
  tcgen05_fence_before();
  __syncthreads();

  uint32_t it = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
    const uint32_t ph = it & 1u;
    const int h = item % nhead;
    const int b = PAIR ? 2 * (item / nhead) + th : item / nhead;      // this row's trajectory
    const int next = item + gridDim.x;
    QX_TRACE(10);    // loop top

    // ---- S = (scaled Q) K^T ---------------------------------------------------------------------------------------
    if (QK16 && warp == 0) {
      mbar_wait(&s.bar_qk, ph);          // the tiles are MMA-ready as TMA wrote them (async proxy both ways: no fence)
      __syncwarp();
      tcgen05_fence_after();
      const uint64_t dq = umma_desc_sw128(s.q_hi), dk = umma_desc_sw128(s.k_hi);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k) {
          const uint64_t hi = (uint64_t)((k * 16 * 2) >> 4), lo = hi + (uint64_t)((kHd * 2) >> 4);   // 16-B units
          mma_f16(tmem, dq + lo, dk + hi, idesc_s, k ? 1u : 0u);
          mma_f16(tmem, dq + hi, dk + lo, idesc_s, 1u);
          mma_f16(tmem, dq + hi, dk + hi, idesc_s, 1u);
        }
        tcgen05_commit(&s.bar_s);
      }
      __syncwarp();
    } else if (warp == 0) {
      tcgen05_fence_after();
      const uint64_t dq_hi = umma_desc_sw128(s.q_hi), dq_lo = umma_desc_sw128(s.q_lo);
      const uint64_t dk_hi = umma_desc_sw128(s.k_hi), dk_lo = umma_desc_sw128(s.k_lo);
      const uint32_t aq_hi = tmem + kColQ, aq_lo = aq_hi + kHd;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kHd / 8; ++k) {
          const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
          if (PAIR) {
            mma_tf32(tmem, dq_lo + adv, dk_hi + adv, idesc_s, k ? 1u : 0u);
            mma_tf32(tmem, dq_hi + adv, dk_lo + adv, idesc_s, 1u);
            mma_tf32(tmem, dq_hi + adv, dk_hi + adv, idesc_s, 1u);
          } else {
            mma_tf32_ta(tmem, aq_lo + 8 * k, dk_hi + adv, idesc_s, k ? 1u : 0u);
            mma_tf32_ta(tmem, aq_hi + 8 * k, dk_lo + adv, idesc_s, 1u);
            mma_tf32_ta(tmem, aq_hi + 8 * k, dk_hi + adv, idesc_s, 1u);
          }
        }
        tcgen05_commit(&s.bar_s);
      }
      __syncwarp();
    }
    QX_TRACE(0);     // S issue
    // V is split while the tensor core works on S
    mbar_wait(&s.bar_v, ph);
    QX_TRACE(1);     // wait V
    if (PAIR) {
      if (!(dbg & 8)) split_tile_row(s.v_hi, s.v_lo, row, half);
    } else {
      // v_hi / v_lo are free (every thread saw the previous item's P V complete before its epilogue) once the output
      // store that staged in this pair's v_hi rows has read them.  The landing tile goes back to TMA when all eight warps
      // have consumed their rows: the stores of the split depend on every load.
      if (half == 0 && lane == 0) bulk_wait_read<0>();
      pair_sync(wq);
      if (row < kMaxKeys && !(dbg & 8)) split_raw_row(s.v_raw, s.v_hi, s.v_lo, row, half);
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.bar_vfree);
      if (tid == 0 && next < n_items) {
        mbar_wait(&s.bar_vfree, ph);
        load_v(next);
      }
    }
    fence_proxy_async();
    QX_TRACE(2);     // V split (+ request of the next V)

    mbar_wait(&s.bar_s, ph);
    QX_TRACE(3);     // wait S
    tcgen05_fence_after();
    if (tid == 0 && next < n_items) load_qk(next);          // Q, K tiles are free: request the next item's
    __syncwarp();                // reconverge (per-lane spins, thread 0's detour): the tcgen05.ld / st below are .sync.aligned
    // ---- softmax (two threads per row) --------------------------------------------------------------------------
    // a quarter works when it holds a token (< L) that is asked for (>= row_lo: only the last layer asks for fewer)
    const bool active = qrev * 32 < L && qrev * 32 + 31 >= row_lo;
    if (PAIR && active) {
      const float l = qrev == 0 ? softmax_pair_row<0>(tmem_row, 64u * th, lrev, half, wq)
                                : softmax_pair_row<1>(tmem_row, 64u * th, lrev, half, wq);
      tmem_st1(tmem_row + kPairX + 2 + half, l);
      if (qrev == 0) {                                        // keys 32-63 are beyond this quarter's causal range: P = 0
        uint32_t z[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) z[t] = 0u;
        TmemRow<16>::st(tmem_row + 32 + 16 * half, z);
        TmemRow<16>::st(tmem_row + kPairPlo + 32 + 16 * half, z);
      }
      tmem_wait_st();
    } else if (active && !(dbg & 4)) {
      const bool wide = NK - 32 * qrev >= 32;
      float l;
      switch (qrev * 2 + (wide ? 1 : 0)) {
        case 0: l = softmax_half_row<0, 16>(tmem_row, lrev, half, wq, pmax, tid); break;
        case 1: l = softmax_half_row<0, 32>(tmem_row, lrev, half, wq, pmax, tid); break;
        case 2: l = softmax_half_row<1, 16>(tmem_row, lrev, half, wq, pmax, tid); break;
        case 3: l = softmax_half_row<1, 32>(tmem_row, lrev, half, wq, pmax, tid); break;
        case 4: l = softmax_half_row<2, 16>(tmem_row, lrev, half, wq, pmax, tid); break;
        case 5: l = softmax_half_row<2, 32>(tmem_row, lrev, half, wq, pmax, tid); break;
        case 6: l = softmax_half_row<3, 16>(tmem_row, lrev, half, wq, pmax, tid); break;
        default: l = softmax_half_row<3, 32>(tmem_row, lrev, half, wq, pmax, tid); break;
      }
      psum[tid] = l;
      for (int c = qrev + 1; 32 * c < NK; ++c) {             // keys beyond the quarter's causal range: P = 0
        if (NK - 32 * c >= 32) chunk_zero<16>(tmem_row, 32 * c + 16 * half);
        else chunk_zero<8>(tmem_row, 32 * c + 8 * half);
      }
      tmem_wait_st();
    }
    QX_TRACE(4);     // softmax
    tcgen05_fence_before();
    __syncthreads();
    QX_TRACE(5);     // CTA barrier
    // ---- O = P V ----------------------------------------------------------------------------------------------------
    if (warp == 0) {
      tcgen05_fence_after();
      if (PAIR) {
        if (elect_one()) {
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t a_hi = tmem + 8 * ks, a_lo = tmem + kPairPlo + 8 * ks;
#pragma unroll
            for (int t = 0; t < 2; ++t) {                     // O_t = P V_t: valid in the rows of trajectory t
              const float* vh = s.v_hi + (64 * t + ks * 8) * kHd;
              const float* vl = s.v_lo + (64 * t + ks * 8) * kHd;
              const uint64_t dv_hi = umma_desc_mn_sw128_32b(vh), dv_lo = umma_desc_mn_sw128_32b(vl);
              const uint32_t acc = tmem + kPairO + 32 * t;
              mma_tf32_ta(acc, a_lo, dv_hi, idesc_o, ks ? 1u : 0u);
              mma_tf32_ta(acc, a_hi, dv_lo, idesc_o, 1u);
              mma_tf32_ta(acc, a_hi, dv_hi, idesc_o, 1u);
            }
          }
          tcgen05_commit(&s.bar_o);
        }
      } else if (elect_one()) {
        for (int ks = 0; ks < ((dbg & 2) ? 0 : nks); ++ks) {
          const uint64_t dv_hi = umma_desc_mn_sw128_32b(s.v_hi + ks * 8 * kHd), dv_lo = umma_desc_mn_sw128_32b(s.v_lo + ks * 8 * kHd);
          const uint32_t a_hi = tmem + 8 * ks, a_lo = tmem + kColPlo + 8 * ks;
          mma_tf32_ta(tmem + kColO, a_lo, dv_hi, idesc_o, ks ? 1u : 0u);
          mma_tf32_ta(tmem + kColO, a_hi, dv_lo, idesc_o, 1u);
          mma_tf32_ta(tmem + kColO, a_hi, dv_hi, idesc_o, 1u);
        }
        tcgen05_commit(&s.bar_o);
      }
      __syncwarp();
    }
    QX_TRACE(12);    // O issue
    // the next item's Q, K are split while the tensor core works on O
    if (!QK16 && next < n_items) split_qk(ph ^ 1u);
    QX_TRACE(6);     // split of the next Q, K
    mbar_wait(&s.bar_o, ph);
    QX_TRACE(7);     // wait O
    tcgen05_fence_after();
    if (!QK16 && !PAIR && next < n_items && !(dbg & 8)) split_q_tmem();     // P_lo is dead: the next item's Q halves move in
    if (PAIR && tid == 0 && next < n_items) load_v(next);
    __syncwarp();
    // ---- epilogue: each thread 16 of the row's 32 output dims -----------------------------------------------------
    const uint32_t col_o = PAIR ? kPairO + 32 * th : kColO;
    auto row_sum = [&]() -> float {      // the partial sums of the row's two threads, lower half first
      if (PAIR) {
        const float l0 = tmem_ld1(tmem_row + kPairX + 2);
        return l0 + tmem_ld1(tmem_row + kPairX + 3);
      }
      return psum[tid & 127] + psum[(tid & 127) + 128];
    };
    if (active && row_lo == 0 && !(dbg & 16)) {
      uint32_t r[16];
      TmemRow<16>::ld(tmem_row + col_o + 16 * half, r);
      if (PAIR) {
        if (half == 0 && lane == 0) bulk_wait_read<0>();     // the previous item's store has read the staging chunk
        pair_sync(wq);
      }
      tmem_wait_ld();
      const float inv = 1.f / row_sum();
      float4 (*stg)[8] = PAIR ? s.stage[wq] : reinterpret_cast<float4 (*)[8]>(s.v_hi + wq * 32 * kHd);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float4 o;
        o.x = __uint_as_float(r[4 * jj + 0]) * inv;
        o.y = __uint_as_float(r[4 * jj + 1]) * inv;
        o.z = __uint_as_float(r[4 * jj + 2]) * inv;
        o.w = __uint_as_float(r[4 * jj + 3]) * inv;
        stg[lrev][(4 * half + jj) ^ (lrev & 7)] = o;
      }
      fence_proxy_async();
      pair_sync(wq);
      if (half == 0 && lane == 0 && b < n_traj) {
        tma_store_3d(&tmOut, stg, h * kHd, qrev * 32, b);
        bulk_commit();
      }
    } else if (active) {
      // last layer: only the rows row_lo .. L-1 are wanted, compactly ([B][L - row_lo][d]); the quarter that straddles
      // row_lo would need a store box starting at a negative row - which the hardware rejects (illegal instruction) - so
      // these few rows go out as plain 16-B stores
      uint32_t r[16];
      TmemRow<16>::ld(tmem_row + col_o + 16 * half, r);
      tmem_wait_ld();
      const int tok = qrev * 32 + lrev;
      const float rsum = row_sum();
      if (tok >= row_lo && tok < L && b < n_traj) {
        const float inv = 1.f / rsum;
        float4* dst = reinterpret_cast<float4*>(out_rows + ((size_t)b * (L - row_lo) + (tok - row_lo)) * d + h * kHd + 16 * half);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float4 o;
          o.x = __uint_as_float(r[4 * jj + 0]) * inv;
          o.y = __uint_as_float(r[4 * jj + 1]) * inv;
          o.z = __uint_as_float(r[4 * jj + 2]) * inv;
          o.w = __uint_as_float(r[4 * jj + 3]) * inv;
          dst[jj] = o;
        }
      }
    }
    QX_TRACE(8);     // epilogue
    tcgen05_fence_before();      // the next item's MMAs overwrite S and O; its Q / K halves must be visible to them
    __syncthreads();
    QX_TRACE(9);     // CTA barrier
  }
  QX_TRACE_DUMP;
  if (half == 0 && lane == 0) bulk_wait_read<0>();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
  }
}

// =====================================================================================================================
// P16: q, k AND v arrive as fp16 (hi, lo) pairs (QKV projection, kGemmPackQk | kGemmPackV; q carries the softmax scale,
// v a factor 2^4 that keeps its lo halves out of the fp16 subnormals) and the probabilities are split into fp16 pairs
// too (times 2^8, folded into the exponent: P = 2^(s - m + 8)), so that BOTH products are kind::f16 and no operand is
// split by a thread any more:
//   S = Q K^T   six MMAs (K = 16), Q and K tiles exactly as TMA landed them (see QK16 above);
//   O = P V     A = P from tensor memory - a column holds two consecutive keys, P_hi in columns [0, NK/2), P_lo behind
//               them, both INSIDE the columns of the scores they replace (a row's two threads have loaded all their
//               scores before either stores: the partial maxima are exchanged in between) - and B = the V tile as an
//               MN-major operand in the plain 128-B swizzle: row = key, 64 fp16 along MN = hi dims | lo dims, so the hi
//               and lo halves are the same tile entered 0 / 64 B into the row; K = 16 keys per MMA, 3 x NK/16 MMAs.
//   epilogue    O / (rowsum * 2^4): rowsum is the sum of the scaled probabilities (2^8 P), the accumulator holds 2^12 O.
// V is double buffered (no split means the landing tile IS the operand): the next item's V is requested when this
// item's has arrived.  Shared memory 77 KB (single trajectory) / 81 KB (PAIR) per CTA, two CTAs per SM as before.
namespace {

template <bool PAIR>
struct AttnSmem16 {
  static constexpr int kKeyRows = PAIR ? kRows : kMaxKeys;
  float q[kTile];
  float k[kKeyRows * kHd];
  float v[2][kKeyRows * kHd];
  float4 stage[4][32][8];                // per quarter 32 rows x 32 dims in the TMA 128-B swizzle layout
  float pmax[256], psum[256];
  uint64_t bar_qk, bar_v, bar_s, bar_o;
  uint32_t tmem_base;
};
static_assert(2 * (sizeof(AttnSmem16<false>) + 1024) <= 228 * 1024 && 2 * (sizeof(AttnSmem16<true>) + 1024) <= 228 * 1024,
              "two attention CTAs must fit the shared memory of an SM");

constexpr uint32_t kPlo16 = kMaxKeys / 2;      // single trajectory: P_hi [0,56) | P_lo [56,112) | O [224,256)
constexpr uint32_t kPairPlo16 = 32;            // PAIR: P_hi [0,32) | P_lo [32,64) | O_a, O_b, exchange as in the fp32 kernel
constexpr float kPShift = 8.f;                 // P is kept times 2^8
constexpr float kPvUnscale = 1.f / 16.f;       // ... and V times 2^4: O = acc / (2^4 * sum of 2^8 P)

template <int W>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t* r);
template <>
__device__ __forceinline__ void tmem_st_cols<8>(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_st_cols<4>(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}

// n (16 or 8) probabilities of one thread -> n / 2 packed columns of P_hi and P_lo.  sc: the scores (registers of the
// sweep), m: row maximum minus the shift, first_masked: index of the first masked key among these n (n: none)
template <int N>
__device__ __forceinline__ void p16_chunk(const uint32_t* sc, float m, int first_masked, float (&l4)[4], uint32_t taddr_hi,
                                          uint32_t taddr_lo) {
  uint32_t hi[N / 2], lo[N / 2];
#pragma unroll
  for (int t = 0; t < N; t += 2) {
    float p0 = ex2_approx(__uint_as_float(sc[t]) - m);
    float p1 = ex2_approx(__uint_as_float(sc[t + 1]) - m);
    if (t >= first_masked) p0 = 0.f;
    if (t + 1 >= first_masked) p1 = 0.f;
    l4[t & 3] += p0;
    l4[(t + 1) & 3] += p1;
    f16_split2(p0, p1, hi[t >> 1], lo[t >> 1]);
  }
  tmem_st_cols<N / 2>(taddr_hi, hi);
  tmem_st_cols<N / 2>(taddr_lo, lo);
}

// softmax_half_row with fp16-pair probabilities (see there for the sweep); returns the partial sum of 2^8 P
template <int Q, int WD>
__device__ __forceinline__ float softmax_half_row16(uint32_t tmem_row, int lrev, int half, int wq, float* pmax, int tid) {
  constexpr int HW = WD / 2;
  uint32_t r[Q + 1][16];
#pragma unroll
  for (int c = 0; c < Q; ++c) TmemRow<16>::ld(tmem_row + 32 * c + 16 * half, r[c]);
  TmemRow<HW>::ld(tmem_row + 32 * Q + HW * half, r[Q]);
  tmem_wait_ld();
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int c = 0; c < Q; ++c) {
#pragma unroll
    for (int t = 0; t < 16; ++t) m4[t & 3] = fmaxf(m4[t & 3], __uint_as_float(r[c][t]));
  }
#pragma unroll
  for (int t = 0; t < HW; ++t) {
    if (HW * half + t <= lrev) m4[t & 3] = fmaxf(m4[t & 3], __uint_as_float(r[Q][t]));
  }
  float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
  pmax[tid] = m;
  pair_sync(wq);                            // both threads of the row have loaded ALL their scores: P may overwrite them
  m = fmaxf(m, pmax[tid ^ 128]) - kPShift;
  float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < Q; ++c) {
    const uint32_t col = 16 * c + 8 * half;
    p16_chunk<16>(r[c], m, 16, l4, tmem_row + col, tmem_row + kPlo16 + col);
  }
  {
    const uint32_t col = 16 * Q + (HW / 2) * half;
    int fm = lrev + 1 - HW * half;          // keys HW half + t with t > lrev - HW half are masked
    fm = fm < 0 ? 0 : fm;
    p16_chunk<HW>(r[Q], m, fm, l4, tmem_row + col, tmem_row + kPlo16 + col);
  }
  return (l4[0] + l4[1]) + (l4[2] + l4[3]);
}

template <int Q>
__device__ __forceinline__ float softmax_pair_row16(uint32_t tmem_row, uint32_t s_col, int lrev, int half, int wq) {
  uint32_t r[Q + 1][16];
#pragma unroll
  for (int c = 0; c <= Q; ++c) TmemRow<16>::ld(tmem_row + s_col + 32 * c + 16 * half, r[c]);
  tmem_wait_ld();
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int c = 0; c < Q; ++c) {
#pragma unroll
    for (int t = 0; t < 16; ++t) m4[t & 3] = fmaxf(m4[t & 3], __uint_as_float(r[c][t]));
  }
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    if (16 * half + t <= lrev) m4[t & 3] = fmaxf(m4[t & 3], __uint_as_float(r[Q][t]));
  }
  float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
  tmem_st1(tmem_row + kPairX + half, m);
  tmem_wait_st();
  tcgen05_fence_before();
  pair_sync(wq);
  tcgen05_fence_after();
  m = fmaxf(m, tmem_ld1(tmem_row + kPairX + (half ^ 1))) - kPShift;
  float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < Q; ++c) {
    const uint32_t col = 16 * c + 8 * half;
    p16_chunk<16>(r[c], m, 16, l4, tmem_row + col, tmem_row + kPairPlo16 + col);
  }
  {
    const uint32_t col = 16 * Q + 8 * half;
    int fm = lrev + 1 - 16 * half;
    fm = fm < 0 ? 0 : fm;
    p16_chunk<16>(r[Q], m, fm, l4, tmem_row + col, tmem_row + kPairPlo16 + col);
  }
  return (l4[0] + l4[1]) + (l4[2] + l4[3]);
}

}  // namespace

template <bool PAIR>
__global__ void __launch_bounds__(256, 2)
tf_attention_p16_kernel(const __grid_constant__ CUtensorMap tmQkv, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmOut,
                        float* __restrict__ out_rows, int n_items, int n_traj, int nhead, int L, int d, int row_lo) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  using Smem = AttnSmem16<PAIR>;
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wq = warp & 3, half = warp >> 2;
  const int NK = PAIR ? kRows : (L + 15) & ~15;
  const int nks = PAIR ? 4 : NK >> 4;               // K = 16 steps of the second product (per trajectory)
  const uint32_t idesc_s = idesc_f16(NK), idesc_o = idesc_f16(kHd) | (1u << 16);      // bit 16: B is MN-major
  const int th = PAIR ? (wq >> 1) : 0;
  const int qrev = PAIR ? (wq & 1) : wq, lrev = lane;

  if (tid == 0) {
    if (smem_u32(smem_raw) & 1023u) __trap();
    mbar_init(&s.bar_qk, 1);
    mbar_init(&s.bar_v, 1);
    mbar_init(&s.bar_s, 1);
    mbar_init(&s.bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "n"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s.tmem_base;
  const uint32_t tmem_row = tmem + ((uint32_t)(wq * 32) << 16);

  auto load_qk = [&](int item) {
    const int b = (PAIR ? 2 : 1) * (item / nhead), h = item % nhead;
    mbar_arrive_expect_tx(&s.bar_qk, (kTile + Smem::kKeyRows * kHd) * 4);
    tma_load_3d(&tmQkv, &s.bar_qk, s.q, h * kHd, 0, b);
    tma_load_3d(&tmK, &s.bar_qk, s.k, d + h * kHd, 0, b);
  };
  auto load_v = [&](int item, uint32_t buf) {
    const int b = (PAIR ? 2 : 1) * (item / nhead), h = item % nhead;
    mbar_arrive_expect_tx(&s.bar_v, Smem::kKeyRows * kHd * 4);
    tma_load_3d(&tmV, &s.bar_v, s.v[buf], 2 * d + h * kHd, 0, b);
  };
  float* const pmax = s.pmax;
  float* const psum = s.psum;
  if ((int)blockIdx.x < n_items && tid == 0) {
    load_qk(blockIdx.x);
    load_v(blockIdx.x, 0);
  }

  uint32_t it = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
    const uint32_t ph = it & 1u;
    const int h = item % nhead;
    const int b = PAIR ? 2 * (item / nhead) + th : item / nhead;
    const int next = item + gridDim.x;

    // ---- S = (scaled Q) K^T: both tiles as TMA wrote them ------------------------------------------------------------
    if (warp == 0) {
      mbar_wait(&s.bar_qk, ph);
      __syncwarp();
      tcgen05_fence_after();
      const uint64_t dq = umma_desc_sw128(s.q), dk = umma_desc_sw128(s.k);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k) {
          const uint64_t hi = (uint64_t)((k * 16 * 2) >> 4), lo = hi + (uint64_t)((kHd * 2) >> 4);
          mma_f16(tmem, dq + lo, dk + hi, idesc_s, k ? 1u : 0u);
          mma_f16(tmem, dq + hi, dk + lo, idesc_s, 1u);
          mma_f16(tmem, dq + hi, dk + hi, idesc_s, 1u);
        }
        tcgen05_commit(&s.bar_s);
      }
      __syncwarp();
    }
    mbar_wait(&s.bar_s, ph);
    tcgen05_fence_after();
    if (tid == 0 && next < n_items) load_qk(next);          // Q, K tiles are free
    __syncwarp();
    // ---- softmax ----------------------------------------------------------------------------------------------------
    const bool active = qrev * 32 < L && qrev * 32 + 31 >= row_lo;
    if (PAIR && active) {
      const float l = qrev == 0 ? softmax_pair_row16<0>(tmem_row, 64u * th, lrev, half, wq)
                                : softmax_pair_row16<1>(tmem_row, 64u * th, lrev, half, wq);
      tmem_st1(tmem_row + kPairX + 2 + half, l);
      if (qrev == 0) {                                        // keys 32-63 are beyond this quarter's causal range: P = 0
        const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st_cols<8>(tmem_row + 16 + 8 * half, z);
        tmem_st_cols<8>(tmem_row + kPairPlo16 + 16 + 8 * half, z);
      }
      tmem_wait_st();
    } else if (active) {
      const bool wide = NK - 32 * qrev >= 32;
      float l;
      switch (qrev * 2 + (wide ? 1 : 0)) {
        case 0: l = softmax_half_row16<0, 16>(tmem_row, lrev, half, wq, pmax, tid); break;
        case 1: l = softmax_half_row16<0, 32>(tmem_row, lrev, half, wq, pmax, tid); break;
        case 2: l = softmax_half_row16<1, 16>(tmem_row, lrev, half, wq, pmax, tid); break;
        case 3: l = softmax_half_row16<1, 32>(tmem_row, lrev, half, wq, pmax, tid); break;
        case 4: l = softmax_half_row16<2, 16>(tmem_row, lrev, half, wq, pmax, tid); break;
        case 5: l = softmax_half_row16<2, 32>(tmem_row, lrev, half, wq, pmax, tid); break;
        case 6: l = softmax_half_row16<3, 16>(tmem_row, lrev, half, wq, pmax, tid); break;
        default: l = softmax_half_row16<3, 32>(tmem_row, lrev, half, wq, pmax, tid); break;
      }
      psum[tid] = l;
      const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      for (int c = qrev + 1; 32 * c < NK; ++c) {             // keys beyond the quarter's causal range: P = 0
        if (NK - 32 * c >= 32) {
          tmem_st_cols<8>(tmem_row + 16 * c + 8 * half, z);
          tmem_st_cols<8>(tmem_row + kPlo16 + 16 * c + 8 * half, z);
        } else {
          tmem_st_cols<4>(tmem_row + 16 * c + 4 * half, z);
          tmem_st_cols<4>(tmem_row + kPlo16 + 16 * c + 4 * half, z);
        }
      }
      tmem_wait_st();
    }
    tcgen05_fence_before();
    __syncthreads();
    // ---- O = P V ----------------------------------------------------------------------------------------------------
    if (warp == 0) {
      mbar_wait(&s.bar_v, ph);
      if (lane == 0 && next < n_items) load_v(next, ph ^ 1u);   // the other buffer: its last reader was item - 1's P V
      __syncwarp();
      tcgen05_fence_after();
      const float* vt = s.v[ph];
      if (elect_one()) {
        for (int ks = 0; ks < nks; ++ks) {
          if (PAIR) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {                     // O_t = P V_t: valid in the rows of trajectory t
              const uint64_t dv = umma_desc_mn_sw128_f16(vt + (64 * t + ks * 16) * kHd);
              const uint32_t acc = tmem + kPairO + 32 * t, a_hi = tmem + 8 * ks, a_lo = tmem + kPairPlo16 + 8 * ks;
              mma_f16_ta(acc, a_lo, dv, idesc_o, ks ? 1u : 0u);
              mma_f16_ta(acc, a_hi, dv + 4, idesc_o, 1u);             // + 64 B: the lo halves of the dims
              mma_f16_ta(acc, a_hi, dv, idesc_o, 1u);
            }
          } else {
            const uint64_t dv = umma_desc_mn_sw128_f16(vt + ks * 16 * kHd);
            const uint32_t a_hi = tmem + 8 * ks, a_lo = tmem + kPlo16 + 8 * ks;
            mma_f16_ta(tmem + kColO, a_lo, dv, idesc_o, ks ? 1u : 0u);
            mma_f16_ta(tmem + kColO, a_hi, dv + 4, idesc_o, 1u);
            mma_f16_ta(tmem + kColO, a_hi, dv, idesc_o, 1u);
          }
        }
        tcgen05_commit(&s.bar_o);
      }
      __syncwarp();
    }
    mbar_wait(&s.bar_o, ph);
    tcgen05_fence_after();
    __syncwarp();
    // ---- epilogue: each thread 16 of the row's 32 output dims -----------------------------------------------------
    const uint32_t col_o = PAIR ? kPairO + 32 * th : kColO;
    auto row_sum = [&]() -> float {
      if (PAIR) {
        const float l0 = tmem_ld1(tmem_row + kPairX + 2);
        return l0 + tmem_ld1(tmem_row + kPairX + 3);
      }
      return psum[tid & 127] + psum[(tid & 127) + 128];
    };
    if (active && row_lo == 0) {
      uint32_t r[16];
      TmemRow<16>::ld(tmem_row + col_o + 16 * half, r);
      if (half == 0 && lane == 0) bulk_wait_read<0>();       // the previous item's store has read the staging chunk
      pair_sync(wq);
      tmem_wait_ld();
      const float inv = kPvUnscale / row_sum();
      float4 (*stg)[8] = s.stage[wq];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float4 o;
        o.x = __uint_as_float(r[4 * jj + 0]) * inv;
        o.y = __uint_as_float(r[4 * jj + 1]) * inv;
        o.z = __uint_as_float(r[4 * jj + 2]) * inv;
        o.w = __uint_as_float(r[4 * jj + 3]) * inv;
        stg[lrev][(4 * half + jj) ^ (lrev & 7)] = o;
      }
      fence_proxy_async();
      pair_sync(wq);
      if (half == 0 && lane == 0 && b < n_traj) {
        tma_store_3d(&tmOut, stg, h * kHd, qrev * 32, b);
        bulk_commit();
      }
    } else if (active) {
      uint32_t r[16];
      TmemRow<16>::ld(tmem_row + col_o + 16 * half, r);
      tmem_wait_ld();
      const int tok = qrev * 32 + lrev;
      const float rsum = row_sum();
      if (tok >= row_lo && tok < L && b < n_traj) {
        const float inv = kPvUnscale / rsum;
        float4* dst = reinterpret_cast<float4*>(out_rows + ((size_t)b * (L - row_lo) + (tok - row_lo)) * d + h * kHd + 16 * half);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float4 o;
          o.x = __uint_as_float(r[4 * jj + 0]) * inv;
          o.y = __uint_as_float(r[4 * jj + 1]) * inv;
          o.z = __uint_as_float(r[4 * jj + 2]) * inv;
          o.w = __uint_as_float(r[4 * jj + 3]) * inv;
          dst[jj] = o;
        }
      }
    }
    tcgen05_fence_before();      // the next item's MMAs overwrite S and O
    __syncthreads();
  }
  if (half == 0 && lane == 0) bulk_wait_read<0>();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
  }
}

// Returns QX_EUNSUPPORTED when the shape is outside the kernel's range (the caller then falls back to the mma.sync
// kernel of qx_tf_attn.cu).  qkv: [B*L][3d] packed rows, out: [B][L - row_lo][d] = the rows row_lo .. L-1 of every
// trajectory (row_lo = 0: all of them).
// qk16: 0 = fp32 q, k, v; 1 = q, k as fp16 (hi, lo) pairs; 2 = q, k, v as pairs (tf_attention_p16_kernel)
int tf_attention_umma(const float* qkv, float* out, int B, int L, int d, int nhead, int row_lo, cudaStream_t st, int qk16) {
  if (d != nhead * kHd || L > kMaxKeys || L < 1 || B < 1 || row_lo < 0 || row_lo >= L) return QX_EUNSUPPORTED;
  // QX_TF_ATTN_PAIR=0 keeps one trajectory per tile whatever L is (A/B runs and the equality test)
  const char* env_pair = getenv("QX_TF_ATTN_PAIR");
  const bool pair = L <= 64 && !(env_pair != nullptr && env_pair[0] == '0');
  CUtensorMap mq, mk, mv, mo;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)(3 * d), (cuuint64_t)L, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)(3 * d) * 4, (cuuint64_t)L * 3 * d * 4};
    const cuuint32_t box1[3] = {(cuuint32_t)kHd, (cuuint32_t)kRows, 1}, box2[3] = {(cuuint32_t)kHd, 64, 2};
    const cuuint32_t boxk[3] = {(cuuint32_t)kHd, (cuuint32_t)kMaxKeys, 1};       // key / value rows of a single-trajectory tile
    int rc = make_map_nd(&mq, qkv, 3, dims, strides, pair ? box2 : box1);
    if (rc) return rc;
    rc = make_map_nd(&mk, qkv, 3, dims, strides, pair ? box2 : boxk);
    if (rc) return rc;
    // V is an MN-major operand of the second product: 32-B swizzle atoms (see umma_desc_mn_sw128_32b)
    rc = make_map_nd(&mv, qkv, 3, dims, strides, pair ? box2 : boxk,
                     qk16 == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  }
  {
    // out holds rows row_lo .. L-1 of every trajectory, compactly: [B][L - row_lo][d]
    const cuuint64_t dims[3] = {(cuuint64_t)d, (cuuint64_t)(L - row_lo), (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)d * 4, (cuuint64_t)(L - row_lo) * d * 4};
    const cuuint32_t box[3] = {(cuuint32_t)kHd, 32, 1};
    const int rc = make_map_nd(&mo, out, 3, dims, strides, box);
    if (rc) return rc;
  }
  static bool cfg[kMaxDevices] = {};
  if (!cfg[current_device_slot()]) {
    cudaFuncSetAttribute(tf_attention_umma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttnSmemT<false>));
    cudaFuncSetAttribute(tf_attention_umma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttnSmemT<true>));
    cudaFuncSetAttribute(tf_attention_umma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttnSmemT<false>));
    cudaFuncSetAttribute(tf_attention_umma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttnSmemT<true>));
    cudaFuncSetAttribute(tf_attention_p16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttnSmem16<false>));
    cudaFuncSetAttribute(tf_attention_p16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttnSmem16<true>));
    cfg[current_device_slot()] = true;
  }
  const int num_sms = device_sm_count();
  const int n_items = (pair ? (B + 1) / 2 : B) * nhead;
  // QX_TF_ATTN_DBG (timing experiments only, results are then wrong on purpose): bit 1 no P V MMAs, bit 2 no softmax,
  // bit 3 no operand splits, bit 4 no epilogue
  const char* env_dbg = getenv("QX_TF_ATTN_DBG");
  const int dbg = env_dbg ? atoi(env_dbg) : 0;
  const int grid = n_items < 2 * num_sms ? n_items : 2 * num_sms;
  if (qk16 == 2 && pair) tf_attention_p16_kernel<true><<<grid, 256, sizeof(AttnSmem16<true>), st>>>(mq, mk, mv, mo, out, n_items, B, nhead, L, d, row_lo);
  else if (qk16 == 2) tf_attention_p16_kernel<false><<<grid, 256, sizeof(AttnSmem16<false>), st>>>(mq, mk, mv, mo, out, n_items, B, nhead, L, d, row_lo);
  else if (pair && qk16) tf_attention_umma_kernel<true, true><<<grid, 256, sizeof(AttnSmemT<true>), st>>>(mq, mk, mv, mo, out, n_items, B, nhead, L, d, row_lo, dbg);
  else if (pair) tf_attention_umma_kernel<true, false><<<grid, 256, sizeof(AttnSmemT<true>), st>>>(mq, mk, mv, mo, out, n_items, B, nhead, L, d, row_lo, dbg);
  else if (qk16) tf_attention_umma_kernel<false, true><<<grid, 256, sizeof(AttnSmemT<false>), st>>>(mq, mk, mv, mo, out, n_items, B, nhead, L, d, row_lo, dbg);
  else tf_attention_umma_kernel<false, false><<<grid, 256, sizeof(AttnSmemT<false>), st>>>(mq, mk, mv, mo, out, n_items, B, nhead, L, d, row_lo, dbg);
  QX_CHECK_LAUNCH("tf_attention_umma_kernel");
#ifdef QX_ATTN_TRACE
  {
    static const char* names[13] = {"S issue", "wait V", "V split", "wait S", "softmax", "barrier", "next Q,K split", "wait O", "epilogue", "barrier",
                                    "loop top", "wait next Q,K", "O issue"};
    long long h[2][16];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_attn_trace, sizeof(h));
    const int per_cta = (n_items + grid - 1) / grid;
    for (int i = 0; i < 13; ++i)
      fprintf(stderr, "attn trace %-20s warp 0: %8.1f  warp 7: %8.1f cycles / item\n", names[i], (double)h[0][i] / per_cta, (double)h[1][i] / per_cta);
  }
#endif
  return QX_OK;
}

}  // namespace qx
