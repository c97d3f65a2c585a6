// Quattro transformer — tensor-core GEMM path (tcgen05 + TMEM + TMA), sm_100a.
//
// The four dense contractions of every encoder layer (transformer_model.py:112-113 of the
// reference -> nn.TransformerEncoderLayer: in_proj, out_proj, linear1, linear2) are
// C[M][N] = act(A[M][K] W[N][K]^T + bias) with M = B*L token rows.  Parity with the reference's
// fp32 CPU inference needs ~1e-6 relative accuracy, which a single TF32 pass (1e-3) does not give,
// so each product is evaluated as a 3xTF32 split:  A = A_hi + A_lo, W = W_hi + W_lo with *_hi
// holding the top 19 bits (what kind::tf32 reads) and *_lo the exact remainder;
// A W^T ~= A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T, accumulated in fp32 in TMEM.
//
// Persistent kernel, one CTA per SM looping over 128 x 128 output tiles (n fastest, so the CTAs
// that share an A tile run together and it is read from HBM once):
//   warp 0    TMA producer: per 32-wide K block, three 2-D tiled loads (A raw, W_hi, W_lo; 128-B
//             swizzle) into a 3-stage shared-memory ring, completion on "full" mbarriers
//   warps 2-5 splitter: turn the raw A tile into A_hi (masked in place) and A_lo, then
//             fence.proxy.async + arrive on "ready"
//   warp 1    allocates 256 TMEM columns (two accumulators); one elected lane issues 12 tcgen05.mma
//             (kind::tf32, M=128, N=128, K=8) per K block, tcgen05.commit's the stage back to the
//             producer and, after the last K block, the accumulator to the epilogue
//   warps 6-9 epilogue: tcgen05.ld (32x32b.x32) TMEM -> registers, + bias, optional ReLU, 16-B
//             stores; runs concurrently with the next tile's MMAs (double-buffered TMEM)
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "qx_tc05.cuh"
#include "qx_tf_common.cuh"

namespace qx {

constexpr int kBM = 128, kBN = 128, kBK = 32, kGemmStages = 3;
constexpr int kTileBytes = kBM * kBK * 4;          // 16 KB, one 128-row x 128-B swizzled tile

struct GemmStage {
  float a_hi[kBM * kBK];
  float a_lo[kBM * kBK];
  float w_hi[kBN * kBK];
  float w_lo[kBN * kBK];
};
struct GemmSmem {
  GemmStage st[kGemmStages];
  float4 epi[4][32][8];          // per epilogue warp: 32 rows x 32 columns, 16-B chunks XOR-swizzled by row
  uint64_t full[kGemmStages], ready[kGemmStages], empty[kGemmStages], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};
constexpr size_t kGemmSmemBytes = sizeof(GemmSmem);

// D = F32, A = B = TF32, both K-major, M = 128, N = 128
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

template <bool RELU>
__global__ void __launch_bounds__(320, 1)
tf_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWhi,
                  const __grid_constant__ CUtensorMap tmWlo, const float* __restrict__ bias, float* __restrict__ C,
                  int M, int N, int K) {
  // No integer round trip on the pointer: the compiler must KNOW these are shared-memory accesses (LDS / STS).  As
  // generic LD.E they are not ordered against the mbarrier arrive that hands a buffer back to the TMA producer.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  GemmSmem& s = *reinterpret_cast<GemmSmem*>(smem_raw);
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u)) __trap();   // the swizzled tiles need 1 KB alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = K / kBK;
  const int n_tiles_n = N / kBN;
  const int n_tiles = n_tiles_n * ceil_div(M, kBM);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kGemmStages; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.ready[i], 128);
      mbar_init(&s.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.acc_full[i], 1);
      mbar_init(&s.acc_empty[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) {   // TMEM: two 128-column fp32 accumulators (double buffered across tiles)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "n"(2 * kBN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s.tmem_base;

  if (warp == 0) {
    // ---- TMA producer ------------------------------------------------------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int n0 = (tile % n_tiles_n) * kBN, m0 = (tile / n_tiles_n) * kBM;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int st = it % kGemmStages;
          mbar_wait(&s.empty[st], ((it / kGemmStages) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&s.full[st], 3 * kTileBytes);
          tma_load_2d(&tmA, &s.full[st], s.st[st].a_hi, kb * kBK, m0);
          tma_load_2d(&tmWhi, &s.full[st], s.st[st].w_hi, kb * kBK, n0);
          tma_load_2d(&tmWlo, &s.full[st], s.st[st].w_lo, kb * kBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----------------------------------------------------------------------------------
    if (lane == 0) {
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
        const uint32_t buf = tl & 1u;
        mbar_wait(&s.acc_empty[buf], ((tl >> 1) & 1u) ^ 1u);     // epilogue drained this accumulator
        tcgen05_fence_after();
        const uint32_t acc = tmem + buf * kBN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int st = it % kGemmStages;
          mbar_wait(&s.ready[st], (it / kGemmStages) & 1u);
          tcgen05_fence_after();
          const uint64_t da_hi = umma_desc_sw128(s.st[st].a_hi), da_lo = umma_desc_sw128(s.st[st].a_lo);
          const uint64_t dw_hi = umma_desc_sw128(s.st[st].w_hi), dw_lo = umma_desc_sw128(s.st[st].w_lo);
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);       // 32 bytes per K=8 slice
            mma_tf32(acc, da_lo + adv, dw_hi + adv, kIdescTf32, (kb | k) ? 1u : 0u);
            mma_tf32(acc, da_hi + adv, dw_lo + adv, kIdescTf32, 1u);
            mma_tf32(acc, da_hi + adv, dw_hi + adv, kIdescTf32, 1u);
          }
          tcgen05_commit(&s.empty[st]);          // frees the stage when these MMAs have read it
        }
        tcgen05_commit(&s.acc_full[buf]);
      }
    }
  } else if (warp < 6) {
    // ---- splitter: raw A -> (A_hi, A_lo) -------------------------------------------------------------
    const int tid = threadIdx.x - 64;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < KB; ++kb, ++it) {
        const int st = it % kGemmStages;
        mbar_wait(&s.full[st], (it / kGemmStages) & 1u);
        float4* hi4 = reinterpret_cast<float4*>(s.st[st].a_hi);
        float4* lo4 = reinterpret_cast<float4*>(s.st[st].a_lo);
#pragma unroll
        for (int i = 0; i < kBM * kBK / 4 / 128; ++i) {
          const float4 v = hi4[tid + i * 128];
          float4 h, l;
          tf32_split(v.x, h.x, l.x);
          tf32_split(v.y, h.y, l.y);
          tf32_split(v.z, h.z, l.z);
          tf32_split(v.w, h.w, l.w);
          hi4[tid + i * 128] = h;
          lo4[tid + i * 128] = l;
        }
        fence_proxy_async();                     // generic-proxy writes -> visible to the tensor core
        mbar_arrive(&s.ready[st]);
      }
    }
  } else {
    // ---- epilogue: TMEM -> registers -> global, overlapping the next tile's MMAs -----------------------
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      const int n0 = (tile % n_tiles_n) * kBN, m0 = (tile / n_tiles_n) * kBM;
      const uint32_t buf = tl & 1u;
      mbar_wait(&s.acc_full[buf], (tl >> 1) & 1u);
      __syncwarp();        // reconverge after the per-lane spin: tcgen05.ld / wait::ld are .sync.aligned
      tcgen05_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < kBN; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem + buf * kBN + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // bias / activation in registers, then transpose through shared memory so that every global
        // store instruction writes four full 128-B row segments
        const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c0);
        float4 (*stg)[8] = s.epi[q];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = b4[j];
          float4 o;
          o.x = __uint_as_float(r[4 * j + 0]) + bb.x;
          o.y = __uint_as_float(r[4 * j + 1]) + bb.y;
          o.z = __uint_as_float(r[4 * j + 2]) + bb.z;
          o.w = __uint_as_float(r[4 * j + 3]) + bb.w;
          if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          stg[lane][j ^ (lane & 7)] = o;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rr = 4 * k + (lane >> 3), c4 = lane & 7;
          const int grow = m0 + q * 32 + rr;
          if (grow < M)
            *reinterpret_cast<float4*>(C + (size_t)grow * N + n0 + c0 + 4 * c4) = stg[rr][c4 ^ (rr & 7)];
        }
        __syncwarp();
      }
      tcgen05_fence_before();
      mbar_arrive(&s.acc_empty[buf]);            // 128 arrivals release the accumulator
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * kBN) : "memory");
  }
}


// ---- main kernel: A operand in TENSOR MEMORY, weights parked in (or ringed through) shared memory ---------------
// What bounded the streaming kernel above (46 % tensor-pipe activity whatever its ring depth; tools/gemm_probe.py):
//  * an M128 N128 K8 kind::tf32 MMA with both operands in shared memory reads 4 KB of A and 4 KB of B in its 64
//    cycles - 128 B/clk, ALL the shared-memory bandwidth of the SM - while the splitter (16 KB read + 32 KB written
//    per K block) and the epilogue transpose compete for the same banks;
//  * one diverged lane issuing the MMAs spent ~10 instructions (R2UR round trips) per UTCHMMA, more than the 64
//    cycles an MMA lasts;
//  * 64 KB of st.global per tile from the epilogue warps stalled them for longer than the tile's main loop;
//  * with K = 128 the W tiles (128 KB per output tile) were two thirds of the L2 -> SM traffic.
// Here:
//  * the splitter (thread = row) reads its raw A row from the TMA landing block and writes A_hi / A_lo straight
//    into TMEM (tcgen05.st, lane = row, column = k); the MMAs take A from there (tcgen05.mma [d], [a], b-desc), so
//    shared memory only carries the landing of raw A, one read of it, and the B operand;
//  * the whole MMA warp walks the issue loop (warp-uniform operands -> uniform registers, UTCHMMAs 1-4 instructions
//    apart), one elected lane issues;
//  * the epilogue stages 32 x 32 chunks in the 128-B-swizzle box layout and lets TMA store them
//    (cp.async.bulk.tensor ... global.shared::cta), double buffered; rows beyond M are clipped by the tensor map;
//  * the W_hi / W_lo tiles live in a 4-block ring of 32 KB.  K = 128 is exactly 4 blocks: a CTA owns ONE n tile for
//    its whole life (CTAs b, b+1, .. of a group share the m tile and run in step, so A comes from HBM once), loads
//    its weights once and never frees them.  K > 128 (linear2) cycles the ring, a block being freed by the
//    tcgen05.commit of the MMAs that read it.
// (An L2 prefetch of the next A tile - cp.async.bulk.prefetch.tensor - was measured and removed: it made the
// linear2 / out-proj shapes 10-17 % SLOWER.)
// TMEM: columns 0-255 two accumulators (the epilogue of tile i overlaps the MMAs of tile i+1), 256-511 four A stages
// of (32 hi + 32 lo) columns.  Results are bit-identical to the streaming kernel (same MMAs, same order).
constexpr int kTaW = 4, kTaRaw = 4, kTaA = 4;

// accumulator -> product: the fp16 weights carry a factor 2^8 (exact to undo)
template <bool F16>
__device__ __forceinline__ float unscale(uint32_t acc) {
  return F16 ? __fmul_rn(__uint_as_float(acc), 1.f / 256.f) : __uint_as_float(acc);
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
constexpr int ta_threads(int sets) { return 64 + 128 * sets + 128; }

struct TaSmem {
  float w_hi[kTaW][kBN * kBK];
  float w_lo[kTaW][kBN * kBK];
  float raw[kTaRaw][kBM * kBK];
  float4 epi[4][2][32][8];       // per epilogue warp: two 32 x 32 fp32 chunks in the TMA 128-B swizzle layout
  uint64_t w_full[kTaW], w_empty[kTaW], full[kTaRaw], raw_empty[kTaRaw], a_ready[kTaA], a_empty[kTaA], acc_full[2],
      acc_empty[2], res_full[4][2];
  uint32_t tmem_base;
};
constexpr size_t kTaSmemBytes = sizeof(TaSmem);
static_assert(kTaSmemBytes <= 227 * 1024, "TMEM-A GEMM exceeds the shared memory of an SM");

// kSets sets of four splitter warps take the K blocks in turn.  A set must meet EVERY phase of the barriers it waits
// on (a parity wait cannot tell phase p from phase p + 2), hence ring depths that are multiples of the set count.
// STREAM_W: K > 128, the weight ring cycles (K % 128 == 0 keeps block kb in slot kb % 4).
// dbg (timing experiments only, QX_TF_GEMM_DBG): bit 0 no stores, bit 2 no epilogue work.
// LN (N == 128 only, the CTA owns whole rows): the epilogue computes C = LayerNorm(res + (A W^T + bias)) * gamma + beta
// instead of a plain store - the post-norm residual blocks of nn.TransformerEncoderLayer (norm1 after out_proj, norm2
// after linear2; transformer_model.py:112-113) without the 2 x 512 B per row round trip of a separate add + LN kernel.
// tmR maps the residual stream [M][128] with the (32 x 32) box of the output map; it may be the output itself.
struct LnArgs {
  const float* gamma;
  const float* beta;
};

// F16: the operands are split into fp16 halves instead of TF32 ones.  x = hi + lo with hi = rn16(x), lo = rn16(x - hi)
// carries the same 22 significant bits (the weights are pre-scaled by 2^8, undone exactly in the epilogue, so that their
// lo halves stay out of the fp16 subnormals; activations need |x| < 65 504), the three products hi hi + hi lo + lo hi
// are exact in the fp32 accumulator like the TF32 ones - but kind::f16 covers K = 16 per MMA where kind::tf32 covers 8:
// half the tensor-pipe time, half the weight bytes in shared memory, half the tensor-memory columns for A.  A weight
// block (128 rows x 128 B) then spans 64 k: two A blocks.
constexpr uint32_t kIdescF16 = (1u << 4) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);   // D = F32, A = B = F16
constexpr float kW16Scale = 256.f;
constexpr int kGemmPackQk = 16;      // flag bit of the kernel's `dbg` argument: n tiles 0 and 1 (Q, K) leave as fp16 (hi, lo) pairs
constexpr int kGemmPackV = 32;       // ... and so does n tile 2 (V, times 2^4)

template <bool RELU, bool STREAM_W, int kSets, bool LN = false, bool F16 = false>
__global__ void __launch_bounds__(ta_threads(kSets), 1)
tf_gemm_ta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWhi,
                  const __grid_constant__ CUtensorMap tmWlo, const __grid_constant__ CUtensorMap tmC,
                  const float* __restrict__ bias, int M, int N, int K, int dbg, const __grid_constant__ CUtensorMap tmR,
                  const LnArgs ln) {
  static_assert(kTaRaw % kSets == 0 && kTaA % kSets == 0, "ring depths must be multiples of the splitter set count");
  // No integer round trip on the pointer: the compiler must KNOW these are shared-memory accesses.  With generic
  // LD.E.128 for the splitter's raw-row reads, the mbarrier arrive that hands the raw block back to the producer was
  // issued while the loads were still in flight, and every ~20th launch a few rows picked up the NEXT block's data.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TaSmem& s = *reinterpret_cast<TaSmem*>(smem_raw);
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u)) __trap();   // the swizzled tiles need 1 KB alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = STREAM_W ? K / kBK : kTaW;
  const int n_tiles_n = N / kBN;
  const int groups = gridDim.x / n_tiles_n;           // CTAs of a group share the m tile
  const int n0 = (blockIdx.x % n_tiles_n) * kBN;
  const int g0 = blockIdx.x / n_tiles_n;
  const int m_tiles = ceil_div(M, kBM);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTaW; ++i) {
      mbar_init(&s.w_full[i], 1);
      mbar_init(&s.w_empty[i], 1);
    }
    for (int i = 0; i < kTaRaw; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.raw_empty[i], 128);
    }
    for (int i = 0; i < kTaA; ++i) {
      mbar_init(&s.a_ready[i], 128);
      mbar_init(&s.a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.acc_full[i], 1);
      mbar_init(&s.acc_empty[i], 128);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&s.res_full[i >> 1][i & 1], 1);
    mbar_fence_init();
  }
  if (warp == 1) {   // all 512 columns: 2 accumulators + 4 A stages
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s.tmem_base;
  const uint32_t tmem_a0 = tmem + 2 * kBN;

  if (warp == 0) {
    // ---- TMA producer: weight blocks (once, or ringed) and the raw A blocks of this CTA's m tiles ---------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int mt = g0; mt < m_tiles; mt += groups) {
        for (int kb = 0; kb < KB; ++kb, ++it) {
          if (F16) {
            // weight block wit = it / 2 (the maps see the fp16 rows as K / 2 floats); resident: the first tile's only
            if ((kb & 1) == 0 && (STREAM_W || it < (uint32_t)KB)) {
              const uint32_t wit = it >> 1;
              const int sw = wit % kTaW;
              mbar_wait(&s.w_empty[sw], ((wit / kTaW) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&s.w_full[sw], 2 * kTileBytes);
              tma_load_2d(&tmWhi, &s.w_full[sw], s.w_hi[sw], (kb >> 1) * kBK, n0);
              tma_load_2d(&tmWlo, &s.w_full[sw], s.w_lo[sw], (kb >> 1) * kBK, n0);
            }
          } else if (STREAM_W || it < (uint32_t)kTaW) {
            const int sw = it % kTaW;
            mbar_wait(&s.w_empty[sw], ((it / kTaW) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&s.w_full[sw], 2 * kTileBytes);
            tma_load_2d(&tmWhi, &s.w_full[sw], s.w_hi[sw], kb * kBK, n0);
            tma_load_2d(&tmWlo, &s.w_full[sw], s.w_lo[sw], kb * kBK, n0);
          }
          const int st = it % kTaRaw;
          mbar_wait(&s.raw_empty[st], ((it / kTaRaw) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&s.full[st], kTileBytes);
          tma_load_2d(&tmA, &s.full[st], s.raw[st], kb * kBK, mt * kBM);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: A from TMEM, B from the weight ring -----------------------------------------------------------
    uint32_t it = 0, tl = 0;
    for (int mt = g0; mt < m_tiles; mt += groups, ++tl) {
      const uint32_t buf = tl & 1u;
      mbar_wait(&s.acc_empty[buf], ((tl >> 1) & 1u) ^ 1u);
      tcgen05_fence_after();
      const uint32_t acc = tmem + buf * kBN;
      for (int kb = 0; kb < KB; ++kb, ++it) {
        const uint32_t wit = F16 ? it >> 1 : it;                    // weight block of this A block
        const uint32_t sa = it % kTaA;
        const uint32_t sw = STREAM_W ? wit % kTaW : (uint32_t)(F16 ? kb >> 1 : kb);    // resident: block kb's own slot
        // resident weights: phase 0 of w_full[sw] completes once and the wait stays satisfied
        mbar_wait(&s.w_full[sw], STREAM_W ? ((wit / kTaW) & 1u) : 0u);
        mbar_wait(&s.a_ready[sa], (it / kTaA) & 1u);
        tcgen05_fence_after();
        const uint32_t a_hi = tmem_a0 + sa * (2 * kBK), a_lo = a_hi + (F16 ? kBK / 2 : kBK);
        const uint64_t dw_hi = umma_desc_sw128(s.w_hi[sw]), dw_lo = umma_desc_sw128(s.w_lo[sw]);
        if (elect_one()) {
          if (F16) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // 32 bytes per K = 16 slice, the odd A block in the upper half of the 128-B weight rows; 8 columns of A
              const uint64_t adv = (uint64_t)((((kb & 1) * kBK + k * 16) * 2) >> 4);
              mma_f16_ta(acc, a_lo + 8 * k, dw_hi + adv, kIdescF16, (kb | k) ? 1u : 0u);
              mma_f16_ta(acc, a_hi + 8 * k, dw_lo + adv, kIdescF16, 1u);
              mma_f16_ta(acc, a_hi + 8 * k, dw_hi + adv, kIdescF16, 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {
              const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);       // 32 bytes per K = 8 slice
              mma_tf32_ta(acc, a_lo + 8 * k, dw_hi + adv, kIdescTf32, (kb | k) ? 1u : 0u);
              mma_tf32_ta(acc, a_hi + 8 * k, dw_lo + adv, kIdescTf32, 1u);
              mma_tf32_ta(acc, a_hi + 8 * k, dw_hi + adv, kIdescTf32, 1u);
            }
          }
          tcgen05_commit(&s.a_empty[sa]);
          if (STREAM_W && (!F16 || (kb & 1))) tcgen05_commit(&s.w_empty[sw]);
          if (kb == KB - 1) tcgen05_commit(&s.acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 2 + 4 * kSets) {
    // ---- splitter: thread = row; raw A row (128 B, swizzled chunks) -> registers -> A_hi, A_lo in TMEM -----------
    const int q = warp & 3;                      // TMEM lane quarter of this warp
    const int row = q * 32 + lane;
    const int my_tiles = g0 < m_tiles ? (m_tiles - g0 + groups - 1) / groups : 0;
    for (uint32_t it = (warp - 2) >> 2; it < (uint32_t)(my_tiles * KB); it += kSets) {
      const int st = it % kTaRaw;
      const uint32_t sa = it % kTaA;
      mbar_wait(&s.full[st], (it / kTaRaw) & 1u);
      const float4* src = reinterpret_cast<const float4*>(s.raw[st]) + row * 8;
      uint32_t hi[F16 ? 16 : 32], lo[F16 ? 16 : 32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = src[j ^ (row & 7)];           // 128-B swizzle: 16-B chunk j of row r sits at j ^ (r % 8)
        if constexpr (F16) {
          // a tensor-memory column holds two consecutive k (the lower one in the low half)
          f16_split2(v.x, v.y, hi[2 * j + 0], lo[2 * j + 0]);
          f16_split2(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
        } else {
          float h, l;
          tf32_split(v.x, h, l); hi[4 * j + 0] = __float_as_uint(h); lo[4 * j + 0] = __float_as_uint(l);
          tf32_split(v.y, h, l); hi[4 * j + 1] = __float_as_uint(h); lo[4 * j + 1] = __float_as_uint(l);
          tf32_split(v.z, h, l); hi[4 * j + 2] = __float_as_uint(h); lo[4 * j + 2] = __float_as_uint(l);
          tf32_split(v.w, h, l); hi[4 * j + 3] = __float_as_uint(h); lo[4 * j + 3] = __float_as_uint(l);
        }
      }
      mbar_wait(&s.a_empty[sa], ((it / kTaA) & 1u) ^ 1u);
      __syncwarp();        // the lanes leave the spin loop one by one; tcgen05.st / wait::st are .sync.aligned
      tcgen05_fence_after();
      const uint32_t taddr = tmem_a0 + sa * (2 * kBK) + ((uint32_t)(q * 32) << 16);
      if constexpr (F16) {
        tmem_st_32x32b_x16(taddr, hi);
        tmem_st_32x32b_x16(taddr + kBK / 2, lo);
      } else {
        tmem_st_32x32b_x32(taddr, hi);
        tmem_st_32x32b_x32(taddr + kBK, lo);
      }
      // Hand the raw block back to the producer only now: the stores above consume every loaded value, so the LDS have
      // returned.  (An arrive placed right after the loads is issued while they are in flight - nothing orders a
      // pending shared-memory read against the TMA write the arrive unblocks; that cost a few rows every ~100th launch.)
      mbar_arrive(&s.raw_empty[st]);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tcgen05_fence_before();
      mbar_arrive(&s.a_ready[sa]);
    }
  } else {
    // ---- epilogue: TMEM -> registers (+ bias, ReLU) -> swizzled staging chunk -> TMA store ------------------------
    const int q = warp & 3;
    uint32_t tl = 0, chunk = 0;
    if (LN && lane == 0 && g0 < m_tiles) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        mbar_arrive_expect_tx(&s.res_full[q][c], 32 * 32 * 4);
        tma_load_2d(&tmR, &s.res_full[q][c], s.epi[q][c], 32 * c, g0 * kBM + q * 32);
      }
    }
    for (int mt = g0; mt < m_tiles; mt += groups, ++tl) {
      const int m0 = mt * kBM;
      const uint32_t buf = tl & 1u;
      mbar_wait(&s.acc_full[buf], (tl >> 1) & 1u);
      __syncwarp();        // reconverge after the per-lane spin: tcgen05.ld / wait::ld are .sync.aligned
      tcgen05_fence_after();
      if constexpr (LN) {
        // thread = row (TMEM lane): the whole 128-column row sits in registers after ONE pass over tensor memory (its
        // 64 B/clk read port is shared with the A operand of the running MMAs), so the accumulator goes back to the
        // MMA warp at once.  The residual rows arrive as four 32 x 32 TMA boxes in the staging chunks (two before the
        // accumulator is waited for, two as the first ones are consumed); a = res + (acc + bias), mean, centred second
        // moment, normalise, scale, and the same staging chunks carry the result to the TMA store.
        const uint32_t tacc = tmem + buf * kBN + ((uint32_t)(q * 32) << 16);
        uint32_t r[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tacc + 32 * c, r[c]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tcgen05_fence_before();
        mbar_arrive(&s.acc_empty[buf]);
        // Summation order = tf_add_ln_kernel's (qx_tf_simt.cu), so fused and separate launches agree bit for bit: four
        // columns left to right (its lane), pairwise over the 8 "lanes" of a chunk, pairwise over the 4 chunks.
        float pc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4 (*stg)[8] = s.epi[q][c & 1];
          mbar_wait(&s.res_full[q][c & 1], (uint32_t)(c >> 1));     // two uses of each chunk per tile: parity = c / 2
          const float4* b4 = reinterpret_cast<const float4*>(bias + 32 * c);
          float t[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = b4[j];
            const float4 hh = stg[lane][j ^ (lane & 7)];
            const float a0 = __fadd_rn(hh.x, __fadd_rn(unscale<F16>(r[c][4 * j + 0]), bb.x));
            const float a1 = __fadd_rn(hh.y, __fadd_rn(unscale<F16>(r[c][4 * j + 1]), bb.y));
            const float a2 = __fadd_rn(hh.z, __fadd_rn(unscale<F16>(r[c][4 * j + 2]), bb.z));
            const float a3 = __fadd_rn(hh.w, __fadd_rn(unscale<F16>(r[c][4 * j + 3]), bb.w));
            t[j] = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
            r[c][4 * j + 0] = __float_as_uint(a0); r[c][4 * j + 1] = __float_as_uint(a1);
            r[c][4 * j + 2] = __float_as_uint(a2); r[c][4 * j + 3] = __float_as_uint(a3);
          }
          pc[c] = __fadd_rn(__fadd_rn(__fadd_rn(t[0], t[1]), __fadd_rn(t[2], t[3])),
                            __fadd_rn(__fadd_rn(t[4], t[5]), __fadd_rn(t[6], t[7])));
          // every lane has CONSUMED its values before the chunk is handed to the next bulk copy: the partial sum below
          // depends on each of the shared-memory loads, and the empty asm pins its computation in front of the barrier
          // (an issued-but-pending LDS is not ordered against the async-proxy write the copy performs)
          asm volatile("" ::"f"(pc[c]) : "memory");
          __syncwarp();
          if (c < 2 && lane == 0) {
            mbar_arrive_expect_tx(&s.res_full[q][c & 1], 32 * 32 * 4);
            tma_load_2d(&tmR, &s.res_full[q][c & 1], stg, 32 * (c + 2), m0 + q * 32);
          }
        }
        const float mu = __fmul_rn(__fadd_rn(__fadd_rn(pc[0], pc[1]), __fadd_rn(pc[2], pc[3])), 1.f / 128.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d0 = __fsub_rn(__uint_as_float(r[c][4 * j + 0]), mu), d1 = __fsub_rn(__uint_as_float(r[c][4 * j + 1]), mu);
            const float d2 = __fsub_rn(__uint_as_float(r[c][4 * j + 2]), mu), d3 = __fsub_rn(__uint_as_float(r[c][4 * j + 3]), mu);
            t[j] = __fmaf_rn(d3, d3, __fmaf_rn(d2, d2, __fmaf_rn(d1, d1, __fmul_rn(d0, d0))));
            r[c][4 * j + 0] = __float_as_uint(d0); r[c][4 * j + 1] = __float_as_uint(d1);
            r[c][4 * j + 2] = __float_as_uint(d2); r[c][4 * j + 3] = __float_as_uint(d3);
          }
          pc[c] = __fadd_rn(__fadd_rn(__fadd_rn(t[0], t[1]), __fadd_rn(t[2], t[3])),
                            __fadd_rn(__fadd_rn(t[4], t[5]), __fadd_rn(t[6], t[7])));
        }
        const float rs = rsqrtf(__fmaf_rn(__fadd_rn(__fadd_rn(pc[0], pc[1]), __fadd_rn(pc[2], pc[3])), 1.f / 128.f, 1e-5f));
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (lane == 0) bulk_wait_read<1>();          // the store that last used this staging chunk has read it
          __syncwarp();
          const float4* g4 = reinterpret_cast<const float4*>(ln.gamma + 32 * c);
          const float4* e4 = reinterpret_cast<const float4*>(ln.beta + 32 * c);
          float4 (*stg)[8] = s.epi[q][c & 1];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 g = g4[j], be = e4[j];
            float4 o;
            o.x = __fmaf_rn(__fmul_rn(__uint_as_float(r[c][4 * j + 0]), rs), g.x, be.x);
            o.y = __fmaf_rn(__fmul_rn(__uint_as_float(r[c][4 * j + 1]), rs), g.y, be.y);
            o.z = __fmaf_rn(__fmul_rn(__uint_as_float(r[c][4 * j + 2]), rs), g.z, be.z);
            o.w = __fmaf_rn(__fmul_rn(__uint_as_float(r[c][4 * j + 3]), rs), g.w, be.w);
            stg[lane][j ^ (lane & 7)] = o;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, stg, 32 * c, m0 + q * 32);
            bulk_commit();
          }
        }
        if (lane == 0 && mt + groups < m_tiles) {      // next tile's first two residual chunks: in flight during its MMAs
          bulk_wait_read<0>();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            mbar_arrive_expect_tx(&s.res_full[q][c], 32 * 32 * 4);
            tma_load_2d(&tmR, &s.res_full[q][c], s.epi[q][c], 32 * c, m0 + groups * kBM + q * 32);
          }
        }
        __syncwarp();
        continue;                                       // the accumulator was released above
      } else if (!(dbg & 4)) {
#pragma unroll
        for (int c0 = 0; c0 < kBN; c0 += 32, ++chunk) {
          uint32_t r[32];
          const uint32_t taddr = tmem + buf * kBN + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(taddr));
          if (lane == 0) bulk_wait_read<1>();          // the store that last used this staging chunk has read it
          __syncwarp();
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c0);
          float4 (*stg)[8] = s.epi[q][chunk & 1u];
          if (F16 && !RELU && (dbg & kGemmPackQk) && (n0 < 2 * kBN || (dbg & kGemmPackV))) {
            // Q / K columns of the QKV projection for the tcgen05 attention (qx_tf_attn05.cu, QK16): the 32 floats of a
            // head slice leave as (hi, lo) fp16 halves - 16-B pieces 0-3 hold hi of dims 0-31, pieces 4-7 lo - the same
            // 128 B per row, which the attention's TMA box lands as a K-major fp16 tile of 64 k: no split over there.
            // Q carries the softmax scale log2(e) / sqrt(head dim) already (what split_q_row applied); with kGemmPackV the
            // V columns are packed as well, times 2^4 (their lo halves then stay out of the fp16 subnormals; undone
            // exactly by the attention epilogue).
            const float qs = n0 == 0 ? rsqrtf(32.f) * 1.4426950408889634f : (n0 == 2 * kBN ? 16.f : 1.f);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 b0 = b4[2 * u], b1 = b4[2 * u + 1];
              const float v0 = __fmul_rn(__fadd_rn(unscale<F16>(r[8 * u + 0]), b0.x), qs);
              const float v1 = __fmul_rn(__fadd_rn(unscale<F16>(r[8 * u + 1]), b0.y), qs);
              const float v2 = __fmul_rn(__fadd_rn(unscale<F16>(r[8 * u + 2]), b0.z), qs);
              const float v3 = __fmul_rn(__fadd_rn(unscale<F16>(r[8 * u + 3]), b0.w), qs);
              const float v4 = __fmul_rn(__fadd_rn(unscale<F16>(r[8 * u + 4]), b1.x), qs);
              const float v5 = __fmul_rn(__fadd_rn(unscale<F16>(r[8 * u + 5]), b1.y), qs);
              const float v6 = __fmul_rn(__fadd_rn(unscale<F16>(r[8 * u + 6]), b1.z), qs);
              const float v7 = __fmul_rn(__fadd_rn(unscale<F16>(r[8 * u + 7]), b1.w), qs);
              uint32_t h[4], l[4];
              f16_split2(v0, v1, h[0], l[0]);
              f16_split2(v2, v3, h[1], l[1]);
              f16_split2(v4, v5, h[2], l[2]);
              f16_split2(v6, v7, h[3], l[3]);
              stg[lane][u ^ (lane & 7)] = make_float4(__uint_as_float(h[0]), __uint_as_float(h[1]), __uint_as_float(h[2]), __uint_as_float(h[3]));
              stg[lane][(4 + u) ^ (lane & 7)] = make_float4(__uint_as_float(l[0]), __uint_as_float(l[1]), __uint_as_float(l[2]), __uint_as_float(l[3]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = b4[j];
              float4 o;
              o.x = __fadd_rn(unscale<F16>(r[4 * j + 0]), bb.x);
              o.y = __fadd_rn(unscale<F16>(r[4 * j + 1]), bb.y);
              o.z = __fadd_rn(unscale<F16>(r[4 * j + 2]), bb.z);
              o.w = __fadd_rn(unscale<F16>(r[4 * j + 3]), bb.w);
              if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              stg[lane][j ^ (lane & 7)] = o;
            }
          }
          fence_proxy_async();                          // generic-proxy writes -> visible to the TMA store
          __syncwarp();
          if (lane == 0 && !(dbg & 1)) {
            tma_store_2d(&tmC, stg, n0 + c0, m0 + q * 32);
            bulk_commit();
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&s.acc_empty[buf]);
    }
    if (lane == 0) bulk_wait_read<0>();              // staging must outlive the last stores' reads
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// W -> (W_hi, W_lo), elementwise, done once per weight upload
__global__ void tf_split_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, size_t count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float h, l;
  tf32_split(w[i], h, l);
  hi[i] = h;
  lo[i] = l;
}

// W -> fp16 (hi, lo) halves of 2^8 W (see tf_gemm_ta_kernel<F16>); hi / lo hold `count` halves each
__global__ void tf_split16_kernel(const float* __restrict__ w, __half* __restrict__ hi, __half* __restrict__ lo, size_t count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float v = w[i] * kW16Scale;
  const __half h = __float2half_rn(v);
  hi[i] = h;
  lo[i] = __float2half_rn(v - __half2float(h));
}

// row-major fp32 matrix [rows][cols] -> tiled map with a (32 x 128) box and 128-B swizzle
static int make_map(CUtensorMap* map, const float* base, int rows, int cols, int box_cols = kBK, int box_rows = kBM) {
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  return make_map_nd(map, base, 2, dims, strides, box);
}

// f16: Whi / Wlo hold the fp16 halves tf_split_weights(.., true) wrote (N x K halves each; the maps see K / 2 floats a row)
int tf_gemm_tc(const float* A, const float* Whi, const float* Wlo, const float* bias, float* C, int M, int N, int K,
               bool relu, cudaStream_t st, bool f16, int pack_qk) {
  if (pack_qk && !(f16 && !relu && N == 3 * kBN && K == kTaW * kBK)) {
    set_error("fp16-pair Q / K output needs the fp16-split QKV projection of a d = 128 model (got N=%d K=%d)", N, K);
    return QX_EINVAL;
  }
  if (N % kBN != 0 || K % kBK != 0) {
    set_error("tensor-core GEMM needs N %% 128 == 0 and K %% 32 == 0 (got N=%d K=%d)", N, K);
    return QX_EINVAL;
  }
  CUtensorMap ma, mh, ml;
  int rc;
  const int Kw = f16 ? K / 2 : K;
  if ((rc = make_map(&ma, A, M, K)) || (rc = make_map(&mh, Whi, N, Kw)) || (rc = make_map(&ml, Wlo, N, Kw))) return rc;
  const int dslot = current_device_slot();
  static bool configured[kMaxDevices] = {};
  if (!configured[dslot]) {
    cudaFuncSetAttribute(tf_gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes);
    cudaFuncSetAttribute(tf_gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes);
    configured[dslot] = true;
  }
  const int num_sms = device_sm_count();
  // QX_TF_GEMM_STREAM=1: the first-generation streaming kernel (A/B measurements and the bitwise-equality test; read
  // per call).  QX_TF_GEMM_VARIANT=1: two splitter warp sets.  QX_TF_GEMM_DBG: timing experiments, see the kernel.
  const char* env_stream = getenv("QX_TF_GEMM_STREAM");
  const bool stream_only = env_stream != nullptr && env_stream[0] == '1';
  if (f16 && (K % (kTaW * kBK) != 0 || N / kBN > num_sms || stream_only)) {
    set_error("fp16-split GEMM needs K %% 128 == 0 and the tensor-memory kernel (got N=%d K=%d)", N, K);
    return QX_EINVAL;
  }
  if (K % (kTaW * kBK) == 0 && N / kBN <= num_sms && !stream_only) {
    // groups of N/128 CTAs (one per n tile) walk the m tiles together
    const int n_tiles_n = N / kBN, m_tiles = ceil_div(M, kBM);
    int groups = num_sms / n_tiles_n;
    if (groups > m_tiles) groups = m_tiles;
    const int grid = groups * n_tiles_n;
    const char* env_dbg = getenv("QX_TF_GEMM_DBG");
    const int dbg = ((env_dbg ? atoi(env_dbg) : 0) & ~(kGemmPackQk | kGemmPackV)) | (pack_qk ? kGemmPackQk : 0) |
                    (pack_qk == 2 ? kGemmPackV : 0);
    const char* env_var = getenv("QX_TF_GEMM_VARIANT");
    const int variant = env_var ? atoi(env_var) : 0;
    CUtensorMap mc;
    if ((rc = make_map(&mc, C, M, N, 32, 32))) return rc;
#define QX_TA_LAUNCH(RELU_, STREAM_, SETS, ...)                                                                         \
  do {                                                                                                                  \
    static bool cfgd[kMaxDevices] = {};                                                                                 \
    if (!cfgd[dslot]) {                                                                                                 \
      cudaFuncSetAttribute(tf_gemm_ta_kernel<RELU_, STREAM_, SETS __VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                           (int)kTaSmemBytes);                                                                          \
      cfgd[dslot] = true;                                                                                               \
    }                                                                                                                   \
    tf_gemm_ta_kernel<RELU_, STREAM_, SETS __VA_ARGS__><<<grid, ta_threads(SETS), kTaSmemBytes, st>>>(ma, mh, ml, mc, bias, M, N, K, dbg, mc, LnArgs{}); \
  } while (0)
#define QX_COMMA ,
    const bool stream_w = K != kTaW * kBK;
    if (f16) {
      if (relu && stream_w) QX_TA_LAUNCH(true, true, 1, QX_COMMA false QX_COMMA true);
      else if (relu) QX_TA_LAUNCH(true, false, 1, QX_COMMA false QX_COMMA true);
      else if (stream_w) QX_TA_LAUNCH(false, true, 1, QX_COMMA false QX_COMMA true);
      else QX_TA_LAUNCH(false, false, 1, QX_COMMA false QX_COMMA true);
    } else if (variant == 1) {
      if (relu && stream_w) QX_TA_LAUNCH(true, true, 2);
      else if (relu) QX_TA_LAUNCH(true, false, 2);
      else if (stream_w) QX_TA_LAUNCH(false, true, 2);
      else QX_TA_LAUNCH(false, false, 2);
    } else {
      if (relu && stream_w) QX_TA_LAUNCH(true, true, 1);
      else if (relu) QX_TA_LAUNCH(true, false, 1);
      else if (stream_w) QX_TA_LAUNCH(false, true, 1);
      else QX_TA_LAUNCH(false, false, 1);
    }
#undef QX_TA_LAUNCH
#undef QX_COMMA
    QX_CHECK_LAUNCH("tf_gemm_ta_kernel");
    return QX_OK;
  }
  const int tiles = (N / kBN) * ceil_div(M, kBM);
  const int grid = tiles < num_sms ? tiles : num_sms;          // persistent: one CTA per SM
  if (relu) tf_gemm_tc_kernel<true><<<grid, 320, kGemmSmemBytes, st>>>(ma, mh, ml, bias, C, M, N, K);
  else tf_gemm_tc_kernel<false><<<grid, 320, kGemmSmemBytes, st>>>(ma, mh, ml, bias, C, M, N, K);
  QX_CHECK_LAUNCH("tf_gemm_tc_kernel");
  return QX_OK;
}

// C[M][128] = LayerNorm(res + (A[M][K] W[128][K]^T + bias)) * gamma + beta in one launch (the LN variant of the TMEM-A
// kernel).  C may be res itself (a tile's residual rows are consumed before its stores are issued, and no other CTA
// touches them) but must not be A.  QX_EUNSUPPORTED when the shape does not fit.
int tf_gemm_ln_tc(const float* A, const float* Whi, const float* Wlo, const float* bias, const float* res, const float* gamma,
                  const float* beta, float* C, int M, int K, cudaStream_t st, bool f16) {
  const int N = kBN;
  if (K % (kTaW * kBK) != 0) return QX_EUNSUPPORTED;
  CUtensorMap ma, mh, ml, mc, mr;
  int rc;
  const int Kw = f16 ? K / 2 : K;
  if ((rc = make_map(&ma, A, M, K)) || (rc = make_map(&mh, Whi, N, Kw)) || (rc = make_map(&ml, Wlo, N, Kw)) ||
      (rc = make_map(&mc, C, M, N, 32, 32)) || (rc = make_map(&mr, res, M, N, 32, 32)))
    return rc;
  const int num_sms = device_sm_count();
  const int m_tiles = ceil_div(M, kBM);
  const int grid = num_sms < m_tiles ? num_sms : m_tiles;
  const LnArgs ln{gamma, beta};
  static bool cfgd[kMaxDevices] = {};
  if (!cfgd[current_device_slot()]) {
    cudaFuncSetAttribute(tf_gemm_ta_kernel<false, false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTaSmemBytes);
    cudaFuncSetAttribute(tf_gemm_ta_kernel<false, true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTaSmemBytes);
    cudaFuncSetAttribute(tf_gemm_ta_kernel<false, false, 1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTaSmemBytes);
    cudaFuncSetAttribute(tf_gemm_ta_kernel<false, true, 1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTaSmemBytes);
    cfgd[current_device_slot()] = true;
  }
  if (f16 && K != kTaW * kBK)
    tf_gemm_ta_kernel<false, true, 1, true, true><<<grid, ta_threads(1), kTaSmemBytes, st>>>(ma, mh, ml, mc, bias, M, N, K, 0, mr, ln);
  else if (f16)
    tf_gemm_ta_kernel<false, false, 1, true, true><<<grid, ta_threads(1), kTaSmemBytes, st>>>(ma, mh, ml, mc, bias, M, N, K, 0, mr, ln);
  else if (K != kTaW * kBK)
    tf_gemm_ta_kernel<false, true, 1, true><<<grid, ta_threads(1), kTaSmemBytes, st>>>(ma, mh, ml, mc, bias, M, N, K, 0, mr, ln);
  else
    tf_gemm_ta_kernel<false, false, 1, true><<<grid, ta_threads(1), kTaSmemBytes, st>>>(ma, mh, ml, mc, bias, M, N, K, 0, mr, ln);
  QX_CHECK_LAUNCH("tf_gemm_ta_kernel<LN>");
  return QX_OK;
}


// ---- the feed-forward block of an encoder layer as ONE kernel ------------------------------------------------------
//   x <- LayerNorm(x + linear2(relu(linear1(x))))      (transformer_model.py:112-113: nn.TransformerEncoderLayer, post-norm)
// As two launches the ff-wide intermediate makes 3.4 of the 4.6 GB the block moves per layer (8 192 quadrotor
// trajectories) and both launches run at the HBM roof.  Here a CTA owns 128 token rows and walks the hidden units in
// chunks of 128: H_j = relu(X W1_j^T + b1_j) is accumulated in tensor memory, read once by the converter warps (bias, ReLU,
// fp16 split) and written back to tensor memory as the A operand of Y += H_j W2_j^T; HBM sees X once (plus the residual
// boxes of the LayerNorm epilogue, L2 hits) and the result once.  fp16 split only (tf_gemm_ta_kernel<F16>); every product,
// its order of accumulation and every rounding are those of the two-launch path, so the results are the same bits.
//   tensor memory   Y [0,128) | H_j [128,256) | X hi [256,320) lo [320,384) | H_j hi [384,448) lo [448,512)   (fp16: 2 k / column)
//   shared memory   4 weight slots (128 rows x 64 k, hi + lo: 32 KB) | X landing: four 32-float blocks | LN staging
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 split X / convert H_j (two threads per row), warps 10-13 LayerNorm epilogue
// The MMA warp runs one product ahead: G1(0) G1(1) G2(0) G1(2) G2(1) .. G2(C-1), so that the conversion of chunk j has the
// time of G1(j+1) to finish before the tensor pipe would wait for it; the weight blocks come in that order.
constexpr uint32_t kFfnY = 0, kFfnH = 128, kFfnX = 256, kFfnH16 = 384;
struct FfnSmem {
  float w_hi[kTaW][kBN * kBK];
  float w_lo[kTaW][kBN * kBK];
  float raw[4][kBM * kBK];
  float4 epi[4][2][32][8];
  uint64_t w_full[kTaW], w_empty[kTaW], raw_full[4], raw_empty[4], x_ready, x_free, hacc_full, hacc_empty, h16_ready, h16_free,
      y_full, y_empty, res_full[4][2];
  uint32_t tmem_base;
};
static_assert(sizeof(FfnSmem) <= 227 * 1024, "fused FFN exceeds the shared memory of an SM");

__global__ void __launch_bounds__(448, 1)
tf_ffn_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1hi,
              const __grid_constant__ CUtensorMap tmW1lo, const __grid_constant__ CUtensorMap tmW2hi,
              const __grid_constant__ CUtensorMap tmW2lo, const __grid_constant__ CUtensorMap tmC,
              const __grid_constant__ CUtensorMap tmR, const float* __restrict__ b1, const float* __restrict__ b2,
              const LnArgs ln, int M, int ff) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  FfnSmem& s = *reinterpret_cast<FfnSmem*>(smem_raw);
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u)) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = ff / 128;                                // hidden chunks
  const int m_tiles = ceil_div(M, kBM);
  const int stride = gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTaW; ++i) {
      mbar_init(&s.w_full[i], 1);
      mbar_init(&s.w_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s.raw_full[i], 1);
      mbar_init(&s.raw_empty[i], 128);
    }
    mbar_init(&s.x_ready, 256);
    mbar_init(&s.x_free, 1);
    mbar_init(&s.hacc_full, 1);
    mbar_init(&s.hacc_empty, 256);
    mbar_init(&s.h16_ready, 256);
    mbar_init(&s.h16_free, 1);
    mbar_init(&s.y_full, 1);
    mbar_init(&s.y_empty, 128);
    for (int i = 0; i < 8; ++i) mbar_init(&s.res_full[i >> 1][i & 1], 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s.tmem_base;

  if (warp == 0) {
    // ---- TMA producer -----------------------------------------------------------------------------------------------
    if (lane == 0) {
      uint32_t wit = 0, rit = 0;
      // the two 64-k weight blocks of product G1(j) (rows = hidden units of chunk j) or G2(j) (k = hidden units of chunk j)
      auto w_blocks = [&](bool second, int j) {
        for (int kb = 0; kb < 2; ++kb, ++wit) {
          const int sw = wit % kTaW;
          mbar_wait(&s.w_empty[sw], ((wit / kTaW) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&s.w_full[sw], 2 * kTileBytes);
          if (second) {
            tma_load_2d(&tmW2hi, &s.w_full[sw], s.w_hi[sw], (2 * j + kb) * kBK, 0);
            tma_load_2d(&tmW2lo, &s.w_full[sw], s.w_lo[sw], (2 * j + kb) * kBK, 0);
          } else {
            tma_load_2d(&tmW1hi, &s.w_full[sw], s.w_hi[sw], kb * kBK, j * 128);
            tma_load_2d(&tmW1lo, &s.w_full[sw], s.w_lo[sw], kb * kBK, j * 128);
          }
        }
      };
      for (int mt = blockIdx.x; mt < m_tiles; mt += stride) {
        for (int kb = 0; kb < 4; ++kb, ++rit) {
          const int st = rit & 3;
          mbar_wait(&s.raw_empty[st], ((rit >> 2) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&s.raw_full[st], kTileBytes);
          tma_load_2d(&tmX, &s.raw_full[st], s.raw[st], kb * kBK, mt * kBM);
        }
        w_blocks(false, 0);
        for (int j = 1; j < C; ++j) {
          w_blocks(false, j);
          w_blocks(true, j - 1);
        }
        w_blocks(true, C - 1);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ---------------------------------------------------------------------------------------------------
    uint32_t wit = 0, tl = 0, n1 = 0, n2 = 0;
    // one product of a chunk: acc (+)= A[tensor memory: 64 columns hi, 64 columns lo] W^T over k = 0..127
    auto product = [&](uint32_t acc, uint32_t a_base, bool fresh) {
      for (int kb = 0; kb < 2; ++kb, ++wit) {
        const uint32_t sw = wit % kTaW;
        mbar_wait(&s.w_full[sw], (wit / kTaW) & 1u);
        tcgen05_fence_after();
        const uint64_t dw_hi = umma_desc_sw128(s.w_hi[sw]), dw_lo = umma_desc_sw128(s.w_lo[sw]);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
            const uint32_t a_hi = a_base + kb * 32 + k * 8, a_lo = a_hi + 64;
            mma_f16_ta(acc, a_lo, dw_hi + adv, kIdescF16, (fresh && kb == 0 && k == 0) ? 0u : 1u);
            mma_f16_ta(acc, a_hi, dw_lo + adv, kIdescF16, 1u);
            mma_f16_ta(acc, a_hi, dw_hi + adv, kIdescF16, 1u);
          }
          tcgen05_commit(&s.w_empty[sw]);
        }
        __syncwarp();
      }
    };
    auto g1 = [&](int j) {
      mbar_wait(&s.hacc_empty, (n1 & 1u) ^ 1u);            // the converters have read the previous chunk's accumulator
      tcgen05_fence_after();
      product(tmem + kFfnH, tmem + kFfnX, true);
      if (elect_one()) {
        tcgen05_commit(&s.hacc_full);
        if (j == C - 1) tcgen05_commit(&s.x_free);         // every product that reads this tile's X halves is issued
      }
      __syncwarp();
      ++n1;
    };
    auto g2 = [&](int j) {
      mbar_wait(&s.h16_ready, n2 & 1u);
      if (j == 0) mbar_wait(&s.y_empty, (tl & 1u) ^ 1u);   // the epilogue has read the previous tile's result
      tcgen05_fence_after();
      product(tmem + kFfnY, tmem + kFfnH16, j == 0);
      if (elect_one()) {
        tcgen05_commit(&s.h16_free);
        if (j == C - 1) tcgen05_commit(&s.y_full);
      }
      __syncwarp();
      ++n2;
    };
    for (int mt = blockIdx.x; mt < m_tiles; mt += stride, ++tl) {
      mbar_wait(&s.x_ready, tl & 1u);
      tcgen05_fence_after();
      g1(0);
      for (int j = 1; j < C; ++j) {
        g1(j);
        g2(j - 1);
      }
      g2(C - 1);
    }
  } else if (warp < 10) {
    // ---- split X, then convert every hidden chunk: two threads per row (warps w and w + 4 share a lane quarter and take
    //      64 of the 128 k each) - the conversion of a chunk has to fit into the time of one product -----------------
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t tl = 0, ct = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += stride, ++tl) {
      mbar_wait(&s.x_free, (tl & 1u) ^ 1u);                // the previous tile's first products are done with its X
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int kb = 2 * part + kk;                      // landing block = ring slot: four blocks per tile, four slots
        mbar_wait(&s.raw_full[kb], tl & 1u);
        const float4* src = reinterpret_cast<const float4*>(s.raw[kb]) + row * 8;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = src[j ^ (row & 7)];
          f16_split2(v.x, v.y, hi[2 * j + 0], lo[2 * j + 0]);
          f16_split2(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
        }
        __syncwarp();
        tcgen05_fence_after();
        tmem_st_32x32b_x16(trow + kFfnX + 16 * kb, hi);
        tmem_st_32x32b_x16(trow + kFfnX + 64 + 16 * kb, lo);
        mbar_arrive(&s.raw_empty[kb]);                     // behind the stores, which consume every loaded value
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tcgen05_fence_before();
      mbar_arrive(&s.x_ready);
      for (int j = 0; j < C; ++j, ++ct) {
        mbar_wait(&s.hacc_full, ct & 1u);
        __syncwarp();
        tcgen05_fence_after();
        uint32_t r[2][32];
#pragma unroll
        for (int c = 0; c < 2; ++c) tmem_ld_32x32b_x32(trow + kFfnH + 32 * (2 * part + c), r[c]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tcgen05_fence_before();
        mbar_arrive(&s.hacc_empty);
        const float* bj = b1 + j * 128 + 64 * part;
        bool waited = false;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t hi[16], lo[16];
          const float4* b4 = reinterpret_cast<const float4*>(bj + 32 * c);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 bb = b4[e];
            const float v0 = fmaxf(__fadd_rn(unscale<true>(r[c][4 * e + 0]), bb.x), 0.f);
            const float v1 = fmaxf(__fadd_rn(unscale<true>(r[c][4 * e + 1]), bb.y), 0.f);
            const float v2 = fmaxf(__fadd_rn(unscale<true>(r[c][4 * e + 2]), bb.z), 0.f);
            const float v3 = fmaxf(__fadd_rn(unscale<true>(r[c][4 * e + 3]), bb.w), 0.f);
            f16_split2(v0, v1, hi[2 * e + 0], lo[2 * e + 0]);
            f16_split2(v2, v3, hi[2 * e + 1], lo[2 * e + 1]);
          }
          if (!waited) {
            mbar_wait(&s.h16_free, (ct & 1u) ^ 1u);        // the previous chunk's second product has read its halves
            __syncwarp();
            tcgen05_fence_after();
            waited = true;
          }
          tmem_st_32x32b_x16(trow + kFfnH16 + 16 * (2 * part + c), hi);
          tmem_st_32x32b_x16(trow + kFfnH16 + 64 + 16 * (2 * part + c), lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tcgen05_fence_before();
        mbar_arrive(&s.h16_ready);
      }
    }
  } else {
    // ---- epilogue: x <- LayerNorm(x + (Y / 2^8 + b2)), as in tf_gemm_ta_kernel<LN> ---------------------------------
    const int q = warp & 3;
    uint32_t tl = 0;
    if (lane == 0 && (int)blockIdx.x < m_tiles) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        mbar_arrive_expect_tx(&s.res_full[q][c], 32 * 32 * 4);
        tma_load_2d(&tmR, &s.res_full[q][c], s.epi[q][c], 32 * c, blockIdx.x * kBM + q * 32);
      }
    }
    for (int mt = blockIdx.x; mt < m_tiles; mt += stride, ++tl) {
      const int m0 = mt * kBM;
      mbar_wait(&s.y_full, tl & 1u);
      __syncwarp();
      tcgen05_fence_after();
      const uint32_t tacc = tmem + kFfnY + ((uint32_t)(q * 32) << 16);
      uint32_t r[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tacc + 32 * c, r[c]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tcgen05_fence_before();
      mbar_arrive(&s.y_empty);
      float pc[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 (*stg)[8] = s.epi[q][c & 1];
        mbar_wait(&s.res_full[q][c & 1], (uint32_t)(c >> 1));
        const float4* b4 = reinterpret_cast<const float4*>(b2 + 32 * c);
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = b4[j];
          const float4 hh = stg[lane][j ^ (lane & 7)];
          const float a0 = __fadd_rn(hh.x, __fadd_rn(unscale<true>(r[c][4 * j + 0]), bb.x));
          const float a1 = __fadd_rn(hh.y, __fadd_rn(unscale<true>(r[c][4 * j + 1]), bb.y));
          const float a2 = __fadd_rn(hh.z, __fadd_rn(unscale<true>(r[c][4 * j + 2]), bb.z));
          const float a3 = __fadd_rn(hh.w, __fadd_rn(unscale<true>(r[c][4 * j + 3]), bb.w));
          t[j] = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
          r[c][4 * j + 0] = __float_as_uint(a0); r[c][4 * j + 1] = __float_as_uint(a1);
          r[c][4 * j + 2] = __float_as_uint(a2); r[c][4 * j + 3] = __float_as_uint(a3);
        }
        pc[c] = __fadd_rn(__fadd_rn(__fadd_rn(t[0], t[1]), __fadd_rn(t[2], t[3])),
                          __fadd_rn(__fadd_rn(t[4], t[5]), __fadd_rn(t[6], t[7])));
        asm volatile("" ::"f"(pc[c]) : "memory");         // every lane has consumed the chunk before it is handed on
        __syncwarp();
        if (c < 2 && lane == 0) {
          mbar_arrive_expect_tx(&s.res_full[q][c & 1], 32 * 32 * 4);
          tma_load_2d(&tmR, &s.res_full[q][c & 1], stg, 32 * (c + 2), m0 + q * 32);
        }
      }
      const float mu = __fmul_rn(__fadd_rn(__fadd_rn(pc[0], pc[1]), __fadd_rn(pc[2], pc[3])), 1.f / 128.f);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d0 = __fsub_rn(__uint_as_float(r[c][4 * j + 0]), mu), d1 = __fsub_rn(__uint_as_float(r[c][4 * j + 1]), mu);
          const float d2 = __fsub_rn(__uint_as_float(r[c][4 * j + 2]), mu), d3 = __fsub_rn(__uint_as_float(r[c][4 * j + 3]), mu);
          t[j] = __fmaf_rn(d3, d3, __fmaf_rn(d2, d2, __fmaf_rn(d1, d1, __fmul_rn(d0, d0))));
          r[c][4 * j + 0] = __float_as_uint(d0); r[c][4 * j + 1] = __float_as_uint(d1);
          r[c][4 * j + 2] = __float_as_uint(d2); r[c][4 * j + 3] = __float_as_uint(d3);
        }
        pc[c] = __fadd_rn(__fadd_rn(__fadd_rn(t[0], t[1]), __fadd_rn(t[2], t[3])),
                          __fadd_rn(__fadd_rn(t[4], t[5]), __fadd_rn(t[6], t[7])));
      }
      const float rs = rsqrtf(__fmaf_rn(__fadd_rn(__fadd_rn(pc[0], pc[1]), __fadd_rn(pc[2], pc[3])), 1.f / 128.f, 1e-5f));
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        const float4* g4 = reinterpret_cast<const float4*>(ln.gamma + 32 * c);
        const float4* e4 = reinterpret_cast<const float4*>(ln.beta + 32 * c);
        float4 (*stg)[8] = s.epi[q][c & 1];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g = g4[j], be = e4[j];
          float4 o;
          o.x = __fmaf_rn(__fmul_rn(__uint_as_float(r[c][4 * j + 0]), rs), g.x, be.x);
          o.y = __fmaf_rn(__fmul_rn(__uint_as_float(r[c][4 * j + 1]), rs), g.y, be.y);
          o.z = __fmaf_rn(__fmul_rn(__uint_as_float(r[c][4 * j + 2]), rs), g.z, be.z);
          o.w = __fmaf_rn(__fmul_rn(__uint_as_float(r[c][4 * j + 3]), rs), g.w, be.w);
          stg[lane][j ^ (lane & 7)] = o;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmC, stg, 32 * c, m0 + q * 32);
          bulk_commit();
        }
      }
      if (lane == 0 && mt + stride < m_tiles) {
        bulk_wait_read<0>();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          mbar_arrive_expect_tx(&s.res_full[q][c], 32 * 32 * 4);
          tma_load_2d(&tmR, &s.res_full[q][c], s.epi[q][c], 32 * c, m0 + stride * kBM + q * 32);
        }
      }
      __syncwarp();
    }
    if (lane == 0) bulk_wait_read<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// ---- the whole tail of an encoder layer as ONE kernel ---------------------------------------------------------------
//   h1 = LayerNorm1(x + out_proj(att));   x <- LayerNorm2(h1 + linear2(relu(linear1(h1))))
// tf_ffn_kernel with one more product in front: G0 = att W_out^T goes through the SAME tensor-memory regions (att halves
// where X's will be, its accumulator where the hidden chunks' will be).  norm1 runs on the eight CONVERTER warps (two
// threads per row, partial sums exchanged through shared memory): they read G0's accumulator and the residual rows (a
// TMA load into the landing blocks the att tile has left), and write h1 twice - as fp16 halves straight into the X
// columns of tensor memory, and as fp32 over the residual in the landing blocks, from where one TMA store per half
// puts it into x for the residual of norm2.  The separate out-proj launch (1.27 GB at the HBM roof) is gone and h1 is
// never read from HBM by a product.  A tile's G0 is issued inside the previous tile's last chunk:
//   MMA warp     .. G1(C-1) G2(C-2) | G0' | G2(C-1) | G1'(0) ..         (' = next tile)
//   converters   .. read H(C-1), split att', convert H(C-1), norm1' ..
// (With norm1 on the four thread-per-row LayerNorm warps - ~8 000 cycles per normalisation under the 128-register cap -
// the kernel took 0.93 ms against 0.52 + 0.24 for the two launches: the converters waited for h1 45 % of the time.)
// Same products, accumulation order and roundings as the separate launches: same bits.
// The converter warps are the bottleneck of this kernel (83 % busy): with two hidden chunks a tile (cart-pole, ff = 256)
// it beats out-proj + feed-forward as two launches (step 1.39 -> 1.32 ms), with four (quadrotor, ff = 512: 0.85 ms
// against 0.52 + 0.24) it does not, so the forward pass takes it for ff <= 256 only (QX_TF_TAIL_FUSE=2: always, =0: never).
struct TailSmem {
  float w_hi[kTaW][kBN * kBK];
  float w_lo[kTaW][kBN * kBK];
  float raw[4][kBM * kBK];       // att tile as four 32-float blocks (TMA), then the residual x rows (TMA), then h1 (norm1)
  float4 epi[4][2][32][8];
  float xch[2][2][kBM];          // [sum | centred second moment][column half][row]: the two threads of a row meet here
  uint64_t w_full[kTaW], w_empty[kTaW], att_full[4], att_done, res_in[4], a_ready, o_full, h1_done, x_ready, x_free, hacc_full,
      hacc_empty, h16_ready, h16_free, y_full, y_empty, res_full[4][2];
  uint32_t tmem_base;
  uint32_t h1_cnt;               // h1 stores completed (two per tile): a COUNT, the norm2 warps may be two tiles behind
};
static_assert(sizeof(TailSmem) <= 227 * 1024, "fused layer tail exceeds the shared memory of an SM");

struct TailArgs {
  const float* bo;       // out_proj.bias
  const float* g1;       // norm1
  const float* be1;
  const float* b1;       // linear1.bias
  const float* b2;       // linear2.bias
  const float* g2;       // norm2
  const float* be2;
};

__global__ void __launch_bounds__(448, 1)
tf_tail_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmX,
               const __grid_constant__ CUtensorMap tmWohi, const __grid_constant__ CUtensorMap tmWolo,
               const __grid_constant__ CUtensorMap tmW1hi, const __grid_constant__ CUtensorMap tmW1lo,
               const __grid_constant__ CUtensorMap tmW2hi, const __grid_constant__ CUtensorMap tmW2lo,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const TailArgs a, int M, int ff) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TailSmem& s = *reinterpret_cast<TailSmem*>(smem_raw);
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u)) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = ff / 128;
  const int m_tiles = ceil_div(M, kBM);
  const int stride = gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTaW; ++i) {
      mbar_init(&s.w_full[i], 1);
      mbar_init(&s.w_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s.att_full[i], 1);
      mbar_init(&s.res_in[i], 1);
    }
    mbar_init(&s.att_done, 256);
    mbar_init(&s.a_ready, 256);
    mbar_init(&s.o_full, 1);
    mbar_init(&s.h1_done, 2);
    s.h1_cnt = 0u;
    mbar_init(&s.x_ready, 256);
    mbar_init(&s.x_free, 1);
    mbar_init(&s.hacc_full, 1);
    mbar_init(&s.hacc_empty, 256);
    mbar_init(&s.h16_ready, 256);
    mbar_init(&s.h16_free, 1);
    mbar_init(&s.y_full, 1);
    mbar_init(&s.y_empty, 128);
    for (int i = 0; i < 8; ++i) mbar_init(&s.res_full[i >> 1][i & 1], 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s.tmem_base;
  const bool have = (int)blockIdx.x < m_tiles;

  if (warp == 0) {
    // ---- TMA producer: att tiles, residual rows, and the weight blocks in the order the MMA warp consumes them ------
    if (lane == 0 && have) {
      uint32_t wit = 0, tl = 0;
      // kind 0: W_out (G0), 1: W1 rows of chunk j (G1), 2: W2 columns of chunk j (G2)
      auto w_blocks = [&](int kind, int j) {
        for (int kb = 0; kb < 2; ++kb, ++wit) {
          const int sw = wit % kTaW;
          mbar_wait(&s.w_empty[sw], ((wit / kTaW) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&s.w_full[sw], 2 * kTileBytes);
          if (kind == 0) {
            tma_load_2d(&tmWohi, &s.w_full[sw], s.w_hi[sw], kb * kBK, 0);
            tma_load_2d(&tmWolo, &s.w_full[sw], s.w_lo[sw], kb * kBK, 0);
          } else if (kind == 1) {
            tma_load_2d(&tmW1hi, &s.w_full[sw], s.w_hi[sw], kb * kBK, j * 128);
            tma_load_2d(&tmW1lo, &s.w_full[sw], s.w_lo[sw], kb * kBK, j * 128);
          } else {
            tma_load_2d(&tmW2hi, &s.w_full[sw], s.w_hi[sw], (2 * j + kb) * kBK, 0);
            tma_load_2d(&tmW2lo, &s.w_full[sw], s.w_lo[sw], (2 * j + kb) * kBK, 0);
          }
        }
      };
      auto load_att = [&](int mt) {
        for (int kb = 0; kb < 4; ++kb) {
          mbar_arrive_expect_tx(&s.att_full[kb], kTileBytes);
          tma_load_2d(&tmA, &s.att_full[kb], s.raw[kb], kb * kBK, mt * kBM);
        }
      };
      auto load_res = [&](int mt, uint32_t t) {             // the tile's x rows, once its att blocks have been split
        mbar_wait(&s.att_done, t & 1u);
        for (int kb = 0; kb < 4; ++kb) {
          mbar_arrive_expect_tx(&s.res_in[kb], kTileBytes);
          tma_load_2d(&tmX, &s.res_in[kb], s.raw[kb], kb * kBK, mt * kBM);
        }
      };
      load_att(blockIdx.x);
      w_blocks(0, 0);
      load_res(blockIdx.x, 0);
      for (int mt = blockIdx.x; mt < m_tiles; mt += stride, ++tl) {
        const bool next = mt + stride < m_tiles;
        w_blocks(1, 0);
        if (next) {
          mbar_wait(&s.h1_done, tl & 1u);                  // this tile's h1 has left the landing blocks
          load_att(mt + stride);
        }
        for (int j = 1; j < C; ++j) {
          w_blocks(1, j);
          w_blocks(2, j - 1);
        }
        if (next) w_blocks(0, 0);
        w_blocks(2, C - 1);
        if (next) load_res(mt + stride, tl + 1);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ---------------------------------------------------------------------------------------------------
    if (have) {
      uint32_t wit = 0, tl = 0, n1 = 0, n2 = 0;
      auto product = [&](uint32_t acc, uint32_t a_base, bool fresh) {
        for (int kb = 0; kb < 2; ++kb, ++wit) {
          const uint32_t sw = wit % kTaW;
          mbar_wait(&s.w_full[sw], (wit / kTaW) & 1u);
          tcgen05_fence_after();
          const uint64_t dw_hi = umma_desc_sw128(s.w_hi[sw]), dw_lo = umma_desc_sw128(s.w_lo[sw]);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
              const uint32_t a_hi = a_base + kb * 32 + k * 8, a_lo = a_hi + 64;
              mma_f16_ta(acc, a_lo, dw_hi + adv, kIdescF16, (fresh && kb == 0 && k == 0) ? 0u : 1u);
              mma_f16_ta(acc, a_hi, dw_lo + adv, kIdescF16, 1u);
              mma_f16_ta(acc, a_hi, dw_hi + adv, kIdescF16, 1u);
            }
            tcgen05_commit(&s.w_empty[sw]);
          }
          __syncwarp();
        }
      };
      auto g0 = [&](uint32_t t) {                           // out-proj of tile t into the hidden accumulator's columns
        mbar_wait(&s.a_ready, t & 1u);
        mbar_wait(&s.hacc_empty, (n1 & 1u) ^ 1u);          // the last hidden chunk's accumulator has been read
        tcgen05_fence_after();
        product(tmem + kFfnH, tmem + kFfnX, true);
        if (elect_one()) tcgen05_commit(&s.o_full);
        __syncwarp();
      };
      auto g1 = [&](int j) {
        mbar_wait(&s.hacc_empty, (n1 & 1u) ^ 1u);
        tcgen05_fence_after();
        product(tmem + kFfnH, tmem + kFfnX, true);
        if (elect_one()) {
          tcgen05_commit(&s.hacc_full);
          if (j == C - 1) tcgen05_commit(&s.x_free);
        }
        __syncwarp();
        ++n1;
      };
      auto g2 = [&](int j) {
        mbar_wait(&s.h16_ready, n2 & 1u);
        if (j == 0) mbar_wait(&s.y_empty, (tl & 1u) ^ 1u);
        tcgen05_fence_after();
        product(tmem + kFfnY, tmem + kFfnH16, j == 0);
        if (elect_one()) {
          tcgen05_commit(&s.h16_free);
          if (j == C - 1) tcgen05_commit(&s.y_full);
        }
        __syncwarp();
        ++n2;
      };
      g0(0);
      for (int mt = blockIdx.x; mt < m_tiles; mt += stride, ++tl) {
        mbar_wait(&s.x_ready, tl & 1u);                    // norm1 has read G0's accumulator and written the X halves
        tcgen05_fence_after();
        g1(0);
        for (int j = 1; j < C; ++j) {
          g1(j);
          g2(j - 1);
        }
        if (mt + stride < m_tiles) g0(tl + 1);
        g2(C - 1);
      }
    }
  } else if (warp < 10) {
    // ---- converters: two threads per row (64 of the 128 columns each) ---------------------------------------------------
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    auto pair_bar = [&]() {          // the two warps of a lane quarter (named barriers 1-4)
      switch (q) {
        case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
      }
    };
    auto split_att = [&](uint32_t t) {
      mbar_wait(&s.att_full[2 * part], t & 1u);
      mbar_wait(&s.att_full[2 * part + 1], t & 1u);
      __syncwarp();
      tcgen05_fence_after();
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int kb = 2 * part + kk;
        const float4* src = reinterpret_cast<const float4*>(s.raw[kb]) + row * 8;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = src[j ^ (row & 7)];
          f16_split2(v.x, v.y, hi[2 * j + 0], lo[2 * j + 0]);
          f16_split2(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
        }
        __syncwarp();
        tmem_st_32x32b_x16(trow + kFfnX + 16 * kb, hi);
        tmem_st_32x32b_x16(trow + kFfnX + 64 + 16 * kb, lo);
      }
      mbar_arrive(&s.att_done);                            // behind the stores, which consume every loaded value
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tcgen05_fence_before();
      mbar_arrive(&s.a_ready);
    };
    // h1 = LayerNorm1(x + (G0 / 2^8 + bo)) for this thread's 64 columns of its row; arithmetic of tf_gemm_ta_kernel<LN>
    auto norm1 = [&](uint32_t t, int m0) {
      mbar_wait(&s.o_full, t & 1u);
      __syncwarp();
      tcgen05_fence_after();
      uint32_t r[2][32];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) tmem_ld_32x32b_x32(trow + kFfnH + 32 * (2 * part + cc), r[cc]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tcgen05_fence_before();
      mbar_wait(&s.res_in[2 * part], t & 1u);
      mbar_wait(&s.res_in[2 * part + 1], t & 1u);
      float pc[2];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * part + cc;
        const float4* blk = reinterpret_cast<const float4*>(s.raw[c]) + row * 8;
        const float4* b4 = reinterpret_cast<const float4*>(a.bo + 32 * c);
        float tt[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = b4[j];
          const float4 hh = blk[j ^ (row & 7)];
          const float a0 = __fadd_rn(hh.x, __fadd_rn(unscale<true>(r[cc][4 * j + 0]), bb.x));
          const float a1 = __fadd_rn(hh.y, __fadd_rn(unscale<true>(r[cc][4 * j + 1]), bb.y));
          const float a2 = __fadd_rn(hh.z, __fadd_rn(unscale<true>(r[cc][4 * j + 2]), bb.z));
          const float a3 = __fadd_rn(hh.w, __fadd_rn(unscale<true>(r[cc][4 * j + 3]), bb.w));
          tt[j] = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
          r[cc][4 * j + 0] = __float_as_uint(a0); r[cc][4 * j + 1] = __float_as_uint(a1);
          r[cc][4 * j + 2] = __float_as_uint(a2); r[cc][4 * j + 3] = __float_as_uint(a3);
        }
        pc[cc] = __fadd_rn(__fadd_rn(__fadd_rn(tt[0], tt[1]), __fadd_rn(tt[2], tt[3])),
                           __fadd_rn(__fadd_rn(tt[4], tt[5]), __fadd_rn(tt[6], tt[7])));
      }
      // (pc0 + pc1) + (pc2 + pc3): each thread has one of the brackets
      float mine = __fadd_rn(pc[0], pc[1]);
      s.xch[0][part][row] = mine;
      pair_bar();
      float other = s.xch[0][part ^ 1][row];
      const float mu = __fmul_rn(part == 0 ? __fadd_rn(mine, other) : __fadd_rn(other, mine), 1.f / 128.f);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        float tt[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d0 = __fsub_rn(__uint_as_float(r[cc][4 * j + 0]), mu), d1 = __fsub_rn(__uint_as_float(r[cc][4 * j + 1]), mu);
          const float d2 = __fsub_rn(__uint_as_float(r[cc][4 * j + 2]), mu), d3 = __fsub_rn(__uint_as_float(r[cc][4 * j + 3]), mu);
          tt[j] = __fmaf_rn(d3, d3, __fmaf_rn(d2, d2, __fmaf_rn(d1, d1, __fmul_rn(d0, d0))));
          r[cc][4 * j + 0] = __float_as_uint(d0); r[cc][4 * j + 1] = __float_as_uint(d1);
          r[cc][4 * j + 2] = __float_as_uint(d2); r[cc][4 * j + 3] = __float_as_uint(d3);
        }
        pc[cc] = __fadd_rn(__fadd_rn(__fadd_rn(tt[0], tt[1]), __fadd_rn(tt[2], tt[3])),
                           __fadd_rn(__fadd_rn(tt[4], tt[5]), __fadd_rn(tt[6], tt[7])));
      }
      mine = __fadd_rn(pc[0], pc[1]);
      s.xch[1][part][row] = mine;
      pair_bar();
      other = s.xch[1][part ^ 1][row];
      const float rs = rsqrtf(__fmaf_rn(part == 0 ? __fadd_rn(mine, other) : __fadd_rn(other, mine), 1.f / 128.f, 1e-5f));
      tcgen05_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * part + cc;
        float4* blk = reinterpret_cast<float4*>(s.raw[c]) + row * 8;
        const float4* g4 = reinterpret_cast<const float4*>(a.g1 + 32 * c);
        const float4* e4 = reinterpret_cast<const float4*>(a.be1 + 32 * c);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g = g4[j], be = e4[j];
          float4 o;
          o.x = __fmaf_rn(__fmul_rn(__uint_as_float(r[cc][4 * j + 0]), rs), g.x, be.x);
          o.y = __fmaf_rn(__fmul_rn(__uint_as_float(r[cc][4 * j + 1]), rs), g.y, be.y);
          o.z = __fmaf_rn(__fmul_rn(__uint_as_float(r[cc][4 * j + 2]), rs), g.z, be.z);
          o.w = __fmaf_rn(__fmul_rn(__uint_as_float(r[cc][4 * j + 3]), rs), g.w, be.w);
          blk[j ^ (row & 7)] = o;                          // over the residual value this thread has consumed
          f16_split2(o.x, o.y, hi[2 * j + 0], lo[2 * j + 0]);
          f16_split2(o.z, o.w, hi[2 * j + 1], lo[2 * j + 1]);
        }
        __syncwarp();
        tmem_st_32x32b_x16(trow + kFfnX + 16 * c, hi);      // G0 is complete (o_full): its att halves are dead
        tmem_st_32x32b_x16(trow + kFfnX + 64 + 16 * c, lo);
      }
      fence_proxy_async();                                 // h1 in the landing blocks -> visible to the TMA store
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tcgen05_fence_before();
      mbar_arrive(&s.x_ready);
      // this half's two blocks (all 128 rows) go to x as the residual of norm2, then back to the producer
      if (part == 0) asm volatile("bar.sync 5, 128;" ::: "memory");
      else asm volatile("bar.sync 6, 128;" ::: "memory");
      if (q == 0 && lane == 0) {                            // completion is awaited a chunk later (h1_flush): not here, where
        tma_store_2d(&tmX, s.raw[2 * part], 32 * (2 * part), m0);          // the whole warp would stand still for it
        tma_store_2d(&tmX, s.raw[2 * part + 1], 32 * (2 * part + 1), m0);
        bulk_commit();
      }
    };
    auto h1_flush = [&]() {
      if (q == 0 && lane == 0) {
        bulk_wait<0>();
        __threadfence();
        mbar_arrive(&s.h1_done);                           // the producer (never behind): the landing blocks are free
        atomicAdd(&s.h1_cnt, 1u);                          // the norm2 warps: h1 is in memory
      }
    };
    if (have) {
      uint32_t tl = 0, ct = 0;
      split_att(0);
      norm1(0, blockIdx.x * kBM);
      for (int mt = blockIdx.x; mt < m_tiles; mt += stride, ++tl) {
        const bool next = mt + stride < m_tiles;
        for (int j = 0; j < C; ++j, ++ct) {
          mbar_wait(&s.hacc_full, ct & 1u);
          __syncwarp();
          tcgen05_fence_after();
          uint32_t r[2][32];
#pragma unroll
          for (int c = 0; c < 2; ++c) tmem_ld_32x32b_x32(trow + kFfnH + 32 * (2 * part + c), r[c]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tcgen05_fence_before();
          mbar_arrive(&s.hacc_empty);
          if (j == 0) h1_flush();                          // the tile's h1 store was issued a product ago
          if (j == C - 1 && next) {
            mbar_wait(&s.x_free, tl & 1u);                 // the tile's first products are done with its X halves
            split_att(tl + 1);
          }
          const float* bj = a.b1 + j * 128 + 64 * part;
          bool waited = false;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t hi[16], lo[16];
            const float4* b4 = reinterpret_cast<const float4*>(bj + 32 * c);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 bb = b4[e];
              const float v0 = fmaxf(__fadd_rn(unscale<true>(r[c][4 * e + 0]), bb.x), 0.f);
              const float v1 = fmaxf(__fadd_rn(unscale<true>(r[c][4 * e + 1]), bb.y), 0.f);
              const float v2 = fmaxf(__fadd_rn(unscale<true>(r[c][4 * e + 2]), bb.z), 0.f);
              const float v3 = fmaxf(__fadd_rn(unscale<true>(r[c][4 * e + 3]), bb.w), 0.f);
              f16_split2(v0, v1, hi[2 * e + 0], lo[2 * e + 0]);
              f16_split2(v2, v3, hi[2 * e + 1], lo[2 * e + 1]);
            }
            if (!waited) {
              mbar_wait(&s.h16_free, (ct & 1u) ^ 1u);
              __syncwarp();
              tcgen05_fence_after();
              waited = true;
            }
            tmem_st_32x32b_x16(trow + kFfnH16 + 16 * (2 * part + c), hi);
            tmem_st_32x32b_x16(trow + kFfnH16 + 64 + 16 * (2 * part + c), lo);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tcgen05_fence_before();
          mbar_arrive(&s.h16_ready);
        }
        if (next) norm1(tl + 1, (mt + stride) * kBM);
      }
    }
  } else {
    // ---- norm2 (thread = row): x <- LayerNorm(h1 + (Y / 2^8 + b2)), as in tf_ffn_kernel; h1 comes back from x ---------
    const int q = warp & 3;
    uint32_t tl = 0;
    // The first two residual boxes of a tile, once its h1 is in memory.  Not a parity wait: with two hidden chunks a tile
    // is shorter than a normalisation, the converters can be two tiles ahead of these warps, and a barrier that has
    // completed phases t AND t + 1 looks, to a wait for phase t, like one that has not completed it.
    auto res_first = [&](int m0, uint32_t t) {
      uint32_t spins = 0;
      while (*reinterpret_cast<volatile uint32_t*>(&s.h1_cnt) < 2u * (t + 1u)) {
        __nanosleep(64);
        if (++spins > (1u << 26)) __trap();                 // like mbar_wait: a lost hand-off must not hang the GPU
      }
      __threadfence();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        mbar_arrive_expect_tx(&s.res_full[q][c], 32 * 32 * 4);
        tma_load_2d(&tmR, &s.res_full[q][c], s.epi[q][c], 32 * c, m0 + q * 32);
      }
    };
    if (lane == 0 && have) res_first(blockIdx.x * kBM, 0);
    for (int mt = blockIdx.x; mt < m_tiles; mt += stride, ++tl) {
      const int m0 = mt * kBM;
      mbar_wait(&s.y_full, tl & 1u);
      __syncwarp();
      tcgen05_fence_after();
      const uint32_t tacc = tmem + kFfnY + ((uint32_t)(q * 32) << 16);
      uint32_t r[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tacc + 32 * c, r[c]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tcgen05_fence_before();
      mbar_arrive(&s.y_empty);
      float pc[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 (*stg)[8] = s.epi[q][c & 1];
        mbar_wait(&s.res_full[q][c & 1], (uint32_t)(c >> 1));
        const float4* b4 = reinterpret_cast<const float4*>(a.b2 + 32 * c);
        float tt[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = b4[j];
          const float4 hh = stg[lane][j ^ (lane & 7)];
          const float a0 = __fadd_rn(hh.x, __fadd_rn(unscale<true>(r[c][4 * j + 0]), bb.x));
          const float a1 = __fadd_rn(hh.y, __fadd_rn(unscale<true>(r[c][4 * j + 1]), bb.y));
          const float a2 = __fadd_rn(hh.z, __fadd_rn(unscale<true>(r[c][4 * j + 2]), bb.z));
          const float a3 = __fadd_rn(hh.w, __fadd_rn(unscale<true>(r[c][4 * j + 3]), bb.w));
          tt[j] = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
          r[c][4 * j + 0] = __float_as_uint(a0); r[c][4 * j + 1] = __float_as_uint(a1);
          r[c][4 * j + 2] = __float_as_uint(a2); r[c][4 * j + 3] = __float_as_uint(a3);
        }
        pc[c] = __fadd_rn(__fadd_rn(__fadd_rn(tt[0], tt[1]), __fadd_rn(tt[2], tt[3])),
                          __fadd_rn(__fadd_rn(tt[4], tt[5]), __fadd_rn(tt[6], tt[7])));
        asm volatile("" ::"f"(pc[c]) : "memory");
        __syncwarp();
        if (c < 2 && lane == 0) {
          mbar_arrive_expect_tx(&s.res_full[q][c & 1], 32 * 32 * 4);
          tma_load_2d(&tmR, &s.res_full[q][c & 1], stg, 32 * (c + 2), m0 + q * 32);
        }
      }
      const float mu = __fmul_rn(__fadd_rn(__fadd_rn(pc[0], pc[1]), __fadd_rn(pc[2], pc[3])), 1.f / 128.f);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float tt[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d0 = __fsub_rn(__uint_as_float(r[c][4 * j + 0]), mu), d1 = __fsub_rn(__uint_as_float(r[c][4 * j + 1]), mu);
          const float d2 = __fsub_rn(__uint_as_float(r[c][4 * j + 2]), mu), d3 = __fsub_rn(__uint_as_float(r[c][4 * j + 3]), mu);
          tt[j] = __fmaf_rn(d3, d3, __fmaf_rn(d2, d2, __fmaf_rn(d1, d1, __fmul_rn(d0, d0))));
          r[c][4 * j + 0] = __float_as_uint(d0); r[c][4 * j + 1] = __float_as_uint(d1);
          r[c][4 * j + 2] = __float_as_uint(d2); r[c][4 * j + 3] = __float_as_uint(d3);
        }
        pc[c] = __fadd_rn(__fadd_rn(__fadd_rn(tt[0], tt[1]), __fadd_rn(tt[2], tt[3])),
                          __fadd_rn(__fadd_rn(tt[4], tt[5]), __fadd_rn(tt[6], tt[7])));
      }
      const float rs = rsqrtf(__fmaf_rn(__fadd_rn(__fadd_rn(pc[0], pc[1]), __fadd_rn(pc[2], pc[3])), 1.f / 128.f, 1e-5f));
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        const float4* g4 = reinterpret_cast<const float4*>(a.g2 + 32 * c);
        const float4* e4 = reinterpret_cast<const float4*>(a.be2 + 32 * c);
        float4 (*stg)[8] = s.epi[q][c & 1];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g = g4[j], be = e4[j];
          float4 o;
          o.x = __fmaf_rn(__fmul_rn(__uint_as_float(r[c][4 * j + 0]), rs), g.x, be.x);
          o.y = __fmaf_rn(__fmul_rn(__uint_as_float(r[c][4 * j + 1]), rs), g.y, be.y);
          o.z = __fmaf_rn(__fmul_rn(__uint_as_float(r[c][4 * j + 2]), rs), g.z, be.z);
          o.w = __fmaf_rn(__fmul_rn(__uint_as_float(r[c][4 * j + 3]), rs), g.w, be.w);
          stg[lane][j ^ (lane & 7)] = o;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmC, stg, 32 * c, m0 + q * 32);
          bulk_commit();
        }
      }
      if (lane == 0 && mt + stride < m_tiles) {
        bulk_wait_read<0>();
        res_first(m0 + stride * kBM, tl + 1);
      }
      __syncwarp();
    }
    if (lane == 0) bulk_wait_read<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// x[M][128] <- LayerNorm2(h1 + linear2(relu(linear1(h1)))), h1 = LayerNorm1(x + att Wo^T + bo): out-proj, both LayerNorms
// and the feed-forward block of an encoder layer in one launch.  All weight halves as tf_split_weights(.., true) writes them.
int tf_tail_fused(const float* att, const float* Wohi, const float* Wolo, const float* bo, const float* g1, const float* be1,
                  float* x, const float* W1hi, const float* W1lo, const float* b1, const float* W2hi, const float* W2lo,
                  const float* b2, const float* g2, const float* be2, int M, int ff, cudaStream_t st) {
  if (ff % 128 != 0 || ff < 128 || M < 1) return QX_EUNSUPPORTED;
  CUtensorMap ma, mx, moh, mol, m1h, m1l, m2h, m2l, mc, mr;
  int rc;
  if ((rc = make_map(&ma, att, M, kBN)) || (rc = make_map(&mx, x, M, kBN)) || (rc = make_map(&moh, Wohi, kBN, kBN / 2)) ||
      (rc = make_map(&mol, Wolo, kBN, kBN / 2)) || (rc = make_map(&m1h, W1hi, ff, kBN / 2)) || (rc = make_map(&m1l, W1lo, ff, kBN / 2)) ||
      (rc = make_map(&m2h, W2hi, kBN, ff / 2)) || (rc = make_map(&m2l, W2lo, kBN, ff / 2)) ||
      (rc = make_map(&mc, x, M, kBN, 32, 32)) || (rc = make_map(&mr, x, M, kBN, 32, 32)))
    return rc;
  static bool cfgd[kMaxDevices] = {};
  if (!cfgd[current_device_slot()]) {
    cudaFuncSetAttribute(tf_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TailSmem));
    cfgd[current_device_slot()] = true;
  }
  const int num_sms = device_sm_count();
  const int m_tiles = ceil_div(M, kBM);
  const int grid = num_sms < m_tiles ? num_sms : m_tiles;
  tf_tail_kernel<<<grid, 448, sizeof(TailSmem), st>>>(ma, mx, moh, mol, m1h, m1l, m2h, m2l, mc, mr,
                                                     TailArgs{bo, g1, be1, b1, b2, g2, be2}, M, ff);
  QX_CHECK_LAUNCH("tf_tail_kernel");
  return QX_OK;
}

// x[M][128] <- LayerNorm(x + linear2(relu(linear1(x)))) with the fp16 halves of W1 [ff][128] and W2 [128][ff] (as
// tf_split_weights(.., true) writes them).  out may be x.  QX_EUNSUPPORTED when ff is not a multiple of 128.
int tf_ffn_fused(const float* x, const float* W1hi, const float* W1lo, const float* b1, const float* W2hi, const float* W2lo,
                 const float* b2, const float* gamma, const float* beta, float* out, int M, int ff, cudaStream_t st) {
  if (ff % 128 != 0 || ff < 128 || M < 1) return QX_EUNSUPPORTED;
  CUtensorMap mx, m1h, m1l, m2h, m2l, mc, mr;
  int rc;
  if ((rc = make_map(&mx, x, M, kBN)) || (rc = make_map(&m1h, W1hi, ff, kBN / 2)) || (rc = make_map(&m1l, W1lo, ff, kBN / 2)) ||
      (rc = make_map(&m2h, W2hi, kBN, ff / 2)) || (rc = make_map(&m2l, W2lo, kBN, ff / 2)) ||
      (rc = make_map(&mc, out, M, kBN, 32, 32)) || (rc = make_map(&mr, x, M, kBN, 32, 32)))
    return rc;
  static bool cfgd[kMaxDevices] = {};
  if (!cfgd[current_device_slot()]) {
    cudaFuncSetAttribute(tf_ffn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FfnSmem));
    cfgd[current_device_slot()] = true;
  }
  const int num_sms = device_sm_count();
  const int m_tiles = ceil_div(M, kBM);
  const int grid = num_sms < m_tiles ? num_sms : m_tiles;
  tf_ffn_kernel<<<grid, 448, sizeof(FfnSmem), st>>>(mx, m1h, m1l, m2h, m2l, mc, mr, b1, b2, LnArgs{gamma, beta}, M, ff);
  QX_CHECK_LAUNCH("tf_ffn_kernel");
  return QX_OK;
}

int tf_split_weights(const float* w, float* hi, float* lo, size_t count, cudaStream_t st, bool f16) {
  if (f16) tf_split16_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(w, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), count);
  else tf_split_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(w, hi, lo, count);
  QX_CHECK_LAUNCH("tf_split_kernel");
  return QX_OK;
}

}  // namespace qx
