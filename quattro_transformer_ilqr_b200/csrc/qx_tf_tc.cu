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

#include "qx_async.cuh"
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
constexpr size_t kGemmSmemBytes = sizeof(GemmSmem) + 1024;   // + slack to align the ring to 1024 B

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major, 128-B swizzle: rows at 128-B pitch, 8-row groups 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3fff);       // start address
  d |= (uint64_t)1 << 16;                                // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset
  d |= (uint64_t)1 << 46;                                // descriptor version
  d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
  return d;
}
// hi = a rounded to the nearest TF32 (10 explicit mantissa bits), lo = the exact remainder rounded
// to TF32 as well: a*w - (hi*w_hi + hi*w_lo + lo*w_hi) is then ~3 * 2^-24 |a w| (fp32-class accuracy).
__device__ __forceinline__ float tf32_rn(float v) {
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void tf32_split(float v, float& hi, float& lo) {
  hi = tf32_rn(v);
  lo = tf32_rn(v - hi);
}
// D = F32, A = B = TF32, both K-major, M = 128, N = 128
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

template <bool RELU>
__global__ void __launch_bounds__(320, 1)
tf_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWhi,
                  const __grid_constant__ CUtensorMap tmWlo, const float* __restrict__ bias, float* __restrict__ C,
                  int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  GemmSmem& s = *reinterpret_cast<GemmSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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

// W -> (W_hi, W_lo), elementwise, done once per weight upload
__global__ void tf_split_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, size_t count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float h, l;
  tf32_split(w[i], h, l);
  hi[i] = h;
  lo[i] = l;
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major fp32 matrix [rows][cols] -> tiled map with a (32 x 128) box and 128-B swizzle
static int make_map(CUtensorMap* map, const float* base, int rows, int cols) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return QX_ECUDA;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kBM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %d cols %d)", (int)r, rows, cols);
    return QX_ECUDA;
  }
  return QX_OK;
}

int tf_gemm_tc(const float* A, const float* Whi, const float* Wlo, const float* bias, float* C, int M, int N, int K,
               bool relu, cudaStream_t st) {
  if (N % kBN != 0 || K % kBK != 0) {
    set_error("tensor-core GEMM needs N %% 128 == 0 and K %% 32 == 0 (got N=%d K=%d)", N, K);
    return QX_EINVAL;
  }
  CUtensorMap ma, mh, ml;
  int rc;
  if ((rc = make_map(&ma, A, M, K)) || (rc = make_map(&mh, Whi, N, K)) || (rc = make_map(&ml, Wlo, N, K))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(tf_gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes);
    cudaFuncSetAttribute(tf_gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes);
    configured = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int tiles = (N / kBN) * ceil_div(M, kBM);
  const int grid = tiles < num_sms ? tiles : num_sms;          // persistent: one CTA per SM
  if (relu) tf_gemm_tc_kernel<true><<<grid, 320, kGemmSmemBytes, st>>>(ma, mh, ml, bias, C, M, N, K);
  else tf_gemm_tc_kernel<false><<<grid, 320, kGemmSmemBytes, st>>>(ma, mh, ml, bias, C, M, N, K);
  QX_CHECK_LAUNCH("tf_gemm_tc_kernel");
  return QX_OK;
}

int tf_split_weights(const float* w, float* hi, float* lo, size_t count, cudaStream_t st) {
  tf_split_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(w, hi, lo, count);
  QX_CHECK_LAUNCH("tf_split_kernel");
  return QX_OK;
}

}  // namespace qx
