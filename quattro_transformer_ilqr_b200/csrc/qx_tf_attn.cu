// Causal multi-head attention of the Quattro transformer on tensor cores.
//
//   = the self-attention inside nn.TransformerEncoderLayer (transformer_model.py:112-113, :134-136: bool causal
//     mask triu(ones(L, L), 1), 4 heads x 32, softmax(Q K^T / sqrt(32)) V), for every (trajectory, head).
//
// The sequences are short (L = 101 / 61 tokens) and the head dimension is 32, so a (trajectory, head) pair is a
// 101x32 . 32x101 . 101x32 problem: far below a tcgen05 tile (M = 128 rows of ONE operand pair) but a good fit for
// warp-level mma.sync.m16n8k8 TF32 tiles.  One CTA per (trajectory, head):
//   * K and V rows are staged in shared memory as fp32 (30 KB for L = 101, so seven CTAs share an SM) and every
//     fragment is split into TF32 (hi, lo) halves on the fly (3xTF32: hi*hi + lo*hi + hi*lo, fp32-accurate
//     products - the 1e-5 parity bar rules plain TF32 out); row pitch 36 floats makes both B-fragment access
//     patterns bank-conflict free;
//   * a warp owns 16 query rows at a time; causality means query tile q needs key tiles 0..2q+1 only, and warp w
//     takes the tile pair (w, nq-1-w) so that the warps carry equal work;
//   * all score tiles of a query tile stay in registers (<= 14 tiles x 4), so the softmax is the exact
//     max-subtracted one of the reference (no online rescaling);
//   * P feeds the second product straight from the accumulator registers: the k-index of that MMA is relabelled
//     (t <-> key 2t, t+4 <-> key 2t+1), which is just a different row pattern when reading V.
#include "qx_common.cuh"

namespace qx {

namespace {

constexpr int kHd = 32;          // head dimension
constexpr int kPitch = 36;       // shared-memory row pitch (floats)
constexpr int kMaxKt = 14;       // key tiles (of 8) a query tile can need: L <= 112

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const float (&a)[4], float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])),
                 "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

}  // namespace

// grid (nhead, B); block 32 * ceil(nq / 2) threads, nq = ceil(L / 16); dynamic shared memory 2 * Lp * 36 floats
__global__ void __launch_bounds__(128)
tf_attention_mma_kernel(const float* __restrict__ qkv, float* __restrict__ out, int L, int d) {
  extern __shared__ float smf[];
  const int Lp = (L + 7) & ~7;                       // keys padded to whole tiles
  float* Ks = smf;                                   // fp32 rows; the TF32 (hi, lo) split is done per fragment
  float* Vs = Ks + (size_t)Lp * kPitch;
  const int head = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* base = qkv + (size_t)b * L * 3 * d + head * kHd;

  // stage K, V (rows beyond L are zero)
  for (int e = tid; e < Lp * (kHd / 4); e += blockDim.x) {
    const int r = e >> 3, c4 = e & 7;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
    if (r < L) {
      kv = *reinterpret_cast<const float4*>(base + (size_t)r * 3 * d + d + 4 * c4);
      vv = *reinterpret_cast<const float4*>(base + (size_t)r * 3 * d + 2 * d + 4 * c4);
    }
    *reinterpret_cast<float4*>(&Ks[r * kPitch + 4 * c4]) = kv;
    *reinterpret_cast<float4*>(&Vs[r * kPitch + 4 * c4]) = vv;
  }
  __syncthreads();

  const int nq = (L + 15) >> 4;
  const float scale_log2 = rsqrtf((float)kHd) * 1.4426950408889634f;     // softmax in base 2: exp(x) = 2^(x log2 e)
  for (int pass = 0; pass < 2; ++pass) {
    const int qt = (pass == 0) ? warp : nq - 1 - warp;
    if (qt >= nq || (pass == 1 && qt <= warp)) break;                     // the middle tile is done once
    const int r0 = 16 * qt + g, r1 = r0 + 8;                              // this lane's two query rows
    // Q fragments, split: a0 (r0, 8kk+t), a1 (r1, 8kk+t), a2 (r0, 8kk+t+4), a3 (r1, 8kk+t+4)
    float qh[4][4], ql[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float v0 = (r0 < L) ? base[(size_t)r0 * 3 * d + 8 * kk + t] : 0.f;
      const float v1 = (r1 < L) ? base[(size_t)r1 * 3 * d + 8 * kk + t] : 0.f;
      const float v2 = (r0 < L) ? base[(size_t)r0 * 3 * d + 8 * kk + t + 4] : 0.f;
      const float v3 = (r1 < L) ? base[(size_t)r1 * 3 * d + 8 * kk + t + 4] : 0.f;
      const float v[4] = {v0, v1, v2, v3};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        qh[kk][i] = to_tf32(v[i]);
        ql[kk][i] = to_tf32(v[i] - qh[kk][i]);
      }
    }
    const int nkt = min(2 * qt + 2, Lp >> 3);                             // key tiles 0 .. nkt-1 (causal)
    float s[kMaxKt][4];
#pragma unroll
    for (int jt = 0; jt < kMaxKt; ++jt) {
      s[jt][0] = s[jt][1] = s[jt][2] = s[jt][3] = 0.f;
      if (jt < nkt) {
        const float* kp = &Ks[(8 * jt + g) * kPitch + t];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float k0 = kp[8 * kk], k1 = kp[8 * kk + 4];
          const float bh0 = to_tf32(k0), bh1 = to_tf32(k1), bl0 = to_tf32(k0 - bh0), bl1 = to_tf32(k1 - bh1);
          mma_tf32(s[jt], ql[kk], bh0, bh1);
          mma_tf32(s[jt], qh[kk], bl0, bl1);
          mma_tf32(s[jt], qh[kk], bh0, bh1);
        }
      }
    }
    // mask (key > row), scale, row maxima: c0 (r0, 8jt+2t), c1 (r0, +1), c2 (r1, 8jt+2t), c3 (r1, +1)
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int jt = 0; jt < kMaxKt; ++jt) {
      if (jt < nkt) {
        const int c = 8 * jt + 2 * t;
        s[jt][0] = (c <= r0) ? s[jt][0] * scale_log2 : -INFINITY;
        s[jt][1] = (c + 1 <= r0) ? s[jt][1] * scale_log2 : -INFINITY;
        s[jt][2] = (c <= r1) ? s[jt][2] * scale_log2 : -INFINITY;
        s[jt][3] = (c + 1 <= r1) ? s[jt][3] * scale_log2 : -INFINITY;
        m0 = fmaxf(m0, fmaxf(s[jt][0], s[jt][1]));
        m1 = fmaxf(m1, fmaxf(s[jt][2], s[jt][3]));
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    // P = 2^(s - m), row sums, O += P V
    float l0 = 0.f, l1 = 0.f;
    float o[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int jt = 0; jt < kMaxKt; ++jt) {
      if (jt < nkt) {
        const float p0 = exp2f(s[jt][0] - m0), p1 = exp2f(s[jt][1] - m0);
        const float p2 = exp2f(s[jt][2] - m1), p3 = exp2f(s[jt][3] - m1);
        l0 += p0 + p1;
        l1 += p2 + p3;
        // A fragment of the second product with the k-index relabelled: a0 = P(r0, key 2t), a1 = P(r1, key 2t),
        // a2 = P(r0, key 2t+1), a3 = P(r1, key 2t+1); B: b0 = V(key 2t, dim g), b1 = V(key 2t+1, dim g)
        const float pf[4] = {p0, p2, p1, p3};
        float ph[4], pl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ph[i] = to_tf32(pf[i]);
          pl[i] = to_tf32(pf[i] - ph[i]);
        }
        const float* vp = &Vs[(8 * jt + 2 * t) * kPitch + g];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float v0 = vp[8 * nt], v1 = vp[kPitch + 8 * nt];
          const float bh0 = to_tf32(v0), bh1 = to_tf32(v1), bl0 = to_tf32(v0 - bh0), bl1 = to_tf32(v1 - bh1);
          mma_tf32(o[nt], pl, bh0, bh1);
          mma_tf32(o[nt], ph, bl0, bl1);
          mma_tf32(o[nt], ph, bh0, bh1);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    // o[nt]: c0 (r0, 8nt+2t), c1 (r0, +1), c2 (r1, 8nt+2t), c3 (r1, +1)
    float* dst = out + (size_t)b * L * d + head * kHd;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (r0 < L) *reinterpret_cast<float2*>(dst + (size_t)r0 * d + 8 * nt + 2 * t) = make_float2(o[nt][0] * i0, o[nt][1] * i0);
      if (r1 < L) *reinterpret_cast<float2*>(dst + (size_t)r1 * d + 8 * nt + 2 * t) = make_float2(o[nt][2] * i1, o[nt][3] * i1);
    }
  }
}

// Returns QX_EUNSUPPORTED when the shape is outside the kernel's range (the caller then uses the SIMT kernel).
int tf_attention_tc(const float* qkv, float* out, int B, int L, int d, int nhead, cudaStream_t st) {
  if (d != nhead * kHd || L > 8 * kMaxKt || L < 1) return QX_EUNSUPPORTED;
  const int Lp = (L + 7) & ~7, nq = (L + 15) >> 4;
  const size_t smem = (size_t)2 * Lp * kPitch * sizeof(float);
  static bool cfg[kMaxDevices] = {};
  if (!cfg[current_device_slot()]) {
    cudaFuncSetAttribute(tf_attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8 * kMaxKt * kPitch * 4);
    cfg[current_device_slot()] = true;
  }
  tf_attention_mma_kernel<<<dim3(nhead, B), 32 * ((nq + 1) / 2), smem, st>>>(qkv, out, L, d);
  QX_CHECK_LAUNCH("tf_attention_mma_kernel");
  return QX_OK;
}

}  // namespace qx
