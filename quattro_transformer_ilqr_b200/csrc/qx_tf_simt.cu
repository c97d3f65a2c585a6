// Quattro transformer forward — CUDA-core (SIMT fp32) path.
//
//   TransformerILQR.predict          quattro_ilqr_tf/transformer_ilqr.py:311-325
//   TransformerPredictor.forward     quattro_ilqr_tf/transformer_model.py:122-138
//   prompt pack / prediction unpack  quattro_ilqr_tf/quattro_ilqr_tf.py:496-518
// Encoder layer semantics are torch's nn.TransformerEncoderLayer defaults (post-LN, ReLU,
// LayerNorm eps 1e-5, causal -inf mask over ALL L tokens, 1/sqrt(head_dim) scaling).
//
// This file holds the exact-fp32 reference path used for small batches and as the numerics
// anchor of the tensor-core path (qx_tf_tc.cu): embed -> [QKV GEMM -> attention -> out-proj GEMM
// -> add+LN -> FFN1 GEMM+ReLU -> FFN2 GEMM -> add+LN] x layers -> output head + de-normalise +
// unpack into gain records.
#include <stdlib.h>
#include "qx_tf_common.cuh"

namespace qx {

// ---- token embedding ---------------------------------------------------------------------------
// grid (L, B), block d threads.  Token rows: [0, N] states, (N, N+P] prompt, rest targets.
template <int LAYOUT>
__global__ void tf_embed_kernel(const QxTfConfig cfg, const float* __restrict__ w, const double* __restrict__ x,
                                const double* __restrict__ x_shift, const float* __restrict__ gains,
                                float* __restrict__ h, int B) {
  const TfLayout lay(cfg);
  const int tok = blockIdx.x, b = blockIdx.y, ch = threadIdx.x;
  const int N = cfg.horizon, n = lay.n, c = lay.c, d = lay.d, P = lay.P, m = cfg.m;
  __shared__ float in[64];
  float acc;
  if (tok <= N) {
    if (ch < n) {
      // (x - x_ref + state_offset - mean) / std in fp64, then float32  (transformer_ilqr.py:312-315)
      const double v = x[((size_t)tok * n + ch) * B + b] + x_shift[ch];
      in[ch] = (float)((v - (double)w[lay.x_mean + ch]) / (double)w[lay.x_std + ch]);
    }
    __syncthreads();
    acc = w[lay.bs + ch];
    for (int i = 0; i < n; ++i) acc += w[lay.ws + (size_t)ch * n + i] * in[i];
  } else if (tok <= N + P) {
    const int t = N - P + (tok - N - 1);             // absolute time of this prompt row
    if (ch < c) {
      const int G = m + m * n;
      const float g = gains[gain_index<LAYOUT>(t, ch, b, B, G)];
      in[ch] = (float)(((double)g - (double)w[lay.u_mean + ch]) / (double)w[lay.u_std + ch]);
    }
    __syncthreads();
    acc = w[lay.bc + ch];
    for (int i = 0; i < c; ++i) acc += w[lay.wc + (size_t)ch * c + i] * in[i];
  } else {
    acc = w[lay.tgt + (size_t)(tok - N - 1 - P) * d + ch];
  }
  h[((size_t)b * lay.L + tok) * d + ch] = acc + w[lay.pe + (size_t)tok * d + ch];
}

// Same embedding for d == 128, one CTA per TRAJECTORY (the per-token CTAs above spend their time on launch
// overhead: 0.62 ms at 8 192 trajectories against a 0.07 ms write stream).  Warp w takes the tokens tok = w (mod 4),
// a lane owns four channels: its slice of state_embed.weight stays in registers for the 51 state rows, the inputs
// are normalised by the lanes i < n (fp64, like predict) and broadcast by shuffles; the target rows are a copy of
// target_embedding + pe.  The accumulation order per channel is that of tf_embed_kernel (bit-identical output).
template <int LAYOUT>
__global__ void __launch_bounds__(128)
tf_embed128_kernel(const QxTfConfig cfg, const float* __restrict__ w, const double* __restrict__ x,
                   const double* __restrict__ x_shift, const float* __restrict__ gains, float* __restrict__ h, int B) {
  const TfLayout lay(cfg);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = cfg.horizon, n = lay.n, c = lay.c, P = lay.P, m = cfg.m, L = lay.L, T = lay.T;
  const int ch = 4 * lane;
  constexpr unsigned FULL = 0xffffffffu;
  // Persistent CTAs (a few per SM), each walking trajectories b, b + grid, ..: the weight staging below is paid once
  // per CTA instead of once per trajectory.
  // Both embedding weights are staged TRANSPOSED in shared memory ([input][channel]) with coalesced global reads: read
  // in place ([channel][input], a lane's four channels 4 n floats apart) every load touched 32 sectors, 12.7 K of the
  // 16 K L1 sectors a CTA moved, and the kernel ran at 1.2 TB/s of output.
  __shared__ float xin[128 * QX_MAX_N];
  __shared__ __align__(16) float sws[QX_MAX_N * 128];
  __shared__ __align__(16) float swc[64 * 128];
  for (int e = threadIdx.x; e < 128 * n; e += blockDim.x) sws[(e % n) * 128 + e / n] = w[lay.ws + e];
  for (int e = threadIdx.x; e < 128 * c; e += blockDim.x) swc[(e % c) * 128 + e / c] = w[lay.wc + e];
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
  float* hb = h + (size_t)b * L * 128;
  __syncthreads();          // the previous trajectory's rows are done with xin (and, first time, the weights are staged)
  // all (N+1) n inputs of the trajectory: 612 independent strided fp64 loads in flight at once, normalised in fp64
  // like predict (transformer_ilqr.py:312-315) and parked as the fp32 model input
  for (int e = threadIdx.x; e < (N + 1) * n; e += blockDim.x) {
    const int tok = e / n, i = e - tok * n;
    const double mu = (double)w[lay.x_mean + i], sd = (double)w[lay.x_std + i];
    xin[e] = (float)(((x[((size_t)tok * n + i) * B + b] + x_shift[i]) - mu) / sd);
  }
  __syncthreads();
  {   // ---- state rows
    float wreg[4][QX_MAX_N];
#pragma unroll
    for (int i = 0; i < QX_MAX_N; ++i) {
      const float4 wv = (i < n) ? *reinterpret_cast<const float4*>(&sws[i * 128 + ch]) : make_float4(0.f, 0.f, 0.f, 0.f);
      wreg[0][i] = wv.x; wreg[1][i] = wv.y; wreg[2][i] = wv.z; wreg[3][i] = wv.w;
    }
    const float4 bias = *reinterpret_cast<const float4*>(&w[lay.bs + ch]);
    for (int tok = warp; tok <= N; tok += 4) {
      float acc[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
      for (int i = 0; i < QX_MAX_N; ++i) {
        if (i < n) {
          const float xi = xin[tok * n + i];
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[q] += wreg[q][i] * xi;
        }
      }
      const float4 pe = *reinterpret_cast<const float4*>(&w[lay.pe + (size_t)tok * 128 + ch]);
      *reinterpret_cast<float4*>(&hb[(size_t)tok * 128 + ch]) = make_float4(acc[0] + pe.x, acc[1] + pe.y, acc[2] + pe.z, acc[3] + pe.w);
    }
  }
  // ---- prompt rows: the last P Riccati gain records, [k | K row-major] (quattro_ilqr_tf.py:496-502)
  for (int pr = warp; pr < P; pr += 4) {
    const int tok = N + 1 + pr, t = N - P + pr;
    const int G = m + m * n;
    float in[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int o = lane + 32 * k;
      if (o < c) {
        const float g = gains[gain_index<LAYOUT>(t, o, b, B, G)];
        in[k] = (float)(((double)g - (double)w[lay.u_mean + o]) / (double)w[lay.u_std + o]);
      }
    }
    const float4 bias = *reinterpret_cast<const float4*>(&w[lay.bc + ch]);
    float acc[4] = {bias.x, bias.y, bias.z, bias.w};
    for (int i = 0; i < c; ++i) {
      const float xi = __shfl_sync(FULL, (i < 32) ? in[0] : in[1], i & 31);
      const float4 wv = *reinterpret_cast<const float4*>(&swc[i * 128 + ch]);
      acc[0] += wv.x * xi; acc[1] += wv.y * xi; acc[2] += wv.z * xi; acc[3] += wv.w * xi;
    }
    const float4 pe = *reinterpret_cast<const float4*>(&w[lay.pe + (size_t)tok * 128 + ch]);
    *reinterpret_cast<float4*>(&hb[(size_t)tok * 128 + ch]) = make_float4(acc[0] + pe.x, acc[1] + pe.y, acc[2] + pe.z, acc[3] + pe.w);
  }
  // ---- target rows: learnable queries
  for (int k = warp; k < T; k += 4) {
    const int tok = N + 1 + P + k;
    const float4 tg = *reinterpret_cast<const float4*>(&w[lay.tgt + (size_t)k * 128 + ch]);
    const float4 pe = *reinterpret_cast<const float4*>(&w[lay.pe + (size_t)tok * 128 + ch]);
    *reinterpret_cast<float4*>(&hb[(size_t)tok * 128 + ch]) = make_float4(tg.x + pe.x, tg.y + pe.y, tg.z + pe.z, tg.w + pe.w);
  }
  }   // trajectories of this CTA
}

// ---- C[M][N] = act(A[M][K] W[N][K]^T + bias) ----------------------------------------------------
// 64x64 tile, 16-wide K slabs, 256 threads, 4x4 outputs per thread.  N, K multiples of 16 / 64.
template <bool RELU>
__global__ void __launch_bounds__(256)
tf_gemm_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
               float* __restrict__ C, int M, int N, int K) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int r = e >> 4, kk = e & 15;
      const int gm = m0 + r, gn = n0 + r;
      As[kk][r] = (gm < M) ? A[(size_t)gm * K + k0 + kk] : 0.f;
      Ws[kk][r] = (gn < N) ? W[(size_t)gn * K + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) wv[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * wv[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j] + bias[gn];
      if (RELU) v = fmaxf(v, 0.f);
      C[(size_t)gm * N + gn] = v;
    }
  }
}

// ---- causal multi-head attention: grid (nhead, B), block 128 threads (one query row each) ------
// K and V rows (128 B) are read as warp-broadcast 16-B vectors.  Two passes over the keys (row
// maximum, then exp / PV) recompute the 32-term score instead of parking it: 26 KB of shared memory
// per block keeps 8 blocks (32 warps) resident per SM, which is what hides the LDS latency.
template <int HD>
__global__ void __launch_bounds__(128)
tf_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int L, int d) {
  extern __shared__ float4 sm4[];
  float4* Ks = sm4;                                  // [L][HD/4]
  float4* Vs = sm4 + (size_t)L * (HD / 4);
  const int head = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* base = qkv + (size_t)b * L * 3 * d;
  for (int e = tid; e < L * (HD / 4); e += 128) {
    const int r = e / (HD / 4), c4 = e % (HD / 4);
    Ks[e] = *reinterpret_cast<const float4*>(base + (size_t)r * 3 * d + d + head * HD + 4 * c4);
    Vs[e] = *reinterpret_cast<const float4*>(base + (size_t)r * 3 * d + 2 * d + head * HD + 4 * c4);
  }
  __syncthreads();
  const int i = tid;
  const bool live = i < L;
  const float scale = rsqrtf((float)HD);
  float4 q[HD / 4];
#pragma unroll
  for (int c = 0; c < HD / 4; ++c) {
    q[c] = live ? *reinterpret_cast<const float4*>(base + (size_t)i * 3 * d + head * HD + 4 * c) : make_float4(0, 0, 0, 0);
    q[c].x *= scale; q[c].y *= scale; q[c].z *= scale; q[c].w *= scale;
  }
  const int jmax = min(L - 1, (tid | 31));           // warp-uniform loop bound (causal: j <= i)
  auto score = [&](int j) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 4; c += 2) {
      const float4 k0 = Ks[j * (HD / 4) + c], k1 = Ks[j * (HD / 4) + c + 1];
      s0 = fmaf(q[c].x, k0.x, s0); s0 = fmaf(q[c].y, k0.y, s0); s0 = fmaf(q[c].z, k0.z, s0); s0 = fmaf(q[c].w, k0.w, s0);
      s1 = fmaf(q[c + 1].x, k1.x, s1); s1 = fmaf(q[c + 1].y, k1.y, s1); s1 = fmaf(q[c + 1].z, k1.z, s1); s1 = fmaf(q[c + 1].w, k1.w, s1);
    }
    return s0 + s1;
  };
  float mx = -INFINITY;
  for (int j = 0; j <= jmax; ++j) {
    const float sc = score(j);
    if (j <= i) mx = fmaxf(mx, sc);
  }
  float den = 0.f;
  float4 o[HD / 4];
#pragma unroll
  for (int c = 0; c < HD / 4; ++c) o[c] = make_float4(0, 0, 0, 0);
  for (int j = 0; j <= jmax; ++j) {
    const float pj = (j <= i) ? expf(score(j) - mx) : 0.f;
    den += pj;
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float4 v = Vs[j * (HD / 4) + c];
      o[c].x = fmaf(pj, v.x, o[c].x); o[c].y = fmaf(pj, v.y, o[c].y);
      o[c].z = fmaf(pj, v.z, o[c].z); o[c].w = fmaf(pj, v.w, o[c].w);
    }
  }
  if (live) {
    const float inv = 1.f / den;
    float4* dst = reinterpret_cast<float4*>(out + ((size_t)b * L + i) * d + head * HD);
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) dst[c] = make_float4(o[c].x * inv, o[c].y * inv, o[c].z * inv, o[c].w * inv);
  }
}

// ---- h = LayerNorm(h + y) * w + b, one warp per row (d == 128) -------------------------------
// Residual source: row r of the output takes h_src row (r / per) * src_per + src_off + (r % per).  The plain layer
// uses h_src == h with per == src_per == rows and src_off == 0; the pruned last layer gathers the T query rows of
// every trajectory out of the full [B][L] stream into a compact [B][T] one (per = T, src_per = L, src_off = L - T).
__global__ void __launch_bounds__(256)
tf_add_ln_kernel(float* __restrict__ h, const float* __restrict__ h_src, const float* __restrict__ y,
                 const float* __restrict__ w, const float* __restrict__ bvec, int rows, int per, int src_per, int src_off) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const size_t srow = (size_t)(row / per) * src_per + src_off + (row % per);
  float4 a = reinterpret_cast<const float4*>(h_src + srow * 128)[lane];
  const float4 r = reinterpret_cast<const float4*>(y + (size_t)row * 128)[lane];
  // Every rounding below is spelled out (no contraction left to the compiler) and the butterflies run low bit first,
  // so that the thread-per-row epilogue of the fused GEMM + LN kernel (qx_tf_tc.cu) reproduces these bits exactly:
  // groups of 8 lanes = one 32-column chunk of its row, then the four chunk partials.
  a.x = __fadd_rn(a.x, r.x); a.y = __fadd_rn(a.y, r.y); a.z = __fadd_rn(a.z, r.z); a.w = __fadd_rn(a.w, r.w);
  float s = __fadd_rn(__fadd_rn(__fadd_rn(a.x, a.y), a.z), a.w);
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
  const float mu = __fmul_rn(s, 1.f / 128.f);
  const float dx = __fsub_rn(a.x, mu), dy = __fsub_rn(a.y, mu), dz = __fsub_rn(a.z, mu), dw = __fsub_rn(a.w, mu);
  float v = __fmaf_rn(dw, dw, __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  const float rs = rsqrtf(__fmaf_rn(v, 1.f / 128.f, 1e-5f));
  const float4 g = reinterpret_cast<const float4*>(w)[lane];
  const float4 be = reinterpret_cast<const float4*>(bvec)[lane];
  float4 o4;
  o4.x = __fmaf_rn(__fmul_rn(dx, rs), g.x, be.x); o4.y = __fmaf_rn(__fmul_rn(dy, rs), g.y, be.y);
  o4.z = __fmaf_rn(__fmul_rn(dz, rs), g.z, be.z); o4.w = __fmaf_rn(__fmul_rn(dw, rs), g.w, be.w);
  reinterpret_cast<float4*>(h + (size_t)row * 128)[lane] = o4;
}

// ---- output head on the target rows + de-normalise + unpack into gain records ------------------
// grid (T, B), block 64 threads (>= c).  pred[t].reshape(m, 1+n): k = [:,0], K = [:,1:]
template <int LAYOUT>
__global__ void tf_output_kernel(const QxTfConfig cfg, const float* __restrict__ w, const float* __restrict__ h,
                                 float* __restrict__ gains, const int32_t* __restrict__ status, int B) {
  const TfLayout lay(cfg);
  const int t = blockIdx.x, b = blockIdx.y, o = threadIdx.x;
  if (status != nullptr && status[b] != 0) return;
  const int d = lay.d, c = lay.c, m = cfg.m, n = lay.n;
  __shared__ float row[128];
  const float* hr = h + ((size_t)b * lay.L + (lay.L - lay.T + t)) * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) row[i] = hr[i];
  __syncthreads();
  if (o >= c) return;
  float acc = w[lay.bo + o];
  for (int i = 0; i < d; ++i) acc += w[lay.wo + (size_t)o * d + i] * row[i];
  const float val = acc * w[lay.u_std + o] + w[lay.u_mean + o];      // float32 de-normalisation
  const int a = o / (1 + n), j = o % (1 + n);                        // row a of [k_a | K_a,:]
  const int slot = (j == 0) ? a : (m + a * n + (j - 1));
  const int G = m + m * n;
  gains[gain_index<LAYOUT>(t, slot, b, B, G)] = val;
}

// de-normalise + unpack of a precomputed output head y[rows][128] (tensor-core path)
template <int LAYOUT>
__global__ void tf_unpack_kernel(const QxTfConfig cfg, const float* __restrict__ w, const float* __restrict__ y,
                                 float* __restrict__ gains, const int32_t* __restrict__ status, int B) {
  const TfLayout lay(cfg);
  const int t = blockIdx.x, b = blockIdx.y, o = threadIdx.x;
  if (status != nullptr && status[b] != 0) return;
  const int c = lay.c, m = cfg.m, n = lay.n;
  if (o >= c) return;
  const float acc = y[((size_t)b * lay.L + (lay.L - lay.T + t)) * 128 + o];
  const float val = acc * w[lay.u_std + o] + w[lay.u_mean + o];
  const int a = o / (1 + n), j = o % (1 + n);
  const int slot = (j == 0) ? a : (m + a * n + (j - 1));
  const int G = m + m * n;
  gains[gain_index<LAYOUT>(t, slot, b, B, G)] = val;
}

// the same, one CTA per trajectory (c <= 64): warp w takes the target rows t = w (mod 4)
template <int LAYOUT>
__global__ void __launch_bounds__(128)
tf_unpack128_kernel(const QxTfConfig cfg, const float* __restrict__ w, const float* __restrict__ y,
                    float* __restrict__ gains, const int32_t* __restrict__ status, int B, int compact) {
  const TfLayout lay(cfg);
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (status != nullptr && status[b] != 0) return;
  const int c = lay.c, m = cfg.m, n = lay.n, G = m + m * n;
  for (int t = warp; t < lay.T; t += 4) {
    // y holds the head output for all L rows of a trajectory, or (compact) for its T query rows only
    const float* row = compact ? y + ((size_t)b * lay.T + t) * 128 : y + ((size_t)b * lay.L + (lay.L - lay.T + t)) * 128;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int o = lane + 32 * k;
      if (o < c) {
        const float val = row[o] * w[lay.u_std + o] + w[lay.u_mean + o];
        const int a = o / (1 + n), j = o % (1 + n);
        const int slot = (j == 0) ? a : (m + a * n + (j - 1));
        gains[gain_index<LAYOUT>(t, slot, b, B, G)] = val;
      }
    }
  }
}

// the same for 32 trajectories x one target row per CTA, transposed through shared memory: the head rows are read as
// whole 4 c-byte pieces and the gains leave as 512-B lines along the trajectory index (a warp of the per-trajectory
// kernel above scattered 13 sixteen-byte pieces 16 B lines apart: 77 us at 8 192 quadrotor trajectories, 24 % of DRAM)
template <int LAYOUT>
__global__ void __launch_bounds__(128)
tf_unpack_tile_kernel(const QxTfConfig cfg, const float* __restrict__ w, const float* __restrict__ y,
                      float* __restrict__ gains, const int32_t* __restrict__ status, int B, int compact) {
  const TfLayout lay(cfg);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.x * 32, t = blockIdx.y;
  const int c = lay.c, m = cfg.m, n = lay.n, G = m + m * n;
  __shared__ float tile[64][33];
  float v[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int b = b0 + warp + 4 * i;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int o = lane + 32 * k;
      if (b < B && o < c) {
        const float* row = compact ? y + ((size_t)b * lay.T + t) * 128 : y + ((size_t)b * lay.L + (lay.L - lay.T + t)) * 128;
        v[i][k] = row[o];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int o = lane + 32 * k;
    if (o < c) {
      const float sd = w[lay.u_std + o], mu = w[lay.u_mean + o];
      const int a = o / (1 + n), j = o % (1 + n);
      const int slot = (j == 0) ? a : (m + a * n + (j - 1));
#pragma unroll
      for (int i = 0; i < 8; ++i) tile[slot][warp + 4 * i] = v[i][k] * sd + mu;
    }
  }
  __syncthreads();
  if (LAYOUT == QX_LAYOUT_REC) {
    for (int e = threadIdx.x; e < (G >> 2) * 32; e += 128) {
      const int q = e >> 5, bl = e & 31, b = b0 + bl;
      if (b < B && (status == nullptr || status[b] == 0))
        *reinterpret_cast<float4*>(&gains[gain_index<LAYOUT>(t, 4 * q, b, B, G)]) =
            make_float4(tile[4 * q][bl], tile[4 * q + 1][bl], tile[4 * q + 2][bl], tile[4 * q + 3][bl]);
    }
  } else {
    for (int e = threadIdx.x; e < G * 32; e += 128) {
      const int slot = e >> 5, bl = e & 31, b = b0 + bl;
      if (b < B && (status == nullptr || status[b] == 0)) gains[gain_index<LAYOUT>(t, slot, b, B, G)] = tile[slot][bl];
    }
  }
}

// zero-padded copy of the output head: W_out [c][d] -> [128][d], bias [c] -> [128]
__global__ void tf_pad_head_kernel(const float* __restrict__ wo, const float* __restrict__ bo, float* __restrict__ wpad,
                                   float* __restrict__ bpad, int c, int d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 128 * d) wpad[i] = (i / d < c) ? wo[i] : 0.f;
  if (i < 128) bpad[i] = (i < c) ? bo[i] : 0.f;
}

static int gemm(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, bool relu,
                cudaStream_t st) {
  dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
  if (relu) tf_gemm_kernel<true><<<grid, 256, 0, st>>>(A, W, bias, C, M, N, K);
  else tf_gemm_kernel<false><<<grid, 256, 0, st>>>(A, W, bias, C, M, N, K);
  QX_CHECK_LAUNCH("tf_gemm_kernel");
  return QX_OK;
}

// tensor-core GEMM (qx_tf_tc.cu)
int tf_attention_tc(const float* qkv, float* out, int B, int L, int d, int nhead, cudaStream_t st);   // qx_tf_attn.cu
int tf_attention_umma(const float* qkv, float* out, int B, int L, int d, int nhead, int row_lo, cudaStream_t st,
                      int qk16 = 0); // qx_tf_attn05.cu
int tf_gemm_tc(const float* A, const float* Whi, const float* Wlo, const float* bias, float* C, int M, int N, int K,
               bool relu, cudaStream_t st, bool f16, int pack_qk = 0);
int tf_split_weights(const float* w, float* hi, float* lo, size_t count, cudaStream_t st, bool f16);
int tf_gemm_ln_tc(const float* A, const float* Whi, const float* Wlo, const float* bias, const float* res, const float* gamma,
                  const float* beta, float* C, int M, int K, cudaStream_t st, bool f16);
int tf_ffn_fused(const float* x, const float* W1hi, const float* W1lo, const float* b1, const float* W2hi, const float* W2lo,
                 const float* b2, const float* gamma, const float* beta, float* out, int M, int ff, cudaStream_t st);
int tf_tail_fused(const float* att, const float* Wohi, const float* Wolo, const float* bo, const float* g1, const float* be1,
                  float* x, const float* W1hi, const float* W1lo, const float* b1, const float* W2hi, const float* W2lo,
                  const float* b2, const float* g2, const float* be2, int M, int ff, cudaStream_t st);

// offsets of the (hi, lo) copies of the four GEMM weights of one layer inside the split buffer
struct SplitLayout {
  size_t in_hi, in_lo, out_hi, out_lo, w1_hi, w1_lo, w2_hi, w2_lo, stride;
  size_t head_pad, head_hi, head_lo, head_bias, total32;   // after the per-layer blocks
  size_t total;        // the fp16 halves of the weight at float offset X sit at float offset total32 + X / 2
  explicit SplitLayout(const TfLayout& l) {
    const size_t d = l.d, ff = l.ff;
    size_t o = 0;
    in_hi = o; o += 3 * d * d; in_lo = o; o += 3 * d * d;
    out_hi = o; o += d * d; out_lo = o; o += d * d;
    w1_hi = o; o += ff * d; w1_lo = o; o += ff * d;
    w2_hi = o; o += d * ff; w2_lo = o; o += d * ff;
    stride = o;
    size_t g = stride * l.layers;
    head_pad = g; g += 128 * d; head_hi = g; g += 128 * d; head_lo = g; g += 128 * d; head_bias = g; g += 128;
    total32 = g;
    total = total32 + (total32 + 1) / 2;
  }
};
// QX_TF_GEMM_F16=0: the dense layers of the forward pass run on the TF32 split (the standalone operators always do);
// so does QX_TF_GEMM_STREAM=1, the first-generation kernel has no fp16 variant.  Read per call, like the other switches.
// QX_TF_DENSE_F16=1 (experiments, tools/gemm_probe.py): the standalone operators use the fp16 split as well
static bool tf_ops_f16() {
  const char* e = getenv("QX_TF_DENSE_F16");
  return e != nullptr && e[0] == '1';
}
// The fp16 kernel needs K % 128 == 0 for every dense layer (K = d or ff); other shapes stay on the TF32 split, whose
// streaming kernel takes any K % 32 == 0.
static bool tf_use_f16(const TfLayout& lay) {
  const char* e = getenv("QX_TF_GEMM_F16");
  const char* s1 = getenv("QX_TF_GEMM_STREAM");
  return lay.d % 128 == 0 && lay.ff % 128 == 0 && !(e != nullptr && e[0] == '0') && !(s1 != nullptr && s1[0] == '1');
}

int tf_forward_simt(const QxTfConfig* cfg, const float* w, const float* wsplit, const double* x, const double* x_shift,
                    float* gains, int layout, const int32_t* status, float* ws, int B, cudaStream_t st) {
  const TfLayout lay(*cfg);
  const int d = lay.d, L = lay.L, ff = lay.ff, rows = B * L;
  float* h = ws;
  float* qkv = h + (size_t)rows * d;
  float* att = qkv + (size_t)rows * 3 * d;
  float* y = att + (size_t)rows * d;
  float* ffh = y + (size_t)rows * d;
  dim3 ge(L, B);
  const bool per_traj = (d == 128) && lay.n <= QX_MAX_N && lay.c <= 64 && cfg->horizon + 1 <= 128;
  if (per_traj) {
    const int embed_ctas = 5 * device_sm_count();          // 45 KB of shared memory per CTA: five fit an SM
    const int ge128 = B < embed_ctas ? B : embed_ctas;
    if (layout == QX_LAYOUT_REC) tf_embed128_kernel<QX_LAYOUT_REC><<<ge128, 128, 0, st>>>(*cfg, w, x, x_shift, gains, h, B);
    else tf_embed128_kernel<QX_LAYOUT_SOA><<<ge128, 128, 0, st>>>(*cfg, w, x, x_shift, gains, h, B);
  } else {
    if (layout == QX_LAYOUT_REC) tf_embed_kernel<QX_LAYOUT_REC><<<ge, d, 0, st>>>(*cfg, w, x, x_shift, gains, h, B);
    else tf_embed_kernel<QX_LAYOUT_SOA><<<ge, d, 0, st>>>(*cfg, w, x, x_shift, gains, h, B);
  }
  QX_CHECK_LAUNCH("tf_embed_kernel");
  const int hd = d / cfg->nhead;
  const size_t att_smem = (size_t)2 * L * hd * sizeof(float);
  static bool att_cfg[kMaxDevices] = {};
  if (!att_cfg[current_device_slot()]) {
    cudaFuncSetAttribute(tf_attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    att_cfg[current_device_slot()] = true;
  }
  const SplitLayout sp(lay);
  const bool f16 = tf_use_f16(lay);
  // the (hi, lo) halves the tensor-core kernels read for the weight whose TF32 halves sit at p: p itself, or its fp16 twin
  auto wh = [&](const float* p) -> const float* { return f16 ? wsplit + sp.total32 + (size_t)(p - wsplit) / 2 : p; };
  // dense contraction: tcgen05 three-term split when split weights are supplied, exact-fp32 CUDA cores otherwise
  auto dense = [&](const float* A, const float* lw, size_t w_off, size_t b_off, size_t hi_off, size_t lo_off, int l,
                   float* C, int N, int K, bool relu) -> int {
    if (wsplit != nullptr) {
      const float* sw = wsplit + (size_t)l * sp.stride;
      return tf_gemm_tc(A, wh(sw + hi_off), wh(sw + lo_off), lw + b_off, C, rows, N, K, relu, st, f16);
    }
    return gemm(A, lw + w_off, lw + b_off, C, rows, N, K, relu, st);
  };
  // Only the T query rows of the LAST layer reach the output head (transformer_model.py:138), and nothing downstream
  // reads the other rows' attention output / FFN of that layer: with the tensor-core kernels the last layer computes
  // Q K V for all rows, then attention, out-proj, both LayerNorms, the FFN and the head on a compact [B][T] row stream
  // (49 of 101 rows for the quadrotor).  Row-wise arithmetic is unchanged, so the predictions are bit-identical
  // (QX_TF_NO_PRUNE=1 switches it off for the equality test).
  // q, k, v leave the QKV projection as fp16 (hi, lo) pairs, MMA-ready for the tcgen05 attention (qx_tf_attn05.cu), when
  // the dense layers run on the fp16 split and that attention kernel takes the shape.  QX_TF_ATTN_QK16: 2 (default) q, k
  // and v packed, both attention products kind::f16; 1: q and k only (P V stays on the in-kernel TF32 split); 0: fp32
  // q, k, v and the TF32 splits (A/B runs, accuracy test).
  const char* env_qk = getenv("QX_TF_ATTN_QK16");
  const char* env_mma0 = getenv("QX_TF_ATTN_MMA");
  int qk16 = 0;
  if (wsplit != nullptr && f16 && d == 128 && cfg->nhead == 4 && L <= 112 && !(env_mma0 != nullptr && env_mma0[0] == '1'))
    qk16 = env_qk == nullptr ? 2 : (env_qk[0] == '0' ? 0 : (env_qk[0] == '1' ? 1 : 2));
  const char* env_np = getenv("QX_TF_NO_PRUNE");
  const bool prune = wsplit != nullptr && per_traj && !(env_np != nullptr && env_np[0] == '1') && lay.T < L;
  const int rows_c = B * lay.T;
  bool pruned = false;
  const float* hfinal = h;          // the row stream the output head reads
  for (int l = 0; l < lay.layers; ++l) {
    const float* lw = w + lay.layer0 + (size_t)l * lay.layer_stride;
    int rc;
    if (qk16) {
      const float* sw = wsplit + (size_t)l * sp.stride;
      rc = tf_gemm_tc(h, wh(sw + sp.in_hi), wh(sw + sp.in_lo), lw + lay.l_inb, qkv, rows, 3 * d, d, false, st, true, qk16);
    } else {
      rc = dense(h, lw, lay.l_inw, lay.l_inb, sp.in_hi, sp.in_lo, l, qkv, 3 * d, d, false);
    }
    if (rc) return rc;
    // tensor-core path: attention on tcgen05 (qx_tf_attn05.cu; QX_TF_ATTN_MMA=1 selects the mma.sync kernel of
    // qx_tf_attn.cu for A/B runs) when the shape fits, else SIMT
    rc = QX_EUNSUPPORTED;
    if (wsplit != nullptr) {
      const char* env_mma = getenv("QX_TF_ATTN_MMA");
      if (!(env_mma != nullptr && env_mma[0] == '1')) {
        if (prune && l == lay.layers - 1) {
          rc = tf_attention_umma(qkv, att, B, L, d, cfg->nhead, L - lay.T, st, qk16);      // att: compact [B][T][d]
          if (rc == QX_OK) pruned = true;
        }
        if (!pruned) rc = tf_attention_umma(qkv, att, B, L, d, cfg->nhead, 0, st, qk16);
      }
      if (qk16 && rc == QX_EUNSUPPORTED) {      // cannot happen (qk16 is only set for shapes the kernel takes); never fall back on packed Q / K
        set_error("tcgen05 attention refused a shape the fp16-pair Q / K path was selected for (L=%d d=%d)", L, d);
        return QX_EINVAL;
      }
      if (rc == QX_EUNSUPPORTED) rc = tf_attention_tc(qkv, att, B, L, d, cfg->nhead, st);
    }
    if (rc == QX_EUNSUPPORTED) {
      tf_attention_kernel<32><<<dim3(cfg->nhead, B), 128, att_smem, st>>>(qkv, att, L, d);
      QX_CHECK_LAUNCH("tf_attention_kernel");
    } else if (rc) {
      return rc;
    }
    // Residual + LayerNorm ride in the epilogue of the GEMM that produces the sub-layer output (d == 128: a CTA owns
    // whole rows).  QX_TF_FUSE_LN (bit 0: out-proj + norm1, bit 1: linear2 + norm2; default 3) = 0 restores the separate
    // add + LN launches for A/B runs and the equality test.
    const char* env_fl = getenv("QX_TF_FUSE_LN");
    const int fuse = (wsplit != nullptr && d == 128) ? (env_fl != nullptr ? atoi(env_fl) : 3) : 0;
    const float* sw = wsplit != nullptr ? wsplit + (size_t)l * sp.stride : nullptr;
    // sub-layer GEMM (N = d) + residual + LayerNorm: out = LN(res[gather] + A W^T + b); out may alias res, never A
    auto dense_ln = [&](bool fused, const float* A, size_t w_off, size_t b_off, size_t hi_off, size_t lo_off, int K, const float* res,
                        size_t g_off, size_t be_off, float* out, int m_rows, int per, int src_per, int src_off) -> int {
      int rc = QX_EUNSUPPORTED;
      if (fused && per == m_rows && src_off == 0)      // gathered residual rows (pruned norm1) are not a TMA box
        rc = tf_gemm_ln_tc(A, wh(sw + hi_off), wh(sw + lo_off), lw + b_off, res, lw + g_off, lw + be_off, out, m_rows, K, st, f16);
      if (rc != QX_EUNSUPPORTED) return rc;
      if (wsplit != nullptr) rc = tf_gemm_tc(A, wh(sw + hi_off), wh(sw + lo_off), lw + b_off, y, m_rows, d, K, false, st, f16);
      else rc = gemm(A, lw + w_off, lw + b_off, y, m_rows, d, K, false, st);
      if (rc) return rc;
      tf_add_ln_kernel<<<ceil_div(m_rows, 8), 256, 0, st>>>(out, res, y, lw + g_off, lw + be_off, m_rows, per, src_per, src_off);
      QX_CHECK_LAUNCH("tf_add_ln_kernel");
      return QX_OK;
    };
    // The whole feed-forward block in one launch (tf_ffn_kernel: the ff-wide intermediate never leaves the SM) when the
    // fp16 split and the fused norm2 are on; QX_TF_FFN_FUSE=0 brings the two launches back (A/B runs, equality test).
    const char* env_ff = getenv("QX_TF_FFN_FUSE");
    const bool fuse_ffn = f16 && (fuse & 2) && !(env_ff != nullptr && env_ff[0] == '0');
    auto ffn = [&](float* xs, int m_rows) -> int {
      if (!fuse_ffn) return QX_EUNSUPPORTED;
      return tf_ffn_fused(xs, wh(sw + sp.w1_hi), wh(sw + sp.w1_lo), lw + lay.l_b1, wh(sw + sp.w2_hi), wh(sw + sp.w2_lo), lw + lay.l_b2,
                          lw + lay.l_n2w, lw + lay.l_n2b, xs, m_rows, ff, st);
    };
    if (pruned) {
      // compact stream: att -> out-proj -> y; hc = LN(h[query rows] + y) lands in att (its input is consumed);
      // FFN (+ hc, norm2) -> hc
      float* hc = att;
      if ((rc = dense_ln(fuse & 1, att, lay.l_outw, lay.l_outb, sp.out_hi, sp.out_lo, d, h, lay.l_n1w, lay.l_n1b, hc, rows_c, lay.T, L,
                         L - lay.T)))
        return rc;
      if ((rc = ffn(hc, rows_c)) == QX_EUNSUPPORTED) {
        if ((rc = tf_gemm_tc(hc, wh(sw + sp.w1_hi), wh(sw + sp.w1_lo), lw + lay.l_b1, ffh, rows_c, ff, d, true, st, f16))) return rc;
        rc = dense_ln(fuse & 2, ffh, lay.l_w2, lay.l_b2, sp.w2_hi, sp.w2_lo, ff, hc, lay.l_n2w, lay.l_n2b, hc, rows_c, rows_c, rows_c, 0);
      }
      if (rc) return rc;
      hfinal = hc;
      break;
    }
    // out-proj + norm1 + the feed-forward block + norm2 in ONE launch (tf_tail_kernel) when ff <= 256, where it is faster
    // than out-proj + norm1 as a launch of their own (QX_TF_TAIL_FUSE=0: never, =2: whatever ff is; same bits)
    const char* env_tl = getenv("QX_TF_TAIL_FUSE");
    const bool tail = env_tl != nullptr ? (env_tl[0] == '2' || (env_tl[0] == '1' && ff <= 256)) : ff <= 256;
    if (fuse_ffn && (fuse & 1) && tail) {
      rc = tf_tail_fused(att, wh(sw + sp.out_hi), wh(sw + sp.out_lo), lw + lay.l_outb, lw + lay.l_n1w, lw + lay.l_n1b, h,
                         wh(sw + sp.w1_hi), wh(sw + sp.w1_lo), lw + lay.l_b1, wh(sw + sp.w2_hi), wh(sw + sp.w2_lo), lw + lay.l_b2,
                         lw + lay.l_n2w, lw + lay.l_n2b, rows, ff, st);
      if (rc == QX_OK) continue;
      if (rc != QX_EUNSUPPORTED) return rc;
    }
    if ((rc = dense_ln(fuse & 1, att, lay.l_outw, lay.l_outb, sp.out_hi, sp.out_lo, d, h, lay.l_n1w, lay.l_n1b, h, rows, rows, rows, 0)))
      return rc;
    if ((rc = ffn(h, rows)) == QX_EUNSUPPORTED) {
      if ((rc = dense(h, lw, lay.l_w1, lay.l_b1, sp.w1_hi, sp.w1_lo, l, ffh, ff, d, true))) return rc;
      rc = dense_ln(fuse & 2, ffh, lay.l_w2, lay.l_b2, sp.w2_hi, sp.w2_lo, ff, h, lay.l_n2w, lay.l_n2b, h, rows, rows, rows, 0);
    }
    if (rc) return rc;
  }
  dim3 go(lay.T, B);
  if (wsplit != nullptr) {      // output head as one more tensor-core GEMM (N padded to 128) over the rows that are left
    int rc = tf_gemm_tc(hfinal, wh(wsplit + sp.head_hi), wh(wsplit + sp.head_lo), wsplit + sp.head_bias, y, pruned ? rows_c : rows, 128, d,
                        false, st, f16);
    if (rc) return rc;
    if (per_traj) {
      // QX_TF_UNPACK_TILE=0: the per-trajectory kernel (A/B runs; same values)
      const char* env_ut = getenv("QX_TF_UNPACK_TILE");
      const dim3 gt(ceil_div(B, 32), lay.T);
      if (env_ut != nullptr && env_ut[0] == '0') {
        if (layout == QX_LAYOUT_REC) tf_unpack128_kernel<QX_LAYOUT_REC><<<B, 128, 0, st>>>(*cfg, w, y, gains, status, B, pruned ? 1 : 0);
        else tf_unpack128_kernel<QX_LAYOUT_SOA><<<B, 128, 0, st>>>(*cfg, w, y, gains, status, B, pruned ? 1 : 0);
      } else if (layout == QX_LAYOUT_REC) {
        tf_unpack_tile_kernel<QX_LAYOUT_REC><<<gt, 128, 0, st>>>(*cfg, w, y, gains, status, B, pruned ? 1 : 0);
      } else {
        tf_unpack_tile_kernel<QX_LAYOUT_SOA><<<gt, 128, 0, st>>>(*cfg, w, y, gains, status, B, pruned ? 1 : 0);
      }
    } else {
      if (layout == QX_LAYOUT_REC) tf_unpack_kernel<QX_LAYOUT_REC><<<go, 64, 0, st>>>(*cfg, w, y, gains, status, B);
      else tf_unpack_kernel<QX_LAYOUT_SOA><<<go, 64, 0, st>>>(*cfg, w, y, gains, status, B);
    }
    QX_CHECK_LAUNCH("tf_unpack_kernel");
    return QX_OK;
  }
  if (layout == QX_LAYOUT_REC) tf_output_kernel<QX_LAYOUT_REC><<<go, 64, 0, st>>>(*cfg, w, h, gains, status, B);
  else tf_output_kernel<QX_LAYOUT_SOA><<<go, 64, 0, st>>>(*cfg, w, h, gains, status, B);
  QX_CHECK_LAUNCH("tf_output_kernel");
  return QX_OK;
}

}  // namespace qx

extern "C" size_t qx_tf_weight_floats(const QxTfConfig* cfg) {
  if (!cfg) return 0;
  return qx::TfLayout(*cfg).total;
}

extern "C" size_t qx_tf_workspace_bytes(const QxTfConfig* cfg, int B) {
  if (!cfg || B <= 0) return 0;
  const qx::TfLayout lay(*cfg);
  return (size_t)B * lay.L * (6 * (size_t)lay.d + lay.ff) * sizeof(float);
}

extern "C" size_t qx_tf_split_floats(const QxTfConfig* cfg) {
  if (!cfg) return 0;
  const qx::TfLayout lay(*cfg);
  return qx::SplitLayout(lay).total;
}

extern "C" int qx_tf_prepare(const QxTfConfig* cfg, const float* weights, float* wsplit, void* stream) {
  QX_CHECK_ARG(cfg && weights && wsplit, "null pointer");
  const qx::TfLayout lay(*cfg);
  const qx::SplitLayout sp(lay);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t d = lay.d, ff = lay.ff;
  // every weight in both formats: TF32 halves at (hi, lo), fp16 halves at total32 + offset / 2
  auto both = [&](const float* w, float* hi, float* lo, size_t count) -> int {
    const int rc = qx::tf_split_weights(w, hi, lo, count, st, false);
    if (rc) return rc;
    return qx::tf_split_weights(w, wsplit + sp.total32 + (size_t)(hi - wsplit) / 2, wsplit + sp.total32 + (size_t)(lo - wsplit) / 2,
                                count, st, true);
  };
  for (int l = 0; l < lay.layers; ++l) {
    const float* lw = weights + lay.layer0 + (size_t)l * lay.layer_stride;
    float* sw = wsplit + (size_t)l * sp.stride;
    int rc;
    if ((rc = both(lw + lay.l_inw, sw + sp.in_hi, sw + sp.in_lo, 3 * d * d))) return rc;
    if ((rc = both(lw + lay.l_outw, sw + sp.out_hi, sw + sp.out_lo, d * d))) return rc;
    if ((rc = both(lw + lay.l_w1, sw + sp.w1_hi, sw + sp.w1_lo, ff * d))) return rc;
    if ((rc = both(lw + lay.l_w2, sw + sp.w2_hi, sw + sp.w2_lo, d * ff))) return rc;
  }
  qx::tf_pad_head_kernel<<<(unsigned)((128 * d + 255) / 256), 256, 0, st>>>(weights + lay.wo, weights + lay.bo,
                                                                          wsplit + sp.head_pad, wsplit + sp.head_bias,
                                                                          lay.c, (int)d);
  QX_CHECK_LAUNCH("tf_pad_head_kernel");
  return both(wsplit + sp.head_pad, wsplit + sp.head_hi, wsplit + sp.head_lo, 128 * d);
}

extern "C" int qx_tf_forward(const QxTfConfig* cfg, const float* weights, const float* wsplit, const double* x,
                             const double* x_shift, float* gains, int model_id, const int32_t* status, void* workspace,
                             int B, void* stream) {
  QX_CHECK_ARG(cfg && weights && x && x_shift && gains && workspace, "null pointer");
  QX_CHECK_ARG(B > 0, "B must be positive");
  QX_CHECK_ARG(cfg->d_model == 128 && cfg->d_model / cfg->nhead == 32, "kernels are built for d_model 128, head_dim 32");
  QX_CHECK_ARG(cfg->control_dim <= 64 && cfg->state_dim <= 64, "state/control dims must be <= 64");
  QX_CHECK_ARG(cfg->dim_ff % 128 == 0, "dim_feedforward must be a multiple of 128");
  QX_CHECK_ARG(cfg->control_dim == cfg->m * (1 + cfg->state_dim), "control_dim must equal m (1 + n)");
  QX_CHECK_ARG(cfg->horizon == cfg->target_len + cfg->prompt_len, "horizon must equal target_len + prompt_len");
  const int layout = qx_layout(model_id);
  if (layout < 0) return layout;
  return qx::tf_forward_simt(cfg, weights, wsplit, x, x_shift, gains, layout, status, static_cast<float*>(workspace), B,
                             static_cast<cudaStream_t>(stream));
}

// nn.Linear (+ ReLU) as used by the model (transformer_model.py:106-107,116 and the encoder layers) on the
// tensor cores: C[M][N] = act(A[M][K] W[N][K]^T + bias) with W given as its (hi, lo) TF32 halves.
extern "C" int qx_tf_split(const float* w, float* hi, float* lo, size_t count, void* stream) {
  QX_CHECK_ARG(w && hi && lo, "null pointer");
  if (count == 0) return QX_OK;
  return qx::tf_split_weights(w, hi, lo, count, static_cast<cudaStream_t>(stream), qx::tf_ops_f16());
}

extern "C" int qx_tf_dense(const float* A, const float* w_hi, const float* w_lo, const float* bias, float* C, int M,
                           int N, int K, int relu, void* stream) {
  QX_CHECK_ARG(A && w_hi && w_lo && bias && C, "null pointer");
  QX_CHECK_ARG(M >= 0 && N > 0 && K > 0, "bad shape");
  if (M == 0) return QX_OK;
  return qx::tf_gemm_tc(A, w_hi, w_lo, bias, C, M, N, K, relu != 0, static_cast<cudaStream_t>(stream), qx::tf_ops_f16());
}

extern "C" int qx_tf_dense_ln(const float* A, const float* w_hi, const float* w_lo, const float* bias, const float* res,
                              const float* gamma, const float* beta, float* C, int M, int K, void* stream) {
  QX_CHECK_ARG(A && w_hi && w_lo && bias && res && gamma && beta && C, "null pointer");
  QX_CHECK_ARG(M >= 0 && K > 0, "bad shape");
  QX_CHECK_ARG(C != A, "the output may alias the residual, not the A operand");
  if (M == 0) return QX_OK;
  const int rc = qx::tf_gemm_ln_tc(A, w_hi, w_lo, bias, res, gamma, beta, C, M, K, static_cast<cudaStream_t>(stream), qx::tf_ops_f16());
  if (rc == QX_EUNSUPPORTED) qx::set_error("qx_tf_dense_ln needs K %% 128 == 0 (got K=%d)", K);
  return rc;
}

// Causal self-attention core of nn.MultiheadAttention as the model uses it (transformer_model.py:112-113, :134-136):
// qkv [B*L][3d] packed in_proj output -> out [B*L][d] (heads concatenated), head dimension 32.
// impl 0: tcgen05 kernel, 1: mma.sync kernel, 2: exact-fp32 CUDA-core kernel.
extern "C" int qx_tf_attention(const float* qkv, float* out, int B, int L, int d, int nhead, int impl, void* stream) {
  QX_CHECK_ARG(qkv && out, "null pointer");
  QX_CHECK_ARG(B >= 0 && L > 0 && d > 0 && nhead > 0 && d == 32 * nhead, "bad shape (head dimension must be 32)");
  if (B == 0) return QX_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (impl == 0) return qx::tf_attention_umma(qkv, out, B, L, d, nhead, 0, st);
  if (impl == 1) return qx::tf_attention_tc(qkv, out, B, L, d, nhead, st);
  QX_CHECK_ARG(impl == 2, "impl must be 0, 1 or 2");
  static bool att_cfg[qx::kMaxDevices] = {};
  if (!att_cfg[qx::current_device_slot()]) {
    cudaFuncSetAttribute(qx::tf_attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    att_cfg[qx::current_device_slot()] = true;
  }
  qx::tf_attention_kernel<32><<<dim3(nhead, B), 128, (size_t)2 * L * 32 * sizeof(float), st>>>(qkv, out, L, d);
  QX_CHECK_LAUNCH("tf_attention_kernel");
  return QX_OK;
}
