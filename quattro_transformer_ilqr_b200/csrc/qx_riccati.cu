// Backward Riccati pass.
//
//   riccati_*_kernel = iLQR_TF.backward_pass          (quattro_ilqr_tf.py:282-319)
//                    / backward_pass_segment (t_start > 0, :327-366)
// per step t = N-1 .. t_start:
//   Q_x = l_x + A^T V_x          Q_u = l_u + B^T V_x
//   Q_xx = l_xx + A^T V_xx A     Q_ux = l_ux + B^T V_xx A     Q_uu = l_uu + B^T V_xx B
//   k = -inv(Q_uu + reg I) Q_u   K = -inv(Q_uu + reg I) Q_ux          (:304-309)
//   V_x  = Q_x  + K^T Q_uu k + K^T Q_u  + Q_ux^T k   (un-regularised Q_uu, :315)
//   V_xx = Q_xx + K^T Q_uu K + K^T Q_ux + Q_ux^T K ; V_xx = (V_xx + V_xx^T)/2   (:316-317)
//
// Two kernels:
//  * riccati_warp_kernel<12,4> (quadrotor): one warp per trajectory.  The step's 1 664-B
//    expansion record streams HBM -> shared memory through a 3-stage ring of 1-D bulk async
//    copies (TMA; SASS UBLKCP) completed on mbarriers, so the copy of steps t-1..t-3 overlaps the
//    arithmetic of step t.  With F = [A | B] (12x16) the step is W = V F, Q = L + F^T W; each
//    lane owns a 2x4 register tile of W and of Q (one 8-B and one 16-B shared load per 8 FMAs,
//    both warp-broadcast -> one wavefront each), the 4x4 solve is done redundantly per lane with
//    partial pivoting, V_xx / V_x stay in shared memory between steps.
//  * riccati_soa_kernel<4,1> (cart-pole): 16+4+... values fit one thread's registers, so one
//    thread per trajectory over the b-fastest SOA layout (every load coalesced).
#include "qx_async.cuh"
#include "qx_common.cuh"

namespace qx {

// Solve M z = r in place (z returned in r) by Gaussian elimination with partial pivoting.
template <int m>
__device__ __forceinline__ void solve_pp(float (&M)[m][m], float (&r)[m]) {
#pragma unroll
  for (int c = 0; c < m; ++c) {
#pragma unroll
    for (int rr = c + 1; rr < m; ++rr) {
      const bool sw = fabsf(M[rr][c]) > fabsf(M[c][c]);
#pragma unroll
      for (int j = 0; j < m; ++j) {
        const float a = M[c][j], b = M[rr][j];
        M[c][j] = sw ? b : a;
        M[rr][j] = sw ? a : b;
      }
      const float a = r[c], b = r[rr];
      r[c] = sw ? b : a;
      r[rr] = sw ? a : b;
    }
    const float inv = 1.0f / M[c][c];
#pragma unroll
    for (int rr = c + 1; rr < m; ++rr) {
      const float f = M[rr][c] * inv;
#pragma unroll
      for (int j = c + 1; j < m; ++j) M[rr][j] -= f * M[c][j];
      r[rr] -= f * r[c];
    }
  }
#pragma unroll
  for (int c = m - 1; c >= 0; --c) {
    float acc = r[c];
#pragma unroll
    for (int j = c + 1; j < m; ++j) acc -= M[c][j] * r[j];
    r[c] = acc / M[c][c];
  }
}

// ------------------------------------------------------------------------------------------
// Thread-per-trajectory, SOA layout (small systems).
template <int n, int m>
__global__ void __launch_bounds__(128)
riccati_soa_kernel(const float* __restrict__ expansion, const float* __restrict__ terminal,
                   float* __restrict__ gains, const int32_t* __restrict__ status, int B, int N, int t_start,
                   float reg) {
  using R = Rec<n, m>;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (status != nullptr && status[b] != 0) return;
  float Vx[n], V[n][n];
#pragma unroll
  for (int i = 0; i < n; ++i) Vx[i] = terminal[(size_t)i * B + b];
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int j = 0; j < n; ++j) V[i][j] = terminal[(size_t)(n + i * n + j) * B + b];

  for (int t = N - 1; t >= t_start; --t) {
    const float* e = expansion + (size_t)t * R::size * B + b;
    auto ld = [&](int idx) { return e[(size_t)idx * B]; };
    float A[n][n], Bm[n][m];
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int j = 0; j < n; ++j) A[i][j] = ld(R::A + i * n + j);
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int j = 0; j < m; ++j) Bm[i][j] = ld(R::Bm + i * m + j);
    // W = V [A | B]
    float WA[n][n], WB[n][m];
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int j = 0; j < n; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int l = 0; l < n; ++l) acc += V[i][l] * A[l][j];
        WA[i][j] = acc;
      }
#pragma unroll
      for (int j = 0; j < m; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int l = 0; l < n; ++l) acc += V[i][l] * Bm[l][j];
        WB[i][j] = acc;
      }
    }
    float Qx[n], Qu[m], Qxx[n][n], Qux[m][n], Quu[m][m];
#pragma unroll
    for (int i = 0; i < n; ++i) {
      float acc = ld(R::lx + i);
#pragma unroll
      for (int l = 0; l < n; ++l) acc += A[l][i] * Vx[l];
      Qx[i] = acc;
    }
#pragma unroll
    for (int j = 0; j < m; ++j) {
      float acc = ld(R::lu + j);
#pragma unroll
      for (int l = 0; l < n; ++l) acc += Bm[l][j] * Vx[l];
      Qu[j] = acc;
    }
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int j = 0; j < n; ++j) {
        float acc = ld(R::lxx + i * n + j);
#pragma unroll
        for (int l = 0; l < n; ++l) acc += A[l][i] * WA[l][j];
        Qxx[i][j] = acc;
      }
#pragma unroll
    for (int a = 0; a < m; ++a) {
#pragma unroll
      for (int j = 0; j < n; ++j) {
        float acc = ld(R::lux + a * n + j);
#pragma unroll
        for (int l = 0; l < n; ++l) acc += Bm[l][a] * WA[l][j];
        Qux[a][j] = acc;
      }
#pragma unroll
      for (int c = 0; c < m; ++c) {
        float acc = ld(R::luu + a * m + c);
#pragma unroll
        for (int l = 0; l < n; ++l) acc += Bm[l][a] * WB[l][c];
        Quu[a][c] = acc;
      }
    }
    // gains: columns of K and k solve (Q_uu + reg I) z = -rhs
    float K[m][n], k[m];
    {
      float rhs[m];
#pragma unroll
      for (int j = 0; j <= n; ++j) {
        float Mr[m][m];
#pragma unroll
        for (int a = 0; a < m; ++a) {
#pragma unroll
          for (int c = 0; c < m; ++c) Mr[a][c] = Quu[a][c] + (a == c ? reg : 0.f);
          rhs[a] = (j < n) ? Qux[a][j < n ? j : 0] : Qu[a];
        }
        solve_pp<m>(Mr, rhs);
#pragma unroll
        for (int a = 0; a < m; ++a) {
          if (j < n) K[a][j < n ? j : 0] = -rhs[a];
          else k[a] = -rhs[a];
        }
      }
    }
    float* g = gains + (size_t)t * R::gain * B + b;
#pragma unroll
    for (int a = 0; a < m; ++a) g[(size_t)a * B] = k[a];
#pragma unroll
    for (int a = 0; a < m; ++a)
#pragma unroll
      for (int j = 0; j < n; ++j) g[(size_t)(m + a * n + j) * B] = K[a][j];
    // value update
    float Mk[m][n], h[m];
#pragma unroll
    for (int a = 0; a < m; ++a) {
#pragma unroll
      for (int j = 0; j < n; ++j) {
        float acc = Qux[a][j];
#pragma unroll
        for (int c = 0; c < m; ++c) acc += Quu[a][c] * K[c][j];
        Mk[a][j] = acc;                       // Q_uu K + Q_ux
      }
      float acc = Qu[a];
#pragma unroll
      for (int c = 0; c < m; ++c) acc += Quu[a][c] * k[c];
      h[a] = acc;                             // Q_uu k + Q_u
    }
    float Vn[n][n];
#pragma unroll
    for (int i = 0; i < n; ++i) {
      float acc = Qx[i];
#pragma unroll
      for (int a = 0; a < m; ++a) acc += K[a][i] * h[a] + Qux[a][i] * k[a];
      Vx[i] = acc;
#pragma unroll
      for (int j = 0; j < n; ++j) {
        float v = Qxx[i][j];
#pragma unroll
        for (int a = 0; a < m; ++a) v += K[a][i] * Mk[a][j] + Qux[a][i] * K[a][j];
        Vn[i][j] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int j = 0; j < n; ++j) V[i][j] = 0.5f * (Vn[i][j] + Vn[j][i]);
  }
}

// ------------------------------------------------------------------------------------------
// Warp-per-trajectory, REC layout, n = 12, m = 4.
constexpr int kRicStages = 3;
constexpr int kRicWarps = 4;

struct RicWarpSmem {
  alignas(128) float rec[kRicStages][416];   // expansion records (bulk-copy destinations)
  alignas(16) float V[12 * 12];              // V_xx (symmetric)
  alignas(16) float W[12 * 16];              // V [A|B]; reused as the un-symmetrised V_xx
  alignas(16) float Qb[4 * 16];              // rows 12..15 of Q: [Q_ux | Q_uu]
  alignas(16) float Ks[4 * 12];              // K
  alignas(16) float Ms[4 * 12];              // Q_uu K + Q_ux
  alignas(16) float Vx[12];
  alignas(16) float q[16];                   // [Q_x | Q_u]
  alignas(16) float kh[8];                   // [k | Q_uu k + Q_u]
  alignas(8) uint64_t bar[kRicStages];
};

__global__ void __launch_bounds__(32 * kRicWarps)
riccati_warp_kernel(const float* __restrict__ expansion, const float* __restrict__ terminal,
                    float* __restrict__ gains, const int32_t* __restrict__ status, int B, int N, int t_start,
                    float reg) {
  constexpr int n = 12, m = 4;
  using R = Rec<n, m>;
  static_assert(R::size == 416 && R::gain == 52, "layout");
  __shared__ RicWarpSmem smem[kRicWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kRicWarps + warp;
  if (b >= B) return;
  if (status != nullptr && status[b] != 0) return;
  RicWarpSmem& s = smem[warp];
  const int steps = N - t_start;

  if (lane == 0) {
    for (int i = 0; i < kRicStages; ++i) mbar_init(&s.bar[i], 1);
    mbar_fence_init();
  }
  __syncwarp();
  auto issue = [&](int k) {   // lane 0: sequence k <-> t = N-1-k
    const int t = N - 1 - k;
    mbar_arrive_expect_tx(&s.bar[k % kRicStages], R::size * sizeof(float));
    bulk_g2s(s.rec[k % kRicStages], expansion + ((size_t)t * B + b) * R::size, R::size * sizeof(float),
             &s.bar[k % kRicStages]);
  };
  if (lane == 0)
    for (int k = 0; k < kRicStages && k < steps; ++k) issue(k);

  // terminal condition
  {
    const float* tm = terminal + (size_t)b * R::term;
    for (int e = lane; e < R::term; e += 32) {
      const float v = tm[e];
      if (e < n) s.Vx[e] = v; else s.V[e - n] = v;
    }
  }
  __syncwarp();

  const int ri = lane >> 2, cj = lane & 3;   // tile coordinates: rows 2ri, 2ri+1; cols 4cj..4cj+3

  for (int k = 0; k < steps; ++k) {
    const int t = N - 1 - k;
    mbar_wait(&s.bar[k % kRicStages], (k / kRicStages) & 1u);
    const float* rec = s.rec[k % kRicStages];
    const float* A = rec + R::A;
    const float* Bm = rec + R::Bm;

    // ---- phase 1: W = V F (lanes 0..23), q = l + F^T V_x (lanes 24..31) -------------------
    if (lane < 24) {
      float acc[2][4] = {};
      const float* bsrc = (cj < 3) ? (A + 4 * cj) : Bm;
      const int bstride = (cj < 3) ? n : m;
#pragma unroll
      for (int l = 0; l < n; ++l) {
        const float2 a = *reinterpret_cast<const float2*>(&s.V[l * n + 2 * ri]);
        const float4 bv = *reinterpret_cast<const float4*>(bsrc + l * bstride);
        acc[0][0] += a.x * bv.x; acc[0][1] += a.x * bv.y; acc[0][2] += a.x * bv.z; acc[0][3] += a.x * bv.w;
        acc[1][0] += a.y * bv.x; acc[1][1] += a.y * bv.y; acc[1][2] += a.y * bv.z; acc[1][3] += a.y * bv.w;
      }
      *reinterpret_cast<float4*>(&s.W[(2 * ri) * 16 + 4 * cj]) = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
      *reinterpret_cast<float4*>(&s.W[(2 * ri + 1) * 16 + 4 * cj]) = make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]);
    } else {
      const int j0 = 2 * (lane - 24);        // columns j0, j0+1 of F
      const float* fsrc = (j0 < n) ? (A + j0) : (Bm + (j0 - n));
      const int fstride = (j0 < n) ? n : m;
      float q0 = (j0 < n) ? rec[R::lx + j0] : rec[R::lu + j0 - n];
      float q1 = (j0 < n) ? rec[R::lx + j0 + 1] : rec[R::lu + j0 + 1 - n];
#pragma unroll
      for (int l = 0; l < n; ++l) {
        const float2 f = *reinterpret_cast<const float2*>(fsrc + l * fstride);
        const float vx = s.Vx[l];
        q0 += f.x * vx;
        q1 += f.y * vx;
      }
      *reinterpret_cast<float2*>(&s.q[j0]) = make_float2(q0, q1);
    }
    __syncwarp();

    // ---- phase 2: Q = L + F^T W, 2x4 tile per lane over the 16x16 matrix -------------------
    float Q[2][4];
    {
      const int i0 = 2 * ri, j0 = 4 * cj;
      // initialise with the cost Hessian block
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = i0 + r;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = j0 + c;
          float v;
          if (i < n && j < n) v = rec[R::lxx + i * n + j];
          else if (i < n) v = rec[R::lux + (j - n) * n + i];        // Q_xu = Q_ux^T
          else if (j < n) v = rec[R::lux + (i - n) * n + j];
          else v = rec[R::luu + (i - n) * m + (j - n)];
          Q[r][c] = v;
        }
      }
      const float* asrc = (ri < 6) ? (A + i0) : (Bm + (i0 - n));
      const int astride = (ri < 6) ? n : m;
#pragma unroll
      for (int l = 0; l < n; ++l) {
        const float2 a = *reinterpret_cast<const float2*>(asrc + l * astride);
        const float4 w = *reinterpret_cast<const float4*>(&s.W[l * 16 + j0]);
        Q[0][0] += a.x * w.x; Q[0][1] += a.x * w.y; Q[0][2] += a.x * w.z; Q[0][3] += a.x * w.w;
        Q[1][0] += a.y * w.x; Q[1][1] += a.y * w.y; Q[1][2] += a.y * w.z; Q[1][3] += a.y * w.w;
      }
      if (ri >= 6) {
        *reinterpret_cast<float4*>(&s.Qb[(i0 - n) * 16 + j0]) = make_float4(Q[0][0], Q[0][1], Q[0][2], Q[0][3]);
        *reinterpret_cast<float4*>(&s.Qb[(i0 + 1 - n) * 16 + j0]) = make_float4(Q[1][0], Q[1][1], Q[1][2], Q[1][3]);
      }
    }
    __syncwarp();
    // the record is consumed: refill its stage with the record of step k + stages
    if (lane == 0 && k + kRicStages < steps) issue(k + kRicStages);

    // ---- phase 3: gains.  lane j < 12 solves column j of K, lane 12 solves k -----------------
    {
      float Quu[m][m], Mr[m][m], rhs[m];
#pragma unroll
      for (int a = 0; a < m; ++a) {
        const float4 v = *reinterpret_cast<const float4*>(&s.Qb[a * 16 + n]);
        Quu[a][0] = v.x; Quu[a][1] = v.y; Quu[a][2] = v.z; Quu[a][3] = v.w;
      }
      const int col = (lane < n) ? lane : 0;
#pragma unroll
      for (int a = 0; a < m; ++a) {
#pragma unroll
        for (int c = 0; c < m; ++c) Mr[a][c] = Quu[a][c] + (a == c ? reg : 0.f);
        rhs[a] = (lane == n) ? s.q[n + a] : s.Qb[a * 16 + col];
      }
      float src[m];
#pragma unroll
      for (int a = 0; a < m; ++a) src[a] = rhs[a];
      solve_pp<m>(Mr, rhs);
      float z[m], hm[m];
#pragma unroll
      for (int a = 0; a < m; ++a) z[a] = -rhs[a];
#pragma unroll
      for (int a = 0; a < m; ++a) {
        float acc = src[a];
#pragma unroll
        for (int c = 0; c < m; ++c) acc += Quu[a][c] * z[c];
        hm[a] = acc;                         // Q_uu K + Q_ux (column), or Q_uu k + Q_u
      }
      float* g = gains + ((size_t)t * B + b) * R::gain;
      if (lane < n) {
#pragma unroll
        for (int a = 0; a < m; ++a) {
          s.Ks[a * n + lane] = z[a];
          s.Ms[a * n + lane] = hm[a];
          g[m + a * n + lane] = z[a];
        }
      } else if (lane == n) {
#pragma unroll
        for (int a = 0; a < m; ++a) {
          s.kh[a] = z[a];
          s.kh[m + a] = hm[a];
          g[a] = z[a];
        }
      }
    }
    __syncwarp();

    // ---- phase 4: V_xx' = Q_xx + K^T (Q_uu K + Q_ux) + Q_ux^T K ;  V_x' ------------------------
    if (ri < 6 && cj < 3) {
      const int i0 = 2 * ri, j0 = 4 * cj;
#pragma unroll
      for (int a = 0; a < m; ++a) {
        const float2 ka = *reinterpret_cast<const float2*>(&s.Ks[a * n + i0]);
        const float4 mb = *reinterpret_cast<const float4*>(&s.Ms[a * n + j0]);
        const float2 qa = *reinterpret_cast<const float2*>(&s.Qb[a * 16 + i0]);
        const float4 kb = *reinterpret_cast<const float4*>(&s.Ks[a * n + j0]);
        Q[0][0] += ka.x * mb.x + qa.x * kb.x; Q[0][1] += ka.x * mb.y + qa.x * kb.y;
        Q[0][2] += ka.x * mb.z + qa.x * kb.z; Q[0][3] += ka.x * mb.w + qa.x * kb.w;
        Q[1][0] += ka.y * mb.x + qa.y * kb.x; Q[1][1] += ka.y * mb.y + qa.y * kb.y;
        Q[1][2] += ka.y * mb.z + qa.y * kb.z; Q[1][3] += ka.y * mb.w + qa.y * kb.w;
      }
      // un-symmetrised V_xx' goes to the W buffer (pitch 12)
      *reinterpret_cast<float4*>(&s.W[i0 * n + j0]) = make_float4(Q[0][0], Q[0][1], Q[0][2], Q[0][3]);
      *reinterpret_cast<float4*>(&s.W[(i0 + 1) * n + j0]) = make_float4(Q[1][0], Q[1][1], Q[1][2], Q[1][3]);
    }
    float vx_new = 0.f;
    if (lane < n) {
      vx_new = s.q[lane];
#pragma unroll
      for (int a = 0; a < m; ++a) vx_new += s.Ks[a * n + lane] * s.kh[m + a] + s.Qb[a * 16 + lane] * s.kh[a];
    }
    __syncwarp();
    if (lane < n) s.Vx[lane] = vx_new;

    // ---- phase 5: symmetrise ---------------------------------------------------------------------
    if (ri < 6 && cj < 3) {
      const int i0 = 2 * ri, j0 = 4 * cj;
      float o[2][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float2 tr = *reinterpret_cast<const float2*>(&s.W[(j0 + c) * n + i0]);   // V'[j][i0..i0+1]
        o[0][c] = 0.5f * (Q[0][c] + tr.x);
        o[1][c] = 0.5f * (Q[1][c] + tr.y);
      }
      *reinterpret_cast<float4*>(&s.V[i0 * n + j0]) = make_float4(o[0][0], o[0][1], o[0][2], o[0][3]);
      *reinterpret_cast<float4*>(&s.V[(i0 + 1) * n + j0]) = make_float4(o[1][0], o[1][1], o[1][2], o[1][3]);
    }
    __syncwarp();
  }
}

}  // namespace qx

extern "C" int qx_riccati(int model_id, const float* expansion, const float* terminal, float* gains,
                          const int32_t* status, int B, int N, int t_start, float reg, void* stream) {
  QX_CHECK_ARG(expansion && terminal && gains, "null pointer");
  QX_CHECK_ARG(B > 0 && N > 0 && t_start >= 0 && t_start < N, "need B > 0 and 0 <= t_start < N");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (model_id) {
    case QX_MODEL_CARTPOLE:
      qx::riccati_soa_kernel<4, 1><<<qx::ceil_div(B, 128), 128, 0, st>>>(expansion, terminal, gains, status, B, N,
                                                                        t_start, reg);
      QX_CHECK_LAUNCH("riccati_soa_kernel");
      return QX_OK;
    case QX_MODEL_QUADROTOR:
      QX_CHECK_ARG((reinterpret_cast<uintptr_t>(expansion) & 15u) == 0, "expansion must be 16-byte aligned");
      qx::riccati_warp_kernel<<<qx::ceil_div(B, qx::kRicWarps), 32 * qx::kRicWarps, 0, st>>>(
          expansion, terminal, gains, status, B, N, t_start, reg);
      QX_CHECK_LAUNCH("riccati_warp_kernel");
      return QX_OK;
  }
  qx::set_error("unknown model_id %d", model_id);
  return QX_EINVAL;
}
