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
//  * riccati_warp_kernel (quadrotor, n = 12, m = 4): two trajectories per warp, 16 lanes each.
//    The step's 1 664-B expansion record streams HBM -> shared memory through a 2-stage ring of
//    1-D bulk async copies (TMA; SASS UBLKCP) completed on mbarriers, so the copies of steps
//    t-1, t-2 overlap the arithmetic of step t.  With F = [A | B] (12x16, stored interleaved) the
//    step is W = V F, Q = L + F^T W on register tiles (two 16-B shared loads, both broadcast
//    inside the half-warp, per 12-16 FMAs); V_xx / V_x stay in shared memory between steps.
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

  // The record of step t-1 is fetched while step t is computed: with a few warps per SM (one thread per trajectory)
  // nothing else hides the HBM latency of the 46 loads, and a step costs load latency + arithmetic otherwise.
  float nxt[R::size];
  auto fetch = [&](int t) {
    const float* e = expansion + (size_t)t * R::size * B + b;
#pragma unroll
    for (int idx = 0; idx < R::size; ++idx) nxt[idx] = e[(size_t)idx * B];
  };
  if (N - 1 >= t_start) fetch(N - 1);
  for (int t = N - 1; t >= t_start; --t) {
    float rec[R::size];
#pragma unroll
    for (int idx = 0; idx < R::size; ++idx) rec[idx] = nxt[idx];
    if (t > t_start) fetch(t - 1);
    auto ld = [&](int idx) { return rec[idx]; };
    float A[n][n], Bm[n][m];
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int j = 0; j < n; ++j) A[i][j] = ld(R::a(i, j));
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int j = 0; j < m; ++j) Bm[i][j] = ld(R::b(i, j));
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
// Half-warp-per-trajectory (two trajectories per warp), REC layout, n = 12, m = 4.
//
// Lane roles inside a half-warp (hl = lane & 15, ri = hl >> 2, cj = hl & 3):
//   phase 1  W = V F (12x16): 3x4 tile (rows 3ri.., cols 4cj..); V is kept in a padded
//            "p4" layout  Vp[l][4*(i/3) + i%3]  so the three V values are one 16-B load; the spare
//            slot of row group 3 carries V_x, which yields q = [l_x|l_u] + F^T V_x for free
//   phase 2  Q = L + F^T W (16x16) on the TRANSPOSED 4x4 tile (rows 4cj.., cols 4ri..): its F
//            operand is the one phase 1 already holds in registers
//   phase 3  lane hl < 12 solves column hl of K, lane 12 solves k (no-pivot elimination of the
//            SPD matrix, pivoted fallback if a pivot is not positive)
//   phase 4  V' = Q_xx + K^T (Q_uu K + Q_ux) + Q_ux^T K on the nine Q_xx tile lanes
//   phase 5  symmetrise V' and rewrite it in the p4 layout (all 16 lanes, 3 rows of 4 each)
constexpr int kRicStages = 2;
constexpr int kRicWarps = 4;

struct RicSmem {                               // per trajectory
  alignas(64) float rec[kRicStages][416];    // expansion records (bulk-copy destinations)
  alignas(16) float Vp[12 * 16];             // V_xx, p4 layout
  alignas(16) float W[12 * 16];              // V F; reused for the un-symmetrised V_xx' (pitch 12)
  alignas(16) float Qb[4 * 16];              // rows 12..15 of Q: [Q_ux | Q_uu]
  alignas(16) float Ks[4 * 12];              // K
  alignas(16) float Ms[4 * 12];              // Q_uu K + Q_ux
  alignas(16) float Vx[16];
  alignas(16) float q[16];                   // [Q_x | Q_u]
  alignas(16) float kh[8];                   // [k | Q_uu k + Q_u]
  alignas(8) uint64_t bar[kRicStages];
  float pad_[29];                            // sizeof == 64 (mod 128): the two half-warps of a warp work on
};                                           // structs 16 banks apart, so their 32-bit accesses never collide
static_assert(sizeof(RicSmem) % 128 == 64, "half-warp bank de-aliasing");

__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.f), r);       // one Newton step: <= 1 ulp
}

// SPD fast path: elimination without pivoting.  Returns false if a pivot is not positive.
__device__ __forceinline__ bool solve_spd4(const float (&Min)[4][4], float (&r)[4]) {
  float M[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) M[a][c] = Min[a][c];
  float inv[4];
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    ok = ok && (M[c][c] > 0.f);
    inv[c] = fast_rcp(M[c][c]);
#pragma unroll
    for (int rr = c + 1; rr < 4; ++rr) {
      const float f = M[rr][c] * inv[c];
#pragma unroll
      for (int j = c + 1; j < 4; ++j) M[rr][j] = fmaf(-f, M[c][j], M[rr][j]);
      r[rr] = fmaf(-f, r[c], r[rr]);
    }
  }
#pragma unroll
  for (int c = 3; c >= 0; --c) {
    float acc = r[c];
#pragma unroll
    for (int j = c + 1; j < 4; ++j) acc = fmaf(-M[c][j], r[j], acc);
    r[c] = acc * inv[c];
  }
  return ok;
}

__device__ __noinline__ void solve_pp4_slow(const float (&Min)[4][4], float (&r)[4]) {
  float M[4][4];
  for (int a = 0; a < 4; ++a)
    for (int c = 0; c < 4; ++c) M[a][c] = Min[a][c];
  solve_pp<4>(M, r);
}

__global__ void __launch_bounds__(32 * kRicWarps, 4)
riccati_warp_kernel(const float* __restrict__ expansion, const float* __restrict__ terminal,
                    float* __restrict__ gains, const int32_t* __restrict__ status, int B, int N, int t_start,
                    float reg) {
  constexpr int n = 12, m = 4;
  using R = Rec<n, m>;
  static_assert(R::size == 416 && R::gain == 52 && R::pitch == 16, "layout");
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ RicSmem smem[kRicWarps][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, hl = lane & 15;
  const int b_mine = (blockIdx.x * kRicWarps + warp) * 2 + half;
  const bool valid = (b_mine < B) && (status == nullptr || status[b_mine] == 0);
  const unsigned vmask = __ballot_sync(FULL, valid);
  if (vmask == 0u) return;
  // an idle half shadows the other half's trajectory (same loads, no stores) so the warp stays convergent
  const int b = valid ? b_mine : (blockIdx.x * kRicWarps + warp) * 2 + (1 - half);
  RicSmem& s = smem[warp][half];
  const int steps = N - t_start;

  if (hl == 0) {
    for (int i = 0; i < kRicStages; ++i) mbar_init(&s.bar[i], 1);
    mbar_fence_init();
  }
  __syncwarp();
  auto issue = [&](int k) {   // lane hl == 0: sequence k <-> t = N-1-k
    const int t = N - 1 - k;
    mbar_arrive_expect_tx(&s.bar[k % kRicStages], R::size * sizeof(float));
    bulk_g2s(s.rec[k % kRicStages], expansion + ((size_t)t * B + b) * R::size, R::size * sizeof(float),
             &s.bar[k % kRicStages]);
  };
  if (hl == 0)
    for (int k = 0; k < kRicStages && k < steps; ++k) issue(k);

  // terminal condition -> V_x, V_xx (p4 layout)
  {
    const float* tm = terminal + (size_t)b * R::term;
    for (int e = hl; e < R::term; e += 16) {
      const float v = tm[e];
      if (e < n) {
        s.Vp[e * 16 + 15] = v;            // V_x rides in the spare slot of row group 3
      } else {
        const int i = (e - n) / n, j = (e - n) % n;
        s.Vp[i * 16 + 4 * (j / 3) + (j % 3)] = v;
      }
    }
  }
  __syncwarp();

  const int ri = hl >> 2, cj = hl & 3;

  for (int k = 0; k < steps; ++k) {
    const int t = N - 1 - k;
    mbar_wait(&s.bar[k % kRicStages], (k / kRicStages) & 1u);
    const float* rec = s.rec[k % kRicStages];
    const float* F = rec + R::F;

    // ---- phase 1: [W; q^T] = [V; V_x^T] F on 3x4 (+1) tiles -------------------------------------------
    // The p4 layout has a spare 4th slot per row group; slot 3 of group 3 carries V_x, so the lanes
    // with ri == 3 get q[4cj..4cj+3] = [l_x|l_u] + F^T V_x for free.  The lane's four F columns are
    // kept in registers: phase 2 uses the transposed tile (rows 4cj.., cols 4ri..) and needs the same.
    float4 fc[n];
    {
      float acc[4][4] = {};
#pragma unroll
      for (int l = 0; l < n; ++l) {
        const float4 a = *reinterpret_cast<const float4*>(&s.Vp[l * 16 + 4 * ri]);     // V[l][3ri..3ri+2], (V_x[l])
        const float4 f = *reinterpret_cast<const float4*>(F + l * 16 + 4 * cj);
        fc[l] = f;
        acc[0][0] = fmaf(a.x, f.x, acc[0][0]); acc[0][1] = fmaf(a.x, f.y, acc[0][1]);
        acc[0][2] = fmaf(a.x, f.z, acc[0][2]); acc[0][3] = fmaf(a.x, f.w, acc[0][3]);
        acc[1][0] = fmaf(a.y, f.x, acc[1][0]); acc[1][1] = fmaf(a.y, f.y, acc[1][1]);
        acc[1][2] = fmaf(a.y, f.z, acc[1][2]); acc[1][3] = fmaf(a.y, f.w, acc[1][3]);
        acc[2][0] = fmaf(a.z, f.x, acc[2][0]); acc[2][1] = fmaf(a.z, f.y, acc[2][1]);
        acc[2][2] = fmaf(a.z, f.z, acc[2][2]); acc[2][3] = fmaf(a.z, f.w, acc[2][3]);
        acc[3][0] = fmaf(a.w, f.x, acc[3][0]); acc[3][1] = fmaf(a.w, f.y, acc[3][1]);
        acc[3][2] = fmaf(a.w, f.z, acc[3][2]); acc[3][3] = fmaf(a.w, f.w, acc[3][3]);
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
        *reinterpret_cast<float4*>(&s.W[(3 * ri + r) * 16 + 4 * cj]) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      if (ri == 3) {
        const float4 lq = *reinterpret_cast<const float4*>(rec + R::lx + 4 * cj);
        *reinterpret_cast<float4*>(&s.q[4 * cj]) = make_float4(lq.x + acc[3][0], lq.y + acc[3][1], lq.z + acc[3][2], lq.w + acc[3][3]);
      }
    }
    __syncwarp();

    // ---- phase 2: Q = L + F^T W on TRANSPOSED 4x4 tiles: rows 4cj.., cols 4ri.. -----------------------------
    float Q[4][4];
    {
      // cost Hessian block of this tile (the Q_xu tile, cj < 3 && ri == 3, is never used)
      const float* lsrc = (cj < 3) ? (rec + R::lxx + (4 * cj) * n + 4 * ri)
                                   : ((ri < 3) ? (rec + R::lux + 4 * ri) : (rec + R::luu));
      const int lstride = (cj < 3 || ri < 3) ? n : m;
      const bool have_l = !(cj < 3 && ri == 3);
      // two predicated rounds: row groups cj == 0 and cj == 2 of l_xx are 96 words apart (same banks)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (have_l && cj != 2) v = *reinterpret_cast<const float4*>(lsrc + r * lstride);
        if (have_l && cj == 2) v = *reinterpret_cast<const float4*>(lsrc + r * lstride);
        Q[r][0] = v.x; Q[r][1] = v.y; Q[r][2] = v.z; Q[r][3] = v.w;
      }
#pragma unroll
      for (int l = 0; l < n; ++l) {
        const float4 a = fc[l];
        const float4 w = *reinterpret_cast<const float4*>(&s.W[l * 16 + 4 * ri]);
        Q[0][0] = fmaf(a.x, w.x, Q[0][0]); Q[0][1] = fmaf(a.x, w.y, Q[0][1]); Q[0][2] = fmaf(a.x, w.z, Q[0][2]); Q[0][3] = fmaf(a.x, w.w, Q[0][3]);
        Q[1][0] = fmaf(a.y, w.x, Q[1][0]); Q[1][1] = fmaf(a.y, w.y, Q[1][1]); Q[1][2] = fmaf(a.y, w.z, Q[1][2]); Q[1][3] = fmaf(a.y, w.w, Q[1][3]);
        Q[2][0] = fmaf(a.z, w.x, Q[2][0]); Q[2][1] = fmaf(a.z, w.y, Q[2][1]); Q[2][2] = fmaf(a.z, w.z, Q[2][2]); Q[2][3] = fmaf(a.z, w.w, Q[2][3]);
        Q[3][0] = fmaf(a.w, w.x, Q[3][0]); Q[3][1] = fmaf(a.w, w.y, Q[3][1]); Q[3][2] = fmaf(a.w, w.z, Q[3][2]); Q[3][3] = fmaf(a.w, w.w, Q[3][3]);
      }
      if (cj == 3) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
          *reinterpret_cast<float4*>(&s.Qb[r * 16 + 4 * ri]) = make_float4(Q[r][0], Q[r][1], Q[r][2], Q[r][3]);
      }
    }
    __syncwarp();
    // the record is consumed: refill its stage with the record of step k + stages
    if (hl == 0 && k + kRicStages < steps) issue(k + kRicStages);

    // ---- phase 3: gains.  lane hl < 12 solves column hl of K, lane 12 solves k -------------------
    {
      float Quu[m][m], Mr[m][m], rhs[m], src[m];
#pragma unroll
      for (int a = 0; a < m; ++a) {
        const float4 v = *reinterpret_cast<const float4*>(&s.Qb[a * 16 + n]);
        Quu[a][0] = v.x; Quu[a][1] = v.y; Quu[a][2] = v.z; Quu[a][3] = v.w;
      }
      const int col = (hl < n) ? hl : 0;
#pragma unroll
      for (int a = 0; a < m; ++a) {
#pragma unroll
        for (int c = 0; c < m; ++c) Mr[a][c] = Quu[a][c] + (a == c ? reg : 0.f);
        src[a] = (hl == n) ? s.q[n + a] : s.Qb[a * 16 + col];
        rhs[a] = src[a];
      }
      const bool ok = solve_spd4(Mr, rhs);
      if (__any_sync(FULL, !ok)) {           // indefinite Q_uu: the reference's LU still "works" (a-Q(4))
#pragma unroll
        for (int a = 0; a < m; ++a) rhs[a] = src[a];
        solve_pp4_slow(Mr, rhs);
      }
      float z[m], hm[m];
#pragma unroll
      for (int a = 0; a < m; ++a) z[a] = -rhs[a];
#pragma unroll
      for (int a = 0; a < m; ++a) {
        float acc = src[a];
#pragma unroll
        for (int c = 0; c < m; ++c) acc = fmaf(Quu[a][c], z[c], acc);
        hm[a] = acc;                         // Q_uu K + Q_ux (column), or Q_uu k + Q_u
      }
      if (hl < n) {
#pragma unroll
        for (int a = 0; a < m; ++a) {
          s.Ks[a * n + hl] = z[a];
          s.Ms[a * n + hl] = hm[a];
          if (valid) gains[gain_index<QX_LAYOUT_REC>(t, m + a * n + hl, b, B, R::gain)] = z[a];
        }
      } else if (hl == n) {
#pragma unroll
        for (int a = 0; a < m; ++a) {
          s.kh[a] = z[a];
          s.kh[m + a] = hm[a];
          if (valid) gains[gain_index<QX_LAYOUT_REC>(t, a, b, B, R::gain)] = z[a];
        }
      }
    }
    __syncwarp();

    // ---- phase 4: V_xx' = Q_xx + K^T (Q_uu K + Q_ux) + Q_ux^T K ;  V_x' ----------------------------
    if (ri < 3 && cj < 3) {           // this lane holds Q_xx rows 4cj.., cols 4ri..
#pragma unroll
      for (int a = 0; a < m; ++a) {
        const float4 ka = *reinterpret_cast<const float4*>(&s.Ks[a * n + 4 * cj]);
        const float4 mb = *reinterpret_cast<const float4*>(&s.Ms[a * n + 4 * ri]);
        const float4 qa = *reinterpret_cast<const float4*>(&s.Qb[a * 16 + 4 * cj]);
        const float4 kb = *reinterpret_cast<const float4*>(&s.Ks[a * n + 4 * ri]);
        const float kar[4] = {ka.x, ka.y, ka.z, ka.w}, qar[4] = {qa.x, qa.y, qa.z, qa.w};
        const float mbc[4] = {mb.x, mb.y, mb.z, mb.w}, kbc[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) Q[r][c] = fmaf(kar[r], mbc[c], fmaf(qar[r], kbc[c], Q[r][c]));
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)    // un-symmetrised V_xx' -> W buffer, pitch 12
        *reinterpret_cast<float4*>(&s.W[(4 * cj + r) * n + 4 * ri]) = make_float4(Q[r][0], Q[r][1], Q[r][2], Q[r][3]);
    }
    float vx_new = 0.f;
    if (hl < n) {
      vx_new = s.q[hl];
#pragma unroll
      for (int a = 0; a < m; ++a) vx_new += s.Ks[a * n + hl] * s.kh[m + a] + s.Qb[a * 16 + hl] * s.kh[a];
    }
    __syncwarp();
    if (hl < n) s.Vp[hl * 16 + 15] = vx_new;

    // ---- phase 5: V = (V' + V'^T) / 2 into the p4 layout: item (i, g) = V[i][3g..3g+2] -----------------
#pragma unroll
    for (int it = 0; it < 3; ++it) {
      const int item = hl + 16 * it;           // 48 items
      const int i = item >> 2, g = item & 3;
      const float* vr = &s.W[i * n + 3 * g];
      const float* vc = &s.W[(3 * g) * n + i];
      float* dst = &s.Vp[i * 16 + 4 * g];      // slot 3 of group 3 holds V_x: leave the 4th word alone
      dst[0] = 0.5f * (vr[0] + vc[0]);
      dst[1] = 0.5f * (vr[1] + vc[n]);
      dst[2] = 0.5f * (vr[2] + vc[2 * n]);
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
      qx::riccati_warp_kernel<<<qx::ceil_div(B, 2 * qx::kRicWarps), 32 * qx::kRicWarps, 0, st>>>(
          expansion, terminal, gains, status, B, N, t_start, reg);
      QX_CHECK_LAUNCH("riccati_warp_kernel");
      return QX_OK;
  }
  qx::set_error("unknown model_id %d", model_id);
  return QX_EINVAL;
}
