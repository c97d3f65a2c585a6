// Structure-exploiting backward pass for the quadrotor with the Euler map and diagonal cost
// weights - the configuration every driver of the reference uses (quadrotor_sim.py:100,
// quadrotor_mpc.py:40-46).
//
//   linearize_quad_blk_kernel = _compute_dynamics_jacobians + _compute_cost_derivatives
//                               (quattro_ilqr_tf.py:182-275), exact derivatives, COMPACT record
//   riccati_quad_blk_kernel   = backward_pass / backward_pass_segment (:282-366)
//
// With x = [p | v | eta | omega] (four 3-vectors) the discrete Jacobians are block bidiagonal,
//
//        | I  dt I  0   0  |          | 0      |      G  = dt d(vdot)/d(eta)      Hp = I + dt d(etadot)/d(eta)
//    A = | 0   I    G   0  |      B = | d 1^T  |      E  = dt d(etadot)/d(omega)  Sp = I + dt d(omegadot)/d(omega)
//        | 0   0    Hp  E  |          | 0      |      d  = dt/m * thrust direction
//        | 0   0    0   Sp |          | T      |      T  = dt * (constant torque map, +-tr[r])
//
// and l_xx = 2Q, l_uu = 2R + barrier'' are diagonal, l_ux = 0.  So instead of the dense 416-float
// record (1 664 B per step and trajectory through HBM, twice) linearise emits the 36 block entries
// plus d, l_x, l_u, diag(l_uu): 64 floats = 256 B, and the Riccati step works on 3x3 blocks.
//
// Riccati: FOUR lanes per trajectory, lane j owns column block j of V_xx (12x3, registers) and a
// replicated V_x.  Per step (all static register indexing):
//   1  U = [A B]^T V_:j, column by column with the sparse A^T/B^T above (in-lane, exactly as sparse
//      as the math); the same map applied to V_x gives Q_x, Q_u
//   2  P_:j = U_:j X_j + U_:,j-1 Y_j with (X_j, Y_j) = (I,0), (I,dt I), (Hp,G), (Sp,E): the only
//      cross-lane traffic of A^T V A is the left neighbour's block (warp shuffles); rows 12..15
//      of the same product are Q_ux
//   3  Q_uu = l_uu + (B^T V) B from lanes 1 and 3, broadcast by shuffles
//   4  every lane eliminates the 4x4 (Q_uu + reg I) and back-substitutes its own three columns of
//      Q_ux (and Q_u): K_:j, k
//   5  V'_:j = Q_xx_:j + (Q_ux - reg K)^T K_:j,  V_x' = Q_x + (Q_ux - reg K)^T k.  This is the
//      reference's  Q_xx + K^T Q_uu K + K^T Q_ux + Q_ux^T K  (:315-316, un-regularised Q_uu) after
//      substituting Q_uu K + Q_ux = -reg K, which holds exactly for K = -(Q_uu + reg I)^-1 Q_ux and
//      avoids forming that difference by cancellation in fp32.  (Q_ux - reg K) is all-gathered
//      through shared memory (broadcast reads).
//   6  V = (V' + V'^T)/2 (:317) through a padded shared-memory transpose.  Not optional: the
//      antisymmetric rounding residue is an unstable mode of the recursion.
// Records arrive through a 3-stage ring of 1-D bulk async copies (TMA, SASS UBLKCP): the eight
// records of a warp are one contiguous 2 KB block, stored float4-interleaved over the eight
// trajectories so that a group's broadcast read of float4 #q is bank-conflict free and linearise
// writes whole 128-B lines without staging.
#include "qx_async.cuh"
#include "qx_models.cuh"

namespace qx {

constexpr int kBlkRec = 64;              // floats per (t, b) record
constexpr int kBlkOct = 8 * kBlkRec;     // floats per octet of trajectories (one bulk copy)
constexpr int kBlkTerm = 16;             // floats per trajectory: [V_x(12) | pad]
// float4 slots of a record: 0-2 G rows (.w = d), 3-5 Hp rows (.w = l_u[0..2]),
// 6-8 E rows (.w = l_u[3], l_uu[0], l_uu[1]), 9-11 Sp rows (.w = l_uu[2], l_uu[3], 0), 12-14 l_x, 15 pad

struct BlkConst {
  float dt, reg;
  float tr[3];        // dt*arm/Ix, dt*arm/Iy, dt*kyaw/Iz
  float lq[12];       // diag(l_xx) = 2 Q_ii
  float lqf[12];      // diag(V_xx at N) = 2 Qf_ii
};

__global__ void __launch_bounds__(128)
linearize_quad_blk_kernel(const __grid_constant__ QxParams p, const double* __restrict__ x,
                          const double* __restrict__ u, float* __restrict__ expansion, float* __restrict__ terminal,
                          const int32_t* __restrict__ status, int B, int N, int t_start) {
  constexpr int n = 12, m = 4;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (status != nullptr && status[b] != 0) return;
  const int t = t_start + blockIdx.y;
  if (t >= N) {      // terminal gradient 2 Qf (x_N - x_ref); V_xx = 2 Qf is constant
    float4* out = reinterpret_cast<float4*>(terminal + (size_t)b * kBlkTerm);
    float g[n];
#pragma unroll
    for (int i = 0; i < n; ++i)
      g[i] = 2.f * float(p.Qf[i * n + i]) * (float(x[((size_t)N * n + i) * B + b]) - float(p.x_ref[i]));
    out[0] = make_float4(g[0], g[1], g[2], g[3]);
    out[1] = make_float4(g[4], g[5], g[6], g[7]);
    out[2] = make_float4(g[8], g[9], g[10], g[11]);
    out[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float xs[n], us[m];
#pragma unroll
  for (int i = 0; i < n; ++i) xs[i] = float(x[((size_t)t * n + i) * B + b]);
#pragma unroll
  for (int j = 0; j < m; ++j) us[j] = float(u[((size_t)t * m + j) * B + b]);
  const float dt = float(p.dt);
  const float mass = float(p.phys[0]), Ix = float(p.phys[1]), Iy = float(p.phys[2]), Iz = float(p.phys[3]);
  const float c1 = (Iy - Iz) / Ix, c2 = (Iz - Ix) / Iy, c3 = (Ix - Iy) / Iz, im = 1.f / mass;
  float sphi, cphi, sth, cth, spsi, cpsi;
  sincosf(xs[6], &sphi, &cphi);
  sincosf(xs[7], &sth, &cth);
  sincosf(xs[8], &spsi, &cpsi);
  const float pp = xs[9], q = xs[10], r = xs[11];
  const float Tm = (us[0] + us[1] + us[2] + us[3]) / mass;
  const float sec = 1.f / cth, tth = sth * sec;
  const float dirx = spsi * sphi + cpsi * sth * cphi;
  const float diry = cpsi * sphi - spsi * sth * cphi;
  const float dirz = cth * cphi;
  const float w = q * sphi + r * cphi, v = q * cphi - r * sphi;
  float gx[n], gu[m], bg[m], bh[m];
#pragma unroll
  for (int i = 0; i < n; ++i) gx[i] = 2.f * float(p.Q[i * n + i]) * (xs[i] - float(p.x_ref[i]));
#pragma unroll
  for (int j = 0; j < m; ++j) gu[j] = 2.f * float(p.R[j * m + j]) * us[j];
  Quadrotor::barrier_d<float>(us, p, bg, bh);
  float lu[m], luu[m];
#pragma unroll
  for (int j = 0; j < m; ++j) {
    lu[j] = gu[j] + bg[j];
    luu[j] = 2.f * float(p.R[j * m + j]) + bh[j];
  }
  // same expressions as Quadrotor::fc_jac (qx_models.cuh), scaled by dt
  float4 rec[16];
  rec[0] = make_float4(dt * (Tm * (spsi * cphi - cpsi * sth * sphi)), dt * (Tm * (cpsi * cth * cphi)),
                       dt * (Tm * diry), dt * (im * dirx));
  rec[1] = make_float4(dt * (Tm * (cpsi * cphi + spsi * sth * sphi)), dt * (Tm * (-spsi * cth * cphi)),
                       dt * (-Tm * dirx), dt * (im * diry));
  rec[2] = make_float4(dt * (-Tm * cth * sphi), dt * (-Tm * sth * cphi), 0.f, dt * (im * dirz));
  rec[3] = make_float4(1.f + dt * (v * tth), dt * (w * sec * sec), 0.f, lu[0]);
  rec[4] = make_float4(dt * (-w), 1.f, 0.f, lu[1]);
  rec[5] = make_float4(dt * (v * sec), dt * (w * sth * sec * sec), 1.f, lu[2]);
  rec[6] = make_float4(dt, dt * (sphi * tth), dt * (cphi * tth), lu[3]);
  rec[7] = make_float4(0.f, dt * cphi, dt * (-sphi), luu[0]);
  rec[8] = make_float4(0.f, dt * (sphi * sec), dt * (cphi * sec), luu[1]);
  rec[9] = make_float4(1.f, dt * (c1 * r), dt * (c1 * q), luu[2]);
  rec[10] = make_float4(dt * (c2 * r), 1.f, dt * (c2 * pp), luu[3]);
  rec[11] = make_float4(dt * (c3 * q), dt * (c3 * pp), 1.f, 0.f);
  rec[12] = make_float4(gx[0], gx[1], gx[2], gx[3]);
  rec[13] = make_float4(gx[4], gx[5], gx[6], gx[7]);
  rec[14] = make_float4(gx[8], gx[9], gx[10], gx[11]);
  rec[15] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int n_oct = ceil_div(B, 8);
  float4* out = reinterpret_cast<float4*>(expansion + ((size_t)t * n_oct + (b >> 3)) * kBlkOct) + (b & 7);
#pragma unroll
  for (int qd = 0; qd < 16; ++qd) out[qd * 8] = rec[qd];     // a warp writes 4 full 128-B lines per slot
}

// ------------------------------------------------------------------------------------------------
constexpr int kBlkStages = 3;
constexpr int kBlkWarps = 2;
constexpr int kBufPitch = 20;                  // row pitch of the per-trajectory exchange buffer (floats)
constexpr int kBufGroup = 12 * kBufPitch;      // 240 = 16 (mod 32): the two groups of a quarter-warp never collide

struct BlkSmem {
  alignas(128) float ring[kBlkWarps][kBlkStages][kBlkOct];
  alignas(16) float buf[kBlkWarps * 8][kBufGroup];
  alignas(8) uint64_t bar[kBlkWarps][kBlkStages];
};

__device__ __forceinline__ float blk_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.f), r);       // one Newton step: <= 1 ulp
}

// pivoted fallback for an indefinite Q_uu (the reference's LU inverse still "works", SURVEY a-Q(4))
__device__ __noinline__ void blk_solve_pp(const float* Min, float* r) {
  float M[4][4];
  for (int a = 0; a < 4; ++a)
    for (int c = 0; c < 4; ++c) M[a][c] = Min[a * 4 + c];
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    for (int rr = c + 1; rr < 4; ++rr)
      if (fabsf(M[rr][c]) > fabsf(M[piv][c])) piv = rr;
    if (piv != c) {
      for (int jj = 0; jj < 4; ++jj) { const float tmp = M[c][jj]; M[c][jj] = M[piv][jj]; M[piv][jj] = tmp; }
      const float tmp = r[c]; r[c] = r[piv]; r[piv] = tmp;
    }
    const float inv = 1.0f / M[c][c];
    for (int rr = c + 1; rr < 4; ++rr) {
      const float f = M[rr][c] * inv;
      for (int jj = c + 1; jj < 4; ++jj) M[rr][jj] -= f * M[c][jj];
      r[rr] -= f * r[c];
    }
  }
  for (int c = 3; c >= 0; --c) {
    float acc = r[c];
    for (int jj = c + 1; jj < 4; ++jj) acc -= M[c][jj] * r[jj];
    r[c] = acc / M[c][c];
  }
}

__global__ void __launch_bounds__(32 * kBlkWarps, 7)
riccati_quad_blk_kernel(const __grid_constant__ BlkConst cst, const float* __restrict__ expansion,
                        const float* __restrict__ terminal, float* __restrict__ gains,
                        const int32_t* __restrict__ status, int B, int N, int t_start) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int G = 52;
  __shared__ BlkSmem sm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 2, j = lane & 3, base = lane & ~3;
  const int n_oct = ceil_div(B, 8);
  const int oct = blockIdx.x * kBlkWarps + warp;
  if (oct >= n_oct) return;
  const int b = oct * 8 + grp;
  const bool valid = (b < B) && (status == nullptr || status[b] == 0);
  if (!__any_sync(FULL, valid)) return;
  const int b_rd = min(b, B - 1);
  const int steps = N - t_start;
  uint64_t* bar = sm.bar[warp];
  float* bufg = sm.buf[warp * 8 + grp];

  if (lane == 0) {
    for (int i = 0; i < kBlkStages; ++i) mbar_init(&bar[i], 1);
    mbar_fence_init();
  }
  __syncwarp();
  auto issue = [&](int k) {       // lane 0: sequence k <-> t = N-1-k
    const int t = N - 1 - k;
    mbar_arrive_expect_tx(&bar[k % kBlkStages], kBlkOct * sizeof(float));
    bulk_g2s(sm.ring[warp][k % kBlkStages], expansion + ((size_t)t * n_oct + oct) * kBlkOct, kBlkOct * sizeof(float),
             &bar[k % kBlkStages]);
  };
  if (lane == 0)
    for (int k = 0; k < kBlkStages && k < steps; ++k) issue(k);

  const float dt = cst.dt, reg = cst.reg;
  const float tr0 = cst.tr[0], tr1 = cst.tr[1], tr2 = cst.tr[2];
  float lq[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) lq[c] = cst.lq[3 * j + c];

  // V_xx(N) = 2 Qf (diagonal), V_x(N) from linearise
  float Vc[12][3], vx[12];
#pragma unroll
  for (int r = 0; r < 12; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Vc[r][c] = (r == 3 * j + c) ? cst.lqf[r] : 0.f;
  {
    const float4* tm = reinterpret_cast<const float4*>(terminal + (size_t)b_rd * kBlkTerm);
    const float4 a0 = tm[0], a1 = tm[1], a2 = tm[2];
    vx[0] = a0.x; vx[1] = a0.y; vx[2] = a0.z; vx[3] = a0.w;
    vx[4] = a1.x; vx[5] = a1.y; vx[6] = a1.z; vx[7] = a1.w;
    vx[8] = a2.x; vx[9] = a2.y; vx[10] = a2.z; vx[11] = a2.w;
  }

  for (int k = 0; k < steps; ++k) {
    const int t = N - 1 - k;
    mbar_wait(&bar[k % kBlkStages], (k / kBlkStages) & 1u);
    const float4* rq = reinterpret_cast<const float4*>(sm.ring[warp][k % kBlkStages]) + grp;
    float Gm[3][3], Hp[3][3], Em[3][3], Sp[3][3], d[3], lu[4], luu[4], qx[12];
    {
      float4 v[12];
#pragma unroll
      for (int qd = 0; qd < 12; ++qd) v[qd] = rq[qd * 8];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        Gm[r][0] = v[r].x; Gm[r][1] = v[r].y; Gm[r][2] = v[r].z; d[r] = v[r].w;
        Hp[r][0] = v[3 + r].x; Hp[r][1] = v[3 + r].y; Hp[r][2] = v[3 + r].z;
        Em[r][0] = v[6 + r].x; Em[r][1] = v[6 + r].y; Em[r][2] = v[6 + r].z;
        Sp[r][0] = v[9 + r].x; Sp[r][1] = v[9 + r].y; Sp[r][2] = v[9 + r].z;
      }
      lu[0] = v[3].w; lu[1] = v[4].w; lu[2] = v[5].w; lu[3] = v[6].w;
      luu[0] = v[7].w; luu[1] = v[8].w; luu[2] = v[9].w; luu[3] = v[10].w;
      const float4 l0 = rq[12 * 8], l1 = rq[13 * 8], l2 = rq[14 * 8];
      qx[0] = l0.x; qx[1] = l0.y; qx[2] = l0.z; qx[3] = l0.w;
      qx[4] = l1.x; qx[5] = l1.y; qx[6] = l1.z; qx[7] = l1.w;
      qx[8] = l2.x; qx[9] = l2.y; qx[10] = l2.z; qx[11] = l2.w;
      // every one of the 15 loads has RETURNED before the stage is handed to the next bulk copy: nothing else orders
      // a shared-memory read still in flight against the async-proxy write that refills the stage
      asm volatile("" ::"f"(v[0].x), "f"(v[1].x), "f"(v[2].x), "f"(v[3].x), "f"(v[4].x), "f"(v[5].x), "f"(v[6].x), "f"(v[7].x),
                   "f"(v[8].x), "f"(v[9].x), "f"(v[10].x), "f"(v[11].x), "f"(l0.x), "f"(l1.x), "f"(l2.x)
                   : "memory");
    }
    __syncwarp();
    // the record is in registers: refill its stage with the record of step k + stages
    if (lane == 0 && k + kBlkStages < steps) issue(k + kBlkStages);

    // ---- 1: y = A^T z, w = B^T z for the three columns of V_:j and for V_x ---------------------
    auto apply = [&](const float* z, float* y, float* w) {
      y[0] = z[0]; y[1] = z[1]; y[2] = z[2];
      y[3] = fmaf(dt, z[0], z[3]); y[4] = fmaf(dt, z[1], z[4]); y[5] = fmaf(dt, z[2], z[5]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float a = Gm[0][c] * z[3];
        a = fmaf(Gm[1][c], z[4], a); a = fmaf(Gm[2][c], z[5], a);
        a = fmaf(Hp[0][c], z[6], a); a = fmaf(Hp[1][c], z[7], a); a = fmaf(Hp[2][c], z[8], a);
        y[6 + c] = a;
        float e = Em[0][c] * z[6];
        e = fmaf(Em[1][c], z[7], e); e = fmaf(Em[2][c], z[8], e);
        e = fmaf(Sp[0][c], z[9], e); e = fmaf(Sp[1][c], z[10], e); e = fmaf(Sp[2][c], z[11], e);
        y[9 + c] = e;
      }
      const float s = fmaf(d[2], z[5], fmaf(d[1], z[4], d[0] * z[3]));
      const float e0 = tr0 * z[9], e1 = tr1 * z[10], e2 = tr2 * z[11];
      w[0] = (s - e0) + (e1 + e2);
      w[1] = (s + e0) + (e1 - e2);
      w[2] = (s + e0) - (e1 - e2);
      w[3] = (s - e0) - (e1 + e2);
    };
    float Uu[4][3], qu[4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float z[12], y[12], w[4];
#pragma unroll
      for (int r = 0; r < 12; ++r) z[r] = Vc[r][c];
      apply(z, y, w);
#pragma unroll
      for (int r = 0; r < 12; ++r) Vc[r][c] = y[r];
#pragma unroll
      for (int a = 0; a < 4; ++a) Uu[a][c] = w[a];
    }
    {
      float y[12], w[4];
      apply(vx, y, w);
#pragma unroll
      for (int r = 0; r < 12; ++r) qx[r] += y[r];
#pragma unroll
      for (int a = 0; a < 4; ++a) qu[a] = lu[a] + w[a];
    }

    // ---- 3: Q_uu = l_uu + (B^T V)_:1 d 1^T + (B^T V)_:3 T --------------------------------------
    float Quu[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float wd = fmaf(Uu[a][2], d[2], fmaf(Uu[a][1], d[1], Uu[a][0] * d[0]));
      const float f0 = Uu[a][0] * tr0, f1 = Uu[a][1] * tr1, f2 = Uu[a][2] * tr2;
      const float ut0 = (f1 + f2) - f0, ut1 = (f0 + f1) - f2, ut2 = (f0 - f1) + f2, ut3 = -(f0 + f1) - f2;
      const float wds = __shfl_sync(FULL, wd, base + 1);
      Quu[a][0] = wds + __shfl_sync(FULL, ut0, base + 3);
      Quu[a][1] = wds + __shfl_sync(FULL, ut1, base + 3);
      Quu[a][2] = wds + __shfl_sync(FULL, ut2, base + 3);
      Quu[a][3] = wds + __shfl_sync(FULL, ut3, base + 3);
      Quu[a][a] += luu[a];
    }

    // ---- 2: [P; Q_ux]_:j = U_:j X_j + U_:,j-1 Y_j ---------------------------------------------
    {
      float X[3][3], Y[3][3];
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          X[s][c] = (j == 2) ? Hp[s][c] : (j == 3) ? Sp[s][c] : (s == c ? 1.f : 0.f);
          Y[s][c] = (j == 2) ? Gm[s][c] : (j == 3) ? Em[s][c] : ((j == 1 && s == c) ? dt : 0.f);
        }
      const int src = base + ((j + 3) & 3);
      auto row = [&](float* uu) {
        const float n0 = __shfl_sync(FULL, uu[0], src), n1 = __shfl_sync(FULL, uu[1], src),
                    n2 = __shfl_sync(FULL, uu[2], src);
        float o[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float a = uu[0] * X[0][c];
          a = fmaf(uu[1], X[1][c], a); a = fmaf(uu[2], X[2][c], a);
          a = fmaf(n0, Y[0][c], a); a = fmaf(n1, Y[1][c], a); a = fmaf(n2, Y[2][c], a);
          o[c] = a;
        }
        uu[0] = o[0]; uu[1] = o[1]; uu[2] = o[2];
      };
#pragma unroll
      for (int r = 0; r < 12; ++r) row(Vc[r]);
#pragma unroll
      for (int a = 0; a < 4; ++a) row(Uu[a]);        // now Q_ux[:, 3j..3j+2]
    }

    // ---- 4: gains.  Eliminate (Q_uu + reg I) once, back-substitute 3 columns of Q_ux and Q_u --
    float Kc[4][3], kk[4];
    {
      float M[4][4], inv[4], f[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) M[a][c] = Quu[a][c] + (a == c ? reg : 0.f);
      bool ok = true;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ok = ok && (M[c][c] > 0.f);
        inv[c] = blk_rcp(M[c][c]);
#pragma unroll
        for (int rr = c + 1; rr < 4; ++rr) {
          f[rr][c] = M[rr][c] * inv[c];
#pragma unroll
          for (int jj = c + 1; jj < 4; ++jj) M[rr][jj] = fmaf(-f[rr][c], M[c][jj], M[rr][jj]);
        }
      }
      auto solve = [&](float* r) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int rr = c + 1; rr < 4; ++rr) r[rr] = fmaf(-f[rr][c], r[c], r[rr]);
#pragma unroll
        for (int c = 3; c >= 0; --c) {
          float acc = r[c];
#pragma unroll
          for (int jj = c + 1; jj < 4; ++jj) acc = fmaf(-M[c][jj], r[jj], acc);
          r[c] = acc * inv[c];
        }
      };
      float rhs[4][4];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int a = 0; a < 4; ++a) rhs[c][a] = Uu[a][c];
#pragma unroll
      for (int a = 0; a < 4; ++a) rhs[3][a] = qu[a];
      if (ok) {
#pragma unroll
        for (int c = 0; c < 4; ++c) solve(rhs[c]);
      } else {
        float Mr[16];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) Mr[a * 4 + c] = Quu[a][c] + (a == c ? reg : 0.f);
        for (int c = 0; c < 4; ++c) blk_solve_pp(Mr, rhs[c]);
      }
      __syncwarp();
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Kc[a][c] = -rhs[c][a];
        kk[a] = -rhs[3][a];
      }
    }
    if (valid) {
      float* gr = gains + ((size_t)t * B + b) * G;
      if (j == 0) *reinterpret_cast<float4*>(gr) = make_float4(kk[0], kk[1], kk[2], kk[3]);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) gr[4 + 12 * a + 3 * j + c] = Kc[a][c];
    }

    // ---- 5: V' = Q_xx + (Q_ux - reg K)^T K,  V_x' = Q_x + (Q_ux - reg K)^T k -------------------
#pragma unroll
    for (int a = 0; a < 4; ++a)
      *reinterpret_cast<float4*>(&bufg[a * kBufPitch + 4 * j]) =
          make_float4(fmaf(-reg, Kc[a][0], Uu[a][0]), fmaf(-reg, Kc[a][1], Uu[a][1]), fmaf(-reg, Kc[a][2], Uu[a][2]), 0.f);
    // l_xx = 2Q sits on the diagonal: rows 3j+c of column c
#pragma unroll
    for (int rb = 0; rb < 4; ++rb)
      if (j == rb) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Vc[3 * rb + c][c] += lq[c];
      }
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float ra[12];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 v = *reinterpret_cast<const float4*>(&bufg[a * kBufPitch + 4 * jj]);
        ra[3 * jj] = v.x; ra[3 * jj + 1] = v.y; ra[3 * jj + 2] = v.z;
      }
#pragma unroll
      for (int r = 0; r < 12; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Vc[r][c] = fmaf(ra[r], Kc[a][c], Vc[r][c]);
        qx[r] = fmaf(ra[r], kk[a], qx[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < 12; ++r) vx[r] = qx[r];
    __syncwarp();

    // ---- 6: V = (V' + V'^T) / 2 ----------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < 12; ++r)
      *reinterpret_cast<float4*>(&bufg[r * kBufPitch + 4 * j]) = make_float4(Vc[r][0], Vc[r][1], Vc[r][2], 0.f);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* rowp = &bufg[(3 * j + c) * kBufPitch];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 v = *reinterpret_cast<const float4*>(rowp + 4 * jj);
        Vc[3 * jj][c] = 0.5f * (Vc[3 * jj][c] + v.x);
        Vc[3 * jj + 1][c] = 0.5f * (Vc[3 * jj + 1][c] + v.y);
        Vc[3 * jj + 2][c] = 0.5f * (Vc[3 * jj + 2][c] + v.z);
      }
    }
    __syncwarp();
  }
}

}  // namespace qx

// ---- C ABI -----------------------------------------------------------------------------------------
static bool blk_ok(const QxParams* p) {
  return p != nullptr && p->model_id == QX_MODEL_QUADROTOR && p->integrator == QX_EULER && p->diag_cost != 0;
}

extern "C" int qx_blk_supported(const QxParams* p) { return blk_ok(p) ? 1 : 0; }

extern "C" size_t qx_blk_expansion_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return (size_t)N * (size_t)qx::ceil_div(B, 8) * qx::kBlkOct * sizeof(float);
}

extern "C" size_t qx_blk_terminal_bytes(int B) { return B > 0 ? (size_t)B * qx::kBlkTerm * sizeof(float) : 0; }

extern "C" int qx_linearize_blk(const QxParams* p, const double* x, const double* u, float* expansion, float* terminal,
                                const int32_t* status, int B, int N, int t_start, void* stream) {
  QX_CHECK_ARG(p && x && u && expansion && terminal, "null pointer");
  QX_CHECK_ARG(blk_ok(p), "the block path needs the quadrotor, the Euler map and diagonal Q, R, Qf");
  QX_CHECK_ARG(B > 0 && N > 0 && t_start >= 0 && t_start < N, "need B > 0 and 0 <= t_start < N");
  QX_CHECK_ARG((reinterpret_cast<uintptr_t>(expansion) & 127u) == 0 && (reinterpret_cast<uintptr_t>(terminal) & 15u) == 0,
               "expansion must be 128-byte and terminal 16-byte aligned");
  const dim3 grid(qx::ceil_div(B, 128), N - t_start + 1);
  qx::linearize_quad_blk_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p, x, u, expansion, terminal,
                                                                                   status, B, N, t_start);
  QX_CHECK_LAUNCH("linearize_quad_blk_kernel");
  return QX_OK;
}

extern "C" int qx_riccati_blk(const QxParams* p, const float* expansion, const float* terminal, float* gains,
                              const int32_t* status, int B, int N, int t_start, float reg, void* stream) {
  QX_CHECK_ARG(p && expansion && terminal && gains, "null pointer");
  QX_CHECK_ARG(blk_ok(p), "the block path needs the quadrotor, the Euler map and diagonal Q, R, Qf");
  QX_CHECK_ARG(B > 0 && N > 0 && t_start >= 0 && t_start < N, "need B > 0 and 0 <= t_start < N");
  QX_CHECK_ARG((reinterpret_cast<uintptr_t>(expansion) & 127u) == 0 && (reinterpret_cast<uintptr_t>(gains) & 15u) == 0,
               "expansion must be 128-byte and gains 16-byte aligned");
  qx::BlkConst c;
  c.dt = (float)p->dt;
  c.reg = reg;
  c.tr[0] = (float)(p->dt * p->phys[4] / p->phys[1]);
  c.tr[1] = (float)(p->dt * p->phys[4] / p->phys[2]);
  c.tr[2] = (float)(p->dt * p->phys[6] / p->phys[3]);
  for (int i = 0; i < 12; ++i) {
    c.lq[i] = (float)(2.0 * p->Q[i * 12 + i]);
    c.lqf[i] = (float)(2.0 * p->Qf[i * 12 + i]);
  }
  const int n_oct = qx::ceil_div(B, 8);
  qx::riccati_quad_blk_kernel<<<qx::ceil_div(n_oct, qx::kBlkWarps), 32 * qx::kBlkWarps, 0,
                                static_cast<cudaStream_t>(stream)>>>(c, expansion, terminal, gains, status, B, N, t_start);
  QX_CHECK_LAUNCH("riccati_quad_blk_kernel");
  return QX_OK;
}
