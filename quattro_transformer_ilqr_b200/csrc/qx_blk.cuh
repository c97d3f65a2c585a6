// Shared pieces of the quadrotor block path (Euler map, diagonal weights): record layout, host-prepared constants
// and the per-(t, b) linearisation.
// (Tried and dropped in round 2: building the record inside the Riccati thread instead of a separate linearise launch.
// The fused loop body is 2 500 instructions = 40 KB, past the 32 KB instruction cache: ncu showed 1.2 cycles of
// "no instruction" stall per issued instruction and 230 us instead of 65 + 89 us for the two launches.
// Also tried and dropped: emitting the records from the line-search rollouts (stages A and W call quad_blk_record on
// the iterate they produce, so the next backward pass needs no linearise launch).  Bit-identical records, but the
// rollout threads run alone on their scheduler and every instruction added to them is serial time: +73 us per
// rollout pass against the 59 us of the stand-alone kernel, which spreads the same work over 50 x more threads.)
#pragma once
#include "qx_models.cuh"

namespace qx {

constexpr int kBlkSlots = 12;            // float4 slots per (t, b) record
constexpr int kBlkRec = 4 * kBlkSlots;   // 48 floats = 192 B: exactly the entries of the expansion that vary
constexpr int kBlkOct = 8 * kBlkRec;     // floats per octet of trajectories (twelve 128-B lines)
constexpr int kBlkTerm = 16;             // floats per trajectory: [V_x(12) | pad]
// The blocks have fixed entries besides the variable ones,
//   G = |g00 g01 g02|  Hp = |h00 h01 0|  E = |dt e01 e02|  Sp = |1   s01 s02|
//       |g10 g11 g12|       |h10 1   0|      |0  e11 e12|       |s10 1   s12|
//       |g20 g21 0  |       |h20 h21 1|      |0  e21 e22|       |s20 s21 1  |
// and only the variable ones are stored.  float4 slots of a record:
//   0: g00 g01 g02 d0 | 1: g10 g11 g12 d1 | 2: g20 g21 d2 h00 | 3: h01 h10 h20 h21 | 4: e01 e02 e11 e12
//   5: e21 e22 s01 s02 | 6: s10 s12 s20 s21 | 7: l_u | 8: diag(l_uu) | 9-11: l_x

// constants of the Riccati recursion
struct BlkConst {
  float dt, reg;
  float tr[3];        // dt*arm/Ix, dt*arm/Iy, dt*kyaw/Iz
  float lq[12];       // diag(l_xx) = 2 Q_ii
  float lqf[12];      // diag(V_xx at N) = 2 Qf_ii
};

// constants of the linearisation, prepared on the host with the float operations the kernels used to do per thread
// (a __grid_constant__ parameter: fp32 instructions take constant-bank operands, so none of this costs a register)
struct LinConst {
  float dt, mass, c1, c2, c3, im;
  float q2[12], qf2[12], xr[12], r2[4];
  float beta, alpha_b;
};

inline BlkConst blk_consts(const QxParams& p, float reg) {
  BlkConst c;
  c.dt = (float)p.dt;
  c.reg = reg;
  c.tr[0] = (float)(p.dt * p.phys[4] / p.phys[1]);
  c.tr[1] = (float)(p.dt * p.phys[4] / p.phys[2]);
  c.tr[2] = (float)(p.dt * p.phys[6] / p.phys[3]);
  for (int i = 0; i < 12; ++i) {
    c.lq[i] = (float)(2.0 * p.Q[i * 12 + i]);
    c.lqf[i] = (float)(2.0 * p.Qf[i * 12 + i]);
  }
  return c;
}

inline LinConst lin_consts(const QxParams& p) {
  LinConst c;
  c.dt = (float)p.dt;
  c.mass = (float)p.phys[0];
  const float Ix = (float)p.phys[1], Iy = (float)p.phys[2], Iz = (float)p.phys[3];
  c.c1 = (Iy - Iz) / Ix;
  c.c2 = (Iz - Ix) / Iy;
  c.c3 = (Ix - Iy) / Iz;
  c.im = 1.f / c.mass;
  for (int i = 0; i < 12; ++i) {
    c.q2[i] = 2.f * (float)p.Q[i * 12 + i];
    c.qf2[i] = 2.f * (float)p.Qf[i * 12 + i];
    c.xr[i] = (float)p.x_ref[i];
  }
  for (int j = 0; j < 4; ++j) c.r2[j] = 2.f * (float)p.R[j * 4 + j];
  c.beta = (float)p.barrier_beta;
  c.alpha_b = (float)p.barrier_alpha;
  return c;
}

// first and second derivative of alpha * softplus(-u; beta)^2 (quadrotor_mpc.py:80, :91-92), fp32: the expressions of
// Quadrotor::barrier_d<float> (qx_models.cuh)
__device__ __forceinline__ void blk_barrier_d(const float* u, const LinConst& k, float* g, float* h) {
  const float beta = k.beta, al = k.alpha_b;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float z = -u[i];
    const float bz = beta * z;
    const float s = (fmaxf(bz, 0.f) + log1pf(expf(-fabsf(bz)))) / beta;
    const float sg = 1.f / (1.f + expf(-beta * z));   // softplus'
    const float s2 = beta * sg * (1.f - sg);          // softplus''
    g[i] = -2.f * al * s * sg;
    h[i] = 2.f * al * (sg * sg + s * s2);
  }
}

// _compute_dynamics_jacobians + _compute_cost_derivatives (quattro_ilqr_tf.py:182-275) at one (x_t, u_t), exact
// derivatives, as the 15 float4 slots of the block record (same expressions as Quadrotor::fc_jac, scaled by dt)
__device__ __forceinline__ void quad_blk_record(const LinConst& k, const float* xs, const float* us, float4* rec) {
  const float dt = k.dt, mass = k.mass, c1 = k.c1, c2 = k.c2, c3 = k.c3, im = k.im;
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
  float gx[12], bg[4], bh[4], lu[4], luu[4];
#pragma unroll
  for (int i = 0; i < 12; ++i) gx[i] = k.q2[i] * (xs[i] - k.xr[i]);
  blk_barrier_d(us, k, bg, bh);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    lu[j] = k.r2[j] * us[j] + bg[j];
    luu[j] = k.r2[j] + bh[j];
  }
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
}

// the 15 logical slots of quad_blk_record -> the 12 stored slots (drops the entries that are 0, 1 or dt by construction)
__device__ __forceinline__ void quad_blk_pack(const float4* r, float4* p) {
  p[0] = r[0];
  p[1] = r[1];
  p[2] = make_float4(r[2].x, r[2].y, r[2].w, r[3].x);
  p[3] = make_float4(r[3].y, r[4].x, r[5].x, r[5].y);
  p[4] = make_float4(r[6].y, r[6].z, r[7].y, r[7].z);
  p[5] = make_float4(r[8].y, r[8].z, r[9].y, r[9].z);
  p[6] = make_float4(r[10].x, r[10].z, r[11].x, r[11].y);
  p[7] = make_float4(r[3].w, r[4].w, r[5].w, r[6].w);
  p[8] = make_float4(r[7].w, r[8].w, r[9].w, r[10].w);
  p[9] = r[12];
  p[10] = r[13];
  p[11] = r[14];
}

// terminal gradient V_x(N) = 2 Qf (x_N - x_ref); V_xx(N) = 2 Qf is constant
__device__ __forceinline__ void quad_blk_terminal(const LinConst& k, const float* xN, float* g) {
#pragma unroll
  for (int i = 0; i < 12; ++i) g[i] = k.qf2[i] * (xN[i] - k.xr[i]);
}

}  // namespace qx
