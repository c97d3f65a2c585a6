// Structure-exploiting backward pass for the quadrotor with the Euler map and diagonal cost
// weights - the configuration every driver of the reference uses (quadrotor_sim.py:100,
// quadrotor_mpc.py:40-46).
//
//   linearize_quad_blk_kernel = _compute_dynamics_jacobians + _compute_cost_derivatives
//                               (quattro_ilqr_tf.py:182-275), exact derivatives, COMPACT record
//   riccati_quad_thr_kernel   = backward_pass / backward_pass_segment (:282-366)    (qx_blk_riccati.cu)
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
// Records are stored float4-interleaved over eight trajectories ([t][b/8][16 slots][8][4 floats]), so that
// linearise writes whole 128-B lines without staging and the Riccati thread of trajectory b copies its sixteen
// 16-B slots with a warp touching four full lines per slot.
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

}  // namespace qx

namespace qx {
int blk_launch_riccati(const BlkConst& c, const float* expansion, const float* terminal, float* gains,
                        const int32_t* status, int B, int N, int t_start, cudaStream_t st);
}

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
  return qx::blk_launch_riccati(c, expansion, terminal, gains, status, B, N, t_start, static_cast<cudaStream_t>(stream));
}
