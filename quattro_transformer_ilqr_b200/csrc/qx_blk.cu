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
// record (1 664 B per step and trajectory through HBM, twice) linearise emits the 25 block entries that vary
// (the blocks also hold fixed 0 / 1 / dt entries, qx_blk.cuh) plus d, l_x, l_u, diag(l_uu): 48 floats = 192 B,
// and the Riccati step works on 3x3 blocks.
// Records are stored float4-interleaved over eight trajectories ([t][b/8][12 slots][8][4 floats]), so that
// linearise writes whole 128-B lines without staging and the Riccati thread of trajectory b copies its twelve
// 16-B slots with a warp touching four full lines per slot.
#include "qx_blk.cuh"

namespace qx {

__global__ void __launch_bounds__(128)
linearize_quad_blk_kernel(const __grid_constant__ LinConst k, const double* __restrict__ x,
                          const double* __restrict__ u, float* __restrict__ expansion, float* __restrict__ terminal,
                          const int32_t* __restrict__ status, int B, int N, int t_start) {
  constexpr int n = 12, m = 4;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (status != nullptr && status[b] != 0) return;
  const int t = t_start + blockIdx.y;
  float xs[n];
#pragma unroll
  for (int i = 0; i < n; ++i) xs[i] = float(x[((size_t)min(t, N) * n + i) * B + b]);
  if (t >= N) {
    float4* out = reinterpret_cast<float4*>(terminal + (size_t)b * kBlkTerm);
    float g[n];
    quad_blk_terminal(k, xs, g);
    out[0] = make_float4(g[0], g[1], g[2], g[3]);
    out[1] = make_float4(g[4], g[5], g[6], g[7]);
    out[2] = make_float4(g[8], g[9], g[10], g[11]);
    out[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float us[m];
#pragma unroll
  for (int j = 0; j < m; ++j) us[j] = float(u[((size_t)t * m + j) * B + b]);
  float4 rec[15], pk[kBlkSlots];
  quad_blk_record(k, xs, us, rec);
  quad_blk_pack(rec, pk);
  const int n_oct = ceil_div(B, 8);
  float4* out = reinterpret_cast<float4*>(expansion + ((size_t)t * n_oct + (b >> 3)) * kBlkOct) + (b & 7);
#pragma unroll
  for (int qd = 0; qd < kBlkSlots; ++qd) out[qd * 8] = pk[qd];     // a warp writes 4 full 128-B lines per slot
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
  qx::linearize_quad_blk_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(qx::lin_consts(*p), x, u, expansion,
                                                                                   terminal, status, B, N, t_start);
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
  const qx::BlkConst c = qx::blk_consts(*p, reg);
  return qx::blk_launch_riccati(c, expansion, terminal, gains, status, B, N, t_start, static_cast<cudaStream_t>(stream));
}
