/*
 * quattro_b200.h — C ABI of libquattro_sm100a.so, the B200 (sm_100a) batched iLQR hot path.
 *
 * The reference (salemon/quattro-transformer-ilqr) has no native layer: its boundary is the
 * Python class iLQR_TF (quattro_ilqr_tf/quattro_ilqr_tf.py:50-612) and the duck-typed
 * predictor TransformerILQR.predict (quattro_ilqr_tf/transformer_ilqr.py:311-325).  Each entry
 * point below replaces the body of one reference method for a BATCH of B independent
 * trajectories; the Python host (quattro_transformer_ilqr_b200/) keeps the reference's class
 * API on top of these calls.  All citations are relative to /root/reference.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller
 *    (PyTorch tensors' data_ptr()); the library never allocates, frees or synchronises;
 *  - work is enqueued on `stream` (a cudaStream_t passed as void*);
 *  - return 0 on success, a negative QX_E* code otherwise; qx_last_error() gives the text
 *    (thread-local); nothing is thrown across the boundary;
 *  - no global mutable state; re-entrant across streams.
 *
 * Device data layout (B = batch, N = horizon, n = state dim, m = control dim, b fastest)
 *   x      double [N+1][n][B]      trajectories (fp64: rollouts and costs are fp64)
 *   u      double [N][m][B]
 *   cost   double [B]
 *   status int32  [B]              0 = active, 1 = finished (converged or no alpha accepted)
 *   iters  int32  [B]              iterations executed so far
 *   alpha  int32  [B]              index into {1,.5,.25,.1,.05,.01} accepted last, -1 = none
 *   expansion / terminal / gains: float, layout depends on the model's "record" kind:
 *     QX_LAYOUT_REC (quadrotor; generic dense records, one contiguous record per (t,b))
 *        expansion float [N][B][REC]   REC = [F n*(n+m) | lx n | lu m | lxx n*n | luu m*m | lux m*n],
 *                                      F = [A | B] interleaved: row i = A[i][0..n) then B[i][0..m)
 *        terminal  float [B][n + n*n]  [Vx | Vxx]
 *        gains     float4 [N][(m + m*n)/4][B]  (b fastest): float4 #0 = k, #1.. = K row-major, four entries each;
 *                  element (t, ch, b) at (((t * 13 + ch / 4) * B + b) * 4 + ch % 4 floats (m + m*n = 52)
 *     QX_LAYOUT_SOA (cart-pole; thread-per-trajectory Riccati, b fastest)
 *        expansion float [N][REC][B], terminal float [n + n*n][B], gains float [N][m + m*n][B]
 */
#ifndef QUATTRO_B200_H
#define QUATTRO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QX_ABI_VERSION 1

#define QX_MODEL_CARTPOLE 0   /* examples/cartpole/cartpole_dynamics.py:48-105, cartpole_mpc.py:187-269 */
#define QX_MODEL_QUADROTOR 1  /* examples/quadrotor/quadrotor_dynamics.py:64-195, quadrotor_mpc.py:40-100 */

#define QX_EULER 0
#define QX_RK4 1

#define QX_LAYOUT_SOA 0
#define QX_LAYOUT_REC 1
#define QX_LAYOUT_BLK 2   /* compact block record of the quadrotor/Euler/diagonal-cost fast path, see qx_linearize_blk */

#define QX_MAX_N 12
#define QX_MAX_M 4
#define QX_NUM_ALPHA 6

#define QX_OK 0
#define QX_EINVAL (-1)   /* bad argument (null pointer, unknown model, size out of range) */
#define QX_ECUDA (-2)    /* a CUDA runtime call failed; see qx_last_error() */
#define QX_EUNSUPPORTED (-3)

/* flags of qx_linesearch */
#define QX_LS_FORCE_ACTIVE 1u   /* fixed-iteration (benchmark) mode: never mark a trajectory finished */

/* Model parameters, refreshed by the host from the MPC object at every optimize() because the
 * reference's callers mutate them between calls ("Config / flags"). */
typedef struct QxParams {
  int32_t model_id;      /* QX_MODEL_* */
  int32_t integrator;    /* QX_EULER | QX_RK4  (discrete_dynamics, cartpole_dynamics.py:94-105) */
  int32_t n, m;
  int32_t diag_cost;     /* 1 when Q, R, Qf are diagonal (fast path; same values) */
  int32_t _pad;
  double dt;
  double phys[8];        /* cart-pole: m_cart, m_pole, length, gravity;
                            quadrotor: mass, Ix, Iy, Iz, arm, gravity, k_yaw */
  double Q[QX_MAX_N * QX_MAX_N];   /* row-major n x n (running_cost, cartpole_mpc.py:255, quadrotor_mpc.py:86-93) */
  double R[QX_MAX_M * QX_MAX_M];
  double Qf[QX_MAX_N * QX_MAX_N];  /* final_cost, cartpole_mpc.py:268, quadrotor_mpc.py:99 */
  double x_ref[QX_MAX_N];
  double barrier_alpha;  /* quadrotor soft barrier weight (quadrotor_mpc.py:45); 0 for cart-pole */
  double barrier_beta;   /* softplus sharpness (quadrotor_mpc.py:46) */
} QxParams;

/* Transformer hyper-parameters (TransformerPredictor.__init__, transformer_model.py:97-116). */
typedef struct QxTfConfig {
  int32_t state_dim;     /* n */
  int32_t control_dim;   /* c = m (1 + n) */
  int32_t d_model, nhead, num_layers, dim_ff;
  int32_t target_len, prompt_len, horizon;   /* tokens: (horizon+1) states | prompt | targets */
  int32_t m;             /* control dimension of the system (for the K/k unpack) */
} QxTfConfig;

int qx_abi_version(void);
const char* qx_last_error(void);

/* Layout queries (pure host functions). */
int qx_layout(int model_id);                       /* QX_LAYOUT_* for the model, or QX_EINVAL */
int qx_record_floats(int model_id);                /* REC */
size_t qx_expansion_bytes(int model_id, int B, int N);
size_t qx_terminal_bytes(int model_id, int B);
size_t qx_gains_bytes(int model_id, int B, int N);

/* simulate + compute_total_cost  (quattro_ilqr_tf.py:127-143).
 * x0 double [n][B]; writes x[N+1][n][B] and cost[B].  status may be NULL (all trajectories). */
int qx_rollout_cost(const QxParams* p, const double* x0, const double* u, double* x, double* cost,
                    const int32_t* status, int B, int N, void* stream);

/* _compute_dynamics_jacobians + _compute_cost_derivatives for t in [t_start, N) and the terminal
 * V_x, V_xx (quattro_ilqr_tf.py:149-275), exact derivatives, one launch for the whole horizon. */
int qx_linearize(const QxParams* p, const double* x, const double* u, float* expansion, float* terminal,
                 const int32_t* status, int B, int N, int t_start, void* stream);

/* backward_pass (t_start = 0, quattro_ilqr_tf.py:282-319) / backward_pass_segment (t_start > 0,
 * :327-366): gains for t in [t_start, N) written at their absolute t. reg = 1e-6 in the reference. */
int qx_riccati(int model_id, const float* expansion, const float* terminal, float* gains,
               const int32_t* status, int B, int N, int t_start, float reg, void* stream);

/* ---- Block fast path of the backward pass: quadrotor + Euler map + diagonal Q, R, Qf (what every driver
 * of the reference runs: quadrotor_sim.py:100, quadrotor_mpc.py:40-46).  There A_t is block bidiagonal in
 * x = [p | v | eta | omega], B_t has two non-zero row blocks, l_xx / l_uu are diagonal and l_ux = 0, so
 * linearise emits 48 floats per (t, b) instead of 416 and the Riccati step works on 3x3 blocks with one
 * thread per trajectory (qx_blk.cu, qx_blk_riccati.cu).  Same results as qx_linearize + qx_riccati (quattro_ilqr_tf.py:182-366).
 *   expansion float [N][ceil(B/8)][12 slots][8 trajectories][4]   (QX_LAYOUT_BLK, 48 floats = 192 B per (t, b)):
 *              only the entries that vary are stored.  With G = dt dv'/deta, Hp = I + dt deta'/deta,
 *              E = dt deta'/domega, Sp = I + dt domega'/domega (G[2][2] = 0; Hp[:,2] = e3, Hp[1][1] = 1;
 *              E[:,0] = dt e1; diag(Sp) = 1) and d = dt/m * thrust direction, the float4 slots are
 *              0: g00 g01 g02 d0 | 1: g10 g11 g12 d1 | 2: g20 g21 d2 h00 | 3: h01 h10 h20 h21 | 4: e01 e02 e11 e12
 *              5: e21 e22 s01 s02 | 6: s10 s12 s20 s21 | 7: l_u | 8: diag(l_uu) | 9-11: l_x
 *   terminal  float [B][16]  = [V_x(12) | pad]   (V_xx(N) = 2 Qf comes from the parameters)
 *   gains     as QX_LAYOUT_REC: float4 [N][13][B], [k | K row-major] in groups of four
 * qx_blk_supported() tells whether the parameter block qualifies; the generic entry points serve the rest. */
int qx_blk_supported(const QxParams* p);
size_t qx_blk_expansion_bytes(int B, int N);
size_t qx_blk_terminal_bytes(int B);
int qx_linearize_blk(const QxParams* p, const double* x, const double* u, float* expansion, float* terminal,
                     const int32_t* status, int B, int N, int t_start, void* stream);
int qx_riccati_blk(const QxParams* p, const float* expansion, const float* terminal, float* gains,
                   const int32_t* status, int B, int N, int t_start, float reg, void* stream);

/* The line search of optimize (quattro_ilqr_tf.py:440-451) with forward_pass (:377-390) inside:
 * first alpha in {1,.5,.25,.1,.05,.01} whose candidate cost <= cost[b]; all in one launch.
 * On return (x_out,u_out,cost) hold the next iterate of every ACTIVE trajectory (the accepted
 * candidate, or an unchanged copy when no alpha is accepted); status/iters/alpha are updated with
 * the reference's stopping rule (:472): finished iff no alpha accepted or |cost - new| < tol.
 * A trajectory that finishes has its final iterate in BOTH (x_in,u_in) and (x_out,u_out), so the
 * caller can ping-pong the buffers.  n_active (device int32, may be NULL) is incremented by the
 * number of trajectories still active afterwards.
 * workspace (qx_linesearch_workspace_bytes(), may be NULL): with it the candidates are evaluated concurrently
 * instead of one alpha after the other - quadrotor (diagonal weights): alpha = 1 for every trajectory (stage A, the
 * candidate goes straight into x_out/u_out), then the five smaller alphas of the rejected ones side by side, costs
 * only (stage B), then the first accepted one rolled out again into x_out/u_out (stage W; a trajectory without an
 * accepted alpha rolls alpha = 0, i.e. its nominal), then the final iterates of the finished trajectories mirrored
 * into x_in/u_in; the workspace holds the queue, two int32 flags, the nominal's cost and five probe costs per
 * trajectory.  Cart-pole: the six alphas side by side (costs only; alpha = 1 also stored), then the winners other
 * than alpha = 1 re-rolled and stored (the workspace holds one int32 hint per trajectory).  Results are bit-identical
 * with and without it.  A candidate whose partial cost already exceeds cost[b] is abandoned early (quadrotor, when
 * Q, R, Qf and the barrier weight are non-negative: the costs are then sums of non-negative terms). */
size_t qx_linesearch_workspace_bytes(int model_id, int B, int N);
int qx_linesearch(const QxParams* p, double* x_in, double* u_in, const float* gains, double* cost,
                  int32_t* status, int32_t* iters, int32_t* alpha, double* x_out, double* u_out,
                  int32_t* n_active, void* workspace, double tol, uint32_t flags, int B, int N, void* stream);

/* Measurement hook (no reference counterpart): four caller-owned cudaEvent_t handles that the staged quadrotor line
 * search records on its stream in front of stage A and behind stages A, B and W, so that bench.py can time the three
 * launches of one qx_linesearch call separately.  Per host thread; count = 0 clears the hook.  Never set it while a
 * stream is being captured into a CUDA graph. */
int qx_linesearch_stage_events(void* const* events, int count);

/* forward_pass with a caller-chosen alpha (quattro_ilqr_tf.py:377-390): candidate trajectory and its
 * total cost for every trajectory, no accept test.  (x_in,u_in) is the nominal, gains as above. */
int qx_forward_pass(const QxParams* p, const double* x_in, const double* u_in, const float* gains, double alpha,
                    double* x_out, double* u_out, double* cost_out, int B, int N, void* stream);

/* ---- Quattro transformer: TransformerILQR.predict for a batch (transformer_ilqr.py:311-325,
 * transformer_model.py:122-138) fused with the prompt pack / prediction unpack of optimize
 * (quattro_ilqr_tf.py:496-518).  Weights are one packed float buffer laid out by
 * qx_tf_weight_offsets(); scratch is qx_tf_workspace_bytes().  Reads x (trajectories) and the
 * last prompt_len gain records, writes gain records for t in [0, target_len). */
size_t qx_tf_weight_floats(const QxTfConfig* cfg);
size_t qx_tf_workspace_bytes(const QxTfConfig* cfg, int B);
/* Tensor-core path: the four dense weights of every layer (and the padded output head) split once into (hi, lo)
 * halves for three-term fp32-accurate products hi hi + hi lo + lo hi - as fp16 halves of 2^8 W (kind::f16, what the
 * forward pass uses) and as TF32 halves (kind::tf32; QX_TF_GEMM_F16=0).  wsplit has qx_tf_split_floats() floats. */
size_t qx_tf_split_floats(const QxTfConfig* cfg);
int qx_tf_prepare(const QxTfConfig* cfg, const float* weights, float* wsplit, void* stream);
/* wsplit == NULL selects the exact-fp32 CUDA-core GEMMs; otherwise the dense layers run on tcgen05 (TMEM accumulators
 * and A operands, TMA-staged weights), the feed-forward block of a layer as one launch whose ff-wide intermediate
 * never leaves the SM (QX_TF_FFN_FUSE=0: two launches, same bits), attention as kind::tf32 products. */
int qx_tf_forward(const QxTfConfig* cfg, const float* weights, const float* wsplit, const double* x,
                  const double* x_shift, float* gains, int model_id, const int32_t* status, void* workspace, int B,
                  void* stream);

/* One dense layer of the model on the tensor cores (nn.Linear [+ ReLU], transformer_model.py:106-116 and
 * the in_proj / out_proj / linear1 / linear2 of nn.TransformerEncoderLayer): C[M][N] = act(A[M][K] W[N][K]^T + bias),
 * row-major fp32, W passed as the (hi, lo) TF32 halves qx_tf_split() produces.  N % 128 == 0, K % 32 == 0,
 * all pointers 128-B aligned.  K == 128 runs the resident-weight kernel, other K the streaming kernel. */
int qx_tf_split(const float* w, float* hi, float* lo, size_t count, void* stream);
int qx_tf_dense(const float* A, const float* w_hi, const float* w_lo, const float* bias, float* C, int M, int N, int K,
                int relu, void* stream);
/* A sub-layer's output GEMM with the residual add and the LayerNorm of the post-norm encoder block in its epilogue
 * (nn.TransformerEncoderLayer, transformer_model.py:112-113: x = norm1(x + out_proj(attn)), x = norm2(x + linear2(..))):
 * C[M][128] = LayerNorm(res[M][128] + A[M][K] W[128][K]^T + bias) * gamma + beta, eps = 1e-5, row-major fp32.
 * K % 128 == 0 (else QX_EUNSUPPORTED); C may be res (in-place update of the residual stream), never A.
 * Bit-identical to qx_tf_dense followed by the separate add + LayerNorm kernel of qx_tf_forward. */
int qx_tf_dense_ln(const float* A, const float* w_hi, const float* w_lo, const float* bias, const float* res,
                   const float* gamma, const float* beta, float* C, int M, int K, void* stream);
/* The causal self-attention core of the encoder layers (nn.MultiheadAttention inside nn.TransformerEncoderLayer,
 * transformer_model.py:112-113; mask :134-135): qkv [B*L][3d] = packed in_proj output, out [B*L][d], 32-wide heads,
 * softmax(Q K^T / sqrt(32) + causal mask) V.  impl 0: tcgen05 (L <= 112; for L <= 64 two trajectories share one
 * M = 128 tile, bit-identical to one per tile), 1: mma.sync (L <= 112), 2: exact fp32. */
int qx_tf_attention(const float* qkv, float* out, int B, int L, int d, int nhead, int impl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QUATTRO_B200_H */
