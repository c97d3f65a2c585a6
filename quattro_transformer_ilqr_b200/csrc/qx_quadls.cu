// Quadrotor rollouts and line search, FOUR lanes per trajectory with the state DISTRIBUTED over the lanes.
//
//   rollout_qd_kernel        = iLQR_TF.simulate + compute_total_cost             (quattro_ilqr_tf.py:127-143)
//   ls_qd_phaseA/B/C_kernel  = the alpha loop of optimize (:440-451) around forward_pass (:377-390)
//                              plus the stopping rule (:472)
//
// Used when Q, R, Qf are diagonal (the reference's QuadrotorMPC, quadrotor_mpc.py:40-46); dense weights
// take the generic thread-per-trajectory kernels of qx_rollout.cu.
//
// x = [p | v | eta | omega]: lane g of a 4-lane group owns block g (three fp64 registers), control g
// (its gain row, its barrier term) and the sincos of angle min(g, 2).  Per step the lanes exchange
//   dx (all-gather, 12 values)        for   du_g = k_g + K[g,:] dx        (:381-382)
//   u  (all-gather, 4 values)         for   thrust and torques
//   the three angles, their sin/cos, 1/cos(theta), and the right neighbour's block (v for p', omega for eta')
// with warp shuffles; each lane integrates and costs only its own block.  Compared with a full state copy
// per lane this halves the fp64 work per step and needs 128 registers instead of ~230, so four warps per
// scheduler hide the latency of the sequential recursion.
//
// The file is compiled with -fmad=false and every fused multiply-add is written out (fma()), so the nominal
// rollout, the line search and its retries evaluate f and L with identical bits for identical inputs - like
// the reference's single Python implementation does; the accept test compares such costs with "<=".
//
// Line search = three stages, so that the first accepted alpha is found without rolling out all six:
//   A   alpha = 1 for every active trajectory, candidate written straight into (x_out, u_out)
//   B1  alpha = .5, .25 of the rejected ones concurrently (queue compacted by A)    -> C1 picks / re-queues
//   B2  alpha = .1, .05, .01 of the still rejected ones                              -> C2 picks / gives up
// Candidates of B are parked in a scratch [plane][row][b]; C copies the winner out with thread-per-trajectory
// (coalesced) row loops.  Queues are appended per CTA in trajectory order, so the eight groups of a B warp
// touch neighbouring columns.
#include "qx_common.cuh"
#include "qx_fp64math.cuh"

namespace qx {

namespace {

constexpr unsigned FULL = 0xffffffffu;

struct QdConst {
  double inv_mass, grav, arm, kyaw, invIx, invIy, invIz, c1, c2, c3, dt, beta, inv_beta, alpha_b;
};

struct QdLane {
  QdConst k;
  double qd[3], qfd[3], xr[3], rd;
  int g, base;
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }

__device__ __forceinline__ void qd_lane_init(const QxParams& p, int lane, QdLane& c) {
  constexpr int n = 12, m = 4;
  c.g = lane & 3;
  c.base = lane & ~3;
  QdConst& k = c.k;
  k.inv_mass = 1.0 / p.phys[0];
  k.invIx = 1.0 / p.phys[1]; k.invIy = 1.0 / p.phys[2]; k.invIz = 1.0 / p.phys[3];
  k.arm = p.phys[4]; k.grav = p.phys[5]; k.kyaw = p.phys[6];
  k.c1 = (p.phys[2] - p.phys[3]) / p.phys[1];
  k.c2 = (p.phys[3] - p.phys[1]) / p.phys[2];
  k.c3 = (p.phys[1] - p.phys[2]) / p.phys[3];
  k.dt = p.dt; k.beta = p.barrier_beta; k.inv_beta = 1.0 / p.barrier_beta; k.alpha_b = p.barrier_alpha;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    c.qd[r] = p.Q[(3 * c.g + r) * n + 3 * c.g + r];
    c.qfd[r] = p.Qf[(3 * c.g + r) * n + 3 * c.g + r];
    c.xr[r] = p.x_ref[3 * c.g + r];
  }
  c.rd = p.R[c.g * m + c.g];
}

// Continuous dynamics of the lane's own block (quadrotor_dynamics.py:64-164) for NA independent states at once;
// every lane of the warp must call it.  Straight-line code: the four block formulas are all evaluated and the
// lane's one selected (divergent branches would cost the same issue slots and stop the compiler from
// interleaving the NA chains).
template <int NA>
__device__ __forceinline__ void qd_fc_n(const double (*xg)[3], const double (*ua)[4], const QdLane& c, double (*dg)[3]) {
  const QdConst& k = c.k;
  const int g = c.g, base = c.base;
  double ang[NA], s[NA], cs[NA];
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    // the angles live in lane 2; lane g < 3 evaluates angle g (lane 3 repeats psi)
    const double a0 = shfl_d(xg[a][0], base + 2), a1 = shfl_d(xg[a][1], base + 2), a2 = shfl_d(xg[a][2], base + 2);
    ang[a] = (g == 0) ? a0 : (g == 1) ? a1 : a2;
  }
  fast_sincos_n<NA>(ang, s, cs);
  const int nb = base + ((g + 1) & 3);       // right neighbour's block: v for lane 0, omega for lane 2
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    const double sec_own = fast_rcp(cs[a]);
    const double n0 = shfl_d(xg[a][0], nb), n1 = shfl_d(xg[a][1], nb), n2 = shfl_d(xg[a][2], nb);
    const double sphi = shfl_d(s[a], base), cphi = shfl_d(cs[a], base);
    const double sth = shfl_d(s[a], base + 1), cth = shfl_d(cs[a], base + 1), sec = shfl_d(sec_own, base + 1);
    const double spsi = shfl_d(s[a], base + 2), cpsi = shfl_d(cs[a], base + 2);
    const double* u = ua[a];
    // block 1: v' = T/m * R e3 - g e3
    const double Tm = ((u[0] + u[1]) + (u[2] + u[3])) * k.inv_mass;
    const double sc = sth * cphi;
    const double v0 = Tm * fma(cpsi, sc, spsi * sphi);
    const double v1 = Tm * fma(-spsi, sc, cpsi * sphi);
    const double v2 = fma(Tm, cth * cphi, -k.grav);
    // block 2: eta' from omega = (n0, n1, n2)
    const double tth = sth * sec;
    const double w = fma(n1, sphi, n2 * cphi);          // q sin(phi) + r cos(phi)
    const double e0 = fma(w, tth, n0);
    const double e1 = fma(n1, cphi, -(n2 * sphi));
    const double e2 = w * sec;
    // block 3: omega' from the own block
    const double pp = xg[a][0], q = xg[a][1], r = xg[a][2];
    const double o0 = fma(k.c1, q * r, k.arm * ((u[1] + u[2]) - (u[0] + u[3])) * k.invIx);
    const double o1 = fma(k.c2, pp * r, k.arm * ((u[0] + u[1]) - (u[2] + u[3])) * k.invIy);
    const double o2 = fma(k.c3, pp * q, k.kyaw * ((u[0] - u[1]) + (u[2] - u[3])) * k.invIz);
    // block 0: p' = v = (n0, n1, n2)
    dg[a][0] = (g == 0) ? n0 : (g == 1) ? v0 : (g == 2) ? e0 : o0;
    dg[a][1] = (g == 0) ? n1 : (g == 1) ? v1 : (g == 2) ? e1 : o1;
    dg[a][2] = (g == 0) ? n2 : (g == 1) ? v2 : (g == 2) ? e2 : o2;
  }
}

// discrete map of the own block (quadrotor_dynamics.py:186-195); the integrator is a template parameter so
// that the Euler kernels do not carry the registers of the RK4 stages
template <int INTEG, int NA>
__device__ __forceinline__ void qd_step_n(double (*xg)[3], const double (*ua)[4], const QdLane& c) {
  const double dt = c.k.dt;
  double k1[NA][3];
  qd_fc_n<NA>(xg, ua, c, k1);
  if (INTEG == QX_EULER) {
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int i = 0; i < 3; ++i) xg[a][i] = fma(dt, k1[a][i], xg[a][i]);
    return;
  }
  double k2[NA][3], k3[NA][3], k4[NA][3], xs[NA][3];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int i = 0; i < 3; ++i) xs[a][i] = fma(0.5 * dt, k1[a][i], xg[a][i]);
  qd_fc_n<NA>(xs, ua, c, k2);
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int i = 0; i < 3; ++i) xs[a][i] = fma(0.5 * dt, k2[a][i], xg[a][i]);
  qd_fc_n<NA>(xs, ua, c, k3);
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int i = 0; i < 3; ++i) xs[a][i] = fma(dt, k3[a][i], xg[a][i]);
  qd_fc_n<NA>(xs, ua, c, k4);
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int i = 0; i < 3; ++i)
      xg[a][i] = fma(dt / 6.0, (k1[a][i] + 2.0 * k2[a][i]) + (2.0 * k3[a][i] + k4[a][i]), xg[a][i]);
}

// The lane's share of L(x, u) for NA candidates: its state block and its control (quadrotor_mpc.py:86-93) with
// the barrier alpha_b * softplus(-u; beta)^2 (:80, :92).  log1p(e) by its series when e is tiny (u well on the
// feasible side), by fast_log1p_unit otherwise - evaluated for all NA only if some lane of the warp needs it.
template <int NA>
__device__ __forceinline__ void qd_running_n(const double (*xg)[3], const double* uh, const QdLane& c, double* cst) {
  const QdConst& k = c.k;
  double e[NA], bz[NA], l1p[NA];
  bool need = false;
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    bz[a] = -k.beta * uh[a];
    e[a] = fast_exp_neg(-fabs(bz[a]));
    l1p[a] = e[a] * fma(-e[a], fma(-e[a], fma(-0.25, e[a], 1.0 / 3.0), 0.5), 1.0);
    need = need || !(e[a] < 1e-4);
  }
  if (__any_sync(FULL, need)) {
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const double full = fast_log1p_unit(e[a]);
      l1p[a] = (e[a] < 1e-4) ? l1p[a] : full;
    }
  }
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double dxv = xg[a][r] - c.xr[r];
      acc = fma(c.qd[r] * dxv, dxv, acc);
    }
    acc = fma(c.rd * uh[a], uh[a], acc);
    const double sp = (fmax(bz[a], 0.0) + l1p[a]) * k.inv_beta;
    cst[a] += acc + k.alpha_b * sp * sp;
  }
}

__device__ __forceinline__ double qd_final(const double* xg, const QdLane& c) {
  double acc = 0.0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double dxv = xg[r] - c.xr[r];
    acc = fma(c.qfd[r] * dxv, dxv, acc);
  }
  return acc;
}

__device__ __forceinline__ double qd_group_sum(double v) {
  v += __shfl_xor_sync(FULL, v, 1);
  v += __shfl_xor_sync(FULL, v, 2);      // identical bits in the four lanes: the butterfly adds commute
  return v;
}

// NA candidate rollouts (forward_pass :377-390, step sizes alpha[0..NA)) of ONE trajectory by a 4-lane group,
// interleaved: the nominal and the gains are loaded once and the NA independent recursions give the scheduler
// NA-fold instruction-level parallelism (a single rollout is a chain of dependent fp64 operations).
// Nominal and gains of trajectory b_src; candidate a is written at column b_dst of (x_dst[a], u_dst[a]) (row
// stride B).  All 32 lanes of the warp must call it; lanes with run == false only take part in the shuffles.
// cst[a] returns the total cost (same bits in the four lanes of the group).
template <int INTEG, int NA>
__device__ __forceinline__ void qd_candidates(const QdLane& c, bool run, const double* __restrict__ x_in,
                                              const double* __restrict__ u_in, const float* __restrict__ gains,
                                              int b_src, const double* alpha, double* const* x_dst, double* const* u_dst,
                                              int b_dst, int B, int N, double* cst) {
  constexpr int n = 12, m = 4, G = 52;
  const int g = c.g, base = c.base;
  // per-lane cursors, advanced by one time step per iteration (no index arithmetic in the loop)
  const size_t rowB = (size_t)B;
  const double* xin = x_in + (size_t)(3 * g) * rowB + b_src;
  const double* uin = u_in + (size_t)g * rowB + b_src;
  const float* rec = gains + (size_t)b_src * G;
  size_t xo = (size_t)(3 * g) * rowB + b_dst, uo = (size_t)g * rowB + b_dst;     // offsets into x_dst[a], u_dst[a]
  double xg[NA][3], xn_nx[3] = {0.0, 0.0, 0.0}, un_nx = 0.0;
  float gk_nx = 0.f;
  float4 gK_nx[3] = {};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double v = run ? xin[(size_t)i * rowB] : 0.0;
#pragma unroll
    for (int a = 0; a < NA; ++a) xg[a][i] = v;
  }
  auto fetch = [&]() {
#pragma unroll
    for (int i = 0; i < 3; ++i) xn_nx[i] = xin[(size_t)i * rowB];
    un_nx = *uin;
    gk_nx = rec[g];
    const float4* row = reinterpret_cast<const float4*>(rec + m + g * n);
    gK_nx[0] = row[0]; gK_nx[1] = row[1]; gK_nx[2] = row[2];
    xin += n * rowB;
    uin += m * rowB;
    rec += rowB * G;
  };
  if (run) fetch();
#pragma unroll
  for (int a = 0; a < NA; ++a) cst[a] = 0.0;
  for (int t = 0; t < N; ++t) {
    const double xn[3] = {xn_nx[0], xn_nx[1], xn_nx[2]};
    const double un = un_nx;
    const double gk = (double)gk_nx;
    const double kr[12] = {(double)gK_nx[0].x, (double)gK_nx[0].y, (double)gK_nx[0].z, (double)gK_nx[0].w,
                           (double)gK_nx[1].x, (double)gK_nx[1].y, (double)gK_nx[1].z, (double)gK_nx[1].w,
                           (double)gK_nx[2].x, (double)gK_nx[2].y, (double)gK_nx[2].z, (double)gK_nx[2].w};
    if (run && t + 1 < N) fetch();
    double uh[NA], ua[NA][4];
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      double dx[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) dx[i] = xg[a][i] - xn[i];
      // du_g = k_g + K[g,:] (x_hat - x): block q of dx comes from lane q
      double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int i = 0; i < 3; ++i) acc[i] = fma(kr[3 * q + i], shfl_d(dx[i], base + q), acc[i]);
      const double kd = (acc[0] + acc[1]) + acc[2];
      uh[a] = fma(alpha[a], gk + kd, un);                   // alpha scales the feedback term too (:382-383)
#pragma unroll
      for (int q = 0; q < 4; ++q) ua[a][q] = shfl_d(uh[a], base + q);
    }
    if (run) {
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        u_dst[a][uo] = uh[a];
#pragma unroll
        for (int i = 0; i < 3; ++i) x_dst[a][xo + (size_t)i * rowB] = xg[a][i];
      }
    }
    qd_running_n<NA>(xg, uh, c, cst);       // lanes with run == false accumulate a cost nobody reads
    xo += n * rowB;
    uo += m * rowB;
    qd_step_n<INTEG, NA>(xg, ua, c);
  }
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    if (run) {
#pragma unroll
      for (int i = 0; i < 3; ++i) x_dst[a][xo + (size_t)i * rowB] = xg[a][i];
      cst[a] += qd_final(xg[a], c);
    }
    cst[a] = qd_group_sum(cst[a]);
  }
}

template <int INTEG>
__device__ __forceinline__ double qd_candidate(const QdLane& c, bool run, const double* __restrict__ x_in,
                                               const double* __restrict__ u_in, const float* __restrict__ gains,
                                               int b_src, double alpha, double* __restrict__ x_dst,
                                               double* __restrict__ u_dst, int b_dst, int B, int N) {
  double cst[1];
  const double al[1] = {alpha};
  double* const xd[1] = {x_dst};
  double* const ud[1] = {u_dst};
  qd_candidates<INTEG, 1>(c, run, x_in, u_in, gains, b_src, al, xd, ud, b_dst, B, N, cst);
  return cst[0];
}

// Append the flagged groups of this CTA to a queue, in trajectory order.  Must be reached by all threads.
__device__ __forceinline__ void qd_ordered_append(bool flag_lane, int value, int32_t* __restrict__ queue,
                                                  int32_t* __restrict__ counter) {
  __shared__ int s_cnt[32];
  __shared__ int s_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = (blockDim.x + 31) >> 5;
  const unsigned bal = __ballot_sync(FULL, flag_lane);
  if (lane == 0) s_cnt[warp] = __popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < nwarps; ++w) { const int cw = s_cnt[w]; s_cnt[w] = tot; tot += cw; }
    s_base = tot ? atomicAdd(counter, tot) : 0;
  }
  __syncthreads();
  if (flag_lane) queue[s_base + s_cnt[warp] + __popc(bal & ((1u << lane) - 1u))] = value;
}

}  // namespace

// Workspace of the staged search (device memory, qx_linesearch_workspace_bytes()).
struct QdWork {
  int32_t* ctl;        // [0] = |queue1|, [1] = |queue2|
  int32_t *queue1, *queue2, *pend1, *pend2;
  double* cur;         // cost of the nominal, snapshot taken by phase A
  double* cand;        // [5][B] candidate costs, plane = alpha index - 1
  double *xs, *us;     // [5][(N+1) n][B], [5][N m][B]
};

__host__ __device__ inline size_t qd_align16(size_t v) { return (v + 15) / 16 * 16; }

size_t qd_workspace_bytes(int B, int N) {
  constexpr size_t n = 12, m = 4, A = QX_NUM_ALPHA - 1;
  return 32 + 4 * qd_align16((size_t)B * 4) + (size_t)B * 8 + A * (size_t)B * 8 +
         A * ((size_t)(N + 1) * n + (size_t)N * m) * (size_t)B * 8;
}

static QdWork qd_carve(void* workspace, int B, int N) {
  constexpr size_t n = 12, m = 4, A = QX_NUM_ALPHA - 1;
  uint8_t* w = static_cast<uint8_t*>(workspace);
  const size_t qb = qd_align16((size_t)B * 4);
  QdWork k;
  k.ctl = reinterpret_cast<int32_t*>(w);
  k.queue1 = reinterpret_cast<int32_t*>(w + 32);
  k.queue2 = reinterpret_cast<int32_t*>(w + 32 + qb);
  k.pend1 = reinterpret_cast<int32_t*>(w + 32 + 2 * qb);
  k.pend2 = reinterpret_cast<int32_t*>(w + 32 + 3 * qb);
  k.cur = reinterpret_cast<double*>(w + 32 + 4 * qb);
  k.cand = k.cur + B;
  k.xs = k.cand + A * B;
  k.us = k.xs + A * (size_t)(N + 1) * n * B;
  return k;
}

// ------------------------------------------------------------------------------------------------
template <int INTEG>
__global__ void __launch_bounds__(128)
rollout_qd_kernel(const __grid_constant__ QxParams p, const double* __restrict__ x0, const double* __restrict__ u,
                  double* __restrict__ x, double* __restrict__ cost, const int32_t* __restrict__ status, int B, int N) {
  constexpr int n = 12, m = 4;
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const bool run = (b < B) && (status == nullptr || status[b] == 0);
  if (!__any_sync(FULL, run)) return;
  QdLane c;
  qd_lane_init(p, lane, c);
  const int g = c.g, base = c.base;
  double xg[1][3], cst[1] = {0.0};
#pragma unroll
  for (int i = 0; i < 3; ++i) xg[0][i] = run ? x0[(size_t)(3 * g + i) * B + b] : 0.0;
  double u_nx = run ? u[(size_t)g * B + b] : 0.0;
  for (int t = 0; t < N; ++t) {
    const double uh[1] = {u_nx};
    if (run && t + 1 < N) u_nx = u[((size_t)(t + 1) * m + g) * B + b];
    double ua[1][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) ua[0][q] = shfl_d(uh[0], base + q);
    if (run) {
#pragma unroll
      for (int i = 0; i < 3; ++i) x[((size_t)t * n + 3 * g + i) * B + b] = xg[0][i];
    }
    qd_running_n<1>(xg, uh, c, cst);
    qd_step_n<INTEG, 1>(xg, ua, c);
  }
  if (run) {
#pragma unroll
    for (int i = 0; i < 3; ++i) x[((size_t)N * n + 3 * g + i) * B + b] = xg[0][i];
    cst[0] += qd_final(xg[0], c);
  }
  const double tot = qd_group_sum(cst[0]);
  if (run && g == 0) cost[b] = tot;
}

// Bookkeeping after the line search of one trajectory (stopping rule :472), executed by its 4-lane group.
// Leaves the next iterate in (x_out,u_out); a finished trajectory has it in both buffers.
__device__ __forceinline__ bool qd_finalize4(int g, int b, int idx, double cur, double new_cost, double tol,
                                             uint32_t flags, double* __restrict__ x_in, double* __restrict__ u_in,
                                             double* __restrict__ x_out, double* __restrict__ u_out,
                                             double* __restrict__ cost, int32_t* __restrict__ status,
                                             int32_t* __restrict__ iters, int32_t* __restrict__ alpha_out, int B, int N) {
  constexpr int n = 12, m = 4;
  const bool found = idx >= 0;
  bool done = (!found) || (fabs(cur - new_cost) < tol);
  if (flags & QX_LS_FORCE_ACTIVE) done = false;
  if (!found || done) {
    // no alpha: (x_out,u_out) <- nominal;  finished with an update: (x_in,u_in) <- final iterate
    const double* xs = found ? x_out : x_in;
    const double* us = found ? u_out : u_in;
    double* xd = found ? x_in : x_out;
    double* ud = found ? u_in : u_out;
    for (int t = 0; t <= N; ++t)
#pragma unroll
      for (int r = 0; r < 3; ++r) xd[((size_t)t * n + 3 * g + r) * B + b] = xs[((size_t)t * n + 3 * g + r) * B + b];
    for (int t = 0; t < N; ++t) ud[((size_t)t * m + g) * B + b] = us[((size_t)t * m + g) * B + b];
  }
  if (g == 0) {
    cost[b] = new_cost;
    alpha_out[b] = idx;
    iters[b] += 1;
    if (done) status[b] = 1;
  }
  return !done;
}

// MODE 0: the whole alpha list sequentially with early exit (single launch, no workspace).
// MODE 1: phase A of the staged search: alpha = 1 only; rejected trajectories are queued.
template <int MODE, int INTEG>
__global__ void __launch_bounds__(128)
ls_qd_phaseA_kernel(const __grid_constant__ QxParams p, double* __restrict__ x_in, double* __restrict__ u_in,
                    const float* __restrict__ gains, double* __restrict__ cost, int32_t* __restrict__ status,
                    int32_t* __restrict__ iters, int32_t* __restrict__ alpha_out, double* __restrict__ x_out,
                    double* __restrict__ u_out, int32_t* __restrict__ n_active, QdWork ws, double tol, uint32_t flags,
                    int B, int N) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const bool mine = (b < B) && (status[b] == 0);
  QdLane c;
  qd_lane_init(p, lane, c);
  const double cur = mine ? cost[b] : 0.0;
  bool accepted = !mine;
  int idx = -1;
  double new_cost = cur;
  if (__any_sync(FULL, mine)) {
    constexpr int n_alpha = (MODE == 0) ? QX_NUM_ALPHA : 1;
    for (int a = 0; a < n_alpha; ++a) {
      if (!__any_sync(FULL, !accepted)) break;
      const bool run = !accepted;
      const double cst = qd_candidate<INTEG>(c, run, x_in, u_in, gains, b, kAlphas[a], x_out, u_out, b, B, N);
      if (run && cst <= cur) {       // first accepted alpha, "<=" (quattro_ilqr_tf.py:444)
        accepted = true;
        idx = a;
        new_cost = cst;
      }
    }
  }
  bool still_active = false;
  const bool pending = (MODE == 1) && mine && idx < 0;
  if (mine && !pending)
    still_active = qd_finalize4(c.g, b, idx, cur, new_cost, tol, flags, x_in, u_in, x_out, u_out, cost, status, iters,
                                alpha_out, B, N) && c.g == 0;
  if (MODE == 1) {
    if (c.g == 0 && b < B) {
      ws.pend1[b] = pending ? 1 : 0;
      if (pending) ws.cur[b] = cur;
    }
    qd_ordered_append(pending && c.g == 0, b, ws.queue1, &ws.ctl[0]);
  }
  const unsigned act = __ballot_sync(FULL, still_active);
  if (lane == 0 && n_active != nullptr && act != 0u) atomicAdd(n_active, __popc(act));
}

// Phase B: one 4-lane work item per queued trajectory; the NA alphas of the stage are rolled out interleaved.
// One-warp CTAs, at most 168 registers: the register file is 4 x 16 K per SM (one slice per scheduler), so 168
// registers x 32 lanes lets three warps share a scheduler where 178 would leave two.
template <int INTEG, int NA>
__global__ void __launch_bounds__(32, 12)
ls_qd_phaseB_kernel(const __grid_constant__ QxParams p, const double* __restrict__ x_in, const double* __restrict__ u_in,
                    const float* __restrict__ gains, const int32_t* __restrict__ queue, const int32_t* __restrict__ nq_ptr,
                    QdWork ws, int a_first, int B, int N) {
  constexpr int n = 12, m = 4;
  const int lane = threadIdx.x & 31;
  const int nq = *nq_ptr;
  const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const bool run = item < nq;
  if (!__any_sync(FULL, run)) return;
  QdLane c;
  qd_lane_init(p, lane, c);
  const int b = run ? queue[item] : 0;
  double al[NA], cst[NA];
  double* xd[NA];
  double* ud[NA];
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    const int plane = a_first + a - 1;
    al[a] = kAlphas[a_first + a];
    xd[a] = ws.xs + (size_t)plane * (N + 1) * n * B;
    ud[a] = ws.us + (size_t)plane * N * m * B;
  }
  qd_candidates<INTEG, NA>(c, run, x_in, u_in, gains, b, al, xd, ud, b, B, N, cst);
  if (run && c.g == 0) {
#pragma unroll
    for (int a = 0; a < NA; ++a) ws.cand[(size_t)(a_first + a - 1) * B + b] = cst[a];
  }
}

// Phase C: one thread per trajectory (coalesced rows).  A pending trajectory takes the first accepted alpha of
// the stage: its candidate is copied out of the scratch and the stopping rule applied; otherwise it is queued
// for the next stage, or (FINAL) keeps its nominal and finishes.  blockIdx.y splits the rows of the copy;
// slice 0 does the bookkeeping.
template <bool FINAL>
__global__ void __launch_bounds__(128)
ls_qd_phaseC_kernel(double* __restrict__ x_in, double* __restrict__ u_in, double* __restrict__ cost,
                    int32_t* __restrict__ status, int32_t* __restrict__ iters, int32_t* __restrict__ alpha_out,
                    double* __restrict__ x_out, double* __restrict__ u_out, int32_t* __restrict__ n_active, QdWork ws,
                    const int32_t* __restrict__ pend_in, int32_t* __restrict__ pend_out, int32_t* __restrict__ queue_out,
                    int32_t* __restrict__ nq_out, int a_first, int n_alpha, double tol, uint32_t flags, int B, int N) {
  constexpr int n = 12, m = 4;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool mine = (b < B) && (pend_in[b] != 0);
  const bool book = blockIdx.y == 0;
  int idx = -1;
  double cur = 0.0, new_cost = 0.0;
  bool done = false;
  if (mine) {
    cur = ws.cur[b];
    new_cost = cur;
    for (int a = 0; a < n_alpha; ++a) {
      const double cc = ws.cand[(size_t)(a_first + a - 1) * B + b];
      if (cc <= cur) { idx = a_first + a; new_cost = cc; break; }
    }
    const bool found = idx >= 0;
    done = ((!found) || (fabs(cur - new_cost) < tol)) && !(flags & QX_LS_FORCE_ACTIVE);
    // rows [r0, r1) of the concatenated [x rows | u rows] space belong to this slice
    const int xrows = (N + 1) * n, rows = xrows + N * m;
    const int per = (rows + (int)gridDim.y - 1) / (int)gridDim.y;
    const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    // one source / up to two destinations per row, loads batched eight deep to keep HBM busy
    const double* xsrc = nullptr;
    const double* usrc = nullptr;
    double *xd0 = nullptr, *ud0 = nullptr, *xd1 = nullptr, *ud1 = nullptr;
    if (found) {
      xsrc = ws.xs + (size_t)(idx - 1) * xrows * B;
      usrc = ws.us + (size_t)(idx - 1) * (N * m) * B;
      xd0 = x_out; ud0 = u_out;
      if (done) { xd1 = x_in; ud1 = u_in; }     // finished: the final iterate also goes to the other buffer
    } else if (FINAL) {                          // no alpha accepted: keep the nominal
      xsrc = x_in; usrc = u_in;
      xd0 = x_out; ud0 = u_out;
    }
    if (xsrc != nullptr) {
      constexpr int U = 8;
      for (int r = r0; r < r1; r += U) {
        double v[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const int rr = r + q;
          if (rr < r1) v[q] = (rr < xrows) ? xsrc[(size_t)rr * B + b] : usrc[(size_t)(rr - xrows) * B + b];
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const int rr = r + q;
          if (rr < r1) {
            if (rr < xrows) {
              xd0[(size_t)rr * B + b] = v[q];
              if (xd1 != nullptr) xd1[(size_t)rr * B + b] = v[q];
            } else {
              ud0[(size_t)(rr - xrows) * B + b] = v[q];
              if (ud1 != nullptr) ud1[(size_t)(rr - xrows) * B + b] = v[q];
            }
          }
        }
      }
    }
  }
  if (!book) return;
  const bool settle = mine && (idx >= 0 || FINAL);
  const bool requeue = mine && !settle;
  if (settle) {
    cost[b] = new_cost;
    alpha_out[b] = idx;
    iters[b] += 1;
    if (done) status[b] = 1;
  }
  if (!FINAL) {
    if (b < B) pend_out[b] = requeue ? 1 : 0;
    qd_ordered_append(requeue, b, queue_out, nq_out);
  }
  const unsigned act = __ballot_sync(FULL, settle && !done);
  if ((threadIdx.x & 31) == 0 && n_active != nullptr && act != 0u) atomicAdd(n_active, __popc(act));
}

// forward_pass with a caller-chosen alpha (quattro_ilqr_tf.py:377-390): the same candidate arithmetic as the
// line search, so qx_forward_pass(alpha) reproduces the candidate the search evaluated bit for bit
template <int INTEG>
__global__ void __launch_bounds__(128)
forward_pass_qd_kernel(const __grid_constant__ QxParams p, const double* __restrict__ x_in, const double* __restrict__ u_in,
                       const float* __restrict__ gains, double alpha, double* __restrict__ x_out, double* __restrict__ u_out,
                       double* __restrict__ cost_out, int B, int N) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const bool run = b < B;
  if (!__any_sync(FULL, run)) return;
  QdLane c;
  qd_lane_init(p, lane, c);
  const double cst = qd_candidate<INTEG>(c, run, x_in, u_in, gains, run ? b : 0, alpha, x_out, u_out, run ? b : 0, B, N);
  if (run && c.g == 0) cost_out[b] = cst;
}

int qd_launch_forward_pass(const QxParams* p, const double* x_in, const double* u_in, const float* gains, double alpha,
                           double* x_out, double* u_out, double* cost_out, int B, int N, cudaStream_t st) {
  const int grid = ceil_div(4 * B, 128);
  if (p->integrator == QX_EULER)
    forward_pass_qd_kernel<QX_EULER><<<grid, 128, 0, st>>>(*p, x_in, u_in, gains, alpha, x_out, u_out, cost_out, B, N);
  else
    forward_pass_qd_kernel<QX_RK4><<<grid, 128, 0, st>>>(*p, x_in, u_in, gains, alpha, x_out, u_out, cost_out, B, N);
  QX_CHECK_LAUNCH("forward_pass_qd_kernel");
  return QX_OK;
}

// ---- host side: called by qx_rollout_cost / qx_linesearch (qx_rollout.cu) for the quadrotor with diagonal weights
int qd_launch_rollout(const QxParams* p, const double* x0, const double* u, double* x, double* cost,
                      const int32_t* status, int B, int N, cudaStream_t st) {
  if (p->integrator == QX_EULER)
    rollout_qd_kernel<QX_EULER><<<ceil_div(4 * B, 128), 128, 0, st>>>(*p, x0, u, x, cost, status, B, N);
  else
    rollout_qd_kernel<QX_RK4><<<ceil_div(4 * B, 128), 128, 0, st>>>(*p, x0, u, x, cost, status, B, N);
  QX_CHECK_LAUNCH("rollout_qd_kernel");
  return QX_OK;
}

template <int INTEG>
static int qd_linesearch_t(const QxParams* p, double* x_in, double* u_in, const float* gains, double* cost,
                           int32_t* status, int32_t* iters, int32_t* alpha, double* x_out, double* u_out,
                           int32_t* n_active, void* workspace, double tol, uint32_t flags, int B, int N,
                           cudaStream_t st) {
  const int grid = ceil_div(4 * B, 128);
  if (workspace == nullptr) {   // single launch, alphas tried one after the other
    ls_qd_phaseA_kernel<0, INTEG><<<grid, 128, 0, st>>>(*p, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out,
                                                        n_active, QdWork{}, tol, flags, B, N);
    QX_CHECK_LAUNCH("ls_qd_phaseA_kernel<0>");
    return QX_OK;
  }
  QX_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "workspace must be 16-byte aligned");
  const QdWork ws = qd_carve(workspace, B, N);
  cudaError_t e = cudaMemsetAsync(ws.ctl, 0, 32, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(ctl)");
  ls_qd_phaseA_kernel<1, INTEG><<<grid, 128, 0, st>>>(*p, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out,
                                                      n_active, ws, tol, flags, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseA_kernel<1>");
  const dim3 gridC(ceil_div(B, 128), 16);
  // stage 1: alpha = .5, .25
  ls_qd_phaseB_kernel<INTEG, 2><<<ceil_div(4 * B, 32), 32, 0, st>>>(*p, x_in, u_in, gains, ws.queue1, &ws.ctl[0], ws, 1, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseB_kernel<1>");
  ls_qd_phaseC_kernel<false><<<gridC, 128, 0, st>>>(x_in, u_in, cost, status, iters, alpha, x_out, u_out, n_active, ws,
                                                    ws.pend1, ws.pend2, ws.queue2, &ws.ctl[1], 1, 2, tol, flags, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseC_kernel<false>");
  // stage 2: alpha = .1, .05, .01
  ls_qd_phaseB_kernel<INTEG, 3><<<ceil_div(4 * B, 32), 32, 0, st>>>(*p, x_in, u_in, gains, ws.queue2, &ws.ctl[1], ws, 3, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseB_kernel<2>");
  ls_qd_phaseC_kernel<true><<<gridC, 128, 0, st>>>(x_in, u_in, cost, status, iters, alpha, x_out, u_out, n_active, ws,
                                                   ws.pend2, nullptr, nullptr, nullptr, 3, 3, tol, flags, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseC_kernel<true>");
  return QX_OK;
}

int qd_launch_linesearch(const QxParams* p, double* x_in, double* u_in, const float* gains, double* cost,
                         int32_t* status, int32_t* iters, int32_t* alpha, double* x_out, double* u_out,
                         int32_t* n_active, void* workspace, double tol, uint32_t flags, int B, int N,
                         cudaStream_t st) {
  if (p->integrator == QX_EULER)
    return qd_linesearch_t<QX_EULER>(p, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out, n_active, workspace,
                                     tol, flags, B, N, st);
  return qd_linesearch_t<QX_RK4>(p, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out, n_active, workspace, tol,
                                 flags, B, N, st);
}

}  // namespace qx
