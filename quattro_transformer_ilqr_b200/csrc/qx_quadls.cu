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
// per lane this halves the fp64 work per step and needs ~100 registers instead of ~230, so several warps
// per scheduler hide the latency of the sequential recursion.
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

namespace qx {

namespace {

constexpr unsigned FULL = 0xffffffffu;

struct QdConst {
  double inv_mass, grav, arm, kyaw, invIx, invIy, invIz, c1, c2, c3, dt, beta, inv_beta, alpha_b;
  int integrator;
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
  k.integrator = p.integrator;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    c.qd[r] = p.Q[(3 * c.g + r) * n + 3 * c.g + r];
    c.qfd[r] = p.Qf[(3 * c.g + r) * n + 3 * c.g + r];
    c.xr[r] = p.x_ref[3 * c.g + r];
  }
  c.rd = p.R[c.g * m + c.g];
}

// continuous dynamics of the lane's own block (quadrotor_dynamics.py:64-164); every lane of the warp must call it
__device__ __forceinline__ void qd_fc(const double* xg, const double* ua, const QdLane& c, double* dg) {
  const QdConst& k = c.k;
  const int g = c.g, base = c.base;
  // the angles live in lane 2; lane g < 3 evaluates angle g (lane 3 repeats psi)
  const double a0 = shfl_d(xg[0], base + 2), a1 = shfl_d(xg[1], base + 2), a2 = shfl_d(xg[2], base + 2);
  const double ang = (g == 0) ? a0 : (g == 1) ? a1 : a2;
  double s, cs;
  ::sincos(ang, &s, &cs);
  const double sec_own = 1.0 / cs;
  // right neighbour's block: v for lane 0, omega for lane 2
  const int nb = base + ((g + 1) & 3);
  const double n0 = shfl_d(xg[0], nb), n1 = shfl_d(xg[1], nb), n2 = shfl_d(xg[2], nb);
  const double sphi = shfl_d(s, base), cphi = shfl_d(cs, base);
  const double sth = shfl_d(s, base + 1), cth = shfl_d(cs, base + 1), sec = shfl_d(sec_own, base + 1);
  const double spsi = shfl_d(s, base + 2), cpsi = shfl_d(cs, base + 2);
  if (g == 0) {
    dg[0] = n0; dg[1] = n1; dg[2] = n2;
  } else if (g == 1) {
    const double Tm = ((ua[0] + ua[1]) + (ua[2] + ua[3])) * k.inv_mass;
    const double sc = sth * cphi;
    dg[0] = Tm * fma(cpsi, sc, spsi * sphi);
    dg[1] = Tm * fma(-spsi, sc, cpsi * sphi);
    dg[2] = fma(Tm, cth * cphi, -k.grav);
  } else if (g == 2) {
    const double tth = sth * sec;
    const double w = fma(n1, sphi, n2 * cphi);          // q sin(phi) + r cos(phi)
    dg[0] = fma(w, tth, n0);
    dg[1] = fma(n1, cphi, -(n2 * sphi));
    dg[2] = w * sec;
  } else {
    const double pp = xg[0], q = xg[1], r = xg[2];
    dg[0] = fma(k.c1, q * r, k.arm * ((ua[1] + ua[2]) - (ua[0] + ua[3])) * k.invIx);
    dg[1] = fma(k.c2, pp * r, k.arm * ((ua[0] + ua[1]) - (ua[2] + ua[3])) * k.invIy);
    dg[2] = fma(k.c3, pp * q, k.kyaw * ((ua[0] - ua[1]) + (ua[2] - ua[3])) * k.invIz);
  }
}

// discrete map of the own block (quadrotor_dynamics.py:186-195)
__device__ __forceinline__ void qd_step(double* xg, const double* ua, const QdLane& c) {
  const double dt = c.k.dt;
  double k1[3];
  qd_fc(xg, ua, c, k1);
  if (c.k.integrator == QX_EULER) {
#pragma unroll
    for (int i = 0; i < 3; ++i) xg[i] = fma(dt, k1[i], xg[i]);
    return;
  }
  double k2[3], k3[3], k4[3], xs[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) xs[i] = fma(0.5 * dt, k1[i], xg[i]);
  qd_fc(xs, ua, c, k2);
#pragma unroll
  for (int i = 0; i < 3; ++i) xs[i] = fma(0.5 * dt, k2[i], xg[i]);
  qd_fc(xs, ua, c, k3);
#pragma unroll
  for (int i = 0; i < 3; ++i) xs[i] = fma(dt, k3[i], xg[i]);
  qd_fc(xs, ua, c, k4);
#pragma unroll
  for (int i = 0; i < 3; ++i) xg[i] = fma(dt / 6.0, (k1[i] + 2.0 * k2[i]) + (2.0 * k3[i] + k4[i]), xg[i]);
}

// alpha_b * softplus(-u; beta)^2 (quadrotor_mpc.py:80, :92); log1p(e) by its series when e is tiny
__device__ __forceinline__ double qd_barrier(double u, const QdConst& k) {
  const double bz = -k.beta * u;
  const double e = ::exp(-fabs(bz));
  const double l1p = (e < 1e-4) ? e * fma(-e, fma(-e, fma(-0.25, e, 1.0 / 3.0), 0.5), 1.0) : ::log1p(e);
  const double sp = (fmax(bz, 0.0) + l1p) * k.inv_beta;
  return k.alpha_b * sp * sp;
}

// the lane's share of L(x, u): its state block and its control (quadrotor_mpc.py:86-93)
__device__ __forceinline__ double qd_running(const double* xg, double uh, const QdLane& c) {
  double acc = 0.0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double dxv = xg[r] - c.xr[r];
    acc = fma(c.qd[r] * dxv, dxv, acc);
  }
  acc = fma(c.rd * uh, uh, acc);
  return acc + qd_barrier(uh, c.k);
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

// One candidate rollout (forward_pass :377-390) by a 4-lane group: nominal and gains of trajectory b_src,
// candidate written at column b_dst of (x_dst, u_dst) (row stride B).  All 32 lanes of the warp must call it;
// lanes with run == false only take part in the shuffles.  Returns the total cost (same bits in the group).
__device__ __forceinline__ double qd_candidate(const QdLane& c, bool run, const double* __restrict__ x_in,
                                               const double* __restrict__ u_in, const float* __restrict__ gains,
                                               int b_src, double alpha, double* __restrict__ x_dst,
                                               double* __restrict__ u_dst, int b_dst, int B, int N) {
  constexpr int n = 12, m = 4, G = 52;
  const int g = c.g, base = c.base;
  double xg[3], xn_nx[3] = {0.0, 0.0, 0.0}, un_nx = 0.0;
  float gk_nx = 0.f;
  float4 gK_nx[3] = {};
#pragma unroll
  for (int i = 0; i < 3; ++i) xg[i] = run ? x_in[(size_t)(3 * g + i) * B + b_src] : 0.0;
  auto fetch = [&](int t) {
#pragma unroll
    for (int i = 0; i < 3; ++i) xn_nx[i] = x_in[((size_t)t * n + 3 * g + i) * B + b_src];
    un_nx = u_in[((size_t)t * m + g) * B + b_src];
    const float* rec = gains + ((size_t)t * B + b_src) * G;
    gk_nx = rec[g];
    const float4* row = reinterpret_cast<const float4*>(rec + m + g * n);
    gK_nx[0] = row[0]; gK_nx[1] = row[1]; gK_nx[2] = row[2];
  };
  if (run) fetch(0);
  double cst = 0.0;
  for (int t = 0; t < N; ++t) {
    double dx[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) dx[i] = xg[i] - xn_nx[i];
    const double un = un_nx;
    const double gk = (double)gk_nx;
    const float kr[12] = {gK_nx[0].x, gK_nx[0].y, gK_nx[0].z, gK_nx[0].w, gK_nx[1].x, gK_nx[1].y,
                          gK_nx[1].z, gK_nx[1].w, gK_nx[2].x, gK_nx[2].y, gK_nx[2].z, gK_nx[2].w};
    if (run && t + 1 < N) fetch(t + 1);
    // du_g = k_g + K[g,:] (x_hat - x): block q of dx comes from lane q
    double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < 3; ++i) acc[i] = fma((double)kr[3 * q + i], shfl_d(dx[i], base + q), acc[i]);
    const double kd = (acc[0] + acc[1]) + acc[2];
    const double uh = fma(alpha, gk + kd, un);              // alpha scales the feedback term too (:382-383)
    double ua[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) ua[q] = shfl_d(uh, base + q);
    if (run) {
      u_dst[((size_t)t * m + g) * B + b_dst] = uh;
#pragma unroll
      for (int i = 0; i < 3; ++i) x_dst[((size_t)t * n + 3 * g + i) * B + b_dst] = xg[i];
      cst += qd_running(xg, uh, c);
    }
    qd_step(xg, ua, c);
  }
  if (run) {
#pragma unroll
    for (int i = 0; i < 3; ++i) x_dst[((size_t)N * n + 3 * g + i) * B + b_dst] = xg[i];
    cst += qd_final(xg, c);
  }
  return qd_group_sum(cst);
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
  double xg[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) xg[i] = run ? x0[(size_t)(3 * g + i) * B + b] : 0.0;
  double cst = 0.0;
  double u_nx = run ? u[(size_t)g * B + b] : 0.0;
  for (int t = 0; t < N; ++t) {
    const double uh = u_nx;
    if (run && t + 1 < N) u_nx = u[((size_t)(t + 1) * m + g) * B + b];
    double ua[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) ua[q] = shfl_d(uh, base + q);
    if (run) {
#pragma unroll
      for (int i = 0; i < 3; ++i) x[((size_t)t * n + 3 * g + i) * B + b] = xg[i];
      cst += qd_running(xg, uh, c);
    }
    qd_step(xg, ua, c);
  }
  if (run) {
#pragma unroll
    for (int i = 0; i < 3; ++i) x[((size_t)N * n + 3 * g + i) * B + b] = xg[i];
    cst += qd_final(xg, c);
  }
  cst = qd_group_sum(cst);
  if (run && g == 0) cost[b] = cst;
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
template <int MODE>
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
      const double cst = qd_candidate(c, run, x_in, u_in, gains, b, kAlphas[a], x_out, u_out, b, B, N);
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

// Phase B: every queued trajectory x every alpha of the stage is one 4-lane work item (alpha-major, so the
// eight groups of a warp share the alpha and work on neighbouring trajectories).
__global__ void __launch_bounds__(128)
ls_qd_phaseB_kernel(const __grid_constant__ QxParams p, const double* __restrict__ x_in, const double* __restrict__ u_in,
                    const float* __restrict__ gains, const int32_t* __restrict__ queue, const int32_t* __restrict__ nq_ptr,
                    QdWork ws, int a_first, int n_alpha, int B, int N) {
  constexpr int n = 12, m = 4;
  const int lane = threadIdx.x & 31;
  const int nq = *nq_ptr;
  const long item = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const bool run = item < (long)n_alpha * nq;
  if (!__any_sync(FULL, run)) return;
  QdLane c;
  qd_lane_init(p, lane, c);
  const int a = run ? (int)(item / nq) : 0;
  const int b = run ? queue[item % nq] : 0;
  const int plane = a_first + a - 1;
  double* xd = ws.xs + (size_t)plane * (N + 1) * n * B;
  double* ud = ws.us + (size_t)plane * N * m * B;
  const double cst = qd_candidate(c, run, x_in, u_in, gains, b, kAlphas[a_first + a], xd, ud, b, B, N);
  if (run && c.g == 0) ws.cand[(size_t)plane * B + b] = cst;
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
    if (found) {
      const double* xsrc = ws.xs + (size_t)(idx - 1) * xrows * B;
      const double* usrc = ws.us + (size_t)(idx - 1) * (N * m) * B;
#pragma unroll 4
      for (int r = r0; r < r1; ++r) {
        const double v = (r < xrows) ? xsrc[(size_t)r * B + b] : usrc[(size_t)(r - xrows) * B + b];
        double* d0 = (r < xrows) ? &x_out[(size_t)r * B + b] : &u_out[(size_t)(r - xrows) * B + b];
        *d0 = v;
        if (done) {        // finished: the final iterate also goes to the other buffer
          double* d1 = (r < xrows) ? &x_in[(size_t)r * B + b] : &u_in[(size_t)(r - xrows) * B + b];
          *d1 = v;
        }
      }
    } else if (FINAL) {    // no alpha accepted: keep the nominal
#pragma unroll 4
      for (int r = r0; r < r1; ++r) {
        if (r < xrows) x_out[(size_t)r * B + b] = x_in[(size_t)r * B + b];
        else u_out[(size_t)(r - xrows) * B + b] = u_in[(size_t)(r - xrows) * B + b];
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

// ---- host side: called by qx_rollout_cost / qx_linesearch (qx_rollout.cu) for the quadrotor with diagonal weights
int qd_launch_rollout(const QxParams* p, const double* x0, const double* u, double* x, double* cost,
                      const int32_t* status, int B, int N, cudaStream_t st) {
  rollout_qd_kernel<<<ceil_div(4 * B, 128), 128, 0, st>>>(*p, x0, u, x, cost, status, B, N);
  QX_CHECK_LAUNCH("rollout_qd_kernel");
  return QX_OK;
}

int qd_launch_linesearch(const QxParams* p, double* x_in, double* u_in, const float* gains, double* cost,
                         int32_t* status, int32_t* iters, int32_t* alpha, double* x_out, double* u_out,
                         int32_t* n_active, void* workspace, double tol, uint32_t flags, int B, int N,
                         cudaStream_t st) {
  const int grid = ceil_div(4 * B, 128);
  if (workspace == nullptr) {   // single launch, alphas tried one after the other
    ls_qd_phaseA_kernel<0><<<grid, 128, 0, st>>>(*p, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out,
                                                 n_active, QdWork{}, tol, flags, B, N);
    QX_CHECK_LAUNCH("ls_qd_phaseA_kernel<0>");
    return QX_OK;
  }
  QX_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "workspace must be 16-byte aligned");
  const QdWork ws = qd_carve(workspace, B, N);
  cudaError_t e = cudaMemsetAsync(ws.ctl, 0, 32, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(ctl)");
  ls_qd_phaseA_kernel<1><<<grid, 128, 0, st>>>(*p, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out, n_active,
                                               ws, tol, flags, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseA_kernel<1>");
  const dim3 gridC(ceil_div(B, 128), 8);
  // stage 1: alpha = .5, .25
  ls_qd_phaseB_kernel<<<ceil_div(4 * 2 * B, 128), 128, 0, st>>>(*p, x_in, u_in, gains, ws.queue1, &ws.ctl[0], ws, 1, 2, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseB_kernel<1>");
  ls_qd_phaseC_kernel<false><<<gridC, 128, 0, st>>>(x_in, u_in, cost, status, iters, alpha, x_out, u_out, n_active, ws,
                                                    ws.pend1, ws.pend2, ws.queue2, &ws.ctl[1], 1, 2, tol, flags, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseC_kernel<false>");
  // stage 2: alpha = .1, .05, .01
  ls_qd_phaseB_kernel<<<ceil_div(4 * 3 * B, 128), 128, 0, st>>>(*p, x_in, u_in, gains, ws.queue2, &ws.ctl[1], ws, 3, 3, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseB_kernel<2>");
  ls_qd_phaseC_kernel<true><<<gridC, 128, 0, st>>>(x_in, u_in, cost, status, iters, alpha, x_out, u_out, n_active, ws,
                                                   ws.pend2, nullptr, nullptr, nullptr, 3, 3, tol, flags, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseC_kernel<true>");
  return QX_OK;
}

}  // namespace qx
