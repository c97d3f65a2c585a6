// Quadrotor rollouts and line search: ONE thread per candidate rollout (trajectory, alpha).
//
//   rollout_qd_kernel        = iLQR_TF.simulate + compute_total_cost             (quattro_ilqr_tf.py:127-143)
//   ls_qd_phaseA/B/C_kernel  = the alpha loop of optimize (:440-451) around forward_pass (:377-390)
//                              plus the stopping rule (:472)
//   forward_pass_qd_kernel   = forward_pass with a caller-chosen alpha (:377-390)
//
// Used when Q, R, Qf are diagonal (the reference's QuadrotorMPC, quadrotor_mpc.py:40-46); dense weights
// take the generic kernels of qx_rollout.cu.
//
// Why a thread per rollout.  A rollout is a chain of N dependent steps of ~400 fp64 operations.  The first
// design of this file spread one rollout over four lanes (state blocks p | v | eta | omega); it paid for that
// with ~70 shuffles and ~25 selects per step and with every lane evaluating all four block formulas: 67 issue
// slots and 24 fp64 instructions per trajectory-step, and the kernels sat at 36 % of the fp64 pipe / 50 % of the
// issue rate (profiles/r1_ncu_full_classic_summary.txt).  One thread per rollout needs no exchange at all:
// 16 issue slots and 13 fp64 instructions per trajectory-step.  A step is one basic block with plenty of
// independent chains (three sincos, four exp, twelve partial sums of K dx), so a single warp per scheduler keeps
// the fp64 pipe busy - which is all a 16 384-trajectory batch provides (512 warps on 592 schedulers).
//
// Early rejection.  L and Lf are sums of non-negative terms (diagonal Q, R, Qf >= 0, barrier >= 0), so the partial
// cost of a candidate never decreases along the horizon - in floating point too, because rounding is monotone
// (a' >= a, b' >= b  =>  fl(a' + b') >= fl(a + b)).  A candidate whose partial cost already exceeds the nominal's
// total is therefore rejected on the spot with certainty; its thread stops loading and storing, and a warp leaves
// the horizon loop once none of its candidates is alive.  On the benchmark workload the rejected alpha = 1, .5, .25
// candidates cross within 2-13 of the 50 steps (they diverge), so most of their HBM traffic disappears.  Accepted
// candidates run to the end with exactly the arithmetic of a plain rollout.
//
// The file is compiled with -fmad=false and every fused multiply-add is written out (fma()), so the nominal
// rollout, the line search and forward_pass evaluate f and L with identical bits for identical inputs - like
// the reference's single Python implementation does; the accept test compares such costs with "<=".  The
// operation order (block sums of the cost, partial sums of K dx) is the one of the four-lane version, whose
// results this file reproduces bit for bit.
//
// Line search = three stages, so that the first accepted alpha is found without rolling out all six:
//   A   alpha = 1 for every active trajectory, candidate written straight into (x_out, u_out)
//   B1  alpha = .5, .25 of the rejected ones, one thread each (queue compacted by A)   -> C1 picks / re-queues
//   B2  alpha = .1, .05, .01 of the still rejected ones                                 -> C2 picks / gives up
// Candidates of B are parked in a scratch [plane][row][b]; C copies the winner out with thread-per-trajectory
// (coalesced) row loops.  Queues are appended per warp in trajectory order, so the threads of a B warp touch
// neighbouring columns.
#include "qx_common.cuh"
#include "qx_fp64math.cuh"

namespace qx {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kN = 12, kM = 4, kG = 52;

// Host-prepared constants.  Passed as a __grid_constant__ kernel parameter: every member is a constant-bank operand
// of the fp64 instructions and costs no register.
struct QdK {
  double inv_mass, grav, arm, kyaw, invIx, invIy, invIz, c1, c2, c3, dt, beta, inv_beta, alpha_b;
  double qd[kN], qfd[kN], xr[kN], rd[kM];
  int32_t early;      // 1: all cost terms are non-negative, partial costs may reject early
  int32_t _pad;
};

QdK qd_consts(const QxParams& p) {
  QdK k;
  k.inv_mass = 1.0 / p.phys[0];
  k.invIx = 1.0 / p.phys[1]; k.invIy = 1.0 / p.phys[2]; k.invIz = 1.0 / p.phys[3];
  k.arm = p.phys[4]; k.grav = p.phys[5]; k.kyaw = p.phys[6];
  k.c1 = (p.phys[2] - p.phys[3]) / p.phys[1];
  k.c2 = (p.phys[3] - p.phys[1]) / p.phys[2];
  k.c3 = (p.phys[1] - p.phys[2]) / p.phys[3];
  k.dt = p.dt; k.beta = p.barrier_beta; k.inv_beta = 1.0 / p.barrier_beta; k.alpha_b = p.barrier_alpha;
  bool nonneg = p.barrier_alpha >= 0.0;
  for (int i = 0; i < kN; ++i) {
    k.qd[i] = p.Q[i * kN + i];
    k.qfd[i] = p.Qf[i * kN + i];
    k.xr[i] = p.x_ref[i];
    nonneg = nonneg && k.qd[i] >= 0.0 && k.qfd[i] >= 0.0;
  }
  for (int j = 0; j < kM; ++j) {
    k.rd[j] = p.R[j * kM + j];
    nonneg = nonneg && k.rd[j] >= 0.0;
  }
  k.early = nonneg ? 1 : 0;
  k._pad = 0;
  return k;
}

// Continuous dynamics (quadrotor_dynamics.py:64-164), x = [p | v | eta | omega].
__device__ __forceinline__ void qd_fc(const double* x, const double* u, const QdK& k, double* d) {
  double s[3], c[3];
  fast_sincos_n<3>(x + 6, s, c);
  const double sphi = s[0], cphi = c[0], sth = s[1], cth = c[1], spsi = s[2], cpsi = c[2];
  const double sec = fast_rcp(cth);
  d[0] = x[3]; d[1] = x[4]; d[2] = x[5];
  // v' = T/m * R e3 - g e3
  const double Tm = ((u[0] + u[1]) + (u[2] + u[3])) * k.inv_mass;
  const double sc = sth * cphi;
  d[3] = Tm * fma(cpsi, sc, spsi * sphi);
  d[4] = Tm * fma(-spsi, sc, cpsi * sphi);
  d[5] = fma(Tm, cth * cphi, -k.grav);
  // eta' from omega
  const double pp = x[9], q = x[10], r = x[11];
  const double tth = sth * sec;
  const double w = fma(q, sphi, r * cphi);            // q sin(phi) + r cos(phi)
  d[6] = fma(w, tth, pp);
  d[7] = fma(q, cphi, -(r * sphi));
  d[8] = w * sec;
  // omega'
  d[9] = fma(k.c1, q * r, k.arm * ((u[1] + u[2]) - (u[0] + u[3])) * k.invIx);
  d[10] = fma(k.c2, pp * r, k.arm * ((u[0] + u[1]) - (u[2] + u[3])) * k.invIy);
  d[11] = fma(k.c3, pp * q, k.kyaw * ((u[0] - u[1]) + (u[2] - u[3])) * k.invIz);
}

// discrete map (quadrotor_dynamics.py:186-195); the integrator is a template parameter so that the Euler kernels
// do not carry the registers of the RK4 stages
template <int INTEG>
__device__ __forceinline__ void qd_step(double* x, const double* u, const QdK& k) {
  const double dt = k.dt;
  double k1[kN];
  qd_fc(x, u, k, k1);
  if (INTEG == QX_EULER) {
#pragma unroll
    for (int i = 0; i < kN; ++i) x[i] = fma(dt, k1[i], x[i]);
    return;
  }
  double k2[kN], k3[kN], k4[kN], xs[kN];
#pragma unroll
  for (int i = 0; i < kN; ++i) xs[i] = fma(0.5 * dt, k1[i], x[i]);
  qd_fc(xs, u, k, k2);
#pragma unroll
  for (int i = 0; i < kN; ++i) xs[i] = fma(0.5 * dt, k2[i], x[i]);
  qd_fc(xs, u, k, k3);
#pragma unroll
  for (int i = 0; i < kN; ++i) xs[i] = fma(dt, k3[i], x[i]);
  qd_fc(xs, u, k, k4);
#pragma unroll
  for (int i = 0; i < kN; ++i) x[i] = fma(dt / 6.0, (k1[i] + 2.0 * k2[i]) + (2.0 * k3[i] + k4[i]), x[i]);
}

// L(x, u) (quadrotor_mpc.py:86-93) added to four partial sums: c[g] takes state block g, control g and its barrier
// term alpha_b * softplus(-u_g; beta)^2 (:80, :92).  log1p(e) by its series when e is tiny (u well on the feasible
// side), by fast_log1p_unit otherwise.
__device__ __forceinline__ void qd_running(const double* x, const double* u, const QdK& k, double* c) {
  double e[kM], bz[kM], l1p[kM];
  bool need = false;
#pragma unroll
  for (int j = 0; j < kM; ++j) {
    bz[j] = -k.beta * u[j];
    e[j] = fast_exp_neg(-fabs(bz[j]));
    l1p[j] = e[j] * fma(-e[j], fma(-e[j], fma(-0.25, e[j], 1.0 / 3.0), 0.5), 1.0);
    need = need || !(e[j] < 1e-4);
  }
  if (need) {
#pragma unroll
    for (int j = 0; j < kM; ++j) {
      const double full = fast_log1p_unit(e[j]);
      l1p[j] = (e[j] < 1e-4) ? l1p[j] : full;
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double dxv = x[3 * g + r] - k.xr[3 * g + r];
      acc = fma(k.qd[3 * g + r] * dxv, dxv, acc);
    }
    acc = fma(k.rd[g] * u[g], u[g], acc);
    const double sp = (fmax(bz[g], 0.0) + l1p[g]) * k.inv_beta;
    c[g] += acc + k.alpha_b * sp * sp;
  }
}

__device__ __forceinline__ void qd_final(const double* x, const QdK& k, double* c) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double dxv = x[3 * g + r] - k.xr[3 * g + r];
      acc = fma(k.qfd[3 * g + r] * dxv, dxv, acc);
    }
    c[g] += acc;
  }
}

__device__ __forceinline__ double qd_total(const double* c) { return (c[0] + c[1]) + (c[2] + c[3]); }

// One candidate rollout (forward_pass :377-390, step size alpha) by one thread.  Nominal at xin / uin (row stride sB),
// gains at rec (float4 [N][13][B], same row stride); the candidate goes to xo / uo (row stride sO).  `limit`: the
// candidate is abandoned (its thread stops loading and storing) once its partial cost exceeds it; +inf = never.
// All 32 lanes of the warp must call it; lanes with run == false only take part in the votes.
// Returns the total cost, or a partial cost > limit for an abandoned candidate.
template <int INTEG>
__device__ __forceinline__ double qd_candidate(const QdK& k, bool run, const double* __restrict__ xin,
                                               const double* __restrict__ uin, const float4* __restrict__ rec, size_t sB,
                                               double alpha, double limit, double* __restrict__ xo,
                                               double* __restrict__ uo, size_t sO, int N) {
  double xh[kN], xn[kN], un[kM];
  float4 g[kG / 4];
#pragma unroll
  for (int i = 0; i < kN; ++i) xh[i] = run ? xin[(size_t)i * sB] : 0.0;
  auto fetch = [&]() {
#pragma unroll
    for (int i = 0; i < kN; ++i) xn[i] = xin[(size_t)i * sB];
#pragma unroll
    for (int j = 0; j < kM; ++j) un[j] = uin[(size_t)j * sB];
#pragma unroll
    for (int q = 0; q < kG / 4; ++q) g[q] = rec[(size_t)q * sB];
    xin += kN * sB;
    uin += kM * sB;
    rec += (kG / 4) * sB;
  };
  if (run) fetch();
  double c[4] = {0.0, 0.0, 0.0, 0.0};
  for (int t = 0; t < N; ++t) {
    if (!__any_sync(FULL, run)) break;
    double uh[kM];
    {
      double dx[kN];
#pragma unroll
      for (int i = 0; i < kN; ++i) dx[i] = xh[i] - xn[i];
      const float* gf = reinterpret_cast<const float*>(g);
#pragma unroll
      for (int j = 0; j < kM; ++j) {
        // du_j = k_j + K[j,:] (x_hat - x)
        double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int i = 0; i < 3; ++i) acc[i] = fma((double)gf[kM + j * kN + 3 * q + i], dx[3 * q + i], acc[i]);
        const double kd = (acc[0] + acc[1]) + acc[2];
        uh[j] = fma(alpha, (double)gf[j] + kd, un[j]);        // alpha scales the feedback term too (:382-383)
      }
    }
    // nominal and gains of this step are consumed: the loads of the next step fly during the rest of this one
    if (run && t + 1 < N) fetch();
    if (run) {
#pragma unroll
      for (int j = 0; j < kM; ++j) uo[(size_t)j * sO] = uh[j];
#pragma unroll
      for (int i = 0; i < kN; ++i) xo[(size_t)i * sO] = xh[i];
    }
    uo += kM * sO;
    xo += kN * sO;
    qd_running(xh, uh, k, c);
    qd_step<INTEG>(xh, uh, k);
    run = run && !(qd_total(c) > limit);
    if (!run) { xh[6] = 0.0; xh[7] = 0.0; xh[8] = 0.0; }      // a dead lane never sends the warp into the slow sincos
  }
  if (run) {
#pragma unroll
    for (int i = 0; i < kN; ++i) xo[(size_t)i * sO] = xh[i];
    qd_final(xh, k, c);
  }
  return qd_total(c);
}

// Append the flagged trajectories of this warp to a queue, in trajectory order.  Must be reached by all 32 lanes.
__device__ __forceinline__ void qd_ordered_append(bool flag, int value, int32_t* __restrict__ queue,
                                                  int32_t* __restrict__ counter) {
  const int lane = threadIdx.x & 31;
  const unsigned bal = __ballot_sync(FULL, flag);
  int base = 0;
  if (lane == 0 && bal != 0u) base = atomicAdd(counter, __popc(bal));
  base = __shfl_sync(FULL, base, 0);
  if (flag) queue[base + __popc(bal & ((1u << lane) - 1u))] = value;
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
__global__ void __launch_bounds__(32)
rollout_qd_kernel(const __grid_constant__ QdK k, const double* __restrict__ x0, const double* __restrict__ u,
                  double* __restrict__ x, double* __restrict__ cost, const int32_t* __restrict__ status, int B, int N) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (status != nullptr && status[b] != 0) return;
  const size_t sB = (size_t)B;
  double xh[kN], uh[kM], un[kM], c[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < kN; ++i) xh[i] = x0[(size_t)i * sB + b];
  const double* up = u + b;
  double* xp = x + b;
#pragma unroll
  for (int j = 0; j < kM; ++j) un[j] = up[(size_t)j * sB];
  for (int t = 0; t < N; ++t) {
#pragma unroll
    for (int j = 0; j < kM; ++j) uh[j] = un[j];
    up += kM * sB;
    if (t + 1 < N) {
#pragma unroll
      for (int j = 0; j < kM; ++j) un[j] = up[(size_t)j * sB];
    }
#pragma unroll
    for (int i = 0; i < kN; ++i) xp[(size_t)i * sB] = xh[i];
    xp += kN * sB;
    qd_running(xh, uh, k, c);
    qd_step<INTEG>(xh, uh, k);
  }
#pragma unroll
  for (int i = 0; i < kN; ++i) xp[(size_t)i * sB] = xh[i];
  qd_final(xh, k, c);
  cost[b] = qd_total(c);
}

// Bookkeeping after the line search of one trajectory (stopping rule :472).  Leaves the next iterate in
// (x_out,u_out); a finished trajectory has it in both buffers.  Returns "still active".
__device__ __forceinline__ bool qd_finalize(int b, int idx, double cur, double new_cost, double tol, uint32_t flags,
                                            double* __restrict__ x_in, double* __restrict__ u_in,
                                            double* __restrict__ x_out, double* __restrict__ u_out,
                                            double* __restrict__ cost, int32_t* __restrict__ status,
                                            int32_t* __restrict__ iters, int32_t* __restrict__ alpha_out, int B, int N) {
  const bool found = idx >= 0;
  bool done = (!found) || (fabs(cur - new_cost) < tol);
  if (flags & QX_LS_FORCE_ACTIVE) done = false;
  if (!found || done) {
    // no alpha: (x_out,u_out) <- nominal;  finished with an update: (x_in,u_in) <- final iterate
    const double* xs = found ? x_out : x_in;
    const double* us = found ? u_out : u_in;
    double* xd = found ? x_in : x_out;
    double* ud = found ? u_in : u_out;
    const int xrows = (N + 1) * kN, urows = N * kM;
    for (int r = 0; r < xrows; ++r) xd[(size_t)r * B + b] = xs[(size_t)r * B + b];
    for (int r = 0; r < urows; ++r) ud[(size_t)r * B + b] = us[(size_t)r * B + b];
  }
  cost[b] = new_cost;
  alpha_out[b] = idx;
  iters[b] += 1;
  if (done) status[b] = 1;
  return !done;
}

// MODE 0: the whole alpha list sequentially with early exit (single launch, no workspace).
// MODE 1: phase A of the staged search: alpha = 1 only; rejected trajectories are queued.
template <int MODE, int INTEG>
__global__ void __launch_bounds__(32)
ls_qd_phaseA_kernel(const __grid_constant__ QdK k, double* __restrict__ x_in, double* __restrict__ u_in,
                    const float* __restrict__ gains, double* __restrict__ cost, int32_t* __restrict__ status,
                    int32_t* __restrict__ iters, int32_t* __restrict__ alpha_out, double* __restrict__ x_out,
                    double* __restrict__ u_out, int32_t* __restrict__ n_active, QdWork ws, double tol, uint32_t flags,
                    int B, int N) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool mine = (b < B) && (status[b] == 0);
  if (!__any_sync(FULL, mine)) {
    if (MODE == 1 && b < B) ws.pend1[b] = 0;
    return;
  }
  const int bs = mine ? b : 0;
  const double cur = mine ? cost[b] : 0.0;
  const double limit = k.early ? cur : __longlong_as_double(0x7ff0000000000000LL);
  bool accepted = !mine;
  int idx = -1;
  double new_cost = cur;
  constexpr int n_alpha = (MODE == 0) ? QX_NUM_ALPHA : 1;
  for (int a = 0; a < n_alpha; ++a) {
    if (!__any_sync(FULL, !accepted)) break;
    const bool run = !accepted;
    const double cst = qd_candidate<INTEG>(k, run, x_in + bs, u_in + bs, reinterpret_cast<const float4*>(gains) + bs,
                                           (size_t)B, kAlphas[a], limit, x_out + bs, u_out + bs, (size_t)B, N);
    if (run && cst <= cur) {       // first accepted alpha, "<=" (quattro_ilqr_tf.py:444)
      accepted = true;
      idx = a;
      new_cost = cst;
    }
  }
  bool still_active = false;
  const bool pending = (MODE == 1) && mine && idx < 0;
  if (mine && !pending)
    still_active = qd_finalize(b, idx, cur, new_cost, tol, flags, x_in, u_in, x_out, u_out, cost, status, iters, alpha_out,
                               B, N);
  if (MODE == 1) {
    if (b < B) {
      ws.pend1[b] = pending ? 1 : 0;
      if (pending) ws.cur[b] = cur;
    }
    qd_ordered_append(pending, b, ws.queue1, &ws.ctl[0]);
  }
  const unsigned act = __ballot_sync(FULL, still_active);
  if (lane == 0 && n_active != nullptr && act != 0u) atomicAdd(n_active, __popc(act));
}

// Phase B: one thread per (queued trajectory, alpha of the stage); blockIdx.y selects the alpha, so the warps of
// the NA candidates of a trajectory read the same nominal and gain lines (L2 hits) and stay coalesced.
template <int INTEG>
__global__ void __launch_bounds__(32)
ls_qd_phaseB_kernel(const __grid_constant__ QdK k, const double* __restrict__ x_in, const double* __restrict__ u_in,
                    const float* __restrict__ gains, const int32_t* __restrict__ queue, const int32_t* __restrict__ nq_ptr,
                    QdWork ws, int a_first, int B, int N) {
  const int nq = *nq_ptr;
  if ((int)(blockIdx.x * blockDim.x) >= nq) return;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  const bool run = item < nq;
  const int b = run ? queue[item] : 0;
  const int a = a_first + blockIdx.y, plane = a - 1;
  const double cur = ws.cur[b];
  const double limit = k.early ? cur : __longlong_as_double(0x7ff0000000000000LL);
  double* xd = ws.xs + (size_t)plane * (N + 1) * kN * B + b;
  double* ud = ws.us + (size_t)plane * N * kM * B + b;
  const double cst = qd_candidate<INTEG>(k, run, x_in + b, u_in + b, reinterpret_cast<const float4*>(gains) + b, (size_t)B,
                                         kAlphas[a], limit, xd, ud, (size_t)B, N);
  if (run) ws.cand[(size_t)plane * B + b] = cst;
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
__global__ void __launch_bounds__(32)
forward_pass_qd_kernel(const __grid_constant__ QdK k, const double* __restrict__ x_in, const double* __restrict__ u_in,
                       const float* __restrict__ gains, double alpha, double* __restrict__ x_out, double* __restrict__ u_out,
                       double* __restrict__ cost_out, int B, int N) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool run = b < B;
  const int bs = run ? b : 0;
  const double cst = qd_candidate<INTEG>(k, run, x_in + bs, u_in + bs, reinterpret_cast<const float4*>(gains) + bs, (size_t)B,
                                         alpha, __longlong_as_double(0x7ff0000000000000LL), x_out + bs, u_out + bs,
                                         (size_t)B, N);
  if (run) cost_out[b] = cst;
}

int qd_launch_forward_pass(const QxParams* p, const double* x_in, const double* u_in, const float* gains, double alpha,
                           double* x_out, double* u_out, double* cost_out, int B, int N, cudaStream_t st) {
  const QdK k = qd_consts(*p);
  const int grid = ceil_div(B, 32);
  if (p->integrator == QX_EULER)
    forward_pass_qd_kernel<QX_EULER><<<grid, 32, 0, st>>>(k, x_in, u_in, gains, alpha, x_out, u_out, cost_out, B, N);
  else
    forward_pass_qd_kernel<QX_RK4><<<grid, 32, 0, st>>>(k, x_in, u_in, gains, alpha, x_out, u_out, cost_out, B, N);
  QX_CHECK_LAUNCH("forward_pass_qd_kernel");
  return QX_OK;
}

// ---- host side: called by qx_rollout_cost / qx_linesearch (qx_rollout.cu) for the quadrotor with diagonal weights
int qd_launch_rollout(const QxParams* p, const double* x0, const double* u, double* x, double* cost,
                      const int32_t* status, int B, int N, cudaStream_t st) {
  const QdK k = qd_consts(*p);
  if (p->integrator == QX_EULER)
    rollout_qd_kernel<QX_EULER><<<ceil_div(B, 32), 32, 0, st>>>(k, x0, u, x, cost, status, B, N);
  else
    rollout_qd_kernel<QX_RK4><<<ceil_div(B, 32), 32, 0, st>>>(k, x0, u, x, cost, status, B, N);
  QX_CHECK_LAUNCH("rollout_qd_kernel");
  return QX_OK;
}

template <int INTEG>
static int qd_linesearch_t(const QxParams* p, double* x_in, double* u_in, const float* gains, double* cost,
                           int32_t* status, int32_t* iters, int32_t* alpha, double* x_out, double* u_out,
                           int32_t* n_active, void* workspace, double tol, uint32_t flags, int B, int N,
                           cudaStream_t st) {
  const QdK k = qd_consts(*p);
  const int grid = ceil_div(B, 32);
  if (workspace == nullptr) {   // single launch, alphas tried one after the other
    ls_qd_phaseA_kernel<0, INTEG><<<grid, 32, 0, st>>>(k, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out,
                                                       n_active, QdWork{}, tol, flags, B, N);
    QX_CHECK_LAUNCH("ls_qd_phaseA_kernel<0>");
    return QX_OK;
  }
  QX_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "workspace must be 16-byte aligned");
  const QdWork ws = qd_carve(workspace, B, N);
  cudaError_t e = cudaMemsetAsync(ws.ctl, 0, 32, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(ctl)");
  ls_qd_phaseA_kernel<1, INTEG><<<grid, 32, 0, st>>>(k, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out,
                                                     n_active, ws, tol, flags, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseA_kernel<1>");
  const dim3 gridC(ceil_div(B, 128), 16);
  // stage 1: alpha = .5, .25
  ls_qd_phaseB_kernel<INTEG><<<dim3(grid, 2), 32, 0, st>>>(k, x_in, u_in, gains, ws.queue1, &ws.ctl[0], ws, 1, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseB_kernel<1>");
  ls_qd_phaseC_kernel<false><<<gridC, 128, 0, st>>>(x_in, u_in, cost, status, iters, alpha, x_out, u_out, n_active, ws,
                                                    ws.pend1, ws.pend2, ws.queue2, &ws.ctl[1], 1, 2, tol, flags, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseC_kernel<false>");
  // stage 2: alpha = .1, .05, .01
  ls_qd_phaseB_kernel<INTEG><<<dim3(grid, 3), 32, 0, st>>>(k, x_in, u_in, gains, ws.queue2, &ws.ctl[1], ws, 3, B, N);
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
