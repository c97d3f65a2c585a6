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
// Line search = three launches, so that the first accepted alpha is found without six sequential rollouts:
//   A   alpha = 1 for every active trajectory, candidate written straight into (x_out, u_out); the rejected
//       trajectories are queued (appended per warp in trajectory order, so neighbours stay neighbours)
//   B   alpha = .5, .25, .1, .05, .01 of the queued ones side by side, one thread each, costs only
//   W   the first accepted alpha of a queued trajectory is rolled out once more, into (x_out, u_out); stopping rule
// B and W exit at once when the queue is empty.
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
// The 28 weights of the running cost, copied to shared memory by every CTA of the candidate kernels.  As kernel
// parameters they are uniform-register operands of the fp64 instructions - and the uniform file holds 63 registers
// against the ~100 fp64 constants of a step, so the compiler parked these in 56 VECTOR registers for the whole loop and
// moved each to a uniform pair at its one use per step (56 R2UR per step).  A broadcast LDS.64 per use costs half the
// instructions and no long-lived register.
struct QdCostW {
  double qd[kN], xr[kN], rd[kM];
};
__device__ __forceinline__ void qd_stage_cost(const QdK& k, QdCostW& w) {     // all 32 lanes of the (one-warp) CTA
  const int lane = threadIdx.x & 31;
  if (lane < kN) { w.qd[lane] = k.qd[lane]; w.xr[lane] = k.xr[lane]; }
  if (lane < kM) w.rd[lane] = k.rd[lane];
  __syncwarp();
}

template <class W>
__device__ __forceinline__ void qd_running(const double* x, const double* u, const QdK& k, const W& w, double* c) {
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
      const double dxv = x[3 * g + r] - w.xr[3 * g + r];
      acc = fma(w.qd[3 * g + r] * dxv, dxv, acc);
    }
    acc = fma(w.rd[g] * u[g], u[g], acc);
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

// Staging buffer of one warp (= one CTA): the nominal and the gains of the NEXT step of every lane, copied
// global -> shared by cp.async (LDGSTS) while the current step is computed.  In registers the same prefetch costs
// 84 of them (the kernels sat at 255 with spills: two warps per scheduler); staged, a candidate needs <= 168 and
// three warps share a scheduler.  [row][lane] layout: conflict free, and a thread only reads what it copied itself.
struct QdStage {
  double xu[2][kN + kM][32];
  float4 g[2][kG / 4][32];
  QdCostW w;
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

// One candidate rollout (forward_pass :377-390, step size alpha) by one thread.  Nominal at xin / uin (row stride sB),
// gains at rec (float4 [N][13][B], same row stride); the candidate goes to xo / uo (row stride sO).  `limit`: the
// candidate is abandoned (its thread stops loading and storing) once its partial cost exceeds it; +inf = never.
// All 32 lanes of the warp must call it; lanes with run == false only take part in the votes.
// Returns the total cost, or a partial cost > limit for an abandoned candidate.
template <int INTEG, bool STORE>
__device__ __forceinline__ double qd_candidate(const QdK& k, QdStage& st, bool run, const double* __restrict__ xin,
                                               const double* __restrict__ uin, const float4* __restrict__ rec, size_t sB,
                                               double alpha, double limit, double* __restrict__ xo,
                                               double* __restrict__ uo, size_t sO, int N) {
  const int lane = threadIdx.x & 31;
  qd_stage_cost(k, st.w);
  double xh[kN];
#pragma unroll
  for (int i = 0; i < kN; ++i) xh[i] = run ? xin[(size_t)i * sB] : 0.0;
  // the copies of step t land in buffer t & 1; one commit group per step (empty when there is nothing to copy), so
  // "all but the newest group" is always "everything up to step t"
  auto fetch = [&](int buf, bool doit) {
    if (doit) {
#pragma unroll
      for (int i = 0; i < kN; ++i) cp_async8(&st.xu[buf][i][lane], xin + (size_t)i * sB);
#pragma unroll
      for (int j = 0; j < kM; ++j) cp_async8(&st.xu[buf][kN + j][lane], uin + (size_t)j * sB);
#pragma unroll
      for (int q = 0; q < kG / 4; ++q) cp_async16(&st.g[buf][q][lane], rec + (size_t)q * sB);
    }
    cp_async_commit();
    xin += kN * sB;
    uin += kM * sB;
    rec += (kG / 4) * sB;
  };
  fetch(0, run);
  fetch(1, run && N > 1);
  double c[4] = {0.0, 0.0, 0.0, 0.0};
  for (int t = 0; t < N; ++t) {
    if (!__any_sync(FULL, run)) break;
    cp_async_wait<1>();
    const int buf = t & 1;
    double uh[kM];
    {
      double dx[kN];
#pragma unroll
      for (int i = 0; i < kN; ++i) dx[i] = xh[i] - st.xu[buf][i][lane];
      float gf[kG];
#pragma unroll
      for (int q = 0; q < kG / 4; ++q) {
        const float4 v = st.g[buf][q][lane];
        gf[4 * q] = v.x; gf[4 * q + 1] = v.y; gf[4 * q + 2] = v.z; gf[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < kM; ++j) {
        // du_j = k_j + K[j,:] (x_hat - x)
        double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int i = 0; i < 3; ++i) acc[i] = fma((double)gf[kM + j * kN + 3 * q + i], dx[3 * q + i], acc[i]);
        const double kd = (acc[0] + acc[1]) + acc[2];
        const double un = st.xu[buf][kN + j][lane];
        // alpha scales the feedback term too (:382-383); alpha = 0 is the nominal itself, whatever the gains hold
        uh[j] = (alpha != 0.0) ? fma(alpha, (double)gf[j] + kd, un) : un;
      }
    }
    // Nominal and gains of this step are consumed - u_hat depends on every staged value, so once it exists all reads
    // of the buffer have returned - and it is refilled with step t + 2: the copies have almost two steps to land.
    asm volatile("" ::"d"(uh[0]), "d"(uh[1]), "d"(uh[2]), "d"(uh[3]) : "memory");
    fetch(buf, run && t + 2 < N);
    if (STORE) {
      if (run) {
#pragma unroll
        for (int j = 0; j < kM; ++j) uo[(size_t)j * sO] = uh[j];
#pragma unroll
        for (int i = 0; i < kN; ++i) xo[(size_t)i * sO] = xh[i];
      }
      uo += kM * sO;
      xo += kN * sO;
    }
    qd_running(xh, uh, k, st.w, c);
    qd_step<INTEG>(xh, uh, k);
    run = run && !(qd_total(c) > limit);
    if (!run) { xh[6] = 0.0; xh[7] = 0.0; xh[8] = 0.0; }      // a dead lane never sends the warp into the slow sincos
  }
  cp_async_wait<0>();
  if (run) {
    if (STORE) {
#pragma unroll
      for (int i = 0; i < kN; ++i) xo[(size_t)i * sO] = xh[i];
    }
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
  int32_t* ctl;        // [0] = |queue|
  int32_t *queue, *pend, *mirror;
  double* cur;         // cost of the nominal, snapshot taken by phase A
  double* cand;        // [5][B] candidate costs, plane = alpha index - 1
};

__host__ __device__ inline size_t qd_align16(size_t v) { return (v + 15) / 16 * 16; }

size_t qd_workspace_bytes(int B, int N) {
  (void)N;
  constexpr size_t A = QX_NUM_ALPHA - 1;
  return 32 + 3 * qd_align16((size_t)B * 4) + (size_t)B * 8 + A * (size_t)B * 8;
}

static QdWork qd_carve(void* workspace, int B) {
  uint8_t* w = static_cast<uint8_t*>(workspace);
  const size_t qb = qd_align16((size_t)B * 4);
  QdWork k;
  k.ctl = reinterpret_cast<int32_t*>(w);
  k.queue = reinterpret_cast<int32_t*>(w + 32);
  k.pend = reinterpret_cast<int32_t*>(w + 32 + qb);
  k.mirror = reinterpret_cast<int32_t*>(w + 32 + 2 * qb);
  k.cur = reinterpret_cast<double*>(w + 32 + 3 * qb);
  k.cand = k.cur + B;
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
    qd_running(xh, uh, k, k, c);
    qd_step<INTEG>(xh, uh, k);
  }
#pragma unroll
  for (int i = 0; i < kN; ++i) xp[(size_t)i * sB] = xh[i];
  qd_final(xh, k, c);
  cost[b] = qd_total(c);
}

// Bookkeeping after the line search of one trajectory (stopping rule :472).  The next iterate must end up in
// (x_out,u_out), and a finished trajectory keeps its final iterate in both buffers.  Returns "still active".
// INLINE_COPY (single-launch search): the copies that this takes - the nominal into (x_out,u_out) when no alpha was
// accepted, the final iterate back into (x_in,u_in) - are made here, by the trajectory's own thread.  Otherwise only
// *mirror is set (1: copy out -> in) and ls_qd_mirror_kernel does it with the whole machine; a trajectory without an
// accepted alpha has already been given its nominal by phase W (alpha = 0 rollout).
template <bool INLINE_COPY>
__device__ __forceinline__ bool qd_finalize(int b, int idx, double cur, double new_cost, double tol, uint32_t flags,
                                            double* __restrict__ x_in, double* __restrict__ u_in,
                                            double* __restrict__ x_out, double* __restrict__ u_out,
                                            double* __restrict__ cost, int32_t* __restrict__ status,
                                            int32_t* __restrict__ iters, int32_t* __restrict__ alpha_out,
                                            int32_t* __restrict__ mirror, int B, int N) {
  const bool found = idx >= 0;
  bool done = (!found) || (fabs(cur - new_cost) < tol);
  if (flags & QX_LS_FORCE_ACTIVE) done = false;
  if (INLINE_COPY) {
    if (!found || done) {
      // no alpha: (x_out,u_out) <- nominal;  finished with an update: (x_in,u_in) <- final iterate
      const double* xs = found ? x_out : x_in;
      const double* us = found ? u_out : u_in;
      double* xd = found ? x_in : x_out;
      double* ud = found ? u_in : u_out;
      const int xrows = (N + 1) * kN, rows = xrows + N * kM;
      constexpr int U = 8;                      // loads batched eight deep: the copy is latency-bound
      for (int r = 0; r < rows; r += U) {
        double v[U];
#pragma unroll
        for (int q = 0; q < U; ++q)
          if (r + q < rows) v[q] = (r + q < xrows) ? xs[(size_t)(r + q) * B + b] : us[(size_t)(r + q - xrows) * B + b];
#pragma unroll
        for (int q = 0; q < U; ++q)
          if (r + q < rows) {
            if (r + q < xrows) xd[(size_t)(r + q) * B + b] = v[q];
            else ud[(size_t)(r + q - xrows) * B + b] = v[q];
          }
      }
    }
  } else {
    *mirror = (found && done) ? 1 : 0;
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
__global__ void __launch_bounds__(32, 8)
ls_qd_phaseA_kernel(const __grid_constant__ QdK k, double* __restrict__ x_in, double* __restrict__ u_in,
                    const float* __restrict__ gains, double* __restrict__ cost, int32_t* __restrict__ status,
                    int32_t* __restrict__ iters, int32_t* __restrict__ alpha_out, double* __restrict__ x_out,
                    double* __restrict__ u_out, int32_t* __restrict__ n_active, QdWork ws, double tol, uint32_t flags,
                    int B, int N) {
  __shared__ QdStage stage;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool mine = (b < B) && (status[b] == 0);
  if (!__any_sync(FULL, mine)) {
    if (MODE == 1 && b < B) { ws.pend[b] = 0; ws.mirror[b] = 0; }
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
    const double cst = qd_candidate<INTEG, true>(k, stage, run, x_in + bs, u_in + bs, reinterpret_cast<const float4*>(gains) + bs,
                                           (size_t)B, kAlphas[a], limit, x_out + bs, u_out + bs, (size_t)B, N);
    if (run && cst <= cur) {       // first accepted alpha, "<=" (quattro_ilqr_tf.py:444)
      accepted = true;
      idx = a;
      new_cost = cst;
    }
  }
  bool still_active = false;
  const bool pending = (MODE == 1) && mine && idx < 0;
  int mir = 0;
  if (mine && !pending)
    still_active = qd_finalize<MODE == 0>(b, idx, cur, new_cost, tol, flags, x_in, u_in, x_out, u_out, cost, status, iters,
                                          alpha_out, &mir, B, N);
  if (MODE == 1) {
    if (b < B) {
      ws.pend[b] = pending ? 1 : 0;
      ws.mirror[b] = mir;
      if (pending) ws.cur[b] = cur;
    }
    qd_ordered_append(pending, b, ws.queue, &ws.ctl[0]);
  }
  const unsigned act = __ballot_sync(FULL, still_active);
  if (lane == 0 && n_active != nullptr && act != 0u) atomicAdd(n_active, __popc(act));
}

// Phase B: one thread per (queued trajectory, alpha in {.5, .25, .1, .05, .01}), COSTS ONLY.  The alpha is the fast
// grid index, so the five candidates of the same 32 trajectories are dispatched together and meet their nominal and
// gain lines in L2.  (Parking the five candidates in HBM and copying the winner out made this stage HBM-bound:
// 1.4 GB per launch at 16 384 trajectories, profiles/r2_*; the winner is re-rolled instead, by phase W.)
template <int INTEG>
__global__ void __launch_bounds__(32, 8)
ls_qd_phaseB_kernel(const __grid_constant__ QdK k, const double* __restrict__ x_in, const double* __restrict__ u_in,
                    const float* __restrict__ gains, QdWork ws, int B, int N) {
  constexpr int NA = QX_NUM_ALPHA - 1;
  __shared__ QdStage stage;
  const int nq = ws.ctl[0];
  const int qb = blockIdx.x / NA;
  if (qb * 32 >= nq) return;
  const int item = qb * 32 + threadIdx.x;
  const bool run = item < nq;
  const int b = run ? ws.queue[item] : 0;
  const int a = 1 + (int)(blockIdx.x % NA);
  const double cur = ws.cur[b];
  const double limit = k.early ? cur : __longlong_as_double(0x7ff0000000000000LL);
  const double cst = qd_candidate<INTEG, false>(k, stage, run, x_in + b, u_in + b, reinterpret_cast<const float4*>(gains) + b,
                                                (size_t)B, kAlphas[a], limit, nullptr, nullptr, 0, N);
  if (run) ws.cand[(size_t)(a - 1) * B + b] = cst;
}

// Phase W: one thread per trajectory.  A pending trajectory takes the first alpha of the stage whose cost passed the
// test (:440-451) and rolls that candidate out again, this time straight into (x_out, u_out) - the same arithmetic
// as the probe, so the same cost - then the stopping rule.  Without an accepted alpha it keeps its nominal (:447-451):
// its thread rolls out alpha = 0, which reproduces the nominal bit for bit (u_hat = u, and x is the rollout of u by
// this very code), so (x_out, u_out) receive it without a separate copy pass.
template <int INTEG>
__global__ void __launch_bounds__(32, 8)
ls_qd_phaseW_kernel(const __grid_constant__ QdK k, double* __restrict__ x_in, double* __restrict__ u_in,
                    const float* __restrict__ gains, double* __restrict__ cost, int32_t* __restrict__ status,
                    int32_t* __restrict__ iters, int32_t* __restrict__ alpha_out, double* __restrict__ x_out,
                    double* __restrict__ u_out, int32_t* __restrict__ n_active, QdWork ws, double tol, uint32_t flags,
                    int B, int N) {
  constexpr int NA = QX_NUM_ALPHA - 1;
  __shared__ QdStage stage;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool mine = (b < B) && (ws.pend[b] != 0);
  if (!__any_sync(FULL, mine)) return;
  const int bs = mine ? b : 0;
  const double cur = mine ? ws.cur[b] : 0.0;
  int idx = -1;
  double alpha = 0.0, new_cost = cur;
  if (mine) {
#pragma unroll
    for (int a = NA; a >= 1; --a) {
      const double cc = ws.cand[(size_t)(a - 1) * B + b];
      if (cc <= cur) {                     // the FIRST accepted alpha, "<=" (quattro_ilqr_tf.py:444)
        idx = a;
        alpha = kAlphas[a];
        new_cost = cc;
      }
    }
  }
  qd_candidate<INTEG, true>(k, stage, mine, x_in + bs, u_in + bs, reinterpret_cast<const float4*>(gains) + bs, (size_t)B, alpha,
                            __longlong_as_double(0x7ff0000000000000LL), x_out + bs, u_out + bs, (size_t)B, N);
  bool still_active = false;
  if (mine) {
    int mir = 0;
    still_active = qd_finalize<false>(b, idx, cur, new_cost, tol, flags, x_in, u_in, x_out, u_out, cost, status, iters,
                                      alpha_out, &mir, B, N);
    ws.mirror[b] = mir;
  }
  const unsigned act = __ballot_sync(FULL, still_active);
  if (lane == 0 && n_active != nullptr && act != 0u) atomicAdd(n_active, __popc(act));
}

// The final iterate of the trajectories that finished in this line search, copied (x_out,u_out) -> (x_in,u_in) so
// that both buffers hold it: thread per trajectory (coalesced rows), blockIdx.y splits the rows, loads batched.
__global__ void __launch_bounds__(128)
ls_qd_mirror_kernel(double* __restrict__ x_in, double* __restrict__ u_in, const double* __restrict__ x_out,
                    const double* __restrict__ u_out, const int32_t* __restrict__ mirror, int B, int N) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || mirror[b] == 0) return;
  const int xrows = (N + 1) * kN, rows = xrows + N * kM;
  const int per = (rows + (int)gridDim.y - 1) / (int)gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  constexpr int U = 8;
  for (int r = r0; r < r1; r += U) {
    double v[U];
#pragma unroll
    for (int q = 0; q < U; ++q)
      if (r + q < r1) v[q] = (r + q < xrows) ? x_out[(size_t)(r + q) * B + b] : u_out[(size_t)(r + q - xrows) * B + b];
#pragma unroll
    for (int q = 0; q < U; ++q)
      if (r + q < r1) {
        if (r + q < xrows) x_in[(size_t)(r + q) * B + b] = v[q];
        else u_in[(size_t)(r + q - xrows) * B + b] = v[q];
      }
  }
}

// forward_pass with a caller-chosen alpha (quattro_ilqr_tf.py:377-390): the same candidate arithmetic as the
// line search, so qx_forward_pass(alpha) reproduces the candidate the search evaluated bit for bit
template <int INTEG>
__global__ void __launch_bounds__(32, 8)
forward_pass_qd_kernel(const __grid_constant__ QdK k, const double* __restrict__ x_in, const double* __restrict__ u_in,
                       const float* __restrict__ gains, double alpha, double* __restrict__ x_out, double* __restrict__ u_out,
                       double* __restrict__ cost_out, int B, int N) {
  __shared__ QdStage stage;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool run = b < B;
  const int bs = run ? b : 0;
  const double cst = qd_candidate<INTEG, true>(k, stage, run, x_in + bs, u_in + bs, reinterpret_cast<const float4*>(gains) + bs, (size_t)B,
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

// Optional probe (qx_linesearch_stage_events): four caller-owned events recorded on the launching stream in front of
// stage A and behind stages A, B, W, so that a benchmark can time the three launches of one call separately.
static thread_local cudaEvent_t g_stage_ev[4] = {nullptr, nullptr, nullptr, nullptr};

void qd_set_stage_events(void* const* ev) {
  for (int i = 0; i < 4; ++i) g_stage_ev[i] = ev ? static_cast<cudaEvent_t>(ev[i]) : nullptr;
}

static inline void qd_stage_mark(int i, cudaStream_t st) {
  if (g_stage_ev[i] != nullptr) cudaEventRecord(g_stage_ev[i], st);
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
  const QdWork ws = qd_carve(workspace, B);
  cudaError_t e = cudaMemsetAsync(ws.ctl, 0, 32, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(ctl)");
  qd_stage_mark(0, st);
  ls_qd_phaseA_kernel<1, INTEG><<<grid, 32, 0, st>>>(k, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out,
                                                     n_active, ws, tol, flags, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseA_kernel<1>");
  qd_stage_mark(1, st);
  ls_qd_phaseB_kernel<INTEG><<<grid * (QX_NUM_ALPHA - 1), 32, 0, st>>>(k, x_in, u_in, gains, ws, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseB_kernel");
  qd_stage_mark(2, st);
  ls_qd_phaseW_kernel<INTEG><<<grid, 32, 0, st>>>(k, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out, n_active, ws,
                                                  tol, flags, B, N);
  QX_CHECK_LAUNCH("ls_qd_phaseW_kernel");
  qd_stage_mark(3, st);
  if (!(flags & QX_LS_FORCE_ACTIVE)) {      // forced mode never finishes a trajectory: nothing to mirror
    ls_qd_mirror_kernel<<<dim3(ceil_div(B, 128), 16), 128, 0, st>>>(x_in, u_in, x_out, u_out, ws.mirror, B, N);
    QX_CHECK_LAUNCH("ls_qd_mirror_kernel");
  }
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
