// Forward rollouts: nominal simulate+cost, and the fused line search.
//
//   rollout_cost_kernel  = iLQR_TF.simulate + compute_total_cost   (quattro_ilqr_tf.py:127-143)
//   linesearch_kernel    = the alpha loop of optimize (:440-451) around forward_pass (:377-390)
//                          plus the stopping rule (:472), one launch for the whole batch.
//
// One thread owns one trajectory (a rollout is a strictly sequential recursion over t); the
// batch index is the fastest-varying index of x/u so every global access of a warp is one
// fully used 256-B (fp64) segment.  All arithmetic is fp64: the accept test compares costs
// whose relative difference is routinely 1e-7..1e-9 near convergence (SURVEY.md hard part 3).
//
// Gains in the REC layout ([t][b][52] floats, written by the warp-per-trajectory Riccati kernel)
// are staged per warp through shared memory with 1-D bulk async copies (TMA, SASS UBLKCP): for a
// fixed t the 32 records of a warp are one contiguous 6656-B chunk; a 4-stage mbarrier ring keeps
// the copy of step t+3 in flight while step t is computed.
#include "qx_async.cuh"
#include "qx_models.cuh"

namespace qx {

template <class Mdl>
__global__ void __launch_bounds__(128)
rollout_cost_kernel(const __grid_constant__ QxParams p, const double* __restrict__ x0,
                    const double* __restrict__ u, double* __restrict__ x, double* __restrict__ cost,
                    const int32_t* __restrict__ status, int B, int N) {
  constexpr int n = Mdl::n, m = Mdl::m;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (status != nullptr && status[b] != 0) return;
  double xs[n], us[m], xn[n];
#pragma unroll
  for (int i = 0; i < n; ++i) xs[i] = x0[(size_t)i * B + b];
  double c = 0.0;
  for (int t = 0; t < N; ++t) {
#pragma unroll
    for (int j = 0; j < m; ++j) us[j] = u[((size_t)t * m + j) * B + b];
#pragma unroll
    for (int i = 0; i < n; ++i) x[((size_t)t * n + i) * B + b] = xs[i];
    c += running_cost<Mdl, double>(xs, us, p);
    discrete_step<Mdl, double>(xs, us, p, xn);
#pragma unroll
    for (int i = 0; i < n; ++i) xs[i] = xn[i];
  }
#pragma unroll
  for (int i = 0; i < n; ++i) x[((size_t)N * n + i) * B + b] = xs[i];
  c += final_cost<Mdl, double>(xs, p);
  cost[b] = c;
}

// ------------------------------------------------------------------------------------------
constexpr int kLsStages = 4;

template <class Mdl>
struct LsSmem {
  static constexpr int G = Rec<Mdl::n, Mdl::m>::gain;
  alignas(128) float ring[kLsStages][32 * G];
  alignas(8) uint64_t bar[kLsStages];
};

template <class Mdl>
__global__ void __launch_bounds__(32)
linesearch_kernel(const __grid_constant__ QxParams p, double* __restrict__ x_in, double* __restrict__ u_in,
                  const float* __restrict__ gains, double* __restrict__ cost, int32_t* __restrict__ status,
                  int32_t* __restrict__ iters, int32_t* __restrict__ alpha_out, double* __restrict__ x_out,
                  double* __restrict__ u_out, int32_t* __restrict__ n_active, double tol, uint32_t flags,
                  int B, int N) {
  constexpr int n = Mdl::n, m = Mdl::m;
  constexpr int G = Rec<n, m>::gain;
  constexpr bool kRec = (Mdl::layout == QX_LAYOUT_REC);
  constexpr unsigned FULL = 0xffffffffu;

  const int lane = threadIdx.x;
  const int b0 = blockIdx.x * 32;
  const int b = b0 + lane;
  const bool mine = (b < B) && (status[b] == 0);
  if (!__any_sync(FULL, mine)) return;
  const int nb = min(32, B - b0);                       // trajectories covered by this warp

  __shared__ LsSmem<Mdl> sm;
  uint32_t seq_issue = 0, seq_wait = 0;                 // ring positions (warp-uniform)
  if (kRec) {
    if (lane == 0) {
      for (int s = 0; s < kLsStages; ++s) mbar_init(&sm.bar[s], 1);
      mbar_fence_init();
    }
    __syncwarp();
  }
  auto issue = [&](int t) {                             // lane 0 only
    const int s = seq_issue % kLsStages;
    const uint32_t bytes = (uint32_t)nb * G * sizeof(float);
    mbar_arrive_expect_tx(&sm.bar[s], bytes);
    bulk_g2s(sm.ring[s], gains + ((size_t)t * B + b0) * G, bytes, &sm.bar[s]);
  };

  const double cur = mine ? cost[b] : 0.0;
  double x_first[n];
#pragma unroll
  for (int i = 0; i < n; ++i) x_first[i] = mine ? x_in[(size_t)i * B + b] : 0.0;

  bool accepted = !mine;          // lanes without work never roll out
  int idx = -1;
  double new_cost = cur;

  for (int a = 0; a < QX_NUM_ALPHA; ++a) {
    if (!__any_sync(FULL, !accepted)) break;
    const bool run = !accepted;
    const double alpha = kAlphas[a];
    double xh[n], xn[n], un[m], uh[m], xnext[n];
    double xn_nx[n], un_nx[m];        // nominal of step t+1, fetched while step t is computed
    float g_nx[kRec ? 1 : G];         // SOA gains of step t+1 (the REC layout comes through the ring)
#pragma unroll
    for (int i = 0; i < n; ++i) xh[i] = x_first[i];
    double c = 0.0;

    if (kRec) {
      if (lane == 0)
        for (int t = 0; t < kLsStages - 1 && t < N; ++t) { issue(t); ++seq_issue; }
      seq_issue = __shfl_sync(FULL, seq_issue, 0);
    }
    auto fetch = [&](int t) {
#pragma unroll
      for (int i = 0; i < n; ++i) xn_nx[i] = x_in[((size_t)t * n + i) * B + b];
#pragma unroll
      for (int j = 0; j < m; ++j) un_nx[j] = u_in[((size_t)t * m + j) * B + b];
      if (!kRec) {
#pragma unroll
        for (int q = 0; q < G; ++q) g_nx[q] = gains[((size_t)t * G + q) * B + b];
      }
    };
    if (run) fetch(0);
    for (int t = 0; t < N; ++t) {
      if (kRec) {
        // all lanes finished reading the stage that is about to be refilled (step t-1's)
        __syncwarp();
        if (t + kLsStages - 1 < N) {
          if (lane == 0) issue(t + kLsStages - 1);
          ++seq_issue;
        }
        mbar_wait(&sm.bar[seq_wait % kLsStages], (seq_wait / kLsStages) & 1u);
      }
      if (run) {
        float g[G];
#pragma unroll
        for (int i = 0; i < n; ++i) xn[i] = xn_nx[i];
#pragma unroll
        for (int j = 0; j < m; ++j) un[j] = un_nx[j];
        if (kRec) {
          const float4* r4 = reinterpret_cast<const float4*>(&sm.ring[seq_wait % kLsStages][lane * G]);
#pragma unroll
          for (int q = 0; q < G / 4; ++q) {
            const float4 v = r4[q];
            g[4 * q + 0] = v.x; g[4 * q + 1] = v.y; g[4 * q + 2] = v.z; g[4 * q + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int q = 0; q < G; ++q) g[q] = g_nx[q];
        }
        if (t + 1 < N) fetch(t + 1);
        // du = k + K (x_hat - x);  u_hat = u + alpha * du   (alpha scales the feedback too)
#pragma unroll
        for (int j = 0; j < m; ++j) {
          double kd = 0.0;
#pragma unroll
          for (int i = 0; i < n; ++i) kd += (double)g[m + j * n + i] * (xh[i] - xn[i]);
          const double du = (double)g[j] + kd;
          uh[j] = un[j] + alpha * du;
        }
#pragma unroll
        for (int i = 0; i < n; ++i) x_out[((size_t)t * n + i) * B + b] = xh[i];
#pragma unroll
        for (int j = 0; j < m; ++j) u_out[((size_t)t * m + j) * B + b] = uh[j];
        c += running_cost<Mdl, double>(xh, uh, p);
        discrete_step<Mdl, double>(xh, uh, p, xnext);
#pragma unroll
        for (int i = 0; i < n; ++i) xh[i] = xnext[i];
      }
      if (kRec) ++seq_wait;
    }
    if (run) {
#pragma unroll
      for (int i = 0; i < n; ++i) x_out[((size_t)N * n + i) * B + b] = xh[i];
      c += final_cost<Mdl, double>(xh, p);
      if (c <= cur) {              // first accepted alpha, "<=" (quattro_ilqr_tf.py:444)
        accepted = true;
        idx = a;
        new_cost = c;
      }
    }
  }

  bool still_active = false;
  if (mine) {
    const bool found = idx >= 0;
    bool done = (!found) || (fabs(cur - new_cost) < tol);        // quattro_ilqr_tf.py:472
    if (flags & QX_LS_FORCE_ACTIVE) done = false;
    if (!found) {
      // keep the nominal: (x_out,u_out) <- (x_in,u_in)
      for (int t = 0; t <= N; ++t)
#pragma unroll
        for (int i = 0; i < n; ++i) x_out[((size_t)t * n + i) * B + b] = x_in[((size_t)t * n + i) * B + b];
      for (int t = 0; t < N; ++t)
#pragma unroll
        for (int j = 0; j < m; ++j) u_out[((size_t)t * m + j) * B + b] = u_in[((size_t)t * m + j) * B + b];
    } else if (done) {
      // finished: leave the final iterate in both buffers so the caller may keep ping-ponging
      for (int t = 0; t <= N; ++t)
#pragma unroll
        for (int i = 0; i < n; ++i) x_in[((size_t)t * n + i) * B + b] = x_out[((size_t)t * n + i) * B + b];
      for (int t = 0; t < N; ++t)
#pragma unroll
        for (int j = 0; j < m; ++j) u_in[((size_t)t * m + j) * B + b] = u_out[((size_t)t * m + j) * B + b];
    }
    cost[b] = new_cost;
    alpha_out[b] = idx;
    iters[b] += 1;
    if (done) status[b] = 1;
    still_active = !done;
  }
  const unsigned act = __ballot_sync(FULL, still_active);
  if (lane == 0 && n_active != nullptr && act != 0u) atomicAdd(n_active, __popc(act));
}

// ------------------------------------------------------------------------------------------
// Quadrotor line search, FOUR lanes per trajectory.
//
// A rollout step is ~1.3 k dependent fp64 instructions (three sincos, a division, four
// exp/log1p of the softplus barrier); with one thread per trajectory a 16 K batch gives every SM
// sub-partition a single warp, so the kernel runs at the latency of that chain.  Here lane g of a
// 4-lane group owns control g (its gain row, its barrier term), the angle min(g,2) (its sincos)
// and three state rows of the cost; u, trig values and cost partials are exchanged with
// warp shuffles and every lane keeps a full copy of the state.  The chain per step shrinks ~6x,
// the batch supplies 4x more warps, and an alpha retry now stalls 8 trajectories instead of 32.
// Used when Q, R, Qf are diagonal (the reference's MPC classes); otherwise linesearch_kernel.
struct QuadConst {
  double inv_mass, g, arm, kyaw, invIx, invIy, invIz, c1, c2, c3, dt, beta, inv_beta, alpha_b;
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// continuous dynamics, cooperative over the 4 lanes of a group (base = first lane of the group)
__device__ __forceinline__ void quad_fc4(const double* x, double u_own, const QuadConst& k, int base, int g,
                                         double* d) {
  double s, c;
  ::sincos(x[6 + (g < 2 ? g : 2)], &s, &c);
  const double sec_own = 1.0 / c;
  const double u0 = shfl_d(u_own, base), u1 = shfl_d(u_own, base + 1), u2 = shfl_d(u_own, base + 2),
               u3 = shfl_d(u_own, base + 3);
  const double sphi = shfl_d(s, base), cphi = shfl_d(c, base);
  const double sth = shfl_d(s, base + 1), cth = shfl_d(c, base + 1), sec = shfl_d(sec_own, base + 1);
  const double spsi = shfl_d(s, base + 2), cpsi = shfl_d(c, base + 2);
  const double pp = x[9], q = x[10], r = x[11];
  const double Tm = (u0 + u1 + u2 + u3) * k.inv_mass;
  const double tth = sth * sec;
  d[0] = x[3];
  d[1] = x[4];
  d[2] = x[5];
  d[3] = Tm * (spsi * sphi + cpsi * sth * cphi);
  d[4] = Tm * (cpsi * sphi - spsi * sth * cphi);
  d[5] = -k.g + Tm * (cth * cphi);
  d[6] = pp + q * sphi * tth + r * cphi * tth;
  d[7] = q * cphi - r * sphi;
  d[8] = (q * sphi + r * cphi) * sec;
  d[9] = k.c1 * (q * r) + k.arm * ((u1 + u2) - (u0 + u3)) * k.invIx;
  d[10] = k.c2 * (pp * r) + k.arm * ((u0 + u1) - (u2 + u3)) * k.invIy;
  d[11] = k.c3 * (pp * q) + k.kyaw * (u0 - u1 + u2 - u3) * k.invIz;
}

__device__ __forceinline__ void quad_step4(const double* x, double u_own, const QuadConst& k, int integrator, int base,
                                           int g, double* xn) {
  double k1[12];
  quad_fc4(x, u_own, k, base, g, k1);
  if (integrator == QX_EULER) {
#pragma unroll
    for (int i = 0; i < 12; ++i) xn[i] = x[i] + k.dt * k1[i];
    return;
  }
  double k2[12], k3[12], k4[12], xs[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) xs[i] = x[i] + 0.5 * k.dt * k1[i];
  quad_fc4(xs, u_own, k, base, g, k2);
#pragma unroll
  for (int i = 0; i < 12; ++i) xs[i] = x[i] + 0.5 * k.dt * k2[i];
  quad_fc4(xs, u_own, k, base, g, k3);
#pragma unroll
  for (int i = 0; i < 12; ++i) xs[i] = x[i] + k.dt * k3[i];
  quad_fc4(xs, u_own, k, base, g, k4);
#pragma unroll
  for (int i = 0; i < 12; ++i) xn[i] = x[i] + (k.dt / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
}

// alpha_b * softplus(-u; beta)^2 ; log1p(e) by its series when e is tiny (u well inside the feasible side)
__device__ __forceinline__ double quad_barrier1(double u, const QuadConst& k) {
  const double bz = -k.beta * u;
  const double e = ::exp(-fabs(bz));
  const double l1p = (e < 1e-4) ? e * (1.0 - e * (0.5 - e * (1.0 / 3.0 - 0.25 * e))) : ::log1p(e);
  const double sp = (fmax(bz, 0.0) + l1p) * k.inv_beta;
  return k.alpha_b * sp * sp;
}

// nominal rollout + cost with exactly the arithmetic of the 4-lane line search, so that a candidate
// with du == 0 reproduces the nominal cost bit for bit (as the reference's shared f/L code does)
__device__ __forceinline__ void quad_consts(const QxParams& p, QuadConst& k) {
  k.inv_mass = 1.0 / p.phys[0];
  k.invIx = 1.0 / p.phys[1]; k.invIy = 1.0 / p.phys[2]; k.invIz = 1.0 / p.phys[3];
  k.arm = p.phys[4]; k.g = p.phys[5]; k.kyaw = p.phys[6];
  k.c1 = (p.phys[2] - p.phys[3]) / p.phys[1];
  k.c2 = (p.phys[3] - p.phys[1]) / p.phys[2];
  k.c3 = (p.phys[1] - p.phys[2]) / p.phys[3];
  k.dt = p.dt; k.beta = p.barrier_beta; k.inv_beta = 1.0 / p.barrier_beta; k.alpha_b = p.barrier_alpha;
}

__global__ void __launch_bounds__(128)
rollout_quad4_kernel(const __grid_constant__ QxParams p, const double* __restrict__ x0, const double* __restrict__ u,
                     double* __restrict__ x, double* __restrict__ cost, const int32_t* __restrict__ status, int B, int N) {
  constexpr int n = 12, m = 4;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, g = lane & 3, base = lane & ~3;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const bool run = (b < B) && (status == nullptr || status[b] == 0);
  if (!__any_sync(FULL, run)) return;
  QuadConst k;
  quad_consts(p, k);
  double qd[3], qfd[3], xr[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    qd[r] = p.Q[(3 * g + r) * n + 3 * g + r];
    qfd[r] = p.Qf[(3 * g + r) * n + 3 * g + r];
    xr[r] = p.x_ref[3 * g + r];
  }
  const double rd = p.R[g * m + g];
  double xh[n], xnx[n];
#pragma unroll
  for (int i = 0; i < n; ++i) xh[i] = run ? x0[(size_t)i * B + b] : 0.0;
  double c = 0.0;
  double u_nx = run ? u[(size_t)g * B + b] : 0.0;
  for (int t = 0; t < N; ++t) {
    const double uh = u_nx;
    if (run) {
      if (t + 1 < N) u_nx = u[((size_t)(t + 1) * m + g) * B + b];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double xv = (g == 0) ? xh[r] : (g == 1) ? xh[3 + r] : (g == 2) ? xh[6 + r] : xh[9 + r];
        x[((size_t)t * n + 3 * g + r) * B + b] = xv;
        const double dxv = xv - xr[r];
        c += qd[r] * dxv * dxv;
      }
      c += rd * uh * uh + quad_barrier1(uh, k);
    }
    quad_step4(xh, uh, k, p.integrator, base, g, xnx);
#pragma unroll
    for (int i = 0; i < n; ++i) xh[i] = xnx[i];
  }
  if (run) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double xv = (g == 0) ? xh[r] : (g == 1) ? xh[3 + r] : (g == 2) ? xh[6 + r] : xh[9 + r];
      x[((size_t)N * n + 3 * g + r) * B + b] = xv;
      const double dxv = xv - xr[r];
      c += qfd[r] * dxv * dxv;
    }
  }
  c += __shfl_xor_sync(FULL, c, 1);
  c += __shfl_xor_sync(FULL, c, 2);
  if (run && g == 0) cost[b] = c;
}

// Per-lane constants of the 4-lane scheme: this lane's share of the diagonal cost.
struct Quad4Lane {
  QuadConst k;
  double qd[3], qfd[3], xr[3], rd;
  int g, base;
};
__device__ __forceinline__ void quad4_lane_init(const QxParams& p, int lane, Quad4Lane& c) {
  constexpr int n = 12, m = 4;
  c.g = lane & 3;
  c.base = lane & ~3;
  quad_consts(p, c.k);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    c.qd[r] = p.Q[(3 * c.g + r) * n + 3 * c.g + r];
    c.qfd[r] = p.Qf[(3 * c.g + r) * n + 3 * c.g + r];
    c.xr[r] = p.x_ref[3 * c.g + r];
  }
  c.rd = p.R[c.g * m + c.g];
}

// One candidate rollout (forward_pass :377-390) by a 4-lane group: nominal and gains of trajectory
// b_src, candidate written at column b_dst of (x_dst, u_dst) (both with row stride B).  All 32 lanes of
// the warp must call it (shuffles); lanes with run == false only take part in the shuffles.
// Returns the total cost (identical bits in the four lanes of a group).
__device__ __forceinline__ double quad4_candidate(const QxParams& p, const Quad4Lane& c, bool run,
                                                  const double* __restrict__ x_in, const double* __restrict__ u_in,
                                                  const float* __restrict__ gains, int b_src, double alpha,
                                                  double* __restrict__ x_dst, double* __restrict__ u_dst, int b_dst,
                                                  int B, int N) {
  constexpr int n = 12, m = 4, G = 52;
  constexpr unsigned FULL = 0xffffffffu;
  const int g = c.g;
  double xh[n], xnx[n], xn_nx[n], un_nx = 0.0;
  float gk_nx = 0.f;
  float4 gK_nx[3];
#pragma unroll
  for (int i = 0; i < n; ++i) xh[i] = run ? x_in[(size_t)i * B + b_src] : 0.0;
  double cst = 0.0;
  auto fetch = [&](int t) {
#pragma unroll
    for (int i = 0; i < n; ++i) xn_nx[i] = x_in[((size_t)t * n + i) * B + b_src];
    un_nx = u_in[((size_t)t * m + g) * B + b_src];
    const float* rec = gains + ((size_t)t * B + b_src) * G;
    gk_nx = rec[g];
    const float4* row = reinterpret_cast<const float4*>(rec + m + g * n);
    gK_nx[0] = row[0]; gK_nx[1] = row[1]; gK_nx[2] = row[2];
  };
  if (run) fetch(0);
  for (int t = 0; t < N; ++t) {
    double uh = 0.0;
    if (run) {
      double xn[n];
#pragma unroll
      for (int i = 0; i < n; ++i) xn[i] = xn_nx[i];
      const double un = un_nx;
      const float gk = gk_nx;
      const float kr[12] = {gK_nx[0].x, gK_nx[0].y, gK_nx[0].z, gK_nx[0].w, gK_nx[1].x, gK_nx[1].y,
                            gK_nx[1].z, gK_nx[1].w, gK_nx[2].x, gK_nx[2].y, gK_nx[2].z, gK_nx[2].w};
      if (t + 1 < N) fetch(t + 1);
      double kd = 0.0;
#pragma unroll
      for (int i = 0; i < n; ++i) kd += (double)kr[i] * (xh[i] - xn[i]);
      uh = un + alpha * ((double)gk + kd);                 // alpha scales the feedback term too
      u_dst[((size_t)t * m + g) * B + b_dst] = uh;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double xv = (g == 0) ? xh[r] : (g == 1) ? xh[3 + r] : (g == 2) ? xh[6 + r] : xh[9 + r];
        x_dst[((size_t)t * n + 3 * g + r) * B + b_dst] = xv;
        const double dxv = xv - c.xr[r];
        cst += c.qd[r] * dxv * dxv;
      }
      cst += c.rd * uh * uh + quad_barrier1(uh, c.k);
    }
    quad_step4(xh, uh, c.k, p.integrator, c.base, g, xnx);
#pragma unroll
    for (int i = 0; i < n; ++i) xh[i] = xnx[i];
  }
  if (run) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double xv = (g == 0) ? xh[r] : (g == 1) ? xh[3 + r] : (g == 2) ? xh[6 + r] : xh[9 + r];
      x_dst[((size_t)N * n + 3 * g + r) * B + b_dst] = xv;
      const double dxv = xv - c.xr[r];
      cst += c.qfd[r] * dxv * dxv;
    }
  }
  // group total (identical bits in the four lanes: the butterfly adds commute)
  cst += __shfl_xor_sync(FULL, cst, 1);
  cst += __shfl_xor_sync(FULL, cst, 2);
  return cst;
}

// Bookkeeping after the line search of one trajectory (stopping rule :472), executed by its 4-lane group.
// Leaves the next iterate in (x_out,u_out); a finished trajectory has it in both buffers.
__device__ __forceinline__ bool quad4_finalize(int g, int b, int idx, double cur, double new_cost, double tol,
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

// Control block of the two-phase line search (device int32[4], zeroed by the caller before phase A).
enum { kCtlActive = 0, kCtlQueue = 1 };

// MODE 0: the whole alpha list sequentially with early exit (single launch).
// MODE 1: phase A of the two-phase search: alpha = 1 only; rejected trajectories are queued.
template <int MODE>
__global__ void __launch_bounds__(128)
linesearch_quad4_kernel(const __grid_constant__ QxParams p, double* __restrict__ x_in, double* __restrict__ u_in,
                        const float* __restrict__ gains, double* __restrict__ cost, int32_t* __restrict__ status,
                        int32_t* __restrict__ iters, int32_t* __restrict__ alpha_out, double* __restrict__ x_out,
                        double* __restrict__ u_out, int32_t* __restrict__ n_active, int32_t* __restrict__ queue,
                        int32_t* __restrict__ ctl, double tol, uint32_t flags, int B, int N) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const bool mine = (b < B) && (status[b] == 0);
  if (!__any_sync(FULL, mine)) return;
  Quad4Lane c;
  quad4_lane_init(p, lane, c);
  const double cur = mine ? cost[b] : 0.0;
  bool accepted = !mine;
  int idx = -1;
  double new_cost = cur;
  constexpr int n_alpha = (MODE == 0) ? QX_NUM_ALPHA : 1;
  for (int a = 0; a < n_alpha; ++a) {
    if (!__any_sync(FULL, !accepted)) break;
    const bool run = !accepted;
    const double cst = quad4_candidate(p, c, run, x_in, u_in, gains, b, kAlphas[a], x_out, u_out, b, B, N);
    if (run && cst <= cur) {       // first accepted alpha, "<=" (quattro_ilqr_tf.py:444)
      accepted = true;
      idx = a;
      new_cost = cst;
    }
  }
  bool still_active = false;
  if (mine) {
    if (MODE == 1 && idx < 0) {
      if (c.g == 0) queue[atomicAdd(&ctl[kCtlQueue], 1)] = b;      // phase B tries the other alphas in parallel
    } else {
      still_active = quad4_finalize(c.g, b, idx, cur, new_cost, tol, flags, x_in, u_in, x_out, u_out, cost, status,
                                    iters, alpha_out, B, N) && c.g == 0;
    }
  }
  const unsigned act = __ballot_sync(FULL, still_active);
  if (lane == 0 && n_active != nullptr && act != 0u) atomicAdd(n_active, __popc(act));
}

// Phase B: every queued trajectory x every remaining alpha is one 4-lane work item (alpha-major, so the
// eight groups of a warp share the alpha and write eight consecutive scratch columns).
__global__ void __launch_bounds__(128)
linesearch_quad4_phaseB_kernel(const __grid_constant__ QxParams p, const double* __restrict__ x_in,
                               const double* __restrict__ u_in, const float* __restrict__ gains,
                               const int32_t* __restrict__ queue, const int32_t* __restrict__ ctl,
                               double* __restrict__ xs, double* __restrict__ us, double* __restrict__ cand_cost, int B,
                               int N) {
  constexpr int n = 12, m = 4;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int nrej = ctl[kCtlQueue];
  const long item = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const bool run = item < (long)(QX_NUM_ALPHA - 1) * nrej;
  if (!__any_sync(FULL, run)) return;
  Quad4Lane c;
  quad4_lane_init(p, lane, c);
  const int a = run ? (int)(item / nrej) : 0;            // 0..4 -> alpha index a + 1
  const int slot = run ? (int)(item % nrej) : 0;
  const int b = run ? queue[slot] : 0;
  double* xd = xs + (size_t)a * (N + 1) * n * B;
  double* ud = us + (size_t)a * N * m * B;
  const double cst = quad4_candidate(p, c, run, x_in, u_in, gains, b, kAlphas[a + 1], xd, ud, slot, B, N);
  if (run && c.g == 0) cand_cost[(size_t)a * B + slot] = cst;
}

// Phase C: first accepted alpha per queued trajectory, copy its candidate out of the scratch, bookkeeping.
__global__ void __launch_bounds__(128)
linesearch_quad4_phaseC_kernel(double* __restrict__ x_in, double* __restrict__ u_in, double* __restrict__ cost,
                               int32_t* __restrict__ status, int32_t* __restrict__ iters,
                               int32_t* __restrict__ alpha_out, double* __restrict__ x_out, double* __restrict__ u_out,
                               int32_t* __restrict__ n_active, const int32_t* __restrict__ queue,
                               const int32_t* __restrict__ ctl, const double* __restrict__ xs,
                               const double* __restrict__ us, const double* __restrict__ cand_cost, double tol,
                               uint32_t flags, int B, int N) {
  constexpr int n = 12, m = 4;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, g = lane & 3;
  const int nrej = ctl[kCtlQueue];
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const bool mine = slot < nrej;
  if (!__any_sync(FULL, mine)) return;
  bool still_active = false;
  if (mine) {
    const int b = queue[slot];
    const double cur = cost[b];
    int idx = -1;
    double new_cost = cur;
    for (int a = 0; a < QX_NUM_ALPHA - 1; ++a) {
      const double cc = cand_cost[(size_t)a * B + slot];
      if (cc <= cur) { idx = a + 1; new_cost = cc; break; }
    }
    if (idx > 0) {
      const double* xsrc = xs + (size_t)(idx - 1) * (N + 1) * n * B;
      const double* usrc = us + (size_t)(idx - 1) * N * m * B;
      for (int t = 0; t <= N; ++t)
#pragma unroll
        for (int r = 0; r < 3; ++r)
          x_out[((size_t)t * n + 3 * g + r) * B + b] = xsrc[((size_t)t * n + 3 * g + r) * B + slot];
      for (int t = 0; t < N; ++t) u_out[((size_t)t * m + g) * B + b] = usrc[((size_t)t * m + g) * B + slot];
    }
    still_active = quad4_finalize(g, b, idx, cur, new_cost, tol, flags, x_in, u_in, x_out, u_out, cost, status, iters,
                                  alpha_out, B, N) && g == 0;
  }
  const unsigned act = __ballot_sync(FULL, still_active);
  if (lane == 0 && n_active != nullptr && act != 0u) atomicAdd(n_active, __popc(act));
}

// ------------------------------------------------------------------------------------------
// One candidate rollout with a caller-chosen alpha: iLQR_TF.forward_pass (quattro_ilqr_tf.py:377-390)
// as a stand-alone entry point (API completeness; the solver itself uses the fused line search).
template <class Mdl>
__global__ void __launch_bounds__(128)
forward_pass_kernel(const __grid_constant__ QxParams p, const double* __restrict__ x_in, const double* __restrict__ u_in,
                    const float* __restrict__ gains, double alpha, double* __restrict__ x_out,
                    double* __restrict__ u_out, double* __restrict__ cost_out, int B, int N) {
  constexpr int n = Mdl::n, m = Mdl::m, G = Rec<n, m>::gain;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double xh[n], xn[n], uh[m], xnext[n];
#pragma unroll
  for (int i = 0; i < n; ++i) xh[i] = x_in[(size_t)i * B + b];
  double c = 0.0;
  for (int t = 0; t < N; ++t) {
#pragma unroll
    for (int i = 0; i < n; ++i) xn[i] = x_in[((size_t)t * n + i) * B + b];
#pragma unroll
    for (int j = 0; j < m; ++j) {
      auto gv = [&](int q) {
        return (double)((Mdl::layout == QX_LAYOUT_REC) ? gains[((size_t)t * B + b) * G + q]
                                                       : gains[((size_t)t * G + q) * B + b]);
      };
      double kd = 0.0;
#pragma unroll
      for (int i = 0; i < n; ++i) kd += gv(m + j * n + i) * (xh[i] - xn[i]);
      uh[j] = u_in[((size_t)t * m + j) * B + b] + alpha * (gv(j) + kd);
      u_out[((size_t)t * m + j) * B + b] = uh[j];
    }
#pragma unroll
    for (int i = 0; i < n; ++i) x_out[((size_t)t * n + i) * B + b] = xh[i];
    c += running_cost<Mdl, double>(xh, uh, p);
    discrete_step<Mdl, double>(xh, uh, p, xnext);
#pragma unroll
    for (int i = 0; i < n; ++i) xh[i] = xnext[i];
  }
#pragma unroll
  for (int i = 0; i < n; ++i) x_out[((size_t)N * n + i) * B + b] = xh[i];
  cost_out[b] = c + final_cost<Mdl, double>(xh, p);
}

// ------------------------------------------------------------------------------------------
template <class Mdl>
static int launch_rollout(const QxParams* p, const double* x0, const double* u, double* x, double* cost,
                          const int32_t* status, int B, int N, cudaStream_t st) {
  const int threads = 128;
  rollout_cost_kernel<Mdl><<<ceil_div(B, threads), threads, 0, st>>>(*p, x0, u, x, cost, status, B, N);
  QX_CHECK_LAUNCH("rollout_cost_kernel");
  return QX_OK;
}

template <class Mdl>
static int launch_linesearch(const QxParams* p, double* x_in, double* u_in, const float* gains, double* cost,
                             int32_t* status, int32_t* iters, int32_t* alpha, double* x_out, double* u_out,
                             int32_t* n_active, double tol, uint32_t flags, int B, int N, cudaStream_t st) {
  linesearch_kernel<Mdl><<<ceil_div(B, 32), 32, 0, st>>>(*p, x_in, u_in, gains, cost, status, iters, alpha,
                                                         x_out, u_out, n_active, tol, flags, B, N);
  QX_CHECK_LAUNCH("linesearch_kernel");
  return QX_OK;
}

// distributed-state 4-lane kernels of qx_quadls.cu (quadrotor, diagonal weights)
size_t qd_workspace_bytes(int B, int N);
int qd_launch_rollout(const QxParams* p, const double* x0, const double* u, double* x, double* cost,
                      const int32_t* status, int B, int N, cudaStream_t st);
int qd_launch_linesearch(const QxParams* p, double* x_in, double* u_in, const float* gains, double* cost,
                         int32_t* status, int32_t* iters, int32_t* alpha, double* x_out, double* u_out,
                         int32_t* n_active, void* workspace, double tol, uint32_t flags, int B, int N,
                         cudaStream_t st);

}  // namespace qx

extern "C" int qx_rollout_cost(const QxParams* p, const double* x0, const double* u, double* x, double* cost,
                               const int32_t* status, int B, int N, void* stream) {
  QX_CHECK_ARG(p && x0 && u && x && cost, "null pointer");
  QX_CHECK_ARG(B > 0 && N > 0, "B and N must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (p->model_id) {
    case QX_MODEL_CARTPOLE: return qx::launch_rollout<qx::CartPole>(p, x0, u, x, cost, status, B, N, st);
    case QX_MODEL_QUADROTOR:
      if (p->diag_cost) return qx::qd_launch_rollout(p, x0, u, x, cost, status, B, N, st);   // same arithmetic as the 4-lane line search
      return qx::launch_rollout<qx::Quadrotor>(p, x0, u, x, cost, status, B, N, st);
  }
  qx::set_error("unknown model_id %d", p->model_id);
  return QX_EINVAL;
}

extern "C" size_t qx_linesearch_workspace_bytes(int model_id, int B, int N) {
  if (model_id != QX_MODEL_QUADROTOR || B <= 0 || N <= 0) return 0;
  return qx::qd_workspace_bytes(B, N);
}

extern "C" int qx_linesearch(const QxParams* p, double* x_in, double* u_in, const float* gains, double* cost,
                             int32_t* status, int32_t* iters, int32_t* alpha, double* x_out, double* u_out,
                             int32_t* n_active, void* workspace, double tol, uint32_t flags, int B, int N,
                             void* stream) {
  QX_CHECK_ARG(p && x_in && u_in && gains && cost && status && iters && alpha && x_out && u_out, "null pointer");
  QX_CHECK_ARG(B > 0 && N > 0, "B and N must be positive");
  QX_CHECK_ARG(x_in != x_out && u_in != u_out, "in and out buffers must differ");
  QX_CHECK_ARG((reinterpret_cast<uintptr_t>(gains) & 15u) == 0, "gains must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (p->model_id) {
    case QX_MODEL_CARTPOLE:
      return qx::launch_linesearch<qx::CartPole>(p, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out,
                                                 n_active, tol, flags, B, N, st);
    case QX_MODEL_QUADROTOR: {
      if (!p->diag_cost)
        return qx::launch_linesearch<qx::Quadrotor>(p, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out,
                                                    n_active, tol, flags, B, N, st);
      return qx::qd_launch_linesearch(p, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out, n_active, workspace,
                                      tol, flags, B, N, st);
    }
  }
  qx::set_error("unknown model_id %d", p->model_id);
  return QX_EINVAL;
}

extern "C" int qx_forward_pass(const QxParams* p, const double* x_in, const double* u_in, const float* gains,
                               double alpha, double* x_out, double* u_out, double* cost_out, int B, int N,
                               void* stream) {
  QX_CHECK_ARG(p && x_in && u_in && gains && x_out && u_out && cost_out, "null pointer");
  QX_CHECK_ARG(B > 0 && N > 0, "B and N must be positive");
  QX_CHECK_ARG(x_in != x_out && u_in != u_out, "in and out buffers must differ");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (p->model_id) {
    case QX_MODEL_CARTPOLE:
      qx::forward_pass_kernel<qx::CartPole><<<qx::ceil_div(B, 128), 128, 0, st>>>(*p, x_in, u_in, gains, alpha, x_out, u_out,
                                                                                 cost_out, B, N);
      break;
    case QX_MODEL_QUADROTOR:
      qx::forward_pass_kernel<qx::Quadrotor><<<qx::ceil_div(B, 128), 128, 0, st>>>(*p, x_in, u_in, gains, alpha, x_out, u_out,
                                                                                  cost_out, B, N);
      break;
    default:
      qx::set_error("unknown model_id %d", p->model_id);
      return QX_EINVAL;
  }
  QX_CHECK_LAUNCH("forward_pass_kernel");
  return QX_OK;
}
