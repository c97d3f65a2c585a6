// Forward rollouts, generic models: nominal simulate+cost, and the fused line search (thread per trajectory).
// The quadrotor with diagonal weights takes the thread-per-candidate kernels of qx_quadls.cu instead.
//
//   rollout_cost_kernel  = iLQR_TF.simulate + compute_total_cost   (quattro_ilqr_tf.py:127-143)
//   linesearch_kernel    = the alpha loop of optimize (:440-451) around forward_pass (:377-390)
//                          plus the stopping rule (:472), one launch for the whole batch.
//
// One thread owns one trajectory (a rollout is a strictly sequential recursion over t); the
// batch index is the fastest-varying index of x, u and the gains, so every global access of a warp is a
// run of whole lines.  All arithmetic is fp64: the accept test compares costs
// whose relative difference is routinely 1e-7..1e-9 near convergence (SURVEY.md hard part 3).
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
// gains of step t for trajectory b into registers (both layouts have b fastest: coalesced)
template <class Mdl>
__device__ __forceinline__ void load_gains(const float* __restrict__ gains, int t, int b, int B, float* g) {
  constexpr int G = Rec<Mdl::n, Mdl::m>::gain;
  if (Mdl::layout == QX_LAYOUT_REC) {
    const float4* g4 = reinterpret_cast<const float4*>(gains) + (size_t)t * (G / 4) * B + b;
#pragma unroll
    for (int q = 0; q < G / 4; ++q) {
      const float4 v = g4[(size_t)q * B];
      g[4 * q + 0] = v.x; g[4 * q + 1] = v.y; g[4 * q + 2] = v.z; g[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int q = 0; q < G; ++q) g[q] = gains[((size_t)t * G + q) * B + b];
  }
}

template <class Mdl>
__global__ void __launch_bounds__(32)
linesearch_kernel(const __grid_constant__ QxParams p, double* __restrict__ x_in, double* __restrict__ u_in,
                  const float* __restrict__ gains, double* __restrict__ cost, int32_t* __restrict__ status,
                  int32_t* __restrict__ iters, int32_t* __restrict__ alpha_out, double* __restrict__ x_out,
                  double* __restrict__ u_out, int32_t* __restrict__ n_active, double tol, uint32_t flags,
                  int B, int N) {
  constexpr int n = Mdl::n, m = Mdl::m;
  constexpr int G = Rec<n, m>::gain;
  constexpr unsigned FULL = 0xffffffffu;

  const int lane = threadIdx.x;
  const int b = blockIdx.x * 32 + lane;
  const bool mine = (b < B) && (status[b] == 0);
  if (!__any_sync(FULL, mine)) return;

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
    float g_nx[G];                    // gains of step t+1
#pragma unroll
    for (int i = 0; i < n; ++i) xh[i] = x_first[i];
    double c = 0.0;

    auto fetch = [&](int t) {
#pragma unroll
      for (int i = 0; i < n; ++i) xn_nx[i] = x_in[((size_t)t * n + i) * B + b];
#pragma unroll
      for (int j = 0; j < m; ++j) un_nx[j] = u_in[((size_t)t * m + j) * B + b];
      load_gains<Mdl>(gains, t, b, B, g_nx);
    };
    if (run) fetch(0);
    for (int t = 0; t < N; ++t) {
      if (run) {
        float g[G];
#pragma unroll
        for (int i = 0; i < n; ++i) xn[i] = xn_nx[i];
#pragma unroll
        for (int j = 0; j < m; ++j) un[j] = un_nx[j];
#pragma unroll
        for (int q = 0; q < G; ++q) g[q] = g_nx[q];
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
// Parallel-alpha line search for the SOA layout (cart-pole).  The kernel above walks the alpha list inside one thread:
// up to six dependent rollouts of N sequential steps on a grid of B/32 warps - a few warps per SM, pure latency
// (0.12 ms at 4 096 cart-pole trajectories when smaller steps are needed).  Here the six candidates of a trajectory
// are rolled out SIDE BY SIDE: CTA = 32 trajectories x 6 warps, warp a takes alpha_a and computes the candidate's
// cost; warp 0 (alpha = 1, the usual winner) also stores its candidate.  The first accepted alpha (:440-451, "<=") is
// picked through shared memory.  Trajectories that take alpha = 1 or none are finished on the spot; the others are
// re-rolled with their winner by linesearch_hint_kernel, which stores it: one or two rollout latencies instead of up to
// six.  The file is compiled with -fmad=false and the candidate arithmetic is the sequential kernel's, so costs,
// decisions and trajectories are bit-identical to it (tests: both paths against the oracle and against each other).
template <class Mdl, bool STORE>
__device__ __forceinline__ double ls_candidate_soa(const QxParams& p, const double* __restrict__ x_in,
                                                   const double* __restrict__ u_in, const float* __restrict__ gains,
                                                   double* __restrict__ x_out, double* __restrict__ u_out, double alpha,
                                                   int b, int B, int N) {
  constexpr int n = Mdl::n, m = Mdl::m;
  constexpr int G = Rec<n, m>::gain;
  double xh[n], xn[n], un[m], uh[m], xnext[n], xn_nx[n], un_nx[m];
  float g[G], g_nx[G];
  auto fetch = [&](int t) {
#pragma unroll
    for (int i = 0; i < n; ++i) xn_nx[i] = x_in[((size_t)t * n + i) * B + b];
#pragma unroll
    for (int j = 0; j < m; ++j) un_nx[j] = u_in[((size_t)t * m + j) * B + b];
#pragma unroll
    for (int q = 0; q < G; ++q) g_nx[q] = gains[((size_t)t * G + q) * B + b];
  };
  fetch(0);
#pragma unroll
  for (int i = 0; i < n; ++i) xh[i] = xn_nx[i];          // x_hat_0 = x_0
  double c = 0.0;
  for (int t = 0; t < N; ++t) {
#pragma unroll
    for (int i = 0; i < n; ++i) xn[i] = xn_nx[i];
#pragma unroll
    for (int j = 0; j < m; ++j) un[j] = un_nx[j];
#pragma unroll
    for (int q = 0; q < G; ++q) g[q] = g_nx[q];
    if (t + 1 < N) fetch(t + 1);
#pragma unroll
    for (int j = 0; j < m; ++j) {
      double kd = 0.0;
#pragma unroll
      for (int i = 0; i < n; ++i) kd += (double)g[m + j * n + i] * (xh[i] - xn[i]);
      const double du = (double)g[j] + kd;
      uh[j] = un[j] + alpha * du;
    }
    if (STORE) {
#pragma unroll
      for (int i = 0; i < n; ++i) x_out[((size_t)t * n + i) * B + b] = xh[i];
#pragma unroll
      for (int j = 0; j < m; ++j) u_out[((size_t)t * m + j) * B + b] = uh[j];
    }
    c += running_cost<Mdl, double>(xh, uh, p);
    discrete_step<Mdl, double>(xh, uh, p, xnext);
#pragma unroll
    for (int i = 0; i < n; ++i) xh[i] = xnext[i];
  }
  if (STORE) {
#pragma unroll
    for (int i = 0; i < n; ++i) x_out[((size_t)N * n + i) * B + b] = xh[i];
  }
  return c + final_cost<Mdl, double>(xh, p);
}

// Bookkeeping of one trajectory once its alpha is known (the tail of linesearch_kernel): stopping rule (:472), the
// nominal kept when no alpha was accepted, the final iterate mirrored into both buffers.  Returns "still active".
template <class Mdl>
__device__ __forceinline__ bool ls_finish_soa(double* __restrict__ x_in, double* __restrict__ u_in, double* __restrict__ x_out,
                                              double* __restrict__ u_out, double* __restrict__ cost, int32_t* __restrict__ status,
                                              int32_t* __restrict__ iters, int32_t* __restrict__ alpha_out, int idx, double cur,
                                              double new_cost, double tol, uint32_t flags, int b, int B, int N) {
  constexpr int n = Mdl::n, m = Mdl::m;
  const bool found = idx >= 0;
  bool done = (!found) || (fabs(cur - new_cost) < tol);
  if (flags & QX_LS_FORCE_ACTIVE) done = false;
  if (!found) {
    for (int t = 0; t <= N; ++t)
#pragma unroll
      for (int i = 0; i < n; ++i) x_out[((size_t)t * n + i) * B + b] = x_in[((size_t)t * n + i) * B + b];
    for (int t = 0; t < N; ++t)
#pragma unroll
      for (int j = 0; j < m; ++j) u_out[((size_t)t * m + j) * B + b] = u_in[((size_t)t * m + j) * B + b];
  } else if (done) {
    for (int t = 0; t <= N; ++t)
#pragma unroll
      for (int i = 0; i < n; ++i) x_in[((size_t)t * n + i) * B + b] = x_out[((size_t)t * n + i) * B + b];
    for (int t = 0; t < N; ++t)
#pragma unroll
      for (int j = 0; j < m; ++j) u_in[((size_t)t * m + j) * B + b] = u_out[((size_t)t * m + j) * B + b];
  }
  cost[b] = found ? new_cost : cur;
  alpha_out[b] = idx;
  iters[b] += 1;
  if (done) status[b] = 1;
  return !done;
}

template <class Mdl>
__global__ void __launch_bounds__(32 * QX_NUM_ALPHA)
linesearch_par_kernel(const __grid_constant__ QxParams p, double* x_in, double* u_in, const float* __restrict__ gains,
                      double* __restrict__ cost, int32_t* __restrict__ status, int32_t* __restrict__ iters,
                      int32_t* __restrict__ alpha_out, double* x_out, double* u_out, int32_t* __restrict__ n_active,
                      int32_t* __restrict__ hint, double tol, uint32_t flags, int B, int N) {
  static_assert(Mdl::layout == QX_LAYOUT_SOA, "parallel-alpha search: SOA layout only");
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ double cand[QX_NUM_ALPHA][32];
  const int lane = threadIdx.x, a = threadIdx.y;
  const int b = blockIdx.x * 32 + lane;
  const bool mine = (b < B) && (status[b] == 0);
  if (mine) {
    cand[a][lane] = (a == 0) ? ls_candidate_soa<Mdl, true>(p, x_in, u_in, gains, x_out, u_out, kAlphas[0], b, B, N)
                             : ls_candidate_soa<Mdl, false>(p, x_in, u_in, gains, x_out, u_out, kAlphas[a], b, B, N);
  }
  __syncthreads();
  if (a != 0) return;
  bool still_active = false;
  if (b < B) {
    int pending = 0;
    if (mine) {
      const double cur = cost[b];
      int idx = -1;
#pragma unroll
      for (int q = QX_NUM_ALPHA - 1; q >= 0; --q)
        if (cand[q][lane] <= cur) idx = q;                 // the FIRST accepted alpha, "<=" (quattro_ilqr_tf.py:444)
      if (idx <= 0)
        still_active = ls_finish_soa<Mdl>(x_in, u_in, x_out, u_out, cost, status, iters, alpha_out, idx, cur, cand[0][lane], tol,
                                          flags, b, B, N);
      else
        pending = idx;
    }
    hint[b] = pending;
  }
  const unsigned act = __ballot_sync(FULL, still_active);
  if (lane == 0 && n_active != nullptr && act != 0u) atomicAdd(n_active, __popc(act));
}

template <class Mdl>
__global__ void __launch_bounds__(32)
linesearch_hint_kernel(const __grid_constant__ QxParams p, double* x_in, double* u_in, const float* __restrict__ gains,
                       double* __restrict__ cost, int32_t* __restrict__ status, int32_t* __restrict__ iters,
                       int32_t* __restrict__ alpha_out, double* x_out, double* u_out, int32_t* __restrict__ n_active,
                       const int32_t* __restrict__ hint, double tol, uint32_t flags, int B, int N) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x;
  const int b = blockIdx.x * 32 + lane;
  const int idx = (b < B) ? hint[b] : 0;
  if (!__any_sync(FULL, idx > 0)) return;
  bool still_active = false;
  if (idx > 0) {
    const double cur = cost[b];
    const double c = ls_candidate_soa<Mdl, true>(p, x_in, u_in, gains, x_out, u_out, kAlphas[idx], b, B, N);
    // same arithmetic as the probe, so c <= cur holds again; were it ever not to, the nominal is kept
    still_active = ls_finish_soa<Mdl>(x_in, u_in, x_out, u_out, cost, status, iters, alpha_out, c <= cur ? idx : -1, cur, c, tol,
                                      flags, b, B, N);
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
      auto gv = [&](int q) { return (double)gains[gain_index<Mdl::layout>(t, q, b, B, G)]; };
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

// SOA layout with a workspace (int32 hint per trajectory): the six alphas side by side, then the winners
template <class Mdl>
static int launch_linesearch_par(const QxParams* p, double* x_in, double* u_in, const float* gains, double* cost,
                                 int32_t* status, int32_t* iters, int32_t* alpha, double* x_out, double* u_out,
                                 int32_t* n_active, int32_t* hint, double tol, uint32_t flags, int B, int N, cudaStream_t st) {
  linesearch_par_kernel<Mdl><<<ceil_div(B, 32), dim3(32, QX_NUM_ALPHA), 0, st>>>(*p, x_in, u_in, gains, cost, status, iters, alpha,
                                                                                  x_out, u_out, n_active, hint, tol, flags, B, N);
  QX_CHECK_LAUNCH("linesearch_par_kernel");
  linesearch_hint_kernel<Mdl><<<ceil_div(B, 32), 32, 0, st>>>(*p, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out,
                                                              n_active, hint, tol, flags, B, N);
  QX_CHECK_LAUNCH("linesearch_hint_kernel");
  return QX_OK;
}

// thread-per-candidate kernels of qx_quadls.cu (quadrotor, diagonal weights)
size_t qd_workspace_bytes(int B, int N);
int qd_launch_rollout(const QxParams* p, const double* x0, const double* u, double* x, double* cost,
                      const int32_t* status, int B, int N, cudaStream_t st);
int qd_launch_forward_pass(const QxParams* p, const double* x_in, const double* u_in, const float* gains, double alpha,
                           double* x_out, double* u_out, double* cost_out, int B, int N, cudaStream_t st);
void qd_set_stage_events(void* const* ev);
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
      if (p->diag_cost) return qx::qd_launch_rollout(p, x0, u, x, cost, status, B, N, st);   // same arithmetic as the line search
      return qx::launch_rollout<qx::Quadrotor>(p, x0, u, x, cost, status, B, N, st);
  }
  qx::set_error("unknown model_id %d", p->model_id);
  return QX_EINVAL;
}

extern "C" size_t qx_linesearch_workspace_bytes(int model_id, int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  if (model_id == QX_MODEL_CARTPOLE) return ((size_t)B * sizeof(int32_t) + 127) & ~(size_t)127;   // winner hints
  if (model_id != QX_MODEL_QUADROTOR) return 0;
  return qx::qd_workspace_bytes(B, N);
}

extern "C" int qx_linesearch_stage_events(void* const* events, int count) {
  QX_CHECK_ARG(count == 0 || (count == 4 && events != nullptr), "pass four cudaEvent_t handles, or count = 0 to clear");
  qx::qd_set_stage_events(count == 4 ? events : nullptr);
  return QX_OK;
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
      if (workspace != nullptr)       // six alphas side by side + winners (two launches); NULL: the sequential kernel
        return qx::launch_linesearch_par<qx::CartPole>(p, x_in, u_in, gains, cost, status, iters, alpha, x_out, u_out, n_active,
                                                       static_cast<int32_t*>(workspace), tol, flags, B, N, st);
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
      if (p->diag_cost) return qx::qd_launch_forward_pass(p, x_in, u_in, gains, alpha, x_out, u_out, cost_out, B, N, st);
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
