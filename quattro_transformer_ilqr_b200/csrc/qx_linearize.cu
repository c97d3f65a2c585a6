// Linearisation of a whole horizon in one launch.
//
//   linearize_kernel = _compute_dynamics_jacobians (quattro_ilqr_tf.py:182-204)
//                    + _compute_cost_derivatives   (:217-275)
//                    + _finite_diff_gradient_final / _finite_diff_hessian_final (:149-174)
// with closed-form derivatives instead of the reference's central differences.
//
// One thread per (t, b).  The kernel is a pure HBM-write stream (quadrotor: 3 248 B read,
// 83 824 B written per trajectory), so the only thing that matters is that the stores are full
// 128-B lines:
//   * SOA layout (cart-pole): the record index is slower than b, every store is coalesced as is;
//   * REC layout (quadrotor): a thread produces one contiguous 1 664-B record; each warp stages
//     half a record per lane (208 floats, row pitch 209 -> conflict-free) in shared memory and
//     streams it out as 128-B row segments.
#include "qx_models.cuh"

namespace qx {

// ---- quadrotor, RK4: the chain rule through the four stages on 3x3 BLOCKS --------------------------------------------
// With x = [p | v | eta | omega] the continuous Jacobian is block upper triangular,
//        | 0  I  0   0  |           | 0      |
//    J = | 0  0  Gc  0  |     Gu =  | dc 1^T |      (Gc, Hc, Ec, Sc, dc: the non-zero blocks of Quadrotor::fc_jac)
//        | 0  0  Hc  Ec |           | 0      |
//        | 0  0  0   Sc |           | Tc     |
// and so is every D_s = d k_s / d x = J_s (I + h D_{s-1}): eight blocks, of which D_pv = I always.  The dense
// formulation below (12x12 products on arrays that live in local memory) took 2.5 ms per launch at 16 384 trajectories
// (50 x the Euler kernel); on blocks the same chain rule is ~1 200 FMAs per (t, b) in registers.
__device__ __forceinline__ void quad_fc_blocks(const float* x, const float* u, const QxParams& p, float Gc[3][3],
                                               float Hc[3][3], float Ec[3][3], float Sc[3][3], float* dc) {
  const float mass = float(p.phys[0]), Ix = float(p.phys[1]), Iy = float(p.phys[2]), Iz = float(p.phys[3]);
  float sphi, cphi, sth, cth, spsi, cpsi;
  sincosf(x[6], &sphi, &cphi);
  sincosf(x[7], &sth, &cth);
  sincosf(x[8], &spsi, &cpsi);
  const float pp = x[9], q = x[10], r = x[11];
  const float Tm = (u[0] + u[1] + u[2] + u[3]) / mass, im = 1.f / mass;
  const float sec = 1.f / cth, tth = sth * sec;
  const float dirx = spsi * sphi + cpsi * sth * cphi;
  const float diry = cpsi * sphi - spsi * sth * cphi;
  const float dirz = cth * cphi;
  Gc[0][0] = Tm * (spsi * cphi - cpsi * sth * sphi); Gc[0][1] = Tm * (cpsi * cth * cphi); Gc[0][2] = Tm * diry;
  Gc[1][0] = Tm * (cpsi * cphi + spsi * sth * sphi); Gc[1][1] = Tm * (-spsi * cth * cphi); Gc[1][2] = -Tm * dirx;
  Gc[2][0] = -Tm * cth * sphi; Gc[2][1] = -Tm * sth * cphi; Gc[2][2] = 0.f;
  dc[0] = im * dirx; dc[1] = im * diry; dc[2] = im * dirz;
  const float w = q * sphi + r * cphi, v = q * cphi - r * sphi;
  Hc[0][0] = v * tth; Hc[0][1] = w * sec * sec; Hc[0][2] = 0.f;
  Hc[1][0] = -w; Hc[1][1] = 0.f; Hc[1][2] = 0.f;
  Hc[2][0] = v * sec; Hc[2][1] = w * sth * sec * sec; Hc[2][2] = 0.f;
  Ec[0][0] = 1.f; Ec[0][1] = sphi * tth; Ec[0][2] = cphi * tth;
  Ec[1][0] = 0.f; Ec[1][1] = cphi; Ec[1][2] = -sphi;
  Ec[2][0] = 0.f; Ec[2][1] = sphi * sec; Ec[2][2] = cphi * sec;
  const float c1 = (Iy - Iz) / Ix, c2 = (Iz - Ix) / Iy, c3 = (Ix - Iy) / Iz;
  Sc[0][0] = 0.f; Sc[0][1] = c1 * r; Sc[0][2] = c1 * q;
  Sc[1][0] = c2 * r; Sc[1][1] = 0.f; Sc[1][2] = c2 * pp;
  Sc[2][0] = c3 * q; Sc[2][1] = c3 * pp; Sc[2][2] = 0.f;
}

// Z = X Y (3x3 times 3xC)
template <int C>
__device__ __forceinline__ void mul3(const float X[3][3], const float (*Y)[C], float (*Z)[C]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) Z[i][j] = fmaf(X[i][2], Y[2][j], fmaf(X[i][1], Y[1][j], X[i][0] * Y[0][j]));
}

__device__ __forceinline__ void quad_rk4_jacobians(const float* x, const float* u, const QxParams& p, float* A, float* Bm) {
  constexpr int n = 12, m = 4;
  const float dt = float(p.dt);
  const float aix = float(p.phys[4]) / float(p.phys[1]), aiy = float(p.phys[4]) / float(p.phys[2]);
  const float kiz = float(p.phys[6]) / float(p.phys[3]);
  const float Tc[3][4] = {{-aix, aix, aix, -aix}, {aiy, aiy, -aiy, -aiy}, {kiz, -kiz, kiz, -kiz}};
  // D blocks (d k_s / d x) and their weighted sums; Du blocks (d k_s / d u, 3x4 each) and theirs
  float Dpe[3][3] = {}, Dpw[3][3] = {}, Dve[3][3] = {}, Dvw[3][3] = {}, Dee[3][3] = {}, Dew[3][3] = {}, Dww[3][3] = {};
  float Ape[3][3] = {}, Apw[3][3] = {}, Ave[3][3] = {}, Avw[3][3] = {}, Aee[3][3] = {}, Aew[3][3] = {}, Aww[3][3] = {};
  float Dup[3][4] = {}, Duv[3][4] = {}, Due[3][4] = {}, Duw[3][4] = {};
  float Bp[3][4] = {}, Bv[3][4] = {}, Be[3][4] = {}, Bw[3][4] = {};
  float k[n], xs[n];
#pragma unroll
  for (int i = 0; i < n; ++i) k[i] = 0.f;
  const float wgt[4] = {1.f, 2.f, 2.f, 1.f}, cs[4] = {0.f, 0.5f, 0.5f, 1.f};
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const float h = cs[s] * dt, w = wgt[s];
#pragma unroll
    for (int i = 0; i < n; ++i) xs[i] = x[i] + h * k[i];
    float Gc[3][3], Hc[3][3], Ec[3][3], Sc[3][3], dc[3];
    quad_fc_blocks(xs, u, p, Gc, Hc, Ec, Sc, dc);
    // rows of M = I + h D_{s-1} that J touches
    float Tee[3][3], Tew[3][3], Tww[3][3], Npe[3][3], Npw[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        Npe[i][j] = h * Dve[i][j];
        Npw[i][j] = h * Dvw[i][j];
        Tee[i][j] = (i == j ? 1.f : 0.f) + h * Dee[i][j];
        Tew[i][j] = h * Dew[i][j];
        Tww[i][j] = (i == j ? 1.f : 0.f) + h * Dww[i][j];
      }
    float t1[3][3], t2[3][3];
    mul3<3>(Gc, Tee, Dve);
    mul3<3>(Gc, Tew, Dvw);
    mul3<3>(Hc, Tew, t1);
    mul3<3>(Ec, Tww, t2);
    mul3<3>(Hc, Tee, Dee);
    mul3<3>(Sc, Tww, Dww);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        Dew[i][j] = t1[i][j] + t2[i][j];
        Dpe[i][j] = Npe[i][j];
        Dpw[i][j] = Npw[i][j];
        Ape[i][j] += w * Dpe[i][j]; Apw[i][j] += w * Dpw[i][j];
        Ave[i][j] += w * Dve[i][j]; Avw[i][j] += w * Dvw[i][j];
        Aee[i][j] += w * Dee[i][j]; Aew[i][j] += w * Dew[i][j];
        Aww[i][j] += w * Dww[i][j];
      }
    // d k_s / d u = J_s (h D_{s-1,u}) + Gu
    float He[3][4], Hw[3][4], u1[3][4], u2[3][4], u3[3][4], u4[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        He[i][j] = h * Due[i][j];
        Hw[i][j] = h * Duw[i][j];
        Dup[i][j] = h * Duv[i][j];
      }
    mul3<4>(Gc, He, u1);
    mul3<4>(Hc, He, u2);
    mul3<4>(Ec, Hw, u3);
    mul3<4>(Sc, Hw, u4);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Duv[i][j] = u1[i][j] + dc[i];
        Due[i][j] = u2[i][j] + u3[i][j];
        Duw[i][j] = u4[i][j] + Tc[i][j];
        Bp[i][j] += w * Dup[i][j]; Bv[i][j] += w * Duv[i][j];
        Be[i][j] += w * Due[i][j]; Bw[i][j] += w * Duw[i][j];
      }
    Quadrotor::fc<float>(xs, u, p, k);
  }
  const float s6 = dt / 6.f;
#pragma unroll
  for (int i = 0; i < n * n; ++i) A[i] = 0.f;
#pragma unroll
  for (int i = 0; i < n; ++i) A[i * n + i] = 1.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    A[i * n + 3 + i] = s6 * 6.f;                        // D_pv = I in every stage
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      A[i * n + 6 + j] = s6 * Ape[i][j];
      A[i * n + 9 + j] = s6 * Apw[i][j];
      A[(3 + i) * n + 6 + j] = s6 * Ave[i][j];
      A[(3 + i) * n + 9 + j] = s6 * Avw[i][j];
      A[(6 + i) * n + 6 + j] = (i == j ? 1.f : 0.f) + s6 * Aee[i][j];
      A[(6 + i) * n + 9 + j] = s6 * Aew[i][j];
      A[(9 + i) * n + 9 + j] = (i == j ? 1.f : 0.f) + s6 * Aww[i][j];
    }
#pragma unroll
    for (int j = 0; j < m; ++j) {
      Bm[i * m + j] = s6 * Bp[i][j];
      Bm[(3 + i) * m + j] = s6 * Bv[i][j];
      Bm[(6 + i) * m + j] = s6 * Be[i][j];
      Bm[(9 + i) * m + j] = s6 * Bw[i][j];
    }
  }
}

// A = d f / d x, Bm = d f / d u of the DISCRETE map, row-major, in float.
template <class Mdl, int INTEG>
__device__ __forceinline__ void discrete_jacobians(const float* x, const float* u, const QxParams& p, float* A,
                                                   float* Bm) {
  constexpr int n = Mdl::n, m = Mdl::m;
  const float dt = float(p.dt);
  if (INTEG == QX_RK4 && Mdl::id == QX_MODEL_QUADROTOR) {
    quad_rk4_jacobians(x, u, p, A, Bm);
    return;
  }
  if (INTEG == QX_EULER) {
    Mdl::template fc_jac<float>(x, u, p, A, Bm);
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int j = 0; j < n; ++j) A[i * n + j] = (i == j ? 1.f : 0.f) + dt * A[i * n + j];
#pragma unroll
    for (int i = 0; i < n * m; ++i) Bm[i] = dt * Bm[i];
    return;
  }
  // RK4: chain rule through the four stages.  D_s = d k_s / d(x,u).
  float k[n], xs[n], J[n * n], G[n * m], Dx[n * n], Du[n * m], Tx[n * n], Tu[n * m];
  const float w[4] = {1.f, 2.f, 2.f, 1.f};
  const float c[4] = {0.f, 0.5f, 0.5f, 1.f};
#pragma unroll
  for (int i = 0; i < n * n; ++i) { A[i] = 0.f; Dx[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < n * m; ++i) { Bm[i] = 0.f; Du[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < n; ++i) k[i] = 0.f;
  for (int s = 0; s < 4; ++s) {
    const float h = c[s] * dt;
    for (int i = 0; i < n; ++i) xs[i] = x[i] + h * k[i];
    Mdl::template fc_jac<float>(xs, u, p, J, G);
    // D_s,x = J (I + h D_{s-1},x),  D_s,u = J (h D_{s-1},u) + G
    for (int i = 0; i < n; ++i) {
      for (int j = 0; j < n; ++j) {
        float acc = J[i * n + j];
        for (int l = 0; l < n; ++l) acc += h * J[i * n + l] * Dx[l * n + j];
        Tx[i * n + j] = acc;
      }
      for (int j = 0; j < m; ++j) {
        float acc = G[i * m + j];
        for (int l = 0; l < n; ++l) acc += h * J[i * n + l] * Du[l * m + j];
        Tu[i * m + j] = acc;
      }
    }
    for (int i = 0; i < n * n; ++i) { Dx[i] = Tx[i]; A[i] += w[s] * Tx[i]; }
    for (int i = 0; i < n * m; ++i) { Du[i] = Tu[i]; Bm[i] += w[s] * Tu[i]; }
    Mdl::template fc<float>(xs, u, p, k);
  }
  const float s6 = dt / 6.f;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) A[i * n + j] = (i == j ? 1.f : 0.f) + s6 * A[i * n + j];
  for (int i = 0; i < n * m; ++i) Bm[i] = s6 * Bm[i];
}

// Sinks: where element e of the (t,b) record goes.
struct SoaSink {
  float* base; size_t stride;   // base + e*stride
  __device__ __forceinline__ void operator()(int e, float v) const { base[(size_t)e * stride] = v; }
};
struct StageSink {
  float* row;                   // this lane's staging row
  __device__ __forceinline__ void operator()(int e, float v) const { row[e] = v; }
};

template <class Mdl, int INTEG, class Sink>
__device__ __forceinline__ void emit_dynamics(const float* x, const float* u, const QxParams& p, Sink out) {
  constexpr int n = Mdl::n, m = Mdl::m;
  using R = Rec<n, m>;
  float A[n * n], Bm[n * m];
  discrete_jacobians<Mdl, INTEG>(x, u, p, A, Bm);
#pragma unroll
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < n; ++j) out(R::a(i, j), A[i * n + j]);
#pragma unroll
    for (int j = 0; j < m; ++j) out(R::b(i, j), Bm[i * m + j]);
  }
  float dx[n], g[n];
#pragma unroll
  for (int i = 0; i < n; ++i) dx[i] = x[i] - float(p.x_ref[i]);
  quad_grad<float, n>(p.Q, dx, p.diag_cost, g);
#pragma unroll
  for (int i = 0; i < n; ++i) out(R::lx + i, g[i]);
  float gu[m], bg[m], bh[m];
  quad_grad<float, m>(p.R, u, p.diag_cost, gu);
  Mdl::template barrier_d<float>(u, p, bg, bh);
#pragma unroll
  for (int j = 0; j < m; ++j) out(R::lu + j, gu[j] + bg[j]);
}

// second half of the record: l_xx, l_uu, l_ux (offset so that it starts at 0 when off = R::lxx)
template <class Mdl, class Sink>
__device__ __forceinline__ void emit_hessians(const float* u, const QxParams& p, Sink out, int off) {
  constexpr int n = Mdl::n, m = Mdl::m;
  using R = Rec<n, m>;
  float bg[m], bh[m];
  Mdl::template barrier_d<float>(u, p, bg, bh);
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int j = 0; j < n; ++j) out(R::lxx - off + i * n + j, float(p.Q[i * n + j] + p.Q[j * n + i]));
#pragma unroll
  for (int i = 0; i < m; ++i)
#pragma unroll
    for (int j = 0; j < m; ++j)
      out(R::luu - off + i * m + j, float(p.R[i * m + j] + p.R[j * m + i]) + (i == j ? bh[i] : 0.f));
#pragma unroll
  for (int i = 0; i < m * n; ++i) out(R::lux - off + i, 0.f);
}

template <class Mdl, class Sink>
__device__ __forceinline__ void emit_terminal(const float* x, const QxParams& p, Sink out) {
  constexpr int n = Mdl::n;
  float dx[n], g[n];
#pragma unroll
  for (int i = 0; i < n; ++i) dx[i] = x[i] - float(p.x_ref[i]);
  quad_grad<float, n>(p.Qf, dx, p.diag_cost, g);
#pragma unroll
  for (int i = 0; i < n; ++i) out(i, g[i]);
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int j = 0; j < n; ++j) out(n + i * n + j, float(p.Qf[i * n + j] + p.Qf[j * n + i]));
}

template <class Mdl>
__device__ __forceinline__ void load_xu(const double* __restrict__ x, const double* __restrict__ u, int t, int b,
                                        int B, float* xs, float* us) {
#pragma unroll
  for (int i = 0; i < Mdl::n; ++i) xs[i] = float(x[((size_t)t * Mdl::n + i) * B + b]);
#pragma unroll
  for (int j = 0; j < Mdl::m; ++j) us[j] = float(u[((size_t)t * Mdl::m + j) * B + b]);
}

// ---- SOA: every store already coalesced --------------------------------------------------------
template <class Mdl, int INTEG>
__global__ void __launch_bounds__(128)
linearize_soa_kernel(const __grid_constant__ QxParams p, const double* __restrict__ x, const double* __restrict__ u,
                     float* __restrict__ expansion, float* __restrict__ terminal,
                     const int32_t* __restrict__ status, int B, int N, int t_start) {
  constexpr int n = Mdl::n, m = Mdl::m;
  using R = Rec<n, m>;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = t_start + blockIdx.y;
  if (b >= B) return;
  if (status != nullptr && status[b] != 0) return;
  float xs[n], us[m];
  load_xu<Mdl>(x, u, t, b, B, xs, us);
  SoaSink sink{expansion + (size_t)t * R::size * B + b, (size_t)B};
  emit_dynamics<Mdl, INTEG>(xs, us, p, sink);
  emit_hessians<Mdl>(us, p, sink, 0);
  if (t == N - 1) {
#pragma unroll
    for (int i = 0; i < n; ++i) xs[i] = float(x[((size_t)N * n + i) * B + b]);
    emit_terminal<Mdl>(xs, p, SoaSink{terminal + b, (size_t)B});
  }
}

// ---- REC: stage half records per warp, stream out full lines -----------------------------------
constexpr int kLinWarps = 1;

template <class Mdl, int INTEG>
__global__ void __launch_bounds__(32 * kLinWarps)
linearize_rec_kernel(const __grid_constant__ QxParams p, const double* __restrict__ x, const double* __restrict__ u,
                     float* __restrict__ expansion, float* __restrict__ terminal,
                     const int32_t* __restrict__ status, int B, int N, int t_start) {
  constexpr int n = Mdl::n, m = Mdl::m;
  using R = Rec<n, m>;
  constexpr int HALF = R::lxx;                 // first half = [A | B | lx | lu]
  static_assert(HALF * 2 == R::size, "record must split in two equal halves");
  static_assert(R::term <= HALF, "terminal record must fit the staging row");
  constexpr int PITCH = HALF + 1;              // odd pitch: lane-strided writes hit 32 distinct banks
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ float stage[kLinWarps][32 * PITCH];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = (blockIdx.x * kLinWarps + warp) * 32;
  const int b = b0 + lane;
  const int t = t_start + blockIdx.y;
  const bool mine = (b < B) && (status == nullptr || status[b] == 0);
  const unsigned live = __ballot_sync(FULL, mine);
  if (live == 0u) return;
  float* row = &stage[warp][lane * PITCH];
  float xs[n], us[m];
  if (mine) load_xu<Mdl>(x, u, t, b, B, xs, us);

  auto flush = [&](float* dst0, size_t rec_stride, int count) {
    // dst0: destination of lane 0's chunk; record r of the warp goes to dst0 + r * rec_stride
    __syncwarp();
    for (int r = 0; r < 32; ++r) {
      if (!((live >> r) & 1u)) continue;
      const float* src = &stage[warp][r * PITCH];
      float* dst = dst0 + (size_t)r * rec_stride;
      for (int e = lane; e < count; e += 32) dst[e] = src[e];
    }
    __syncwarp();
  };

  float* rec0 = expansion + ((size_t)t * B + b0) * R::size;
  if (mine) emit_dynamics<Mdl, INTEG>(xs, us, p, StageSink{row});
  flush(rec0, R::size, HALF);
  if (mine) emit_hessians<Mdl>(us, p, StageSink{row}, HALF);
  flush(rec0 + HALF, R::size, HALF);
  if (t == N - 1) {
    if (mine) {
#pragma unroll
      for (int i = 0; i < n; ++i) xs[i] = float(x[((size_t)N * n + i) * B + b]);
      emit_terminal<Mdl>(xs, p, StageSink{row});
    }
    flush(terminal + (size_t)b0 * R::term, R::term, R::term);
  }
}

// ---- REC, quadrotor + Euler: only the entries that change from record to record are recomputed ----
// Of the 416 floats of a record only 59 depend on (x_t, u_t): 27 entries of A = I + dt J, 12 of B,
// [l_x | l_u] and the 4 diagonal entries of l_uu (the barrier curvature).  Each warp owns 32
// trajectories and a run of consecutive time steps: the first half-record [F | l_x | l_u] lives
// in a per-lane shared-memory row whose constant entries are written once, the second half
// [l_xx | l_uu | l_ux] is a per-warp constant held in registers (two 16-B pieces per lane) and
// patched with the four barrier terms on the way out.  Every global store is 512 contiguous bytes.
constexpr int kLinSteps = 10;      // time steps per warp

__global__ void __launch_bounds__(32)
linearize_quad_euler_kernel(const __grid_constant__ QxParams p, const double* __restrict__ x,
                            const double* __restrict__ u, float* __restrict__ expansion, float* __restrict__ terminal,
                            const int32_t* __restrict__ status, int B, int N, int t_start) {
  using Mdl = Quadrotor;
  constexpr int n = 12, m = 4;
  using R = Rec<n, m>;
  constexpr int HALF = R::lxx;                 // 208 floats = 52 float4
  constexpr int PITCH4 = 53;                   // row pitch in float4 (16-B aligned rows)
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ float4 stage4[32 * PITCH4];
  __shared__ float bh_s[32][4];
  const int lane = threadIdx.x;
  const int b0 = blockIdx.x * 32, b = b0 + lane;
  const int t_lo = t_start + blockIdx.y * kLinSteps, t_hi = min(N, t_lo + kLinSteps);
  const bool mine = (b < B) && (status == nullptr || status[b] == 0);
  const unsigned live = __ballot_sync(FULL, mine);
  if (live == 0u) return;
  float* row = reinterpret_cast<float*>(&stage4[lane * PITCH4]);
  const float dt = float(p.dt);

  // second half-record template [l_xx | l_uu | l_ux]: float4 #lane and #(32 + lane)
  auto tmpl_elem = [&](int e) -> float {       // e in [0, 208)
    if (e < n * n) { const int i = e / n, j = e % n; return float(p.Q[i * n + j] + p.Q[j * n + i]); }
    e -= n * n;
    if (e < m * m) { const int i = e / m, j = e % m; return float(p.R[i * m + j] + p.R[j * m + i]); }
    return 0.f;
  };
  float4 tmpl0, tmpl1;
  tmpl0 = make_float4(tmpl_elem(4 * lane), tmpl_elem(4 * lane + 1), tmpl_elem(4 * lane + 2), tmpl_elem(4 * lane + 3));
  {
    const int e = 4 * (32 + lane);
    tmpl1 = (lane < 20) ? make_float4(tmpl_elem(e), tmpl_elem(e + 1), tmpl_elem(e + 2), tmpl_elem(e + 3))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
  }

  // constant entries of the first half: A = I + dt J and B = dt G where J, G do not depend on (x,u)
  if (mine) {
    for (int e = 0; e < R::lx; ++e) row[e] = 0.f;
#pragma unroll
    for (int i = 0; i < n; ++i) row[R::a(i, i)] = 1.f;
    row[R::a(0, 3)] = dt; row[R::a(1, 4)] = dt; row[R::a(2, 5)] = dt;
    row[R::a(6, 9)] = dt;
    const float aix = dt * float(p.phys[4] / p.phys[1]), aiy = dt * float(p.phys[4] / p.phys[2]);
    const float kiz = dt * float(p.phys[6] / p.phys[3]);
    row[R::b(9, 0)] = -aix; row[R::b(9, 1)] = aix; row[R::b(9, 2)] = aix; row[R::b(9, 3)] = -aix;
    row[R::b(10, 0)] = aiy; row[R::b(10, 1)] = aiy; row[R::b(10, 2)] = -aiy; row[R::b(10, 3)] = -aiy;
    row[R::b(11, 0)] = kiz; row[R::b(11, 1)] = -kiz; row[R::b(11, 2)] = kiz; row[R::b(11, 3)] = -kiz;
  }
  const float mass = float(p.phys[0]), Ix = float(p.phys[1]), Iy = float(p.phys[2]), Iz = float(p.phys[3]);
  const float c1 = (Iy - Iz) / Ix, c2 = (Iz - Ix) / Iy, c3 = (Ix - Iy) / Iz, im = 1.f / mass;

  for (int t = t_lo; t < t_hi; ++t) {
    if (mine) {
      float xs[n], us[m];
      load_xu<Mdl>(x, u, t, b, B, xs, us);
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
      // same expressions as Quadrotor::fc_jac (qx_models.cuh), scaled by dt
      row[R::a(3, 6)] = dt * (Tm * (spsi * cphi - cpsi * sth * sphi));
      row[R::a(3, 7)] = dt * (Tm * (cpsi * cth * cphi));
      row[R::a(3, 8)] = dt * (Tm * diry);
      row[R::a(4, 6)] = dt * (Tm * (cpsi * cphi + spsi * sth * sphi));
      row[R::a(4, 7)] = dt * (Tm * (-spsi * cth * cphi));
      row[R::a(4, 8)] = dt * (-Tm * dirx);
      row[R::a(5, 6)] = dt * (-Tm * cth * sphi);
      row[R::a(5, 7)] = dt * (-Tm * sth * cphi);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        row[R::b(3, j)] = dt * (im * dirx);
        row[R::b(4, j)] = dt * (im * diry);
        row[R::b(5, j)] = dt * (im * dirz);
      }
      const float w = q * sphi + r * cphi, v = q * cphi - r * sphi;
      row[R::a(6, 6)] = 1.f + dt * (v * tth);
      row[R::a(6, 7)] = dt * (w * sec * sec);
      row[R::a(6, 10)] = dt * (sphi * tth);
      row[R::a(6, 11)] = dt * (cphi * tth);
      row[R::a(7, 6)] = dt * (-w);
      row[R::a(7, 10)] = dt * cphi;
      row[R::a(7, 11)] = dt * (-sphi);
      row[R::a(8, 6)] = dt * (v * sec);
      row[R::a(8, 7)] = dt * (w * sth * sec * sec);
      row[R::a(8, 10)] = dt * (sphi * sec);
      row[R::a(8, 11)] = dt * (cphi * sec);
      row[R::a(9, 10)] = dt * (c1 * r);
      row[R::a(9, 11)] = dt * (c1 * q);
      row[R::a(10, 9)] = dt * (c2 * r);
      row[R::a(10, 11)] = dt * (c2 * pp);
      row[R::a(11, 9)] = dt * (c3 * q);
      row[R::a(11, 10)] = dt * (c3 * pp);
      float dx[n], gx[n], gu[m], bg[m], bh[m];
#pragma unroll
      for (int i = 0; i < n; ++i) dx[i] = xs[i] - float(p.x_ref[i]);
      quad_grad<float, n>(p.Q, dx, p.diag_cost, gx);
      quad_grad<float, m>(p.R, us, p.diag_cost, gu);
      Mdl::barrier_d<float>(us, p, bg, bh);
#pragma unroll
      for (int i = 0; i < n; ++i) row[R::lx + i] = gx[i];
#pragma unroll
      for (int j = 0; j < m; ++j) row[R::lu + j] = gu[j] + bg[j];
      *reinterpret_cast<float4*>(bh_s[lane]) = make_float4(bh[0], bh[1], bh[2], bh[3]);
    }
    __syncwarp();
    float4* out4 = reinterpret_cast<float4*>(expansion + ((size_t)t * B + b0) * R::size);
    for (int rr = 0; rr < 32; ++rr) {
      if (!((live >> rr) & 1u)) continue;
      const float4* src = &stage4[rr * PITCH4];
      float4* dst = out4 + (size_t)rr * (R::size / 4);
      dst[lane] = src[lane];
      if (lane < 20) dst[32 + lane] = src[32 + lane];
      // second half: template + this record's barrier curvature on the l_uu diagonal
      float4 v1 = tmpl1;
      if (lane >= 4 && lane < 8) {             // float4 #(36 + a) holds l_uu[a][0..3]
        const int a = lane - 4;
        const float add = bh_s[rr][a];
        if (a == 0) v1.x += add; else if (a == 1) v1.y += add; else if (a == 2) v1.z += add; else v1.w += add;
      }
      dst[HALF / 4 + lane] = tmpl0;
      if (lane < 20) dst[HALF / 4 + 32 + lane] = v1;
    }
    __syncwarp();
  }
  if (t_hi == N) {   // terminal V_x, V_xx at x_N
    if (mine) {
      float xs[n];
#pragma unroll
      for (int i = 0; i < n; ++i) xs[i] = float(x[((size_t)N * n + i) * B + b]);
      emit_terminal<Mdl>(xs, p, StageSink{row});
    }
    __syncwarp();
    for (int rr = 0; rr < 32; ++rr) {
      if (!((live >> rr) & 1u)) continue;
      const float* src = reinterpret_cast<const float*>(&stage4[rr * PITCH4]);
      float* dst = terminal + (size_t)(b0 + rr) * R::term;
      for (int e = lane; e < R::term; e += 32) dst[e] = src[e];
    }
  }
}

}  // namespace qx

extern "C" int qx_linearize(const QxParams* p, const double* x, const double* u, float* expansion, float* terminal,
                            const int32_t* status, int B, int N, int t_start, void* stream) {
  QX_CHECK_ARG(p && x && u && expansion && terminal, "null pointer");
  QX_CHECK_ARG(B > 0 && N > 0 && t_start >= 0 && t_start < N, "need B > 0 and 0 <= t_start < N");
  QX_CHECK_ARG(p->integrator == QX_EULER || p->integrator == QX_RK4, "unknown integrator");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int T = N - t_start;
  switch (p->model_id) {
    case QX_MODEL_CARTPOLE: {
      dim3 grid(qx::ceil_div(B, 128), T);
      if (p->integrator == QX_EULER)
        qx::linearize_soa_kernel<qx::CartPole, QX_EULER><<<grid, 128, 0, st>>>(*p, x, u, expansion, terminal, status, B, N, t_start);
      else
        qx::linearize_soa_kernel<qx::CartPole, QX_RK4><<<grid, 128, 0, st>>>(*p, x, u, expansion, terminal, status, B, N, t_start);
      QX_CHECK_LAUNCH("linearize_soa_kernel");
      return QX_OK;
    }
    case QX_MODEL_QUADROTOR: {
      dim3 grid(qx::ceil_div(B, 32 * qx::kLinWarps), T);
      if (p->integrator == QX_EULER) {
        dim3 g2(qx::ceil_div(B, 32), qx::ceil_div(T, qx::kLinSteps));
        qx::linearize_quad_euler_kernel<<<g2, 32, 0, st>>>(*p, x, u, expansion, terminal, status, B, N, t_start);
      } else
        qx::linearize_rec_kernel<qx::Quadrotor, QX_RK4><<<grid, 32 * qx::kLinWarps, 0, st>>>(*p, x, u, expansion, terminal,
                                                                                           status, B, N, t_start);
      QX_CHECK_LAUNCH("linearize_rec_kernel");
      return QX_OK;
    }
  }
  qx::set_error("unknown model_id %d", p->model_id);
  return QX_EINVAL;
}
