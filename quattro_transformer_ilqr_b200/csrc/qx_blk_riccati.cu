// Backward Riccati pass of the quadrotor block path (Euler map, diagonal weights), ONE thread per trajectory.
//
//   riccati_quad_thr_kernel = backward_pass / backward_pass_segment (quattro_ilqr_tf.py:282-366)
//
// Same records, same gains and same mathematics as riccati_quad_blk_kernel (qx_blk.cu, four lanes per trajectory);
// what changes is who does the work.  With four lanes per trajectory every step pays for the exchange: neighbour
// shuffles for A^T V A, an all-gather of (Q_ux - reg K) and a transpose through shared memory for the
// symmetrisation - 1 420 issued instructions per warp-step for eight trajectories, of which ~600 are FMAs
// (profiles/r1_ncu_full_classic_summary.txt: issue 64 %, FMA pipe 42 %).  A thread that owns the whole 12 x 12
// problem needs no exchange and can use the symmetry of V: 78 numbers instead of 144, the upper triangle of
// A^T V A and of the value update only - about 1 800 FMAs per trajectory-step, issued for 32 trajectories at once.
//
//   x = [p | v | eta | omega],   A = | I dtI 0  0  |   B = | 0     |     (qx_blk.cu explains the records)
//                                    | 0  I  G  0  |       | d 1^T |
//                                    | 0  0  Hp E  |       | 0     |
//                                    | 0  0  0  Sp |       | T     |
// Per step, with W = V A restricted to the blocks that are needed:
//   W_a,eta = V_a,v G + V_a,eta Hp,   W_a,om = V_a,eta E + V_a,om Sp          a in {p, v, eta, om}
//   P = A^T W (upper triangle, ten 3x3 blocks), parked in shared memory as the blocks are finished
//   Q_ux = B^T W = 1 d^T W_v,: + T^T W_om,:          Q_uu = l_uu + B^T V B          Q_x, Q_u from V_x
//   K, k from the eliminated (Q_uu + reg I)           (gains :304-309; pivoted fallback if it is not positive)
//   V' = P + l_xx + (Q_ux - reg K)^T K  (upper triangle only: symmetric by construction, which is what the
//        reference's (V + V^T)/2 (:317) enforces),   V_x' = Q_x + (Q_ux - reg K)^T k
// (Q_ux - reg K)^T K is the reference's K^T Q_uu K + K^T Q_ux + Q_ux^T K (:316) after Q_uu K + Q_ux = -reg K.
//
// Registers hold V (78), V_x (12), the record (48) and the running blocks; P lives in shared memory ([78][32] floats,
// lane-contiguous: conflict free).  The record of the next step is copied global -> shared with cp.async
// (LDGSTS, 15 x 16 B per thread) while the current step is computed: no staging registers, no barrier - a thread
// only ever reads what it copied itself.
#include "qx_blk.cuh"

namespace qx {

namespace {

constexpr int kThrTerm = kBlkTerm;

__host__ __device__ constexpr int TRI(int i, int j) {
  return i <= j ? (i * 12 - (i * (i - 1)) / 2 + (j - i)) : (j * 12 - (j * (j - 1)) / 2 + (i - j));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float thr_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.f), r);       // one Newton step: <= 1 ulp
}

// pivoted fallback for an indefinite Q_uu (the reference's LU inverse still "works", SURVEY a-Q(4))
__device__ __noinline__ void thr_solve_pp(const float* Min, float* r) {
  float M[4][4];
  for (int a = 0; a < 4; ++a)
    for (int c = 0; c < 4; ++c) M[a][c] = Min[a * 4 + c];
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    for (int rr = c + 1; rr < 4; ++rr)
      if (fabsf(M[rr][c]) > fabsf(M[piv][c])) piv = rr;
    if (piv != c) {
      for (int jj = 0; jj < 4; ++jj) { const float tmp = M[c][jj]; M[c][jj] = M[piv][jj]; M[piv][jj] = tmp; }
      const float tmp = r[c]; r[c] = r[piv]; r[piv] = tmp;
    }
    const float inv = 1.0f / M[c][c];
    for (int rr = c + 1; rr < 4; ++rr) {
      const float f = M[rr][c] * inv;
      for (int jj = c + 1; jj < 4; ++jj) M[rr][jj] -= f * M[c][jj];
      r[rr] -= f * r[c];
    }
  }
  for (int c = 3; c >= 0; --c) {
    float acc = r[c];
    for (int jj = c + 1; jj < 4; ++jj) acc -= M[c][jj] * r[jj];
    r[c] = acc / M[c][c];
  }
}

// block (a, b) of the symmetric V (upper triangle storage) as a 3x3
__device__ __forceinline__ void vblk(const float* V, int a, int b, float X[3][3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) X[r][c] = V[TRI(3 * a + r, 3 * b + c)];
}

// column of B^T z for z restricted to the v block (through d) and the omega block (through T = sign * tr):
// q[a] = s + sum_r sg[r][a] e_r with sg = [[-1,1,1,-1],[1,1,-1,-1],[1,-1,1,-1]]
__device__ __forceinline__ void bt4(float s, float e0, float e1, float e2, float* q) {
  q[0] = (s - e0) + (e1 + e2);
  q[1] = (s + e0) + (e1 - e2);
  q[2] = (s + e0) - (e1 - e2);
  q[3] = (s - e0) - (e1 + e2);
}

struct ThrSmem {
  float4 rec[2][kBlkSlots][32];
  float P[78][32];
};

// The variable entries of the four record blocks (fixed ones - 0, 1, dt - are applied by the products below):
//   G = |g00 g01 g02|  Hp = |h00 h01 0|  E = |dt e01 e02|  Sp = |1   s01 s02|
//       |g10 g11 g12|       |h10 1   0|      |0  e11 e12|       |s10 1   s12|
//       |g20 g21 0  |       |h20 h21 1|      |0  e21 e22|       |s20 s21 1  |
struct Blocks {
  float g00, g01, g02, g10, g11, g12, g20, g21;
  float h00, h01, h10, h20, h21;
  float e01, e02, e11, e12, e21, e22;
  float s01, s02, s10, s12, s20, s21;
  float dt;
};
// The term order of every sum below is the one of a dense 3x3 product with the fixed entries written out
// (x * 0 dropped, x * 1 as an addition): same bits as the dense formulation, ~10 % fewer instructions.

// row vector times [G; Hp]:  o = x G + z Hp
__device__ __forceinline__ void row_GH(const Blocks& m, const float* x, const float* z, float* o) {
  float a = x[0] * m.g00;
  a = fmaf(x[1], m.g10, a); a = fmaf(x[2], m.g20, a);
  a = fmaf(z[0], m.h00, a); a = fmaf(z[1], m.h10, a); a = fmaf(z[2], m.h20, a);
  o[0] = a;
  a = x[0] * m.g01;
  a = fmaf(x[1], m.g11, a); a = fmaf(x[2], m.g21, a);
  a = fmaf(z[0], m.h01, a); a = a + z[1]; a = fmaf(z[2], m.h21, a);
  o[1] = a;
  a = x[0] * m.g02;
  a = fmaf(x[1], m.g12, a);
  o[2] = a + z[2];
}
// row vector times [E; Sp]:  o = z E + y Sp
__device__ __forceinline__ void row_ES(const Blocks& m, const float* z, const float* y, float* o) {
  float a = z[0] * m.dt;
  a = a + y[0]; a = fmaf(y[1], m.s10, a); a = fmaf(y[2], m.s20, a);
  o[0] = a;
  a = z[0] * m.e01;
  a = fmaf(z[1], m.e11, a); a = fmaf(z[2], m.e21, a);
  a = fmaf(y[0], m.s01, a); a = a + y[1]; a = fmaf(y[2], m.s21, a);
  o[1] = a;
  a = z[0] * m.e02;
  a = fmaf(z[1], m.e12, a); a = fmaf(z[2], m.e22, a);
  a = fmaf(y[0], m.s02, a); a = fmaf(y[1], m.s12, a);
  o[2] = a + y[2];
}
// [G; Hp]^T times a column:  o = G^T x + Hp^T z
__device__ __forceinline__ void col_GH(const Blocks& m, const float* x, const float* z, float* o) {
  float a = m.g00 * x[0];
  a = fmaf(m.g10, x[1], a); a = fmaf(m.g20, x[2], a);
  a = fmaf(m.h00, z[0], a); a = fmaf(m.h10, z[1], a); a = fmaf(m.h20, z[2], a);
  o[0] = a;
  a = m.g01 * x[0];
  a = fmaf(m.g11, x[1], a); a = fmaf(m.g21, x[2], a);
  a = fmaf(m.h01, z[0], a); a = a + z[1]; a = fmaf(m.h21, z[2], a);
  o[1] = a;
  a = m.g02 * x[0];
  a = fmaf(m.g12, x[1], a);
  o[2] = a + z[2];
}
// [E; Sp]^T times a column:  o = E^T z + Sp^T y
__device__ __forceinline__ void col_ES(const Blocks& m, const float* z, const float* y, float* o) {
  float a = m.dt * z[0];
  a = a + y[0]; a = fmaf(m.s10, y[1], a); a = fmaf(m.s20, y[2], a);
  o[0] = a;
  a = m.e01 * z[0];
  a = fmaf(m.e11, z[1], a); a = fmaf(m.e21, z[2], a);
  a = fmaf(m.s01, y[0], a); a = a + y[1]; a = fmaf(m.s21, y[2], a);
  o[1] = a;
  a = m.e02 * z[0];
  a = fmaf(m.e12, z[1], a); a = fmaf(m.e22, z[2], a);
  a = fmaf(m.s02, y[0], a); a = fmaf(m.s12, y[1], a);
  o[2] = a + y[2];
}
// W = X [G; ...] row by row
__device__ __forceinline__ void blk_GH(const Blocks& m, const float X[3][3], const float Z[3][3], float W[3][3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) row_GH(m, X[r], Z[r], W[r]);
}
__device__ __forceinline__ void blk_ES(const Blocks& m, const float Z[3][3], const float Y[3][3], float W[3][3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) row_ES(m, Z[r], Y[r], W[r]);
}

}  // namespace

__global__ void __launch_bounds__(32)
riccati_quad_thr_kernel(const __grid_constant__ BlkConst cst, const float* __restrict__ expansion,
                        const float* __restrict__ terminal, float* __restrict__ gains,
                        const int32_t* __restrict__ status, int B, int N, int t_start) {
  constexpr int G = 52;
  __shared__ ThrSmem sm;
  const int lane = threadIdx.x;
  const int b = blockIdx.x * 32 + lane;
  if (b >= B) return;
  if (status != nullptr && status[b] != 0) return;
  const int n_oct = ceil_div(B, 8);
  const int steps = N - t_start;
  const float4* ebase = reinterpret_cast<const float4*>(expansion) + (size_t)(b >> 3) * (8 * kBlkSlots) + (b & 7);
  const size_t stride_t = (size_t)n_oct * (8 * kBlkSlots);      // float4 per time step
  auto issue = [&](int k) {
    const float4* src = ebase + (size_t)(N - 1 - k) * stride_t;
#pragma unroll
    for (int q = 0; q < kBlkSlots; ++q) cp_async16(&sm.rec[k & 1][q][lane], src + q * 8);
    cp_async_commit();
  };
  issue(0);

  const float dt = cst.dt, reg = cst.reg;
  const float tr0 = cst.tr[0], tr1 = cst.tr[1], tr2 = cst.tr[2];

  // V_xx(N) = 2 Qf (diagonal), V_x(N) from linearise
  float V[78], vx[12];
#pragma unroll
  for (int i = 0; i < 12; ++i)
#pragma unroll
    for (int j = i; j < 12; ++j) V[TRI(i, j)] = (i == j) ? cst.lqf[i] : 0.f;
  {
    const float4* tm = reinterpret_cast<const float4*>(terminal + (size_t)b * kThrTerm);
    const float4 a0 = tm[0], a1 = tm[1], a2 = tm[2];
    vx[0] = a0.x; vx[1] = a0.y; vx[2] = a0.z; vx[3] = a0.w;
    vx[4] = a1.x; vx[5] = a1.y; vx[6] = a1.z; vx[7] = a1.w;
    vx[8] = a2.x; vx[9] = a2.y; vx[10] = a2.z; vx[11] = a2.w;
  }
  float4* g4 = reinterpret_cast<float4*>(gains) + (size_t)(N - 1) * (G / 4) * B + b;      // float4 [N][13][B]
  const size_t gstride = (size_t)(G / 4) * B;

  for (int k = 0; k < steps; ++k) {
    if (k + 1 < steps) {
      issue(k + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    Blocks m;
    float d[3], lu[4], luu[4], qx[12];
    {
      float4 v[kBlkSlots];
#pragma unroll
      for (int q = 0; q < kBlkSlots; ++q) v[q] = sm.rec[k & 1][q][lane];
      m.g00 = v[0].x; m.g01 = v[0].y; m.g02 = v[0].z; d[0] = v[0].w;
      m.g10 = v[1].x; m.g11 = v[1].y; m.g12 = v[1].z; d[1] = v[1].w;
      m.g20 = v[2].x; m.g21 = v[2].y; d[2] = v[2].z; m.h00 = v[2].w;
      m.h01 = v[3].x; m.h10 = v[3].y; m.h20 = v[3].z; m.h21 = v[3].w;
      m.e01 = v[4].x; m.e02 = v[4].y; m.e11 = v[4].z; m.e12 = v[4].w;
      m.e21 = v[5].x; m.e22 = v[5].y; m.s01 = v[5].z; m.s02 = v[5].w;
      m.s10 = v[6].x; m.s12 = v[6].y; m.s20 = v[6].z; m.s21 = v[6].w;
      m.dt = dt;
      lu[0] = v[7].x; lu[1] = v[7].y; lu[2] = v[7].z; lu[3] = v[7].w;
      luu[0] = v[8].x; luu[1] = v[8].y; luu[2] = v[8].z; luu[3] = v[8].w;
      qx[0] = v[9].x; qx[1] = v[9].y; qx[2] = v[9].z; qx[3] = v[9].w;
      qx[4] = v[10].x; qx[5] = v[10].y; qx[6] = v[10].z; qx[7] = v[10].w;
      qx[8] = v[11].x; qx[9] = v[11].y; qx[10] = v[11].z; qx[11] = v[11].w;
    }

    // ---- Q_x = l_x + A^T V_x,  Q_u = l_u + B^T V_x ------------------------------------------------
    float qu[4];
    {
      float ae[3], aw[3];
      col_GH(m, vx + 3, vx + 6, ae);
      col_ES(m, vx + 6, vx + 9, aw);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        qx[c] += vx[c];
        qx[3 + c] += fmaf(dt, vx[c], vx[3 + c]);
        qx[6 + c] += ae[c];
        qx[9 + c] += aw[c];
      }
      const float s = fmaf(d[2], vx[5], fmaf(d[1], vx[4], d[0] * vx[3]));
      float w[4];
      bt4(s, tr0 * vx[9], tr1 * vx[10], tr2 * vx[11], w);
#pragma unroll
      for (int a = 0; a < 4; ++a) qu[a] = lu[a] + w[a];
    }

    // ---- Q_uu = l_uu + B^T V B ---------------------------------------------------------------------
    float Quu[4][4];
    {
      float Vvv[3][3], Vvw[3][3], Vww[3][3];
      vblk(V, 1, 1, Vvv); vblk(V, 1, 3, Vvw); vblk(V, 3, 3, Vww);
      float a0 = 0.f, y[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float t = fmaf(Vvv[r][2], d[2], fmaf(Vvv[r][1], d[1], Vvv[r][0] * d[0]));
        a0 = fmaf(d[r], t, a0);
        y[r] = fmaf(Vvw[2][r], d[2], fmaf(Vvw[1][r], d[1], Vvw[0][r] * d[0]));      // (V_om,v d)_r
      }
      float g[4];
      bt4(a0, tr0 * y[0], tr1 * y[1], tr2 * y[2], g);       // alpha0 + g_a
      // T^T V_om,om T: z_rs = tr_r tr_s V_rs, signs sg[r][a] sg[s][c]
      const float tr[3] = {tr0, tr1, tr2};
      float z[3][3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) z[r][s2] = tr[r] * tr[s2] * Vww[r][s2];
      constexpr float SG[3][4] = {{-1.f, 1.f, 1.f, -1.f}, {1.f, 1.f, -1.f, -1.f}, {1.f, -1.f, 1.f, -1.f}};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = a; c < 4; ++c) {
          float acc = (g[a] - a0) + g[c];
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s2 = 0; s2 < 3; ++s2) acc += (SG[r][a] * SG[s2][c]) * z[r][s2];
          Quu[a][c] = acc;
          Quu[c][a] = acc;
        }
#pragma unroll
      for (int a = 0; a < 4; ++a) Quu[a][a] += luu[a];
    }

    // ---- P = A^T V A (upper triangle, to shared memory) and Q_ux = B^T V A ----------------------------
    float Qux[4][12];
    float* P = &sm.P[0][lane];
    auto putP = [&](int i, int j, float v) { P[TRI(i, j) * 32] = v; };
    {
      // p and v columns of Q_ux need no product with the record blocks
      float qp[3][4];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float s = fmaf(d[2], V[TRI(5, c)], fmaf(d[1], V[TRI(4, c)], d[0] * V[TRI(3, c)]));
        bt4(s, tr0 * V[TRI(9, c)], tr1 * V[TRI(10, c)], tr2 * V[TRI(11, c)], qp[c]);
        const float s2 = fmaf(d[2], V[TRI(5, 3 + c)], fmaf(d[1], V[TRI(4, 3 + c)], d[0] * V[TRI(3, 3 + c)]));
        float qv[4];
        bt4(s2, tr0 * V[TRI(9, 3 + c)], tr1 * V[TRI(10, 3 + c)], tr2 * V[TRI(11, 3 + c)], qv);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          Qux[a][c] = qp[c][a];
          Qux[a][3 + c] = fmaf(dt, qp[c][a], qv[a]);
        }
      }
      // row block p
      float Wph[3][3], Wpw[3][3];
      {
        float Vpv[3][3], Vpe[3][3], Vpw[3][3];
        vblk(V, 0, 1, Vpv); vblk(V, 0, 2, Vpe); vblk(V, 0, 3, Vpw);
        blk_GH(m, Vpv, Vpe, Wph);
        blk_ES(m, Vpe, Vpw, Wpw);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            if (c >= r) putP(r, c, V[TRI(r, c)]);
            putP(r, 3 + c, fmaf(dt, V[TRI(r, c)], Vpv[r][c]));
            putP(r, 6 + c, Wph[r][c]);
            putP(r, 9 + c, Wpw[r][c]);
          }
      }
      // row block v
      float Wvh[3][3], Wvw[3][3];
      {
        float Vvv[3][3], Vve[3][3], Vvw[3][3];
        vblk(V, 1, 1, Vvv); vblk(V, 1, 2, Vve); vblk(V, 1, 3, Vvw);
        blk_GH(m, Vvv, Vve, Wvh);
        blk_ES(m, Vve, Vvw, Wvw);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            if (c >= r)
              putP(3 + r, 3 + c,
                   fmaf(dt * dt, V[TRI(r, c)], fmaf(dt, V[TRI(r, 3 + c)] + V[TRI(c, 3 + r)], Vvv[r][c])));
            putP(3 + r, 6 + c, fmaf(dt, Wph[r][c], Wvh[r][c]));
            putP(3 + r, 9 + c, fmaf(dt, Wpw[r][c], Wvw[r][c]));
          }
      }
      // row block eta
      float Wew[3][3];
      {
        float Vev[3][3], Vee[3][3], Vew[3][3], Weh[3][3];
        vblk(V, 2, 1, Vev); vblk(V, 2, 2, Vee); vblk(V, 2, 3, Vew);
        blk_GH(m, Vev, Vee, Weh);
        blk_ES(m, Vee, Vew, Wew);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float wv[3] = {Wvh[0][c], Wvh[1][c], Wvh[2][c]}, we[3] = {Weh[0][c], Weh[1][c], Weh[2][c]};
          float o[3];
          col_GH(m, wv, we, o);                    // column c of G^T W_v,eta + Hp^T W_eta,eta
#pragma unroll
          for (int r = 0; r <= c; ++r) putP(6 + r, 6 + c, o[r]);
          const float xv[3] = {Wvw[0][c], Wvw[1][c], Wvw[2][c]}, xe[3] = {Wew[0][c], Wew[1][c], Wew[2][c]};
          col_GH(m, xv, xe, o);                    // column c of G^T W_v,om + Hp^T W_eta,om
#pragma unroll
          for (int r = 0; r < 3; ++r) putP(6 + r, 9 + c, o[r]);
        }
      }
      // row block omega
      {
        float Vwv[3][3], Vwe[3][3], Vww[3][3], Wwh[3][3], Www[3][3];
        vblk(V, 3, 1, Vwv); vblk(V, 3, 2, Vwe); vblk(V, 3, 3, Vww);
        blk_GH(m, Vwv, Vwe, Wwh);
        blk_ES(m, Vwe, Vww, Www);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float ze[3] = {Wew[0][c], Wew[1][c], Wew[2][c]}, yw[3] = {Www[0][c], Www[1][c], Www[2][c]};
          float o[3];
          col_ES(m, ze, yw, o);                    // column c of E^T W_eta,om + Sp^T W_om,om
#pragma unroll
          for (int r = 0; r <= c; ++r) putP(9 + r, 9 + c, o[r]);
        }
        // eta and omega columns of Q_ux = 1 d^T W_v,: + T^T W_om,:
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float q[4];
          const float sh = fmaf(d[2], Wvh[2][c], fmaf(d[1], Wvh[1][c], d[0] * Wvh[0][c]));
          bt4(sh, tr0 * Wwh[0][c], tr1 * Wwh[1][c], tr2 * Wwh[2][c], q);
#pragma unroll
          for (int a = 0; a < 4; ++a) Qux[a][6 + c] = q[a];
          const float sw = fmaf(d[2], Wvw[2][c], fmaf(d[1], Wvw[1][c], d[0] * Wvw[0][c]));
          bt4(sw, tr0 * Www[0][c], tr1 * Www[1][c], tr2 * Www[2][c], q);
#pragma unroll
          for (int a = 0; a < 4; ++a) Qux[a][9 + c] = q[a];
        }
      }
    }

    // ---- gains: eliminate (Q_uu + reg I) once, back-substitute the 12 columns of Q_ux and Q_u ----------
    float kk[4];
    {
      float M[4][4], inv[4], f[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) M[a][c] = Quu[a][c] + (a == c ? reg : 0.f);
      bool ok = true;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ok = ok && (M[c][c] > 0.f);
        inv[c] = thr_rcp(M[c][c]);
#pragma unroll
        for (int rr = c + 1; rr < 4; ++rr) {
          f[rr][c] = M[rr][c] * inv[c];
#pragma unroll
          for (int jj = c + 1; jj < 4; ++jj) M[rr][jj] = fmaf(-f[rr][c], M[c][jj], M[rr][jj]);
        }
      }
      auto solve = [&](float* r) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int rr = c + 1; rr < 4; ++rr) r[rr] = fmaf(-f[rr][c], r[c], r[rr]);
#pragma unroll
        for (int c = 3; c >= 0; --c) {
          float acc = r[c];
#pragma unroll
          for (int jj = c + 1; jj < 4; ++jj) acc = fmaf(-M[c][jj], r[jj], acc);
          r[c] = acc * inv[c];
        }
      };
      // K = -X: X overwrites nothing yet (Q_ux is needed below), so solve into a column at a time
      float Kx[4][12];
#pragma unroll
      for (int a = 0; a < 4; ++a) kk[a] = qu[a];
      if (ok) {
#pragma unroll
        for (int c = 0; c < 12; ++c) {
          float r[4] = {Qux[0][c], Qux[1][c], Qux[2][c], Qux[3][c]};
          solve(r);
#pragma unroll
          for (int a = 0; a < 4; ++a) Kx[a][c] = -r[a];
        }
        solve(kk);
      } else {
        float Mr[16];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) Mr[a * 4 + c] = Quu[a][c] + (a == c ? reg : 0.f);
#pragma unroll
        for (int c = 0; c < 12; ++c) {
          float r[4] = {Qux[0][c], Qux[1][c], Qux[2][c], Qux[3][c]};
          thr_solve_pp(Mr, r);
#pragma unroll
          for (int a = 0; a < 4; ++a) Kx[a][c] = -r[a];
        }
        thr_solve_pp(Mr, kk);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) kk[a] = -kk[a];
      // gain record [k | K row-major]
      {
        g4[0] = make_float4(kk[0], kk[1], kk[2], kk[3]);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int q = 0; q < 3; ++q)
            g4[(size_t)(1 + 3 * a + q) * B] = make_float4(Kx[a][4 * q], Kx[a][4 * q + 1], Kx[a][4 * q + 2], Kx[a][4 * q + 3]);
        g4 -= gstride;
      }
      // ---- V' = P + l_xx + (Q_ux - reg K)^T K,  V_x' = Q_x + (Q_ux - reg K)^T k ------------------------
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 12; ++c) Qux[a][c] = fmaf(-reg, Kx[a][c], Qux[a][c]);
#pragma unroll
      for (int i = 0; i < 12; ++i) {
#pragma unroll
        for (int j = i; j < 12; ++j) {
          float acc = P[TRI(i, j) * 32];
          if (i == j) acc += cst.lq[i];
#pragma unroll
          for (int a = 0; a < 4; ++a) acc = fmaf(Qux[a][i], Kx[a][j], acc);
          V[TRI(i, j)] = acc;
        }
        float acc = qx[i];
#pragma unroll
        for (int a = 0; a < 4; ++a) acc = fmaf(Qux[a][i], kk[a], acc);
        vx[i] = acc;
      }
    }
  }
}

int blk_launch_riccati(const BlkConst& c, const float* expansion, const float* terminal, float* gains,
                        const int32_t* status, int B, int N, int t_start, cudaStream_t st) {
  riccati_quad_thr_kernel<<<ceil_div(B, 32), 32, 0, st>>>(c, expansion, terminal, gains, status, B, N, t_start);
  QX_CHECK_LAUNCH("riccati_quad_thr_kernel");
  return QX_OK;
}

}  // namespace qx
