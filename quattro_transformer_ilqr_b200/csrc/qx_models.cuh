// Device models: cart-pole and quadrotor dynamics, costs and their exact derivatives.
//
// Reference formulas (paths relative to /root/reference):
//   cart-pole  dynamics  examples/cartpole/cartpole_dynamics.py:48-71, discretisation :94-105
//              cost      examples/cartpole/cartpole_mpc.py:255-256, :268-269
//   quadrotor  dynamics  examples/quadrotor/quadrotor_dynamics.py:64-164, discretisation :186-195
//              cost      examples/quadrotor/quadrotor_mpc.py:80, :86-93, :99-100
// The reference differentiates these by central finite differences
// (quattro_ilqr_tf/quattro_ilqr_tf.py:149-275); here the derivatives are closed-form.
//
// Everything is templated on the scalar type: rollouts and costs run in double (identical
// line-search decisions need cost differences far below fp32 resolution), the linearisation
// runs in float (its output feeds the fp32 Riccati recursion).
#pragma once
#include "qx_common.cuh"

namespace qx {

template <typename T> struct Mth;
template <> struct Mth<float> {
  static __device__ __forceinline__ void sincos(float a, float* s, float* c) { sincosf(a, s, c); }
  static __device__ __forceinline__ float exp(float a) { return expf(a); }
  static __device__ __forceinline__ float log1p(float a) { return log1pf(a); }
  static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
  static __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
};
template <> struct Mth<double> {
  static __device__ __forceinline__ void sincos(double a, double* s, double* c) { ::sincos(a, s, c); }
  static __device__ __forceinline__ double exp(double a) { return ::exp(a); }
  static __device__ __forceinline__ double log1p(double a) { return ::log1p(a); }
  static __device__ __forceinline__ double abs(double a) { return fabs(a); }
  static __device__ __forceinline__ double max(double a, double b) { return fmax(a, b); }
};

// v^T M v for the cost matrices (dense row-major n x n in QxParams, or its diagonal only).
template <typename T, int n>
__device__ __forceinline__ T quad_form(const double* __restrict__ M, const T* v, bool diag) {
  T acc = T(0);
  if (diag) {
#pragma unroll
    for (int i = 0; i < n; ++i) acc += T(M[i * n + i]) * v[i] * v[i];
  } else {
#pragma unroll
    for (int i = 0; i < n; ++i) {
      T r = T(0);
#pragma unroll
      for (int j = 0; j < n; ++j) r += T(M[i * n + j]) * v[j];
      acc += v[i] * r;
    }
  }
  return acc;
}

// g = (M + M^T) v
template <typename T, int n>
__device__ __forceinline__ void quad_grad(const double* __restrict__ M, const T* v, bool diag, T* g) {
  if (diag) {
#pragma unroll
    for (int i = 0; i < n; ++i) g[i] = T(2) * T(M[i * n + i]) * v[i];
  } else {
#pragma unroll
    for (int i = 0; i < n; ++i) {
      T r = T(0);
#pragma unroll
      for (int j = 0; j < n; ++j) r += T(M[i * n + j] + M[j * n + i]) * v[j];
      g[i] = r;
    }
  }
}

// ------------------------------------------------------------------------------------------
struct CartPole {
  static constexpr int n = 4, m = 1, id = QX_MODEL_CARTPOLE, layout = QX_LAYOUT_SOA;

  template <typename T>
  static __device__ __forceinline__ void fc(const T* x, const T* u, const QxParams& p, T* d) {
    const T M = T(p.phys[0]), mp = T(p.phys[1]), l = T(p.phys[2]), g = T(p.phys[3]);
    T s, c;
    Mth<T>::sincos(x[2], &s, &c);
    const T tm = M + mp;
    const T temp = (u[0] + mp * l * (x[3] * x[3]) * s) / tm;
    const T thdd = (-g * s + c * temp) / (l * (T(4.0 / 3.0) - (mp * c * c) / tm));
    d[0] = x[1];
    d[1] = temp - (mp * l * thdd * c) / tm;
    d[2] = x[3];
    d[3] = thdd;
  }

  // J = d fc / dx (row-major 4x4), G = d fc / du (4x1)
  template <typename T>
  static __device__ __forceinline__ void fc_jac(const T* x, const T* u, const QxParams& p, T* J, T* G) {
    const T M = T(p.phys[0]), mp = T(p.phys[1]), l = T(p.phys[2]), g = T(p.phys[3]);
    T s, c;
    Mth<T>::sincos(x[2], &s, &c);
    const T thd = x[3];
    const T tm = M + mp, ml = mp * l;
    const T temp = (u[0] + ml * thd * thd * s) / tm;
    const T t_th = ml * thd * thd * c / tm;
    const T t_thd = T(2) * ml * thd * s / tm;
    const T t_F = T(1) / tm;
    const T D = l * (T(4.0 / 3.0) - mp * c * c / tm);
    const T D_th = l * (T(2) * mp * c * s / tm);
    const T Nn = -g * s + c * temp;
    const T N_th = -g * c - s * temp + c * t_th;
    const T a = Nn / D;
    const T a_th = (N_th * D - Nn * D_th) / (D * D);
    const T a_thd = c * t_thd / D;
    const T a_F = c * t_F / D;
    const T k = ml / tm;
#pragma unroll
    for (int i = 0; i < 16; ++i) J[i] = T(0);
    J[0 * 4 + 1] = T(1);
    J[1 * 4 + 2] = t_th - k * (a_th * c - a * s);
    J[1 * 4 + 3] = t_thd - k * c * a_thd;
    J[2 * 4 + 3] = T(1);
    J[3 * 4 + 2] = a_th;
    J[3 * 4 + 3] = a_thd;
    G[0] = T(0);
    G[1] = t_F - k * c * a_F;
    G[2] = T(0);
    G[3] = a_F;
  }

  // control-dependent part of the running cost beyond u^T R u (none for the cart-pole)
  template <typename T>
  static __device__ __forceinline__ T barrier(const T*, const QxParams&) { return T(0); }
  template <typename T>
  static __device__ __forceinline__ void barrier_d(const T*, const QxParams&, T* g, T* h) {
    g[0] = T(0);
    h[0] = T(0);
  }
};

// ------------------------------------------------------------------------------------------
struct Quadrotor {
  static constexpr int n = 12, m = 4, id = QX_MODEL_QUADROTOR, layout = QX_LAYOUT_REC;

  template <typename T>
  static __device__ __forceinline__ void fc(const T* x, const T* u, const QxParams& p, T* d) {
    const T mass = T(p.phys[0]), Ix = T(p.phys[1]), Iy = T(p.phys[2]), Iz = T(p.phys[3]);
    const T arm = T(p.phys[4]), g = T(p.phys[5]), kyaw = T(p.phys[6]);
    T sphi, cphi, sth, cth, spsi, cpsi;
    Mth<T>::sincos(x[6], &sphi, &cphi);
    Mth<T>::sincos(x[7], &sth, &cth);
    Mth<T>::sincos(x[8], &spsi, &cpsi);
    const T pp = x[9], q = x[10], r = x[11];
    const T Tt = u[0] + u[1] + u[2] + u[3];
    const T Tm = Tt / mass;
    const T tth = sth / cth;
    d[0] = x[3];
    d[1] = x[4];
    d[2] = x[5];
    d[3] = Tm * (spsi * sphi + cpsi * sth * cphi);
    d[4] = Tm * (cpsi * sphi - spsi * sth * cphi);
    d[5] = -g + Tm * (cth * cphi);
    d[6] = pp + q * sphi * tth + r * cphi * tth;
    d[7] = q * cphi - r * sphi;
    d[8] = (q * sphi + r * cphi) / cth;
    const T tau_phi = arm * ((u[1] + u[2]) - (u[0] + u[3]));
    const T tau_th = arm * ((u[0] + u[1]) - (u[2] + u[3]));
    const T tau_psi = kyaw * (u[0] - u[1] + u[2] - u[3]);
    d[9] = ((Iy - Iz) / Ix) * (q * r) + tau_phi / Ix;
    d[10] = ((Iz - Ix) / Iy) * (pp * r) + tau_th / Iy;
    d[11] = ((Ix - Iy) / Iz) * (pp * q) + tau_psi / Iz;
  }

  template <typename T>
  static __device__ __forceinline__ void fc_jac(const T* x, const T* u, const QxParams& p, T* J, T* G) {
    const T mass = T(p.phys[0]), Ix = T(p.phys[1]), Iy = T(p.phys[2]), Iz = T(p.phys[3]);
    const T arm = T(p.phys[4]), kyaw = T(p.phys[6]);
    T sphi, cphi, sth, cth, spsi, cpsi;
    Mth<T>::sincos(x[6], &sphi, &cphi);
    Mth<T>::sincos(x[7], &sth, &cth);
    Mth<T>::sincos(x[8], &spsi, &cpsi);
    const T pp = x[9], q = x[10], r = x[11];
    const T Tt = u[0] + u[1] + u[2] + u[3];
    const T Tm = Tt / mass, im = T(1) / mass;
    const T sec = T(1) / cth, tth = sth * sec;
#pragma unroll
    for (int i = 0; i < 144; ++i) J[i] = T(0);
#pragma unroll
    for (int i = 0; i < 48; ++i) G[i] = T(0);
    J[0 * 12 + 3] = T(1);
    J[1 * 12 + 4] = T(1);
    J[2 * 12 + 5] = T(1);
    const T dirx = spsi * sphi + cpsi * sth * cphi;
    const T diry = cpsi * sphi - spsi * sth * cphi;
    const T dirz = cth * cphi;
    J[3 * 12 + 6] = Tm * (spsi * cphi - cpsi * sth * sphi);
    J[3 * 12 + 7] = Tm * (cpsi * cth * cphi);
    J[3 * 12 + 8] = Tm * diry;
    J[4 * 12 + 6] = Tm * (cpsi * cphi + spsi * sth * sphi);
    J[4 * 12 + 7] = Tm * (-spsi * cth * cphi);
    J[4 * 12 + 8] = -Tm * dirx;
    J[5 * 12 + 6] = -Tm * cth * sphi;
    J[5 * 12 + 7] = -Tm * sth * cphi;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      G[3 * 4 + j] = im * dirx;
      G[4 * 4 + j] = im * diry;
      G[5 * 4 + j] = im * dirz;
    }
    const T w = q * sphi + r * cphi;
    const T v = q * cphi - r * sphi;
    J[6 * 12 + 6] = v * tth;
    J[6 * 12 + 7] = w * sec * sec;
    J[6 * 12 + 9] = T(1);
    J[6 * 12 + 10] = sphi * tth;
    J[6 * 12 + 11] = cphi * tth;
    J[7 * 12 + 6] = -w;
    J[7 * 12 + 10] = cphi;
    J[7 * 12 + 11] = -sphi;
    J[8 * 12 + 6] = v * sec;
    J[8 * 12 + 7] = w * sth * sec * sec;
    J[8 * 12 + 10] = sphi * sec;
    J[8 * 12 + 11] = cphi * sec;
    const T c1 = (Iy - Iz) / Ix, c2 = (Iz - Ix) / Iy, c3 = (Ix - Iy) / Iz;
    J[9 * 12 + 10] = c1 * r;
    J[9 * 12 + 11] = c1 * q;
    J[10 * 12 + 9] = c2 * r;
    J[10 * 12 + 11] = c2 * pp;
    J[11 * 12 + 9] = c3 * q;
    J[11 * 12 + 10] = c3 * pp;
    const T aix = arm / Ix, aiy = arm / Iy, kiz = kyaw / Iz;
    G[9 * 4 + 0] = -aix; G[9 * 4 + 1] = aix;  G[9 * 4 + 2] = aix;   G[9 * 4 + 3] = -aix;
    G[10 * 4 + 0] = aiy; G[10 * 4 + 1] = aiy; G[10 * 4 + 2] = -aiy; G[10 * 4 + 3] = -aiy;
    G[11 * 4 + 0] = kiz; G[11 * 4 + 1] = -kiz; G[11 * 4 + 2] = kiz; G[11 * 4 + 3] = -kiz;
  }

  // softplus(z; beta) = log1p(exp(beta z)) / beta, evaluated without overflow
  template <typename T>
  static __device__ __forceinline__ T softplus(T z, T beta) {
    const T bz = beta * z;
    return (Mth<T>::max(bz, T(0)) + Mth<T>::log1p(Mth<T>::exp(-Mth<T>::abs(bz)))) / beta;
  }

  // alpha * sum_i softplus(-u_i)^2   (quadrotor_mpc.py:91-92)
  template <typename T>
  static __device__ __forceinline__ T barrier(const T* u, const QxParams& p) {
    const T beta = T(p.barrier_beta);
    T acc = T(0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const T s = softplus<T>(-u[i], beta);
      acc += s * s;
    }
    return T(p.barrier_alpha) * acc;
  }

  // first and second derivative of the barrier w.r.t. each u_i (it is separable)
  template <typename T>
  static __device__ __forceinline__ void barrier_d(const T* u, const QxParams& p, T* g, T* h) {
    const T beta = T(p.barrier_beta), al = T(p.barrier_alpha);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const T z = -u[i];
      const T s = softplus<T>(z, beta);
      const T sg = T(1) / (T(1) + Mth<T>::exp(-beta * z));   // softplus'
      const T s2 = beta * sg * (T(1) - sg);                   // softplus''
      g[i] = -T(2) * al * s * sg;
      h[i] = T(2) * al * (sg * sg + s * s2);
    }
  }
};

// ---- pieces shared by all kernels -------------------------------------------------------------
// running cost L(x,u) = dx^T Q dx + u^T R u + barrier(u)
template <class Mdl, typename T>
__device__ __forceinline__ T running_cost(const T* x, const T* u, const QxParams& p) {
  T dx[Mdl::n];
#pragma unroll
  for (int i = 0; i < Mdl::n; ++i) dx[i] = x[i] - T(p.x_ref[i]);
  T c = quad_form<T, Mdl::n>(p.Q, dx, p.diag_cost) + quad_form<T, Mdl::m>(p.R, u, p.diag_cost);
  c += Mdl::template barrier<T>(u, p);
  return c;
}

template <class Mdl, typename T>
__device__ __forceinline__ T final_cost(const T* x, const QxParams& p) {
  T dx[Mdl::n];
#pragma unroll
  for (int i = 0; i < Mdl::n; ++i) dx[i] = x[i] - T(p.x_ref[i]);
  return quad_form<T, Mdl::n>(p.Qf, dx, p.diag_cost);
}

// x_next = f(x, u): forward Euler or classical RK4 (cartpole_dynamics.py:94-105)
template <class Mdl, typename T>
__device__ __forceinline__ void discrete_step(const T* x, const T* u, const QxParams& p, T* xn) {
  constexpr int n = Mdl::n;
  const T dt = T(p.dt);
  T k1[n];
  Mdl::template fc<T>(x, u, p, k1);
  if (p.integrator == QX_EULER) {
#pragma unroll
    for (int i = 0; i < n; ++i) xn[i] = x[i] + dt * k1[i];
    return;
  }
  T k2[n], k3[n], k4[n], xs[n];
#pragma unroll
  for (int i = 0; i < n; ++i) xs[i] = x[i] + T(0.5) * dt * k1[i];
  Mdl::template fc<T>(xs, u, p, k2);
#pragma unroll
  for (int i = 0; i < n; ++i) xs[i] = x[i] + T(0.5) * dt * k2[i];
  Mdl::template fc<T>(xs, u, p, k3);
#pragma unroll
  for (int i = 0; i < n; ++i) xs[i] = x[i] + dt * k3[i];
  Mdl::template fc<T>(xs, u, p, k4);
#pragma unroll
  for (int i = 0; i < n; ++i) xn[i] = x[i] + (dt / T(6)) * (k1[i] + T(2) * k2[i] + T(2) * k3[i] + k4[i]);
}

}  // namespace qx
