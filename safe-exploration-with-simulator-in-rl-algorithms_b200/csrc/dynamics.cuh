// Per-environment swimmer dynamics as register-resident device functions (one thread = one env).
//
// gym variant  -- same equations as envs/gym_swimmer/swimmer/remy_swimmer_env.py:95-214, but NOT
// the reference's formulation.  The reference carries every joint acceleration/force as an affine
// row over the n+2 unknowns and solves a dense (n+2)x(n+2) system (O(n^3) flops, O(n^2) storage).
// Here the same Newton/Euler equations are reduced analytically:
//   * f_n = 0 sums all segment equations:  Gdd = sum_i Phi_i / (n m)        (Phi_i = friction force)
//   * eliminating thdd_i leaves a block-tridiagonal SPD system in the n-1 interior joint forces
//       Q_{j-1} f_{j-1} + P_j f_j + Q_j f_{j+1} = r_j ,   P_j = 2I + 3(N_{j-1}+N_j), Q_j = 3N_j - I,
//       N_i = n_i n_i^T,  r_j = Phi_j - Phi_{j-1} - w_{j-1} - w_j,
//       w_i = (6 tau_i / l) n_i - (l m thd_i^2 / 2) p_i,   tau_i = k thd_i l^3/12 + u_{i-1} - u_i
//     solved by an unpivoted 2x2-block LDL^T sweep (one reciprocal per joint),
//   * thdd_i = (6/(m l)) n_i.(f_i + f_{i+1}) + 12 tau_i/(m l^2).
// O(n) flops and O(n) registers, no pivoting, no local memory.  Agrees with the reference's
// pivoted dense solve to ~1e-15 relative (tests/test_parity_step.py).
//
// rlglue variant -- rlglue/environment/SwimmerEnvironment.cpp:139-271 including its literal
// column/weight quirks.  The (5n+2) system is reduced algebraically to the
// n+2 unknowns [Gdd_1, thdd_1..n] (the f_j and Gdd_i are affine in those), then solved with a
// partially pivoted LU.
#pragma once
#include <cuda_runtime.h>

namespace swm {

// Physical constants of one model, precomputed on the host (kernel argument -> constant bank).
struct Phys {
  double l, m, k, h, max_u, dirx, diry;
  double kl;         // k*l
  double inv_nm;     // 1/(n*m)
  double tau_c;      // k*l^3/12
  double six_over_l; // 6/l
  double half_lm;    // l*m/2
  double thdd_c;     // 6/(m*l)
  double inv_I;      // 12/(m*l^2)
  double inv_n;      // 1/n
  double half_l;     // l/2
  double I;          // m*l^2/12
};

// ~2-ulp reciprocal for well-scaled positive arguments (the 2x2 block determinants are >= 4):
// hardware seed + two Newton steps, no special-case path.
__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  return x;
}

// ---------------------------------------------------------------------------------------------
// gym variant.  s[i] = sin(th_i), c[i] = cos(th_i).  u has N-1 entries.
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void gym_accelerations(const Phys& P, const double (&s)[N],
                                                  const double (&c)[N], double gdx, double gdy,
                                                  const double (&thd)[N], const double* u,
                                                  double& gddx, double& gddy, double (&thdd)[N]) {
  // segment-centre velocities in the head frame, then shifted to the barycentric frame
  double gx[N], gy[N];
  {
    double ax = 0.0, ay = 0.0, mx = 0.0, my = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double a = P.l * thd[i];
      const double as = a * s[i], ac = a * c[i];
      gx[i] = fma(-0.5, as, ax);
      gy[i] = fma(0.5, ac, ay);
      ax -= as;
      ay += ac;
      mx += gx[i];
      my += gy[i];
    }
    const double sx = fma(-P.inv_n, mx, gdx), sy = fma(-P.inv_n, my, gdy);
#pragma unroll
    for (int i = 0; i < N; ++i) { gx[i] += sx; gy[i] += sy; }
  }
  // friction force Phi_i = F_i n_i,  F_i = -k l (Gdot_i . n_i),  n_i = (-s, c)
  double phx[N], phy[N], tau[N], wx[N], wy[N];
  double sumx = 0.0, sumy = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double F = -P.kl * fma(gy[i], c[i], -gx[i] * s[i]);
    phx[i] = -F * s[i];
    phy[i] = F * c[i];
    sumx += phx[i];
    sumy += phy[i];
    double t = P.tau_c * thd[i];
    if (i >= 1) t += u[i - 1];
    if (i <= N - 2) t -= u[i];
    tau[i] = t;
    const double al = P.six_over_l * t, be = P.half_lm * (thd[i] * thd[i]);
    wx[i] = -fma(al, s[i], be * c[i]);
    wy[i] = fma(al, c[i], -be * s[i]);
  }
  gddx = sumx * P.inv_nm;
  gddy = sumy * P.inv_nm;

  // block-tridiagonal SPD system for the interior joint forces f_1..f_{N-1}
  constexpr int J = N - 1;
  double fx[J + 2], fy[J + 2];  // f_0 .. f_N with f_0 = f_N = 0
  fx[0] = fy[0] = 0.0;
  fx[J + 1] = fy[J + 1] = 0.0;
  if (J >= 1) {
    double Xa[J], Xb[J], Xd[J];  // inverse of the pivot block [[a,b],[b,d]]
    double rx[J], ry[J];         // eliminated right-hand side
    double pa, pb, pd;           // current pivot block
#pragma unroll
    for (int j = 1; j <= J; ++j) {
      // segment j-1 and j meet at joint j
      const double ss0 = s[j - 1] * s[j - 1], sc0 = s[j - 1] * c[j - 1], cc0 = c[j - 1] * c[j - 1];
      const double ss1 = s[j] * s[j], sc1 = s[j] * c[j], cc1 = c[j] * c[j];
      pa = fma(3.0, ss0 + ss1, 2.0);
      pb = -3.0 * (sc0 + sc1);
      pd = fma(3.0, cc0 + cc1, 2.0);
      double r0 = (phx[j] - phx[j - 1]) - (wx[j - 1] + wx[j]);
      double r1 = (phy[j] - phy[j - 1]) - (wy[j - 1] + wy[j]);
      if (j >= 2) {
        // Q_{j-1} = 3 N_{j-1} - I couples f_{j-1} and f_j
        const double qa = fma(3.0, ss0, -1.0), qb = -3.0 * sc0, qd = fma(3.0, cc0, -1.0);
        // T = Q X   (X = inverse of previous pivot)
        const double t00 = fma(qa, Xa[j - 2], qb * Xb[j - 2]);
        const double t01 = fma(qa, Xb[j - 2], qb * Xd[j - 2]);
        const double t10 = fma(qb, Xa[j - 2], qd * Xb[j - 2]);
        const double t11 = fma(qb, Xb[j - 2], qd * Xd[j - 2]);
        pa -= fma(t00, qa, t01 * qb);
        pb -= fma(t00, qb, t01 * qd);
        pd -= fma(t10, qb, t11 * qd);
        r0 -= fma(t00, rx[j - 2], t01 * ry[j - 2]);
        r1 -= fma(t10, rx[j - 2], t11 * ry[j - 2]);
      }
      const double idet = fast_rcp(fma(pa, pd, -pb * pb));
      Xa[j - 1] = pd * idet;
      Xb[j - 1] = -pb * idet;
      Xd[j - 1] = pa * idet;
      rx[j - 1] = r0;
      ry[j - 1] = r1;
    }
#pragma unroll
    for (int j = J; j >= 1; --j) {
      double y0 = rx[j - 1], y1 = ry[j - 1];
      if (j < J) {
        const double qa = fma(3.0, s[j] * s[j], -1.0), qb = -3.0 * (s[j] * c[j]),
                     qd = fma(3.0, c[j] * c[j], -1.0);
        y0 -= fma(qa, fx[j + 1], qb * fy[j + 1]);
        y1 -= fma(qb, fx[j + 1], qd * fy[j + 1]);
      }
      fx[j] = fma(Xa[j - 1], y0, Xb[j - 1] * y1);
      fy[j] = fma(Xb[j - 1], y0, Xd[j - 1] * y1);
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double vx = fx[i] + fx[i + 1], vy = fy[i] + fy[i + 1];
    thdd[i] = fma(P.thdd_c, fma(c[i], vy, -s[i] * vx), tau[i] * P.inv_I);
  }
}

// ---------------------------------------------------------------------------------------------
// rlglue variant (reduced form of SwimmerEnvironment.cpp:139-271).
// Unknowns z = [g1x, g1y, thdd_0..thdd_{N-1}] with g1 = Gdd of the first segment; affine rows have
// N+3 slots (last = constant).
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void rlglue_accelerations(const Phys& P, const double (&s)[N],
                                                     const double (&c)[N], double gdx, double gdy,
                                                     const double (&thd)[N], const double* u,
                                                     double& gddx, double& gddy,
                                                     double (&thdd)[N]) {
  constexpr int M = N + 2, S = N + 3;
  // --- compute_friction (cpp:238-271); n_i = (-s, c) ---
  double g1x = gdx, g1y = gdy;
  {
    // G1_dot -= l/n * sum_{i=1..n} sum_{j<i} e_ij thd_j n_j, e = 0.5 for j==0 or j==i-1 else 1
    double accx = 0.0, accy = 0.0;
#pragma unroll
    for (int i = 1; i <= N; ++i) {
      double sx = 0.0, sy = 0.0;
#pragma unroll
      for (int j = 0; j < i; ++j) {
        const double e = (j == 0 || j == i - 1) ? 0.5 : 1.0;
        sx += e * thd[j] * (-s[j]);
        sy += e * thd[j] * c[j];
      }
      accx += sx;
      accy += sy;
    }
    g1x -= P.l * P.inv_n * accx;
    g1y -= P.l * P.inv_n * accy;
  }
  double Fx[N], Fy[N];
  {
    double cx = g1x, cy = g1y;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (i >= 1) {
        cx += P.half_l * (thd[i - 1] * (-s[i - 1]) + thd[i] * (-s[i]));
        cy += P.half_l * (thd[i - 1] * c[i - 1] + thd[i] * c[i]);
      }
      const double dot = fma(cy, c[i], -cx * s[i]);
      const double F = -P.kl * dot;
      Fx[i] = -F * s[i];
      Fy[i] = F * c[i];
    }
  }
  // --- affine forms: segment accelerations G_i (cpp:187-207 solved for Gdd_{i+1}) and joint
  //     forces f_j = f_{j-1} + m Gdd_j - F_j (cpp:172-181) ---
  double A[M][S];
#pragma unroll
  for (int r = 0; r < M; ++r)
#pragma unroll
    for (int q = 0; q < S; ++q) A[r][q] = 0.0;
  double Gx[S], Gy[S], fxa[S], fya[S], fxp[S], fyp[S];
#pragma unroll
  for (int q = 0; q < S; ++q) { Gx[q] = Gy[q] = fxa[q] = fya[q] = fxp[q] = fyp[q] = 0.0; }
  Gx[0] = 1.0;
  Gy[1] = 1.0;
  // walk segments; after processing segment i (0-based) fxa/fya hold f_{i+1}, fxp/fyp hold f_i.
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (i >= 1) {
      // Gdd_i = Gdd_{i-1} + (l/2)[(thdd n - thd^2 p)_{i-1} + (thdd n - thd^2 p)_i]
      Gx[2 + i - 1] += P.half_l * (-s[i - 1]);
      Gy[2 + i - 1] += P.half_l * c[i - 1];
      Gx[2 + i] += P.half_l * (-s[i]);
      Gy[2 + i] += P.half_l * c[i];
      Gx[S - 1] -= P.half_l * (thd[i - 1] * thd[i - 1] * c[i - 1] + thd[i] * thd[i] * c[i]);
      Gy[S - 1] -= P.half_l * (thd[i - 1] * thd[i - 1] * s[i - 1] + thd[i] * thd[i] * s[i]);
    }
#pragma unroll
    for (int q = 0; q < S; ++q) {
      fxp[q] = fxa[q];
      fyp[q] = fya[q];
      fxa[q] = fma(P.m, Gx[q], fxa[q]);
      fya[q] = fma(P.m, Gy[q], fya[q]);
    }
    fxa[S - 1] -= Fx[i];
    fya[S - 1] -= Fy[i];
    // torque row of segment i-1 (1-based i): I thdd - (l/2) n.(f_i + f_{i+1}) = B   (cpp:152-161,
    // literal column choice f_i, f_{i+1}); available once f_{i+1} is known.
    if (i >= 1) {
      const int sgm = i - 1;
#pragma unroll
      for (int q = 0; q < S; ++q) {
        const double vx = fxp[q] + fxa[q], vy = fyp[q] + fya[q];
        // for sgm == N-2 the second force is f_N, which the system pins to zero (cpp:183-184)
        const double ux = (sgm == N - 2) ? fxp[q] : vx, uy = (sgm == N - 2) ? fyp[q] : vy;
        A[2 + sgm][q] = -P.half_l * fma(c[sgm], uy, -s[sgm] * ux);
      }
    }
  }
  // rows 0,1: f_N = 0
#pragma unroll
  for (int q = 0; q < S; ++q) { A[0][q] = fxa[q]; A[1][q] = fya[q]; }
  // torque row of the last segment: forces f_N (=0) and "f_{N+1}" which aliases Gdd_1 (cpp:155,157)
  {
    const int sgm = N - 1;
    A[2 + sgm][0] = -P.half_l * (-s[sgm]);
    A[2 + sgm][1] = -P.half_l * c[sgm];
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    A[2 + i][2 + i] += P.I;
    double Bi = -P.tau_c * thd[i];  // M_friction (cpp:267)
    if (i >= 1) Bi += u[i - 1];
    if (i <= N - 2) Bi -= u[i];
    A[2 + i][S - 1] -= Bi;
  }
  // --- partially pivoted LU on [A | -const] ---
#pragma unroll
  for (int col = 0; col < M; ++col) {
    int best = col;
    double bv = fabs(A[col][col]);
#pragma unroll
    for (int r = col + 1; r < M; ++r) {
      const double v = fabs(A[r][col]);
      if (v > bv) { bv = v; best = r; }
    }
#pragma unroll
    for (int r = col + 1; r < M; ++r) {
      if (r == best) {
#pragma unroll
        for (int q = col; q < S; ++q) { const double t = A[col][q]; A[col][q] = A[r][q]; A[r][q] = t; }
      }
    }
    const double inv = 1.0 / A[col][col];
#pragma unroll
    for (int r = col + 1; r < M; ++r) {
      const double f = A[r][col] * inv;
#pragma unroll
      for (int q = col + 1; q < S; ++q) A[r][q] = fma(-f, A[col][q], A[r][q]);
    }
  }
  double z[M];
#pragma unroll
  for (int r = M - 1; r >= 0; --r) {
    double acc = -A[r][S - 1];
#pragma unroll
    for (int q = r + 1; q < M; ++q) acc = fma(-A[r][q], z[q], acc);
    z[r] = acc / A[r][r];
  }
#pragma unroll
  for (int i = 0; i < N; ++i) thdd[i] = z[2 + i];
  // Gdd = mean_i Gdd_i (cpp:222-225); Gdd_i = g1 + (l/2) sum_{q} w_iq v_q
  {
    double sx = 0.0, sy = 0.0, cx = z[0], cy = z[1];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (i >= 1) {
        const double vx0 = fma(z[2 + i - 1], -s[i - 1], -(thd[i - 1] * thd[i - 1]) * c[i - 1]);
        const double vy0 = fma(z[2 + i - 1], c[i - 1], -(thd[i - 1] * thd[i - 1]) * s[i - 1]);
        const double vx1 = fma(z[2 + i], -s[i], -(thd[i] * thd[i]) * c[i]);
        const double vy1 = fma(z[2 + i], c[i], -(thd[i] * thd[i]) * s[i]);
        cx += P.half_l * (vx0 + vx1);
        cy += P.half_l * (vy0 + vy1);
      }
      sx += cx;
      sy += cy;
    }
    gddx = sx * P.inv_n;
    gddy = sy * P.inv_n;
  }
}

// One integration step in registers.  VARIANT 0: explicit Euler (remy_swimmer_env.py:88-91),
// VARIANT 1: semi-implicit Euler (SwimmerEnvironment.cpp:228-236).  Returns the reward
// Gdot_new . direction (remy_swimmer_env.py:238-243 / cpp:273-277).
template <int N, int VARIANT>
__device__ __forceinline__ double swimmer_step(const Phys& P, double& gdx, double& gdy,
                                               double (&th)[N], double (&thd)[N],
                                               const double* u) {
  double s[N], c[N];
#pragma unroll
  for (int i = 0; i < N; ++i) sincos(th[i], &s[i], &c[i]);
  double gddx, gddy, thdd[N];
  if (VARIANT == 0) gym_accelerations<N>(P, s, c, gdx, gdy, thd, u, gddx, gddy, thdd);
  else rlglue_accelerations<N>(P, s, c, gdx, gdy, thd, u, gddx, gddy, thdd);
  gdx = fma(P.h, gddx, gdx);
  gdy = fma(P.h, gddy, gdy);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (VARIANT == 0) {
      th[i] = fma(P.h, thd[i], th[i]);
      thd[i] = fma(P.h, thdd[i], thd[i]);
    } else {
      thd[i] = fma(P.h, thdd[i], thd[i]);
      th[i] = fma(P.h, thd[i], th[i]);
    }
  }
  return fma(gdx, P.dirx, gdy * P.diry);
}

}  // namespace swm
