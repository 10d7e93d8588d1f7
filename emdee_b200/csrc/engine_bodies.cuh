// engine_bodies.cuh -- device-resident rigid-body integrator: one thread per body, body state as SoA over bodies.
// Part of the single translation unit engine.cu (included there, in order; not a standalone header).
//
// Reference loops replaced (paths relative to the reference tree):
//   k_body_frame          <- src/EmDeeData.f90:420-439 + src/ArBee.f90:97-130 (update_rigid_bodies, tBody_update: unwrapping,
//                            centre of mass, member offsets, inertia tensor, principal frame, quaternion, body coordinates)
//   k_body_boost          <- src/ArBee.f90:330-357 + src/EmDeeData.f90:864-922 (force_and_torque, boost, kinetic_energies)
//   k_body_move           <- src/EmDeeData.f90:823-860 + src/ArBee.f90:178-313 (move, rotate_no_squish, rotate_exact)
//   k_body_momenta        <- src/ArBee.f90:317-326 (particle_momenta, for EmDee_download "momenta")
//   k_body_take_momenta   <- src/EmDeeData.f90:157-189 (assign_momenta, for EmDee_upload "momenta")
//   k_body_set_omega      <- src/ArBee.f90:347-352 (assign_momenta from angular velocities, for EmDee_random_momenta)
//   k_shadow_*            <- src/EmDeeCode.f90:1150-1209 (pre_force / post_force bookkeeping of EmDee_verlet_step)
//
// Orientation algebra is written with quaternion products instead of the reference's 4x3 matrices:
//   B(q) v = q (x) (0,v)      C(q) v = (0,v) (x) q      Bt(q) p = vec(conj(q) (x) p)      Ct(q) B(q) v = Rot(q) v
#pragma once

namespace emdee {
namespace {

struct BodyView {
  int nb;                 // number of bodies
  const int* first;       // nb+1: members of body b are items first[b] .. first[b+1]-1
  const int* atom;        // item -> atom index (ascending inside a body)
  const double* mItem;    // item -> mass
  double* d;              // 3 per item: member position in the body (principal-axes) frame
  // component c of body b lives at [c*nb + b]
  double *mass, *MoI, *rcm, *pcm, *q, *pi, *omega, *Fb, *tau;
};

struct Rotor {
  double q[4], pi[4], w[3], I[3];
};

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
// q (x) (0, v)
__device__ __forceinline__ void quat_times_vec(const double q[4], const double v[3], double o[4]) {
  o[0] = -(q[1] * v[0] + q[2] * v[1] + q[3] * v[2]);
  o[1] = q[0] * v[0] + (q[2] * v[2] - q[3] * v[1]);
  o[2] = q[0] * v[1] + (q[3] * v[0] - q[1] * v[2]);
  o[3] = q[0] * v[2] + (q[1] * v[1] - q[2] * v[0]);
}
// (0, v) (x) q
__device__ __forceinline__ void vec_times_quat(const double v[3], const double q[4], double o[4]) {
  o[0] = -(q[1] * v[0] + q[2] * v[1] + q[3] * v[2]);
  o[1] = q[0] * v[0] + (v[1] * q[3] - v[2] * q[2]);
  o[2] = q[0] * v[1] + (v[2] * q[1] - v[0] * q[3]);
  o[3] = q[0] * v[2] + (v[0] * q[2] - v[1] * q[1]);
}
// vector part of conj(q) (x) p: the body-frame components conjugate to a quaternion momentum
__device__ __forceinline__ void conj_times_quat_vec(const double q[4], const double p[4], double o[3]) {
  o[0] = q[0] * p[1] - p[0] * q[1] - (q[2] * p[3] - q[3] * p[2]);
  o[1] = q[0] * p[2] - p[0] * q[2] - (q[3] * p[1] - q[1] * p[3]);
  o[2] = q[0] * p[3] - p[0] * q[3] - (q[1] * p[2] - q[2] * p[1]);
}
// vector part of p (x) conj(q)
__device__ __forceinline__ void quat_times_conj_vec(const double p[4], const double q[4], double o[3]) {
  o[0] = q[0] * p[1] - p[0] * q[1] - (p[2] * q[3] - p[3] * q[2]);
  o[1] = q[0] * p[2] - p[0] * q[2] - (p[3] * q[1] - p[1] * q[3]);
  o[2] = q[0] * p[3] - p[0] * q[3] - (p[1] * q[2] - p[2] * q[1]);
}
// body -> space map of q, row-major 3x3 (scaled by |q|^2 like the product Ct(q) B(q) it stands for)
__device__ __forceinline__ void rotation_of(const double q[4], double A[9]) {
  const double a = q[0], b = q[1], c = q[2], d = q[3];
  const double aa = a * a, bb = b * b, cc = c * c, dd = d * d;
  A[0] = aa + bb - cc - dd; A[1] = 2.0 * (b * c - a * d); A[2] = 2.0 * (b * d + a * c);
  A[3] = 2.0 * (b * c + a * d); A[4] = aa - bb + cc - dd; A[5] = 2.0 * (c * d - a * b);
  A[6] = 2.0 * (b * d - a * c); A[7] = 2.0 * (c * d + a * b); A[8] = aa - bb - cc + dd;
}
__device__ __forceinline__ void apply3(const double A[9], const double v[3], double o[3]) {
  o[0] = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
  o[1] = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
  o[2] = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
}
__device__ __forceinline__ double sgn1(double x) { return copysign(1.0, x); }

// omega = (1/2) I^-1 Bt(q) pi
__device__ __forceinline__ void rotor_omega_from_pi(Rotor& r) {
  double t[3];
  conj_times_quat_vec(r.q, r.pi, t);
#pragma unroll
  for (int x = 0; x < 3; ++x) r.w[x] = 0.5 * (1.0 / r.I[x]) * t[x];
}
// pi = B(q) (2 I omega)
__device__ __forceinline__ void rotor_pi_from_omega(Rotor& r) {
  const double v[3] = {2.0 * r.I[0] * r.w[0], 2.0 * r.I[1] * r.w[1], 2.0 * r.I[2] * r.w[2]};
  quat_times_vec(r.q, v, r.pi);
}

// ---- Carlson symmetric elliptic integrals by duplication (reference src/math.f90:505-638: same truncation order and
// tolerance, so both sides converge to the same double) ----------------------------------------------------------
__device__ double carlson_rf(double x, double y, double z) {
  double mu, dx, dy, dz;
  for (;;) {
    mu = (x + y + z) / 3.0;
    dx = 2.0 - (mu + x) / mu;
    dy = 2.0 - (mu + y) / mu;
    dz = 2.0 - (mu + z) / mu;
    if (fmax(fmax(fabs(dx), fabs(dy)), fabs(dz)) < 1e-3) break;
    const double sx = sqrt(x), sy = sqrt(y), sz = sqrt(z);
    const double lam = sx * (sy + sz) + sy * sz;
    x = 0.25 * (x + lam);
    y = 0.25 * (y + lam);
    z = 0.25 * (z + lam);
  }
  const double e2 = dx * dy - dz * dz, e3 = dx * dy * dz;
  return (1.0 + ((1.0 / 24.0) * e2 - 0.1 - (3.0 / 44.0) * e3) * e2 + (1.0 / 14.0) * e3) / sqrt(mu);
}
__device__ double carlson_rc(double x, double y) {
  double mu, s;
  for (;;) {
    mu = (x + y + y) / 3.0;
    s = (y + mu) / mu - 2.0;
    if (fabs(s) < 1e-3) break;
    const double lam = 2.0 * sqrt(x) * sqrt(y) + y;
    x = 0.25 * (x + lam);
    y = 0.25 * (y + lam);
  }
  return (1.0 + s * s * (0.3 + s * ((1.0 / 7.0) + s * (0.375 + s * (9.0 / 22.0))))) / sqrt(mu);
}
__device__ double carlson_rj(double x, double y, double z, double p) {
  double mu, dx, dy, dz, dp, sigma = 0.0, pw = 1.0;
  for (;;) {
    mu = 0.2 * (x + y + z + p + p);
    dx = (mu - x) / mu;
    dy = (mu - y) / mu;
    dz = (mu - z) / mu;
    dp = (mu - p) / mu;
    if (fmax(fmax(fabs(dx), fabs(dy)), fmax(fabs(dz), fabs(dp))) < 1e-3) break;
    const double sx = sqrt(x), sy = sqrt(y), sz = sqrt(z);
    const double lam = sx * (sy + sz) + sy * sz;
    double al = p * (sx + sy + sz) + sx * sy * sz;
    al *= al;
    const double be = p * (p + lam) * (p + lam);
    sigma += pw * carlson_rc(al, be);
    pw *= 0.25;
    x = 0.25 * (x + lam);
    y = 0.25 * (y + lam);
    z = 0.25 * (z + lam);
    p = 0.25 * (p + lam);
  }
  const double c1 = 3.0 / 14.0, c2 = 1.0 / 3.0, c3 = 3.0 / 22.0, c4 = 3.0 / 26.0;
  const double ea = dx * (dy + dz) + dy * dz, eb = dx * dy * dz, ec = dp * dp;
  const double e2 = ea - 3.0 * ec, e3 = eb + 2.0 * dp * (ea - ec);
  const double s1 = 1.0 + e2 * (-c1 + 0.75 * c3 * e2 - 1.5 * c4 * e3);
  const double s2 = eb * (0.5 * c2 + dp * (-c3 - c3 + dp * c4));
  const double s3 = dp * ea * (c2 - dp * c3) - c2 * dp * ec;
  return 3.0 * sigma + pw * (s1 + s2 + s3) / (mu * sqrt(mu));
}

// Jacobi sn, cn, dn (reference src/math.f90:440-501): arithmetic-geometric-mean ladder up, Landen transformation down
__device__ void jacobi_sncndn(double u, double m, double& sn, double& cn, double& dn) {
  const double eps = 2.220446049250313e-16;
  if (fabs(m) < 2.0 * eps) {
    sn = sin(u); cn = cos(u); dn = 1.0;
    return;
  }
  if (fabs(m - 1.0) < 2.0 * eps) {
    sn = tanh(u);
    cn = dn = 1.0 / cosh(u);
    return;
  }
  double am[16], gm[16];
  int n = 0;
  am[0] = 1.0;
  gm[0] = sqrt(1.0 - m);
  while (fabs(am[n] - gm[n]) > 4.0 * eps * fabs(am[n] + gm[n]) && n < 14) {
    am[n + 1] = 0.5 * (am[n] + gm[n]);
    gm[n + 1] = sqrt(am[n] * gm[n]);
    ++n;
  }
  double s, c;
  sincos(u * am[n], &s, &c);
  const bool tangent = fabs(s) < fabs(c);
  double cc = am[n] * (tangent ? s / c : c / s), dd = 1.0;
  while (n > 0) {
    const double r = cc * cc / am[n];
    cc = dd * cc;
    --n;
    dd = (r + gm[n]) / (r + am[n]);
  }
  if (tangent) {
    dn = sqrt(1.0 - m) / dd;
    cn = dn * sgn1(c) / hypot(1.0, cc);
    sn = cn * cc / sqrt(1.0 - m);
  } else {
    dn = dd;
    sn = sgn1(s) / hypot(1.0, cc);
    cn = cc * sn;
  }
}

__device__ __forceinline__ int nearest_step(double x) {   // reference src/math.f90:428-436 (staircase)
  return x > 0.0 ? (int)ceil(x - 0.5) : (int)floor(x + 0.5);
}

// rotation about principal axis k (0-based) by the angle its own momentum dictates (Miller et al. 2002)
__device__ __forceinline__ void rotor_axis(Rotor& r, int k, double dt) {
  double e[3] = {0.0, 0.0, 0.0};
  e[k] = 1.0;
  double Pq[4], Pp[4];
  quat_times_vec(r.q, e, Pq);
  quat_times_vec(r.pi, e, Pp);
  const double ang = dt * (r.pi[0] * Pq[0] + r.pi[1] * Pq[1] + r.pi[2] * Pq[2] + r.pi[3] * Pq[3]) / (4.0 * r.I[k]);
  double s, c;
  sincos(ang, &s, &c);
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    r.q[x] = c * r.q[x] + s * Pq[x];
    r.pi[x] = c * r.pi[x] + s * Pp[x];
  }
}

// reference src/ArBee.f90:178-192; omega is NOT refreshed (the reference leaves it stale until the next boost)
__device__ void rotor_no_squish(Rotor& r, double delta_t, int n) {
  const double dt = delta_t / n, h = 0.5 * dt;
  for (int i = 0; i < n; ++i) {
    rotor_axis(r, 2, h);
    rotor_axis(r, 1, h);
    rotor_axis(r, 0, dt);
    rotor_axis(r, 1, h);
    rotor_axis(r, 2, h);
  }
}

// Torque-free motion of an asymmetric top in closed form (reference src/ArBee.f90:221-313). Returns false on the
// reference's early exit (angular momentum along the first axis: plain uniaxial rotation, member offsets NOT refreshed).
__device__ bool rotor_exact(Rotor& r, double dt) {
  const double eps = 2.220446049250313e-16, PI = 3.14159265358979324;
  const double I1 = r.I[0], I2 = r.I[1], I3 = r.I[2];
  const double w0[3] = {r.w[0], r.w[1], r.w[2]};
  double Lb[3] = {I1 * w0[0], I2 * w0[1], I3 * w0[2]};
  double L2 = Lb[1] * Lb[1] + Lb[2] * Lb[2];
  if (L2 < eps) {
    rotor_axis(r, 0, dt);
    return false;
  }
  L2 = Lb[0] * Lb[0] + L2;
  const double L = sqrt(L2);
  const double twoK = Lb[0] * w0[0] + Lb[1] * w0[1] + Lb[2] * w0[2];
  const double r1 = L2 - twoK * I3, r3 = twoK * I1 - L2;
  const double l1 = (1.0 / (I2 * (I2 - I3))) * r1, l3 = (1.0 / (I2 * (I1 - I2))) * r3;
  const double lmin = fmin(l1, l3);
  double a1 = sgn1(w0[0]) * sqrt((1.0 / (I1 * (I1 - I3))) * r1), a2 = sqrt(lmin),
         a3 = sgn1(w0[2]) * sqrt((1.0 / (I3 * (I1 - I3))) * r3);
  const double m = lmin / fmax(l1, l3);
  const double K = carlson_rf(0.0, 1.0 - m, 1.0), inv2K = 0.5 / K;
  double s0 = w0[1] / a2, c0, u0;
  int i0;
  if (fabs(s0) < 1.0) {
    c0 = (l1 < l3) ? w0[0] / a1 : w0[2] / a3;
    u0 = s0 * carlson_rf(1.0 - s0 * s0, 1.0 - m * s0 * s0, 1.0);
    i0 = nearest_step(u0 * inv2K);
  } else {
    a2 = fabs(w0[1]);
    s0 = sgn1(s0);
    c0 = 0.0;
    u0 = copysign(K, s0);
    i0 = 0;
  }
  const double wp = ((I3 - I1) / I2) * a1 * a3 / a2;
  const double u = wp * dt + u0;
  const int jump = nearest_step(u * inv2K) - i0;
  double sn, cn, dn;
  jacobi_sncndn(u, m, sn, cn, dn);
  const double alpha = I1 * a1 / L;
  double eta = alpha * alpha;
  eta = eta / (1.0 - eta);
  double dF;
  if (l1 < l3) {
    r.w[0] = a1 * cn; r.w[1] = a2 * sn; r.w[2] = a3 * dn;
    const double C = sqrt(m + eta), d0 = w0[2] / a3;
    auto theta = [&](double x) { const double x2 = x * x; return -(1.0 / 3.0) * eta * x * x2 * carlson_rj(1.0 - x2, 1.0 - m * x2, 1.0, 1.0 + eta * x2); };
    dF = u - u0 + sgn1(cn) * theta(sn) - sgn1(c0) * theta(s0) + (alpha / C) * (atan(C * sn / dn) - atan(C * s0 / d0));
    if (jump != 0) dF = dF + jump * 2.0 * theta(1.0);
  } else {
    r.w[0] = a1 * dn; r.w[1] = a2 * sn; r.w[2] = a3 * cn;
    const double ke = m * eta, C = sqrt(1.0 + ke);
    auto theta = [&](double x) { const double x2 = x * x; return -(1.0 / 3.0) * ke * x * x2 * carlson_rj(1.0 - x2, 1.0 - m * x2, 1.0, 1.0 + ke * x2); };
    dF = u - u0 + sgn1(cn) * theta(sn) - sgn1(c0) * theta(s0) + (alpha / C) * (atan(C * sn / cn) - atan(C * s0 / c0));
    if (jump != 0) dF = dF + jump * (2.0 * theta(1.0) + (alpha / C) * PI);
  }
  dF = (eta + 1.0) * dF;
  const double ang = (L2 * (u - u0) + r3 * dF) / (2.0 * L * I1 * wp);
  const double z0[4] = {Lb[2], Lb[1], L - Lb[0], 0.0};
  Lb[0] = I1 * r.w[0]; Lb[1] = I2 * r.w[1]; Lb[2] = I3 * r.w[2];
  double sa, ca;
  sincos(ang, &sa, &ca);
  const double z[4] = {Lb[2] * ca - Lb[1] * sa, Lb[1] * ca + Lb[2] * sa, (L - Lb[0]) * ca, (L - Lb[0]) * sa};
  // q <- normalize( z (z0.q) + C(z) Ct(z0) q )
  double t3[3], t4[4], nq[4];
  quat_times_conj_vec(r.q, z0, t3);     // Ct(z0) q = vec(q (x) conj(z0))
  vec_times_quat(t3, z, t4);            // C(z) t3 = (0,t3) (x) z
  const double z0q = z0[0] * r.q[0] + z0[1] * r.q[1] + z0[2] * r.q[2] + z0[3] * r.q[3];
  double nrm = 0.0;
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    nq[x] = z[x] * z0q + t4[x];
    nrm += nq[x] * nq[x];
  }
  nrm = sqrt(nrm);
#pragma unroll
  for (int x = 0; x < 4; ++x) r.q[x] = nq[x] / nrm;
  const double twoL[3] = {2.0 * Lb[0], 2.0 * Lb[1], 2.0 * Lb[2]};
  quat_times_vec(r.q, twoL, r.pi);
  return true;
}

// ---- 3x3 symmetric eigen-decomposition (reference src/math.f90:244-424, after J. Kopp 2006) --------------------------
// a = upper triangle {a00,a01,a02,a11,a12,a22}; eigenvalues sorted by decreasing magnitude; v[k] = k-th eigenvector.
__device__ void principal_axes(const double s[6], double w[3], double v[3][3]) {
  const double eps = 2.220446049250313e-16;
  double a00 = s[0], a11 = s[3];
  const double a01 = s[1], a02 = s[2], a12 = s[4], a22 = s[5];
  const double de = a01 * a12, dd = a01 * a01, ee = a12 * a12, ff = a02 * a02;
  const double tr = a00 + a11 + a22;
  const double c1 = (a00 * a11 + a00 * a22 + a11 * a22) - (dd + ee + ff);
  const double c0 = 27.0 * (a22 * dd + a00 * ee + a11 * ff - a00 * a11 * a22 - 2.0 * a02 * de);
  double p = tr * tr - 3.0 * c1;
  const double rr = tr * (p - 1.5 * c1) - 0.5 * c0;
  const double sp = sqrt(fabs(p));
  const double ang = (1.0 / 3.0) * atan2(sqrt(fabs(6.75 * c1 * c1 * (p - c1) + c0 * (rr + 0.25 * c0))), rr);
  double sa, ca;
  sincos(ang, &sa, &ca);
  const double c = sp * ca, sq = (1.0 / sqrt(3.0)) * sp * sa;
  p = (1.0 / 3.0) * (tr - c);
  double w1 = p + c, w2 = p - sq, w3 = p + sq, t;
  if (fabs(w1) < fabs(w3)) { t = w1; w1 = w3; w3 = t; }
  if (fabs(w1) < fabs(w2)) { t = w1; w1 = w2; w2 = t; }
  if (fabs(w2) < fabs(w3)) { t = w2; w2 = w3; w3 = t; }
  w[0] = w1; w[1] = w2; w[2] = w3;
  const double tiny8 = 8.0 * eps * fabs(w1), thresh = tiny8 * tiny8;
  const double n1 = a01 * a01 + a02 * a02, n2 = a01 * a01 + a12 * a12;
  const double q0 = a01 * a12 - a02 * a11, q1 = a02 * a01 - a12 * a00, q2 = a01 * a01;

  // eigenvector of eigenvalue `lam` as the cross product of the first two columns of (A - lam), with the
  // degenerate-column fall-backs; b00, b11 are the shifted diagonal entries
  auto column_cross = [&](double lam, double b00, double b11, double o[3]) {
    o[0] = q0 + a02 * lam;
    o[1] = q1 + a12 * lam;
    o[2] = b00 * b11 - q2;
    const double norm = o[0] * o[0] + o[1] * o[1] + o[2] * o[2];
    const double m1 = n1 + b00 * b00, m2 = n2 + b11 * b11;
    if (m1 <= thresh) {
      o[0] = 1.0; o[1] = 0.0; o[2] = 0.0;
    } else if (m2 <= thresh) {
      o[0] = 0.0; o[1] = 1.0; o[2] = 0.0;
    } else if (norm < (64.0 * eps) * (64.0 * eps) * (m1 * m2)) {
      double big = fabs(a01), f = -b00 / a01;
      if (fabs(b11) > big) { big = fabs(b11); f = -a01 / b11; }
      if (fabs(a12) > big) f = -a02 / a12;
      const double nn = 1.0 / sqrt(1.0 + f * f);
      o[0] = nn; o[1] = f * nn; o[2] = 0.0;
    } else {
      const double sc = sqrt(1.0 / norm);
      o[0] *= sc; o[1] *= sc; o[2] *= sc;
    }
  };
  column_cross(w1, a00 - w1, a11 - w1, v[0]);
  const double gap = w1 - w2;
  if (fabs(gap) > tiny8) {
    column_cross(w2, (a00 - w1) + gap, (a11 - w1) + gap, v[1]);
  } else {
    // (near-)degenerate pair: v1 = v0 x column_i(A - w2) for the first usable column
    const double full[3][3] = {{((a00 - w1) + w1), a01, a02}, {a01, ((a11 - w1) + w1), a12}, {a02, a12, a22}};
    bool ok = false;
    for (int i = 0; i < 3 && !ok; ++i) {
      double col[3] = {full[0][i], full[1][i], full[2][i]};
      col[i] -= w2;
      // the reference shifts the diagonal cumulatively (a(i,i) is not restored between trials); a later column
      // only sees its OWN diagonal entry, so trial i is unaffected by the earlier shifts
      const double cn = col[0] * col[0] + col[1] * col[1] + col[2] * col[2];
      ok = cn > thresh;
      if (ok) {
        cross3(v[0], col, v[1]);
        const double norm = v[1][0] * v[1][0] + v[1][1] * v[1][1] + v[1][2] * v[1][2];
        ok = norm > (256.0 * eps) * (256.0 * eps) * cn;
        if (ok) {
          const double sc = sqrt(1.0 / norm);
          v[1][0] *= sc; v[1][1] *= sc; v[1][2] *= sc;
        }
      }
    }
    if (!ok) {
      int i = 0;
      while (v[0][i] == 0.0) ++i;
      const int j = (i + 1) % 3;
      const double nn = 1.0 / sqrt(v[0][i] * v[0][i] + v[0][j] * v[0][j]);
      v[1][i] = v[0][j] * nn;
      v[1][j] = -v[0][i] * nn;
      v[1][(i + 2) % 3] = 0.0;
    }
  }
  cross3(v[0], v[1], v[2]);
}

// rows of A are the principal axes (space -> body); Shepperd's method (reference src/math.f90:189-218)
__device__ void quaternion_of_axes(const double A[3][3], double q[4]) {
  const double t[4] = {1.0 + A[0][0] + A[1][1] + A[2][2], 1.0 + A[0][0] - A[1][1] - A[2][2],
                       1.0 - A[0][0] + A[1][1] - A[2][2], 1.0 - A[0][0] - A[1][1] + A[2][2]};
  int k = 0;
  for (int i = 1; i < 4; ++i)
    if (t[i] > t[k]) k = i;
  const double x = A[1][2] - A[2][1], y = A[2][0] - A[0][2], z = A[0][1] - A[1][0];
  const double xy = A[0][1] + A[1][0], xz = A[0][2] + A[2][0], yz = A[1][2] + A[2][1];
  double v[4];
  if (k == 0) { v[0] = t[0]; v[1] = x; v[2] = y; v[3] = z; }
  else if (k == 1) { v[0] = x; v[1] = t[1]; v[2] = xy; v[3] = xz; }
  else if (k == 2) { v[0] = y; v[1] = xy; v[2] = t[2]; v[3] = yz; }
  else { v[0] = z; v[1] = xz; v[2] = yz; v[3] = t[3]; }
  const double f = 0.5 * sqrt(1.0 / t[k]);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = v[i] * f;
}

// ---- per-body load / store --------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_rotor(const BodyView& v, int b, Rotor& r) {
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    r.q[x] = v.q[(size_t)x * v.nb + b];
    r.pi[x] = v.pi[(size_t)x * v.nb + b];
  }
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    r.w[x] = v.omega[(size_t)x * v.nb + b];
    r.I[x] = v.MoI[(size_t)x * v.nb + b];
  }
}
__device__ __forceinline__ void store_rotor(const BodyView& v, int b, const Rotor& r) {
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    v.q[(size_t)x * v.nb + b] = r.q[x];
    v.pi[(size_t)x * v.nb + b] = r.pi[x];
  }
#pragma unroll
  for (int x = 0; x < 3; ++x) v.omega[(size_t)x * v.nb + b] = r.w[x];
}

// block-wide sum of W values per thread, then the grid-wide last-block fold of engine_common.cuh
template <int W>
__device__ __forceinline__ void reduce_and_finish(const double (&mine)[W], double* __restrict__ partial,
                                                  unsigned int* __restrict__ ticket, double* __restrict__ out) {
  __shared__ double red[TPB / 32][W];
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int x = 0; x < W; ++x) {
    double s = mine[x];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) red[threadIdx.x >> 5][x] = s;
  }
  __syncthreads();
  double tot[W];
#pragma unroll
  for (int x = 0; x < W; ++x) tot[x] = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int x = 0; x < W; ++x)
      for (int w = 0; w < TPB / 32; ++w) tot[x] += red[w][x];
  }
  grid_finish<W>(tot, partial, ticket, out, 1.0);
}

// ---- kernels ----------------------------------------------------------------------------------------------------------
// update_rigid_bodies + tBody_update (reference src/EmDeeData.f90:420-439, src/ArBee.f90:97-130): make the body whole
// around its first member (minimum image, written back to R), centre of mass, member offsets delta = r - rcm (what the
// force kernel's rigid-body virial reads), inertia tensor -> principal frame -> MoI, quaternion, body-frame member
// coordinates. Momenta are left as they are. The unwrapping uses individually rounded operations so that the
// coordinates are the reference's to the last bit.
__global__ void __launch_bounds__(TPB) k_body_frame(BodyView v, double* __restrict__ R, double* __restrict__ delta, double L) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= v.nb) return;
  const int k0 = v.first[b], k1 = v.first[b + 1];
  const double invL = 1.0 / L;
  const size_t a0 = (size_t)v.atom[k0];
  const double r0[3] = {R[3 * a0], R[3 * a0 + 1], R[3 * a0 + 2]};
  double msum = 0.0, c[3] = {0.0, 0.0, 0.0};
  for (int k = k0; k < k1; ++k) {
    const size_t a = (size_t)v.atom[k];
    const double m = v.mItem[k];
    msum += m;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      double r = R[3 * a + x];
      if (k > k0) {
        r = __dsub_rn(r, __dmul_rn(L, round(__dmul_rn(invL, __dsub_rn(r, r0[x])))));
        R[3 * a + x] = r;
      }
      c[x] = __dadd_rn(c[x], __dmul_rn(m, r));
    }
  }
  const double inv = 1.0 / msum;
  v.mass[b] = msum;
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    c[x] = __dmul_rn(c[x], inv);
    v.rcm[(size_t)x * v.nb + b] = c[x];
  }
  double t[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int k = k0; k < k1; ++k) {
    const size_t a = (size_t)v.atom[k];
    const double m = v.mItem[k];
    const double x = __dsub_rn(R[3 * a], c[0]), y = __dsub_rn(R[3 * a + 1], c[1]), z = __dsub_rn(R[3 * a + 2], c[2]);
    delta[3 * a] = x;
    delta[3 * a + 1] = y;
    delta[3 * a + 2] = z;
    t[0] += m * (y * y + z * z);
    t[3] += m * (x * x + z * z);
    t[5] += m * (x * x + y * y);
    t[1] += m * x * y;
    t[2] += m * x * z;
    t[4] += m * y * z;
  }
  t[1] = -t[1]; t[2] = -t[2]; t[4] = -t[4];
  double w[3], ax[3][3], q[4];
  principal_axes(t, w, ax);
  quaternion_of_axes(ax, q);
#pragma unroll
  for (int x = 0; x < 3; ++x) v.MoI[(size_t)x * v.nb + b] = w[x];
#pragma unroll
  for (int x = 0; x < 4; ++x) v.q[(size_t)x * v.nb + b] = q[x];
  for (int k = k0; k < k1; ++k) {
    const size_t a = (size_t)v.atom[k];
    const double dl[3] = {delta[3 * a], delta[3 * a + 1], delta[3 * a + 2]};
#pragma unroll
    for (int x = 0; x < 3; ++x) v.d[3 * (size_t)k + x] = ax[x][0] * dl[0] + ax[x][1] * dl[1] + ax[x][2] * dl[2];
  }
}

// force_and_torque + boost of the centre-of-mass and quaternion momenta (+ the body part of kinetic_energies).
// out[0..2] = sum 1/M pcm^2 per dimension, out[3..5] = sum I omega^2 per principal axis.
// Several GPUs: body state is replicated on every rank, but a rank only holds the forces of the atoms it owns. So the
// kernel runs in two phases around an all-reduce of (F, tau): phase 1 sums the members this rank owns, phase 2 kicks
// every body with the reduced sums -- identical arithmetic on identical inputs, so the replicas stay bit-identical.
// phase 0 (one GPU) does both at once.
__global__ void __launch_bounds__(TPB) k_body_boost(BodyView v, const double* __restrict__ F, const double* __restrict__ delta,
                                                    const unsigned char* __restrict__ owned, int phase,
                                                    double CP, double CF, int translate, int rotate, int want_ke,
                                                    double* __restrict__ partial, unsigned int* __restrict__ ticket,
                                                    double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  double ke[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (b < v.nb) {
    double Fb[3] = {0.0, 0.0, 0.0}, tq[3] = {0.0, 0.0, 0.0};
    if (phase != 2) {
      for (int k = v.first[b]; k < v.first[b + 1]; ++k) {
        const size_t a = (size_t)v.atom[k];
        if (owned != nullptr && !owned[a]) continue;
        const double f[3] = {F[3 * a], F[3 * a + 1], F[3 * a + 2]};
        const double dl[3] = {delta[3 * a], delta[3 * a + 1], delta[3 * a + 2]};
        double c[3];
        cross3(dl, f, c);
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          Fb[x] += f[x];
          tq[x] += c[x];
        }
      }
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        v.Fb[(size_t)x * v.nb + b] = Fb[x];
        v.tau[(size_t)x * v.nb + b] = tq[x];
      }
    } else {
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        Fb[x] = v.Fb[(size_t)x * v.nb + b];
        tq[x] = v.tau[(size_t)x * v.nb + b];
      }
    }
    if (phase != 1) {
      const double invM = 1.0 / v.mass[b];
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        double p = v.pcm[(size_t)x * v.nb + b];
        if (translate) {
          p = CP * p + CF * Fb[x];
          v.pcm[(size_t)x * v.nb + b] = p;
        }
        ke[x] = invM * p * p;
      }
      Rotor r;
      load_rotor(v, b, r);
      if (rotate) {
        const double t3[3] = {2.0 * CF * tq[0], 2.0 * CF * tq[1], 2.0 * CF * tq[2]};
        double t4[4];
        vec_times_quat(t3, r.q, t4);
#pragma unroll
        for (int x = 0; x < 4; ++x) r.pi[x] = CP * r.pi[x] + t4[x];
        rotor_omega_from_pi(r);
        store_rotor(v, b, r);
      }
#pragma unroll
      for (int x = 0; x < 3; ++x) ke[3 + x] = r.I[x] * r.w[x] * r.w[x];
    }
  }
  if (!want_ke || phase == 1) return;
  reduce_and_finish<6>(ke, partial, ticket, out);
}

__global__ void __launch_bounds__(TPB) k_mask_and(int N, const unsigned char* __restrict__ a, const unsigned char* __restrict__ b,
                                                  unsigned char* __restrict__ o) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) o[i] = (unsigned char)(a[i] && b[i]);
}

// move: centre of mass drift + free rotation + member coordinates R = rcm + delta (only when rotating, like the reference)
__global__ void __launch_bounds__(TPB) k_body_move(BodyView v, double* __restrict__ R, double* __restrict__ delta,
                                                   double CR, double CP, double dt, int translate, int rotate, int mode) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= v.nb) return;
  double c[3];
  const double invM = 1.0 / v.mass[b];
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    c[x] = v.rcm[(size_t)x * v.nb + b];
    if (translate) {
      c[x] = CR * c[x] + CP * invM * v.pcm[(size_t)x * v.nb + b];
      v.rcm[(size_t)x * v.nb + b] = c[x];
    }
  }
  if (!rotate) return;
  Rotor r;
  load_rotor(v, b, r);
  bool refresh = true;
  if (mode == 0) refresh = rotor_exact(r, dt);
  else rotor_no_squish(r, dt, mode);
  store_rotor(v, b, r);
  double A[9];
  rotation_of(r.q, A);
  for (int k = v.first[b]; k < v.first[b + 1]; ++k) {
    const size_t a = (size_t)v.atom[k];
    double dl[3];
    if (refresh) {
      const double db[3] = {v.d[3 * (size_t)k], v.d[3 * (size_t)k + 1], v.d[3 * (size_t)k + 2]};
      apply3(A, db, dl);
#pragma unroll
      for (int x = 0; x < 3; ++x) delta[3 * a + x] = dl[x];
    } else {
#pragma unroll
      for (int x = 0; x < 3; ++x) dl[x] = delta[3 * a + x];
    }
#pragma unroll
    for (int x = 0; x < 3; ++x) R[3 * a + x] = c[x] + dl[x];
  }
}

// particle_momenta: P_k = m_k (pcm/M + omega_space x delta_k), written into the member slots of P
__global__ void __launch_bounds__(TPB) k_body_momenta(BodyView v, const double* __restrict__ delta, double* __restrict__ P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= v.nb) return;
  Rotor r;
  load_rotor(v, b, r);
  double A[9], ws[3], vc[3];
  rotation_of(r.q, A);
  apply3(A, r.w, ws);
  const double invM = 1.0 / v.mass[b];
#pragma unroll
  for (int x = 0; x < 3; ++x) vc[x] = invM * v.pcm[(size_t)x * v.nb + b];
  for (int k = v.first[b]; k < v.first[b + 1]; ++k) {
    const size_t a = (size_t)v.atom[k];
    const double dl[3] = {delta[3 * a], delta[3 * a + 1], delta[3 * a + 2]};
    double c[3];
    cross3(ws, dl, c);
#pragma unroll
    for (int x = 0; x < 3; ++x) P[3 * a + x] = v.mItem[k] * (vc[x] + c[x]);
  }
}

// assign_momenta: pcm = sum P_k, pi = C(q) (2 sum delta_k x P_k), omega from pi; kinetic sums as in k_body_boost
__global__ void __launch_bounds__(TPB) k_body_take_momenta(BodyView v, const double* __restrict__ delta,
                                                           const double* __restrict__ P, double* __restrict__ partial,
                                                           unsigned int* __restrict__ ticket, double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  double ke[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (b < v.nb) {
    double pc[3] = {0.0, 0.0, 0.0}, L[3] = {0.0, 0.0, 0.0};
    for (int k = v.first[b]; k < v.first[b + 1]; ++k) {
      const size_t a = (size_t)v.atom[k];
      const double p[3] = {P[3 * a], P[3 * a + 1], P[3 * a + 2]};
      const double dl[3] = {delta[3 * a], delta[3 * a + 1], delta[3 * a + 2]};
      double c[3];
      cross3(dl, p, c);
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        pc[x] += p[x];
        L[x] += c[x];
      }
    }
    const double invM = 1.0 / v.mass[b];
    Rotor r;
    load_rotor(v, b, r);
    const double twoL[3] = {2.0 * L[0], 2.0 * L[1], 2.0 * L[2]};
    vec_times_quat(twoL, r.q, r.pi);
    rotor_omega_from_pi(r);
    store_rotor(v, b, r);
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      v.pcm[(size_t)x * v.nb + b] = pc[x];
      ke[x] = invM * pc[x] * pc[x];
      ke[3 + x] = r.I[x] * r.w[x] * r.w[x];
    }
  }
  reduce_and_finish<6>(ke, partial, ticket, out);
}

// sum over free atoms of p^2/m per dimension (EmDee_upload "momenta", reference assign_momenta src/EmDeeData.f90:168-172);
// 4 wide with the last slot unused so that the grid_finish<3> instantiation of k_boost keeps its shared-memory layout
__global__ void __launch_bounds__(TPB) k_free_kinetic(int N, const unsigned char* __restrict__ isFree, const double* __restrict__ P,
                                                      const double* __restrict__ invMass, double* __restrict__ partial,
                                                      unsigned int* __restrict__ ticket, double* __restrict__ out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (a < N && (isFree == nullptr || isFree[a])) {
    const double im = invMass[a];
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      const double p = P[3 * (size_t)a + x];
      acc[x] = im * p * p;
    }
  }
  reduce_and_finish<4>(acc, partial, ticket, out);
}

// omega (and pcm) were just written by the host (random momenta): derive the quaternion momenta
__global__ void __launch_bounds__(TPB) k_body_set_omega(BodyView v) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= v.nb) return;
  Rotor r;
  load_rotor(v, b, r);
  rotor_pi_from_omega(r);
  store_rotor(v, b, r);
}

// ---- EmDee_verlet_step bookkeeping (shadow Hamiltonian), reference src/EmDeeCode.f90:1150-1209 -------------------------
// orientation after a "virtual" kick-and-rotate of duration t, starting from the current state
__device__ __forceinline__ void virtual_rotation(const BodyView& v, int b, double t, int mode, double q[4]) {
  Rotor r;
  load_rotor(v, b, r);
  const double t3[3] = {t * v.tau[b], t * v.tau[(size_t)v.nb + b], t * v.tau[2 * (size_t)v.nb + b]};
  double t4[4];
  vec_times_quat(t3, r.q, t4);
#pragma unroll
  for (int x = 0; x < 4; ++x) r.pi[x] += t4[x];
  rotor_omega_from_pi(r);
  if (mode != 0) rotor_no_squish(r, t, mode);
  else rotor_exact(r, t);
#pragma unroll
  for (int x = 0; x < 4; ++x) q[x] = r.q[x];
}

__global__ void __launch_bounds__(TPB) k_shadow_pre_bodies(BodyView v, double dt, int mode, double* __restrict__ r0,
                                                           double* __restrict__ q0) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= v.nb) return;
  const double h = 0.5 * dt, invM = 1.0 / v.mass[b];
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    const size_t i = (size_t)x * v.nb + b;
    r0[i] = 2.5 * v.rcm[i] + h * invM * (v.pcm[i] - h * v.Fb[i]);
  }
  double qv[4];
  virtual_rotation(v, b, -dt, mode, qv);
#pragma unroll
  for (int x = 0; x < 4; ++x) q0[(size_t)x * v.nb + b] = 0.5 * qv[x] - 3.0 * v.q[(size_t)x * v.nb + b];
}

// out[0] = Us, out[1] = Ks_t, out[2] = Ks_r (body contributions); the sum is 4 wide (last slot unused) so that this
// kernel does not share the grid_finish<3> instantiation -- and its shared-memory symbols -- with k_boost
__global__ void __launch_bounds__(TPB) k_shadow_post_bodies(BodyView v, double dt, int mode, const double* __restrict__ r0,
                                                            const double* __restrict__ q0, double* __restrict__ partial,
                                                            unsigned int* __restrict__ ticket, double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (b < v.nb) {
    const double h = 0.5 * dt, invM = 1.0 / v.mass[b];
    double Fb[3], tq[3], ff = 0.0, kt = 0.0;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      const size_t i = (size_t)x * v.nb + b;
      Fb[x] = v.Fb[i];
      tq[x] = v.tau[i];
      const double rdot = 2.5 * v.rcm[i] + dt * invM * (v.pcm[i] + h * Fb[x]) - r0[i];
      kt += rdot * v.pcm[i];
      ff += Fb[x] * Fb[x];
    }
    Rotor r;
    load_rotor(v, b, r);
    double qv[4], qd[4], qq = 0.0, kr = 0.0;
    virtual_rotation(v, b, dt, mode, qv);
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      qd[x] = q0[(size_t)x * v.nb + b] + 1.5 * r.q[x] + qv[x];
      qq += qd[x] * r.q[x];
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) kr += (qd[x] - qq * r.q[x]) * r.pi[x];
    double t4[4], tb[3], tt = 0.0;
    vec_times_quat(tq, r.q, t4);          // C(q) tau
    conj_times_quat_vec(r.q, t4, tb);     // Bt(q) C(q) tau: the torque in the body frame
#pragma unroll
    for (int x = 0; x < 3; ++x) tt += (1.0 / r.I[x]) * tb[x] * tb[x];
    acc[0] = invM * ff + tt;
    acc[1] = kt;
    acc[2] = kr;
  }
  reduce_and_finish<4>(acc, partial, ticket, out);
}

__global__ void __launch_bounds__(TPB) k_shadow_pre_atoms(int N, const unsigned char* __restrict__ isFree, double dt,
                                                          const double* __restrict__ R, const double* __restrict__ P,
                                                          const double* __restrict__ F, const double* __restrict__ invMass,
                                                          double* __restrict__ s0) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= N || (isFree != nullptr && !isFree[a])) return;
  const double h = 0.5 * dt, im = invMass[a];
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    const size_t i = 3 * (size_t)a + x;
    s0[i] = 2.5 * R[i] + h * im * (P[i] - h * F[i]);
  }
}

// out[0] = Us, out[1] = Ks_t (free-atom contributions)
__global__ void __launch_bounds__(TPB) k_shadow_post_atoms(int N, const unsigned char* __restrict__ isFree, double dt,
                                                           const double* __restrict__ R, const double* __restrict__ P,
                                                           const double* __restrict__ F, const double* __restrict__ invMass,
                                                           const double* __restrict__ s0, double* __restrict__ partial,
                                                           unsigned int* __restrict__ ticket, double* __restrict__ out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[2] = {0.0, 0.0};
  if (a < N && (isFree == nullptr || isFree[a])) {
    const double h = 0.5 * dt, im = invMass[a];
    double ff = 0.0, kt = 0.0;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      const size_t i = 3 * (size_t)a + x;
      const double rdot = 2.5 * R[i] + dt * im * (P[i] + h * F[i]) - s0[i];
      kt += rdot * P[i];
      ff += F[i] * F[i];
    }
    acc[0] = im * ff;
    acc[1] = kt;
  }
  reduce_and_finish<2>(acc, partial, ticket, out);
}

}  // namespace
}  // namespace emdee
