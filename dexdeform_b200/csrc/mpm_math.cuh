// mpm_math.cuh -- device math shared by every kernel of the MLS-MPM hot path (sm_100a only).
//
// Semantics follow the reference leaf functions (cited per function, paths relative to the DexDeform repo);
// the implementation is written for registers: row-major 3x3 matrices as 9 named floats, no local arrays
// that could spill, explicit fmaf where the order does not matter for parity.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace dd {

#define DD_DEV __device__ __forceinline__

struct V3 { float x, y, z; };
struct Q4 { float w, x, y, z; };           // (w,x,y,z) as mpm/csrc/quat.h:5-12
struct M3 { float a00, a01, a02, a10, a11, a12, a20, a21, a22; };  // row-major, mpm/csrc/mat3.h:8

DD_DEV V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
DD_DEV V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
DD_DEV V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
DD_DEV V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
DD_DEV V3 operator*(float s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
DD_DEV V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
DD_DEV V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
DD_DEV void operator+=(V3 &a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
DD_DEV void operator-=(V3 &a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
DD_DEV float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DD_DEV V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DD_DEV V3 vabs(V3 a) { return v3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
DD_DEV V3 vmax(V3 a, float b) { return v3(fmaxf(a.x, b), fmaxf(a.y, b), fmaxf(a.z, b)); }
DD_DEV V3 vzero() { return v3(0.f, 0.f, 0.f); }

DD_DEV M3 m3(float a00, float a01, float a02, float a10, float a11, float a12, float a20, float a21, float a22) {
  M3 r; r.a00 = a00; r.a01 = a01; r.a02 = a02; r.a10 = a10; r.a11 = a11; r.a12 = a12; r.a20 = a20; r.a21 = a21; r.a22 = a22; return r;
}
DD_DEV M3 mzero() { return m3(0, 0, 0, 0, 0, 0, 0, 0, 0); }
DD_DEV M3 mdiag(V3 d) { return m3(d.x, 0, 0, 0, d.y, 0, 0, 0, d.z); }
DD_DEV M3 mdiag(float d) { return m3(d, 0, 0, 0, d, 0, 0, 0, d); }
DD_DEV M3 transpose(M3 m) { return m3(m.a00, m.a10, m.a20, m.a01, m.a11, m.a21, m.a02, m.a12, m.a22); }
DD_DEV M3 operator+(M3 a, M3 b) { return m3(a.a00 + b.a00, a.a01 + b.a01, a.a02 + b.a02, a.a10 + b.a10, a.a11 + b.a11, a.a12 + b.a12, a.a20 + b.a20, a.a21 + b.a21, a.a22 + b.a22); }
DD_DEV M3 operator-(M3 a, M3 b) { return m3(a.a00 - b.a00, a.a01 - b.a01, a.a02 - b.a02, a.a10 - b.a10, a.a11 - b.a11, a.a12 - b.a12, a.a20 - b.a20, a.a21 - b.a21, a.a22 - b.a22); }
DD_DEV M3 operator*(M3 a, float s) { return m3(a.a00 * s, a.a01 * s, a.a02 * s, a.a10 * s, a.a11 * s, a.a12 * s, a.a20 * s, a.a21 * s, a.a22 * s); }
DD_DEV M3 operator*(float s, M3 a) { return a * s; }
DD_DEV void operator+=(M3 &a, M3 b) { a = a + b; }
// elementwise product (mat3.h:117-129)
DD_DEV M3 hadamard(M3 a, M3 b) { return m3(a.a00 * b.a00, a.a01 * b.a01, a.a02 * b.a02, a.a10 * b.a10, a.a11 * b.a11, a.a12 * b.a12, a.a20 * b.a20, a.a21 * b.a21, a.a22 * b.a22); }
// Packed fp32 pairs (sm_100): one FFMA2 / FMUL2 instruction works on an aligned register pair, and either multiplicand may be
// a single register broadcast to both halves -- "scalar times row" costs two issue slots instead of three for a 3-vector.
// Each half is an ordinary IEEE fma / mul, so results equal the scalar formulation with the same association.
DD_DEV float2 pk(float a, float b) { return make_float2(a, b); }
DD_DEV float2 fma2(float s, float2 v, float2 acc) { return __ffma2_rn(make_float2(s, s), v, acc); }
DD_DEV float2 fma2(float2 a, float2 b, float2 acc) { return __ffma2_rn(a, b, acc); }
DD_DEV float2 mul2(float s, float2 v) { return __fmul2_rn(make_float2(s, s), v); }
DD_DEV float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
// matrix product, k accumulated 0,1,2 (mat3.h:86-96); columns 0,1 of every row as one packed pair
DD_DEV M3 mul(M3 a, M3 b) {
  float2 b0 = pk(b.a00, b.a01), b1 = pk(b.a10, b.a11), b2 = pk(b.a20, b.a21);
  float2 r0 = fma2(a.a02, b2, fma2(a.a01, b1, mul2(a.a00, b0)));
  float2 r1 = fma2(a.a12, b2, fma2(a.a11, b1, mul2(a.a10, b0)));
  float2 r2 = fma2(a.a22, b2, fma2(a.a21, b1, mul2(a.a20, b0)));
  return m3(r0.x, r0.y, fmaf(a.a02, b.a22, fmaf(a.a01, b.a12, a.a00 * b.a02)),
            r1.x, r1.y, fmaf(a.a12, b.a22, fmaf(a.a11, b.a12, a.a10 * b.a02)),
            r2.x, r2.y, fmaf(a.a22, b.a22, fmaf(a.a21, b.a12, a.a20 * b.a02)));
}
// a * b^T and a^T * b without materialising the transpose
DD_DEV M3 mul_nt(M3 a, M3 b) { return mul(a, transpose(b)); }
DD_DEV M3 mul_tn(M3 a, M3 b) { return mul(transpose(a), b); }
DD_DEV V3 mul(M3 a, V3 b) {
  return v3(a.a00 * b.x + a.a01 * b.y + a.a02 * b.z, a.a10 * b.x + a.a11 * b.y + a.a12 * b.z, a.a20 * b.x + a.a21 * b.y + a.a22 * b.z);
}
DD_DEV V3 mul_t(M3 a, V3 b) {  // a^T b
  return v3(a.a00 * b.x + a.a10 * b.y + a.a20 * b.z, a.a01 * b.x + a.a11 * b.y + a.a21 * b.z, a.a02 * b.x + a.a12 * b.y + a.a22 * b.z);
}
DD_DEV M3 outer(V3 a, V3 b) { return m3(a.x * b.x, a.x * b.y, a.x * b.z, a.y * b.x, a.y * b.y, a.y * b.z, a.z * b.x, a.z * b.y, a.z * b.z); }
DD_DEV float trace(M3 a) { return a.a00 + a.a11 + a.a22; }
DD_DEV V3 diag(M3 a) { return v3(a.a00, a.a11, a.a22); }
// sum_ij a_ij b_ij
DD_DEV float ddot(M3 a, M3 b) {
  return a.a00 * b.a00 + a.a01 * b.a01 + a.a02 * b.a02 + a.a10 * b.a10 + a.a11 * b.a11 + a.a12 * b.a12 + a.a20 * b.a20 + a.a21 * b.a21 + a.a22 * b.a22;
}
// M * diag(d)  (scales columns)
DD_DEV M3 mul_diag(M3 a, V3 d) { return m3(a.a00 * d.x, a.a01 * d.y, a.a02 * d.z, a.a10 * d.x, a.a11 * d.y, a.a12 * d.z, a.a20 * d.x, a.a21 * d.y, a.a22 * d.z); }
// diag(d) * M  (scales rows)
DD_DEV M3 diag_mul(V3 d, M3 a) { return m3(a.a00 * d.x, a.a01 * d.x, a.a02 * d.x, a.a10 * d.y, a.a11 * d.y, a.a12 * d.y, a.a20 * d.z, a.a21 * d.z, a.a22 * d.z); }

// AoS access in the reference layouts (vec3 = 12 B, mat3 = 36 B, quat = 16 B; mpm/types.py:25-75)
DD_DEV V3 ld_v3(const float *p, int i) { const float *q = p + 3 * (size_t)i; return v3(q[0], q[1], q[2]); }
DD_DEV void st_v3(float *p, int i, V3 v) { float *q = p + 3 * (size_t)i; q[0] = v.x; q[1] = v.y; q[2] = v.z; }
DD_DEV void add_v3(float *p, int i, V3 v) { float *q = p + 3 * (size_t)i; q[0] += v.x; q[1] += v.y; q[2] += v.z; }
DD_DEV M3 ld_m3(const float *p, int i) { const float *q = p + 9 * (size_t)i; return m3(q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8]); }
DD_DEV void st_m3(float *p, int i, M3 m) {
  float *q = p + 9 * (size_t)i;
  q[0] = m.a00; q[1] = m.a01; q[2] = m.a02; q[3] = m.a10; q[4] = m.a11; q[5] = m.a12; q[6] = m.a20; q[7] = m.a21; q[8] = m.a22;
}
DD_DEV void add_m3(float *p, int i, M3 m) {
  float *q = p + 9 * (size_t)i;
  q[0] += m.a00; q[1] += m.a01; q[2] += m.a02; q[3] += m.a10; q[4] += m.a11; q[5] += m.a12; q[6] += m.a20; q[7] += m.a21; q[8] += m.a22;
}
DD_DEV Q4 ld_q4(const float *p, int i) { float4 t = reinterpret_cast<const float4 *>(p)[i]; Q4 q; q.w = t.x; q.x = t.y; q.y = t.z; q.z = t.w; return q; }

// ---------------------------------------------------------------------------------------------- quaternions
DD_DEV Q4 qconj(Q4 q) { Q4 r; r.w = q.w; r.x = -q.x; r.y = -q.y; r.z = -q.z; return r; }
// v + 2 (w (u x v) + u x (u x v))   (quat.h:14-19)
DD_DEV V3 qrot(Q4 q, V3 v) {
  V3 u = v3(q.x, q.y, q.z);
  V3 uv = cross(u, v);
  V3 uuv = cross(u, uv);
  return v + 2.f * (q.w * uv + uuv);
}
DD_DEV V3 xform(V3 p, Q4 q, V3 pt) { return p + qrot(q, pt); }                // quat.h:89-92
DD_DEV V3 xform_inv(V3 p, Q4 q, V3 pt) { return qrot(qconj(q), pt - p); }     // quat.h:94-97
// adjoint of y = q * v : accumulates into gq, gv  (quat.h:25-47)
DD_DEV void qrot_adj(Q4 q, V3 v, V3 g, Q4 &gq, V3 &gv) {
  V3 u = v3(q.x, q.y, q.z);
  V3 uv = cross(u, v);
  gq.w += dot(uv, g) * 2.f;
  V3 g_uv = (2.f * q.w) * g;
  V3 g_uuv = 2.f * g;
  V3 g_u = cross(uv, g_uuv);
  g_uv = g_uv + cross(g_uuv, u);
  g_u += cross(v, g_uv);
  gv += g + cross(g_uv, u);
  gq.x += g_u.x; gq.y += g_u.y; gq.z += g_u.z;
}
// adjoint of y = p + q * pt  (quat.h:50-62)
DD_DEV void xform_adj(Q4 q, V3 pt, V3 g, V3 &gp, Q4 &gq, V3 &gpt) { gp += g; qrot_adj(q, pt, g, gq, gpt); }
// adjoint of y = q^-1 * (pt - p)  (quat.h:64-80)
DD_DEV void xform_inv_adj(V3 p, Q4 q, V3 pt, V3 g, V3 &gp, Q4 &gq, V3 &gpt) {
  Q4 tq; tq.w = tq.x = tq.y = tq.z = 0.f;
  V3 tp = vzero();
  qrot_adj(qconj(q), pt - p, g, tq, tp);
  gq.w += tq.w; gq.x -= tq.x; gq.y -= tq.y; gq.z -= tq.z;
  gp -= tp;
  gpt += tp;
}

// ---------------------------------------------------------------------------------------------- shapes
// length with the 1e-30 epsilon evaluated in double as in the reference (shape.h:4-9: float + double literal)
#ifdef DD_FLOAT_LENGTH
// engine kernels: the 1e-30 epsilon only matters for |a|^2 < 1e-23, where both forms return ~1e-15; no fp64 sqrt
DD_DEV float length30(V3 a) { return sqrtf(dot(a, a) + 1e-30f); }
#else
DD_DEV float length30(V3 a) { return (float)sqrt((double)dot(a, a) + 1e-30); }
#endif
DD_DEV V3 normalized(V3 a) { return a / length30(a); }                           // shape.h:12-14
DD_DEV V3 normalized_adj(V3 vec, V3 g) {                                         // shape.h:16-25
  float doted = (float)((double)dot(vec, vec) + 1e-30);
  V3 t = g - vec * dot(vec / doted, g);
  float k = (float)(1. / (double)sqrtf(doted));
  return t * k;
}
DD_DEV void abs_adj(V3 gx, V3 &g) {                                              // shape.h:27-31
  if (gx.x < 0) g.x = -g.x;
  if (gx.y < 0) g.y = -g.y;
  if (gx.z < 0) g.z = -g.z;
}
// body descriptor: tfsr = (type, friction, softness, round) and args, both stored as quats (simulator.py:396-397)
DD_DEV int shape_type(Q4 tfsr) { return (int)floorf(tfsr.w + 0.1f); }            // shape.h:33-35
DD_DEV float shape_sdf(Q4 tfsr, Q4 args, V3 gx) {                                // shape.h:37-61
  float sdf;
  if (shape_type(tfsr) == 0) {
    V3 q = vabs(gx) - v3(args.w, args.x, args.y);
    sdf = length30(vmax(q, 0.f)) + fminf(fmaxf(fmaxf(q.x, q.y), q.z), 0.f);
  } else {
    V3 p2 = gx;
    float r = args.w, h = args.x;
    p2.y += h / 2;
    p2.y -= fminf(fmaxf(p2.y, 0.f), h);
    sdf = length30(p2) - r;
  }
  return sdf - tfsr.z;
}
DD_DEV V3 shape_grad(Q4 tfsr, Q4 args, V3 gx) {                                  // shape.h:65-102
  if (shape_type(tfsr) == 0) {
    V3 grad = vzero();
    V3 q = vabs(gx) - v3(args.w, args.x, args.y);
    float inside = fmaxf(fmaxf(q.x, q.y), q.z);
    if (inside <= 0) {
      if (q.x == inside) grad.x += 1;
      if (q.y == inside) grad.y += 1;
      if (q.z == inside) grad.z += 1;
    } else {
      grad = normalized(vmax(q, 0.f));
    }
    abs_adj(gx, grad);
    return grad;
  }
  V3 p2 = gx;
  float h = args.x;
  p2.y += h / 2;
  p2.y -= fminf(fmaxf(p2.y, 0.f), h);
  return normalized(p2);
}
// adjoint of shape_grad w.r.t. gx.  The capsule branch returns exactly zero in the reference because its
// grad_in is never seeded from grad_out (shape.h:133-146); kept bug-for-bug.
DD_DEV V3 shape_grad_adj(Q4 tfsr, Q4 args, V3 gx, V3 gout) {                     // shape.h:104-152
  if (shape_type(tfsr) != 0) return vzero();
  V3 q = vabs(gx) - v3(args.w, args.x, args.y);
  float inside = fmaxf(fmaxf(q.x, q.y), q.z);
  if (inside <= 0) return vzero();
  V3 gin = gout;
  abs_adj(gx, gin);
  gin = normalized_adj(vmax(q, 0.f), gin);
  if (q.x < 0) gin.x = 0;
  if (q.y < 0) gin.y = 0;
  if (q.z < 0) gin.z = 0;
  abs_adj(gx, gin);
  return gin;
}
// contact activation test (integrator.cu:707-710): influence compared as double against 0.1
DD_DEV bool contact_active(float dist, float softness, float &influence) {
  influence = fminf(expf(-dist * softness), 1.f);
  return (softness > 0 && (double)influence > 0.1) || dist <= 0;
}

// ---------------------------------------------------------------------------------------------- B-spline stencil
struct Stencil {
  int bx, by, bz;   // base node
  V3 fx;            // x*inv_dx - base, in [0.5, 1.5)
  V3 w0, w1, w2;    // w_i per axis: w0 = weights of node offset 0 for (x,y,z), ...
};
// integrator.cu:348-352
DD_DEV Stencil make_stencil(V3 x, float inv_dx) {
  Stencil s;
  float tx = x.x * inv_dx - 0.5f, ty = x.y * inv_dx - 0.5f, tz = x.z * inv_dx - 0.5f;
  s.bx = (int)floorf(tx); s.by = (int)floorf(ty); s.bz = (int)floorf(tz);
  s.fx = v3(x.x * inv_dx - (float)s.bx, x.y * inv_dx - (float)s.by, x.z * inv_dx - (float)s.bz);
  V3 a = v3(1.5f, 1.5f, 1.5f) - s.fx, b = s.fx - v3(1.f, 1.f, 1.f), c = s.fx - v3(0.5f, 0.5f, 0.5f);
  s.w0 = 0.5f * (a * a);
  s.w1 = v3(0.75f, 0.75f, 0.75f) - b * b;
  s.w2 = 0.5f * (c * c);
  return s;
}
// d w_i / d x per axis (integrator.cu:493)
DD_DEV void stencil_dw(const Stencil &s, float inv_dx, V3 &d0, V3 &d1, V3 &d2) {
  d0 = (-inv_dx) * (v3(1.5f, 1.5f, 1.5f) - s.fx);
  d1 = inv_dx * ((-2.f) * s.fx + v3(2.f, 2.f, 2.f));
  d2 = (-inv_dx) * (s.fx * (-1.f) + v3(0.5f, 0.5f, 0.5f));
}
DD_DEV float pick(V3 a0, V3 a1, V3 a2, int i, int axis) {
  V3 t = i == 0 ? a0 : (i == 1 ? a1 : a2);
  return axis == 0 ? t.x : (axis == 1 ? t.y : t.z);
}

// ---------------------------------------------------------------------------------------------- 3x3 SVD
// (1) svd3_f64: the reference's algorithm (svd.h:109-413, ericjang/svd3 after McAdams et al.) in double:
//     8 Jacobi sweeps of A^T A with approximate Givens rotations accumulated in a quaternion, sign-preserving
//     column sort, Givens QR of A V.  Used by the ABI-1 compute_svd entry point so that U, V agree with the
//     reference even where they are not unique (repeated singular values).
DD_DEV void approx_givens64(double a11, double a12, double a22, double &ch, double &sh) {
  const double gamma = (double)5.82842712474619f, cstar = (double)0.9238795325112867f, sstar = (double)0.3826834323650897f;
  double c = 2 * (a11 - a22), s = a12;
  bool b = gamma * s * s < c * c;
  double w = rsqrt(c * c + s * s);
  ch = b ? w * c : cstar;
  sh = b ? w * s : sstar;
}
// one Jacobi conjugation on the symmetric matrix (s11,s21,s22,s31,s32,s33) followed by the cyclic relabelling;
// the quaternion components are passed already permuted so that no dynamic indexing is needed
DD_DEV void jacobi64(double &s11, double &s21, double &s22, double &s31, double &s32, double &s33,
                     double &qx, double &qy, double &qz, double &qw) {
  double ch, sh;
  approx_givens64(s11, s21, s22, ch, sh);
  double scale = ch * ch + sh * sh;
  double a = (ch * ch - sh * sh) / scale;
  double b = (2 * sh * ch) / scale;
  double n11 = a * (a * s11 + b * s21) + b * (a * s21 + b * s22);
  double n21 = a * (-b * s11 + a * s21) + b * (-b * s21 + a * s22);
  double n22 = -b * (-b * s11 + a * s21) + a * (-b * s21 + a * s22);
  double n31 = a * s31 + b * s32;
  double n32 = -b * s31 + a * s32;
  double n33 = s33;
  // q <- q * (ch, sh along the current z axis): tmp = q.xyz*sh; sh *= q.w; q *= ch; q[z]+=sh; q.w-=tmp[z]; q[x]+=tmp[y]; q[y]-=tmp[x]
  double tx = qx * sh, ty = qy * sh, tz = qz * sh;
  sh *= qw;
  qx *= ch; qy *= ch; qz *= ch; qw *= ch;
  qz += sh; qw -= tz; qx += ty; qy -= tx;
  s11 = n22; s21 = n32; s22 = n33; s31 = n21; s32 = n31; s33 = n11;
}
DD_DEV void cswap64(bool c, double &x, double &y) { double z = x; x = c ? y : x; y = c ? z : y; }
DD_DEV void cnswap64(bool c, double &x, double &y) { double z = -x; x = c ? y : x; y = c ? z : y; }
DD_DEV void qr_givens64(double a1, double a2, double &ch, double &sh) {
  double r2 = a1 * a1 + a2 * a2;
  double rho = r2 / sqrt(r2);
  sh = rho > 1e-6 ? a2 : 0;
  ch = fabs(a1) + fmax(rho, 1e-6);
  cswap64(a1 < 0, sh, ch);
  double w = rsqrt(ch * ch + sh * sh);
  ch *= w; sh *= w;
}
static __device__ __noinline__ void svd3_f64(const M3 &A, M3 &U, V3 &sig, M3 &Vo) {
  double a11 = A.a00, a12 = A.a01, a13 = A.a02, a21 = A.a10, a22 = A.a11, a23 = A.a12, a31 = A.a20, a32 = A.a21, a33 = A.a22;
  double s11 = a11 * a11 + a21 * a21 + a31 * a31, s21 = a12 * a11 + a22 * a21 + a32 * a31, s22 = a12 * a12 + a22 * a22 + a32 * a32;
  double s31 = a13 * a11 + a23 * a21 + a33 * a31, s32 = a13 * a12 + a23 * a22 + a33 * a32, s33 = a13 * a13 + a23 * a23 + a33 * a33;
  double qx = 0, qy = 0, qz = 0, qw = 1;
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    jacobi64(s11, s21, s22, s31, s32, s33, qx, qy, qz, qw);  // (x,y,z) = (0,1,2)
    jacobi64(s11, s21, s22, s31, s32, s33, qy, qz, qx, qw);  // (1,2,0)
    jacobi64(s11, s21, s22, s31, s32, s33, qz, qx, qy, qw);  // (2,0,1)
  }
  double v11, v12, v13, v21, v22, v23, v31, v32, v33;
  {
    double qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz, qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
    v11 = 1 - 2 * (qyy + qzz); v12 = 2 * (qxy - qwz); v13 = 2 * (qxz + qwy);
    v21 = 2 * (qxy + qwz); v22 = 1 - 2 * (qxx + qzz); v23 = 2 * (qyz - qwx);
    v31 = 2 * (qxz - qwy); v32 = 2 * (qyz + qwx); v33 = 1 - 2 * (qxx + qyy);
  }
  double b11 = a11 * v11 + a12 * v21 + a13 * v31, b12 = a11 * v12 + a12 * v22 + a13 * v32, b13 = a11 * v13 + a12 * v23 + a13 * v33;
  double b21 = a21 * v11 + a22 * v21 + a23 * v31, b22 = a21 * v12 + a22 * v22 + a23 * v32, b23 = a21 * v13 + a22 * v23 + a23 * v33;
  double b31 = a31 * v11 + a32 * v21 + a33 * v31, b32 = a31 * v12 + a32 * v22 + a33 * v32, b33 = a31 * v13 + a32 * v23 + a33 * v33;
  {
    double r1 = b11 * b11 + b21 * b21 + b31 * b31, r2 = b12 * b12 + b22 * b22 + b32 * b32, r3 = b13 * b13 + b23 * b23 + b33 * b33;
    bool c = r1 < r2;
    cnswap64(c, b11, b12); cnswap64(c, v11, v12); cnswap64(c, b21, b22); cnswap64(c, v21, v22); cnswap64(c, b31, b32); cnswap64(c, v31, v32);
    cswap64(c, r1, r2);
    c = r1 < r3;
    cnswap64(c, b11, b13); cnswap64(c, v11, v13); cnswap64(c, b21, b23); cnswap64(c, v21, v23); cnswap64(c, b31, b33); cnswap64(c, v31, v33);
    cswap64(c, r1, r3);
    c = r2 < r3;
    cnswap64(c, b12, b13); cnswap64(c, v12, v13); cnswap64(c, b22, b23); cnswap64(c, v22, v23); cnswap64(c, b32, b33); cnswap64(c, v32, v33);
  }
  double ch1, sh1, ch2, sh2, ch3, sh3, ca, cb;
  qr_givens64(b11, b21, ch1, sh1);
  ca = 1 - 2 * sh1 * sh1; cb = 2 * ch1 * sh1;
  double r11 = ca * b11 + cb * b21, r12 = ca * b12 + cb * b22, r13 = ca * b13 + cb * b23;
  double r21 = -cb * b11 + ca * b21, r22 = -cb * b12 + ca * b22, r23 = -cb * b13 + ca * b23;
  double r31 = b31, r32 = b32, r33 = b33;
  qr_givens64(r11, r31, ch2, sh2);
  ca = 1 - 2 * sh2 * sh2; cb = 2 * ch2 * sh2;
  b11 = ca * r11 + cb * r31; b12 = ca * r12 + cb * r32; b13 = ca * r13 + cb * r33;
  b21 = r21; b22 = r22; b23 = r23;
  b31 = -cb * r11 + ca * r31; b32 = -cb * r12 + ca * r32; b33 = -cb * r13 + ca * r33;
  qr_givens64(b22, b32, ch3, sh3);
  ca = 1 - 2 * sh3 * sh3; cb = 2 * ch3 * sh3;
  r22 = ca * b22 + cb * b32;
  r33 = -cb * b23 + ca * b33;
  double sh12 = sh1 * sh1, sh22 = sh2 * sh2, sh32 = sh3 * sh3;
  U.a00 = (float)((-1 + 2 * sh12) * (-1 + 2 * sh22));
  U.a01 = (float)(4 * ch2 * ch3 * (-1 + 2 * sh12) * sh2 * sh3 + 2 * ch1 * sh1 * (-1 + 2 * sh32));
  U.a02 = (float)(4 * ch1 * ch3 * sh1 * sh3 - 2 * ch2 * (-1 + 2 * sh12) * sh2 * (-1 + 2 * sh32));
  U.a10 = (float)(2 * ch1 * sh1 * (1 - 2 * sh22));
  U.a11 = (float)(-8 * ch1 * ch2 * ch3 * sh1 * sh2 * sh3 + (-1 + 2 * sh12) * (-1 + 2 * sh32));
  U.a12 = (float)(-2 * ch3 * sh3 + 4 * sh1 * (ch3 * sh1 * sh3 + ch1 * ch2 * sh2 * (-1 + 2 * sh32)));
  U.a20 = (float)(2 * ch2 * sh2);
  U.a21 = (float)(2 * ch3 * (1 - 2 * sh22) * sh3);
  U.a22 = (float)((-1 + 2 * sh22) * (-1 + 2 * sh32));
  Vo = m3((float)v11, (float)v12, (float)v13, (float)v21, (float)v22, (float)v23, (float)v31, (float)v32, (float)v33);
  sig = v3((float)b11, (float)r22, (float)r33);
}

// (2) svd3_f32: the production SVD.  Same structure (Jacobi on A^T A with quaternion accumulation, column sort,
//     Givens QR) but entirely in fp32 registers with hardware rsqrt, NSWEEP sweeps, and a re-normalised quaternion.
//     Accuracy is measured in tests (|U S V^T - A|, orthogonality, sigma vs fp64) -- a few fp32 ulp for the
//     well-conditioned deformation gradients of this path.  ~3.5x fewer instructions than (1) and none of them fp64.
#ifdef __CUDA_ARCH__
#define DD_RSQRT(x) rsqrtf(x)
#else
#define DD_RSQRT(x) (1.0f / sqrtf(x))
#endif
#define DD_HD __host__ __device__ __forceinline__
DD_HD void jacobi32(float &s11, float &s21, float &s22, float &s31, float &s32, float &s33, float &qx, float &qy, float &qz, float &qw) {
  // approximate Givens half-angle (McAdams et al. 2011, Alg. 2): (ch, sh) ~ (2(s11 - s22), s21), clamped to pi/8 steps
  float ch = 2.f * (s11 - s22), sh = s21;
  bool b = 5.82842712474619f * sh * sh < ch * ch;
  float w = DD_RSQRT(ch * ch + sh * sh);
  ch = b ? w * ch : 0.9238795325112867f;
  sh = b ? w * sh : 0.3826834323650897f;
  // rotation (a, b) = (cos, sin) of the full angle; (ch, sh) is a unit vector so no rescale is needed
  float a = ch * ch - sh * sh, bb = 2.f * sh * ch;
  float t1 = -bb * s11 + a * s21, t2 = -bb * s21 + a * s22, u1 = a * s11 + bb * s21, u2 = a * s21 + bb * s22;
  float n11 = a * u1 + bb * u2, n21 = a * t1 + bb * t2, n22 = -bb * t1 + a * t2;
  float n31 = a * s31 + bb * s32, n32 = -bb * s31 + a * s32, n33 = s33;
  float tx = qx * sh, ty = qy * sh, tz = qz * sh;
  sh *= qw;
  qx *= ch; qy *= ch; qz *= ch; qw *= ch;
  qz += sh; qw -= tz; qx += ty; qy -= tx;
  s11 = n22; s21 = n32; s22 = n33; s31 = n21; s32 = n31; s33 = n11;
}
DD_HD void cswap32(bool c, float &x, float &y) { float z = x; x = c ? y : x; y = c ? z : y; }
DD_HD void cnswap32(bool c, float &x, float &y) { float z = -x; x = c ? y : x; y = c ? z : y; }
DD_HD void qr_givens32(float a1, float a2, float &ch, float &sh) {
  float r2 = a1 * a1 + a2 * a2;
  float rho = r2 * DD_RSQRT(fmaxf(r2, 1e-37f));
  sh = rho > 1e-6f ? a2 : 0.f;
  ch = fabsf(a1) + fmaxf(rho, 1e-6f);
  cswap32(a1 < 0.f, sh, ch);
  float w = DD_RSQRT(ch * ch + sh * sh);
  ch *= w; sh *= w;
}
template <int NSWEEP = 5>
DD_HD void svd3_f32(const M3 &A, M3 &U, V3 &sig, M3 &Vo) {
  float s11 = A.a00 * A.a00 + A.a10 * A.a10 + A.a20 * A.a20, s21 = A.a01 * A.a00 + A.a11 * A.a10 + A.a21 * A.a20;
  float s22 = A.a01 * A.a01 + A.a11 * A.a11 + A.a21 * A.a21, s31 = A.a02 * A.a00 + A.a12 * A.a10 + A.a22 * A.a20;
  float s32 = A.a02 * A.a01 + A.a12 * A.a11 + A.a22 * A.a21, s33 = A.a02 * A.a02 + A.a12 * A.a12 + A.a22 * A.a22;
  float qx = 0.f, qy = 0.f, qz = 0.f, qw = 1.f;
#pragma unroll
  for (int it = 0; it < NSWEEP; ++it) {
    jacobi32(s11, s21, s22, s31, s32, s33, qx, qy, qz, qw);
    jacobi32(s11, s21, s22, s31, s32, s33, qy, qz, qx, qw);
    jacobi32(s11, s21, s22, s31, s32, s33, qz, qx, qy, qw);
  }
  {
    float n = DD_RSQRT(qx * qx + qy * qy + qz * qz + qw * qw);
    qx *= n; qy *= n; qz *= n; qw *= n;
  }
  float qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz, qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
  float v11 = 1.f - 2.f * (qyy + qzz), v12 = 2.f * (qxy - qwz), v13 = 2.f * (qxz + qwy);
  float v21 = 2.f * (qxy + qwz), v22 = 1.f - 2.f * (qxx + qzz), v23 = 2.f * (qyz - qwx);
  float v31 = 2.f * (qxz - qwy), v32 = 2.f * (qyz + qwx), v33 = 1.f - 2.f * (qxx + qyy);
  float b11 = A.a00 * v11 + A.a01 * v21 + A.a02 * v31, b12 = A.a00 * v12 + A.a01 * v22 + A.a02 * v32, b13 = A.a00 * v13 + A.a01 * v23 + A.a02 * v33;
  float b21 = A.a10 * v11 + A.a11 * v21 + A.a12 * v31, b22 = A.a10 * v12 + A.a11 * v22 + A.a12 * v32, b23 = A.a10 * v13 + A.a11 * v23 + A.a12 * v33;
  float b31 = A.a20 * v11 + A.a21 * v21 + A.a22 * v31, b32 = A.a20 * v12 + A.a21 * v22 + A.a22 * v32, b33 = A.a20 * v13 + A.a21 * v23 + A.a22 * v33;
  {
    float r1 = b11 * b11 + b21 * b21 + b31 * b31, r2 = b12 * b12 + b22 * b22 + b32 * b32, r3 = b13 * b13 + b23 * b23 + b33 * b33;
    bool c = r1 < r2;
    cnswap32(c, b11, b12); cnswap32(c, v11, v12); cnswap32(c, b21, b22); cnswap32(c, v21, v22); cnswap32(c, b31, b32); cnswap32(c, v31, v32);
    cswap32(c, r1, r2);
    c = r1 < r3;
    cnswap32(c, b11, b13); cnswap32(c, v11, v13); cnswap32(c, b21, b23); cnswap32(c, v21, v23); cnswap32(c, b31, b33); cnswap32(c, v31, v33);
    cswap32(c, r1, r3);
    c = r2 < r3;
    cnswap32(c, b12, b13); cnswap32(c, v12, v13); cnswap32(c, b22, b23); cnswap32(c, v22, v23); cnswap32(c, b32, b33); cnswap32(c, v32, v33);
  }
  float ch1, sh1, ch2, sh2, ch3, sh3, ca, cb;
  qr_givens32(b11, b21, ch1, sh1);
  ca = 1.f - 2.f * sh1 * sh1; cb = 2.f * ch1 * sh1;
  float r11 = ca * b11 + cb * b21, r12 = ca * b12 + cb * b22, r13 = ca * b13 + cb * b23;
  float r21 = -cb * b11 + ca * b21, r22 = -cb * b12 + ca * b22, r23 = -cb * b13 + ca * b23;
  float r31 = b31, r32 = b32, r33 = b33;
  (void)r21;
  qr_givens32(r11, r31, ch2, sh2);
  ca = 1.f - 2.f * sh2 * sh2; cb = 2.f * ch2 * sh2;
  b11 = ca * r11 + cb * r31;
  b22 = r22; b23 = r23;
  b32 = -cb * r12 + ca * r32; b33 = -cb * r13 + ca * r33;
  qr_givens32(b22, b32, ch3, sh3);
  ca = 1.f - 2.f * sh3 * sh3; cb = 2.f * ch3 * sh3;
  r22 = ca * b22 + cb * b32;
  r33 = -cb * b23 + ca * b33;
  float sh12 = sh1 * sh1, sh22 = sh2 * sh2, sh32 = sh3 * sh3;
  U.a00 = (-1.f + 2.f * sh12) * (-1.f + 2.f * sh22);
  U.a01 = 4.f * ch2 * ch3 * (-1.f + 2.f * sh12) * sh2 * sh3 + 2.f * ch1 * sh1 * (-1.f + 2.f * sh32);
  U.a02 = 4.f * ch1 * ch3 * sh1 * sh3 - 2.f * ch2 * (-1.f + 2.f * sh12) * sh2 * (-1.f + 2.f * sh32);
  U.a10 = 2.f * ch1 * sh1 * (1.f - 2.f * sh22);
  U.a11 = -8.f * ch1 * ch2 * ch3 * sh1 * sh2 * sh3 + (-1.f + 2.f * sh12) * (-1.f + 2.f * sh32);
  U.a12 = -2.f * ch3 * sh3 + 4.f * sh1 * (ch3 * sh1 * sh3 + ch1 * ch2 * sh2 * (-1.f + 2.f * sh32));
  U.a20 = 2.f * ch2 * sh2;
  U.a21 = 2.f * ch3 * (1.f - 2.f * sh22) * sh3;
  U.a22 = (-1.f + 2.f * sh22) * (-1.f + 2.f * sh32);
  Vo.a00 = v11; Vo.a01 = v12; Vo.a02 = v13; Vo.a10 = v21; Vo.a11 = v22; Vo.a12 = v23; Vo.a20 = v31; Vo.a21 = v32; Vo.a22 = v33;
  sig.x = b11; sig.y = r22; sig.z = r33;
}

// (3) svd3_warm: (2) with a warm start.  `q` (x,y,z,w) is the quaternion of V from the previous substep of the same
//     particle (identity on a cold start); the Jacobi sweeps run on V0^T A^T A V0 and stop as soon as the off-diagonal
//     mass is at rounding level, so a particle whose F changed little converges in 0-1 sweeps instead of 5.  On return
//     q holds the converged quaternion.  Calling it again with that q and the same A reproduces U, sig, V bit for bit
//     with zero sweeps -- which is how the adjoint kernel re-creates the forward factors without re-iterating.
// SORT = false skips the ordering of the singular values (every consumer here is invariant under a consistent
// permutation of the triplets); then `q` is also the quaternion of the returned V.  qu receives the quaternion (x,y,z,w) of U,
// the product of the three Givens rotations of the QR step: U = Rz(t1) Ry(-t2) Rx(t3) with (ch_i, sh_i) = (cos, sin)(t_i / 2).
template <bool SORT = true>
DD_HD int svd3_warm(const M3 &A, float &qx, float &qy, float &qz, float &qw, M3 &U, V3 &sig, M3 &Vo, int max_sweeps = 6, float *qu = nullptr) {
  float v11, v12, v13, v21, v22, v23, v31, v32, v33;
  float b11, b12, b13, b21, b22, b23, b31, b32, b33;
  int sweeps = 0;
  {
    float qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz, qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
    v11 = 1.f - 2.f * (qyy + qzz); v12 = 2.f * (qxy - qwz); v13 = 2.f * (qxz + qwy);
    v21 = 2.f * (qxy + qwz); v22 = 1.f - 2.f * (qxx + qzz); v23 = 2.f * (qyz - qwx);
    v31 = 2.f * (qxz - qwy); v32 = 2.f * (qyz + qwx); v33 = 1.f - 2.f * (qxx + qyy);
  }
  b11 = A.a00 * v11 + A.a01 * v21 + A.a02 * v31; b12 = A.a00 * v12 + A.a01 * v22 + A.a02 * v32; b13 = A.a00 * v13 + A.a01 * v23 + A.a02 * v33;
  b21 = A.a10 * v11 + A.a11 * v21 + A.a12 * v31; b22 = A.a10 * v12 + A.a11 * v22 + A.a12 * v32; b23 = A.a10 * v13 + A.a11 * v23 + A.a12 * v33;
  b31 = A.a20 * v11 + A.a21 * v21 + A.a22 * v31; b32 = A.a20 * v12 + A.a21 * v22 + A.a22 * v32; b33 = A.a20 * v13 + A.a21 * v23 + A.a22 * v33;
  float s11 = b11 * b11 + b21 * b21 + b31 * b31, s21 = b12 * b11 + b22 * b21 + b32 * b31, s22 = b12 * b12 + b22 * b22 + b32 * b32;
  float s31 = b13 * b11 + b23 * b21 + b33 * b31, s32 = b13 * b12 + b23 * b22 + b33 * b32, s33 = b13 * b13 + b23 * b23 + b33 * b33;
  const float tol2 = 1.0e-13f;  // (3.2e-7)^2: off-diagonal mass relative to the trace, at fp32 rounding level
  bool rotated = false;
#pragma unroll 1
  for (; sweeps < max_sweeps; ++sweeps) {
    float off = s21 * s21 + s31 * s31 + s32 * s32, tr = s11 + s22 + s33;
    if (off <= tol2 * tr * tr) break;
    jacobi32(s11, s21, s22, s31, s32, s33, qx, qy, qz, qw);
    jacobi32(s11, s21, s22, s31, s32, s33, qy, qz, qx, qw);
    jacobi32(s11, s21, s22, s31, s32, s33, qz, qx, qy, qw);
    rotated = true;
  }
  if (rotated) {
    float n = DD_RSQRT(qx * qx + qy * qy + qz * qz + qw * qw);
    qx *= n; qy *= n; qz *= n; qw *= n;
    float qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz, qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
    v11 = 1.f - 2.f * (qyy + qzz); v12 = 2.f * (qxy - qwz); v13 = 2.f * (qxz + qwy);
    v21 = 2.f * (qxy + qwz); v22 = 1.f - 2.f * (qxx + qzz); v23 = 2.f * (qyz - qwx);
    v31 = 2.f * (qxz - qwy); v32 = 2.f * (qyz + qwx); v33 = 1.f - 2.f * (qxx + qyy);
    b11 = A.a00 * v11 + A.a01 * v21 + A.a02 * v31; b12 = A.a00 * v12 + A.a01 * v22 + A.a02 * v32; b13 = A.a00 * v13 + A.a01 * v23 + A.a02 * v33;
    b21 = A.a10 * v11 + A.a11 * v21 + A.a12 * v31; b22 = A.a10 * v12 + A.a11 * v22 + A.a12 * v32; b23 = A.a10 * v13 + A.a11 * v23 + A.a12 * v33;
    b31 = A.a20 * v11 + A.a21 * v21 + A.a22 * v31; b32 = A.a20 * v12 + A.a21 * v22 + A.a22 * v32; b33 = A.a20 * v13 + A.a21 * v23 + A.a22 * v33;
  }
  if (SORT) {
    float r1 = b11 * b11 + b21 * b21 + b31 * b31, r2 = b12 * b12 + b22 * b22 + b32 * b32, r3 = b13 * b13 + b23 * b23 + b33 * b33;
    bool c = r1 < r2;
    cnswap32(c, b11, b12); cnswap32(c, v11, v12); cnswap32(c, b21, b22); cnswap32(c, v21, v22); cnswap32(c, b31, b32); cnswap32(c, v31, v32);
    cswap32(c, r1, r2);
    c = r1 < r3;
    cnswap32(c, b11, b13); cnswap32(c, v11, v13); cnswap32(c, b21, b23); cnswap32(c, v21, v23); cnswap32(c, b31, b33); cnswap32(c, v31, v33);
    cswap32(c, r1, r3);
    c = r2 < r3;
    cnswap32(c, b12, b13); cnswap32(c, v12, v13); cnswap32(c, b22, b23); cnswap32(c, v22, v23); cnswap32(c, b32, b33); cnswap32(c, v32, v33);
  }
  float ch1, sh1, ch2, sh2, ch3, sh3, ca, cb;
  qr_givens32(b11, b21, ch1, sh1);
  ca = 1.f - 2.f * sh1 * sh1; cb = 2.f * ch1 * sh1;
  float r11 = ca * b11 + cb * b21, r12 = ca * b12 + cb * b22, r13 = ca * b13 + cb * b23;
  float r22 = -cb * b12 + ca * b22, r23 = -cb * b13 + ca * b23;
  float r31 = b31, r32 = b32, r33 = b33;
  qr_givens32(r11, r31, ch2, sh2);
  ca = 1.f - 2.f * sh2 * sh2; cb = 2.f * ch2 * sh2;
  b11 = ca * r11 + cb * r31;
  b22 = r22; b23 = r23;
  b32 = -cb * r12 + ca * r32; b33 = -cb * r13 + ca * r33;
  qr_givens32(b22, b32, ch3, sh3);
  ca = 1.f - 2.f * sh3 * sh3; cb = 2.f * ch3 * sh3;
  r22 = ca * b22 + cb * b32;
  r33 = -cb * b23 + ca * b33;
  float sh12 = sh1 * sh1, sh22 = sh2 * sh2, sh32 = sh3 * sh3;
  U.a00 = (-1.f + 2.f * sh12) * (-1.f + 2.f * sh22);
  U.a01 = 4.f * ch2 * ch3 * (-1.f + 2.f * sh12) * sh2 * sh3 + 2.f * ch1 * sh1 * (-1.f + 2.f * sh32);
  U.a02 = 4.f * ch1 * ch3 * sh1 * sh3 - 2.f * ch2 * (-1.f + 2.f * sh12) * sh2 * (-1.f + 2.f * sh32);
  U.a10 = 2.f * ch1 * sh1 * (1.f - 2.f * sh22);
  U.a11 = -8.f * ch1 * ch2 * ch3 * sh1 * sh2 * sh3 + (-1.f + 2.f * sh12) * (-1.f + 2.f * sh32);
  U.a12 = -2.f * ch3 * sh3 + 4.f * sh1 * (ch3 * sh1 * sh3 + ch1 * ch2 * sh2 * (-1.f + 2.f * sh32));
  U.a20 = 2.f * ch2 * sh2;
  U.a21 = 2.f * ch3 * (1.f - 2.f * sh22) * sh3;
  U.a22 = (-1.f + 2.f * sh22) * (-1.f + 2.f * sh32);
  Vo.a00 = v11; Vo.a01 = v12; Vo.a02 = v13; Vo.a10 = v21; Vo.a11 = v22; Vo.a12 = v23; Vo.a20 = v31; Vo.a21 = v32; Vo.a22 = v33;
  sig.x = b11; sig.y = r22; sig.z = r33;
  if (qu) {
    float a = ch1 * ch2, bb = sh1 * sh2, c_ = -ch1 * sh2, d = sh1 * ch2;  // Rz(t1) Ry(-t2) as (w, x, y, z) = (a, bb, c_, d)
    qu[3] = a * ch3 - bb * sh3; qu[0] = a * sh3 + bb * ch3; qu[1] = c_ * ch3 + d * sh3; qu[2] = d * ch3 - c_ * sh3;
  }
  return sweeps;
}

// rotation matrix of a unit quaternion (x, y, z, w)
DD_HD M3 quat_to_m3(float qx, float qy, float qz, float qw) {
  float qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz, qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
  M3 m;
  m.a00 = 1.f - 2.f * (qyy + qzz); m.a01 = 2.f * (qxy - qwz); m.a02 = 2.f * (qxz + qwy);
  m.a10 = 2.f * (qxy + qwz); m.a11 = 1.f - 2.f * (qxx + qzz); m.a12 = 2.f * (qyz - qwx);
  m.a20 = 2.f * (qxz - qwy); m.a21 = 2.f * (qyz + qwx); m.a22 = 1.f - 2.f * (qxx + qyy);
  return m;
}

// ---------------------------------------------------------------------------------------------- constitutive model
// von-Mises return mapping in log-strain space (integrator.cu:42-67).  Returns J and writes F_new; `plastic`
// and `ee` (= exp of the projected log strains) are kept for the adjoint.
struct Plastic { bool plastic; V3 eps, eh, ee; float ehn, dg; };
// FAST (fused engine, fp32 SVD): MUFU-based log / exp / reciprocals.  Absolute error of __logf on [0.05, 20] is < 4e-7, i.e. a
// relative error of that size in F_new -- below the error of the fp32 SVD itself; the per-stage ABI keeps the exact forms.
template <bool FAST = false>
DD_DEV float von_mises(const M3 &Ft, const M3 &U, V3 s, const M3 &Vm, float yield, float mu, M3 &outF, Plastic &pl) {
  V3 sn = vmax(s, 0.05f);
  pl.eps = FAST ? v3(__logf(sn.x), __logf(sn.y), __logf(sn.z)) : v3(logf(sn.x), logf(sn.y), logf(sn.z));
  float mean = FAST ? (pl.eps.x + pl.eps.y + pl.eps.z) * (1.f / 3.f) : (pl.eps.x + pl.eps.y + pl.eps.z) / 3.f;
  pl.eh = v3(pl.eps.x - mean, pl.eps.y - mean, pl.eps.z - mean);
  pl.ehn = sqrtf(dot(pl.eh, pl.eh) + 1e-8f);  // norm(), integrator.cu:28-31
  pl.dg = pl.ehn - (FAST ? __fdividef(yield, 2 * mu) : yield / (2 * mu));
  pl.plastic = pl.dg > 0.f;
  if (pl.plastic) {
    V3 e = pl.eps - (FAST ? __fdividef(pl.dg, pl.ehn) : pl.dg / pl.ehn) * pl.eh;
    pl.ee = FAST ? v3(__expf(e.x), __expf(e.y), __expf(e.z)) : v3(expf(e.x), expf(e.y), expf(e.z));
    outF = mul_nt(mul_diag(U, pl.ee), Vm);
    return pl.ee.x * pl.ee.y * pl.ee.z;
  }
  outF = Ft;
  return s.x * s.y * s.z;
}
// Kirchhoff-like stress of the fixed-corotated model as the reference forms it (integrator.cu:364-369)
DD_DEV M3 fixed_corotated(const M3 &nF, const M3 &r, float J, float mu, float lam) {
  return (2.f * mu) * mul_nt(nF - r, nF) + mdiag(lam * J * (J - 1));
}

// clamp of the SVD adjoint (integrator.cu:102-108)
DD_DEV float clamp_eps(float a) { return a >= 0.f ? fmaxf(a, 1e-6f) : fminf(a, -1e-6f); }

// Adjoint through (U,sigma,V) = svd(Ft): returns dL/dFt given dL/dU, dL/dsigma, dL/dV (integrator.cu:131-159,
// without the newF_grad term).
DD_DEV M3 svd_adj(const M3 &u, V3 sigma, const M3 &v, const M3 &gu, V3 gs, const M3 &gv) {
#ifdef DD_FLOAT_LENGTH
  // s_j^2 - s_i^2 = (s_j - s_i)(s_j + s_i): the difference of nearby floats is exact, so this matches the reference's
  // double-precision squares rounded to float to within one ulp, without fp64 multiplies and divides
  float d10 = (sigma.y - sigma.x) * (sigma.y + sigma.x), d20 = (sigma.z - sigma.x) * (sigma.z + sigma.x), d21 = (sigma.z - sigma.y) * (sigma.z + sigma.y);
  M3 K = m3(0.f, 1.f / clamp_eps(d10), 1.f / clamp_eps(d20), 1.f / clamp_eps(-d10), 0.f, 1.f / clamp_eps(d21), 1.f / clamp_eps(-d20), 1.f / clamp_eps(-d21), 0.f);
#else
  double s0 = sigma.x, s1 = sigma.y, s2 = sigma.z;
  s0 = s0 * s0; s1 = s1 * s1; s2 = s2 * s2;
  M3 K = m3(0.f, (float)(1.0 / clamp_eps((float)(s1 - s0))), (float)(1.0 / clamp_eps((float)(s2 - s0))),
            (float)(1.0 / clamp_eps((float)(s0 - s1))), 0.f, (float)(1.0 / clamp_eps((float)(s2 - s1))),
            (float)(1.0 / clamp_eps((float)(s0 - s2))), (float)(1.0 / clamp_eps((float)(s1 - s2))), 0.f);
#endif
  M3 ut_gu = mul_tn(u, gu);
  M3 vt_gv = mul_tn(v, gv);
  M3 u_term = mul_nt(mul(u, mul_diag(hadamard(K, ut_gu - transpose(ut_gu)), sigma)), v);
  M3 v_term = mul(u, diag_mul(sigma, mul_nt(hadamard(K, vt_gv - transpose(vt_gv)), v)));
  M3 sigma_term = mul_nt(mul_diag(u, gs), v);
  return u_term + sigma_term + v_term;
}

}  // namespace dd
