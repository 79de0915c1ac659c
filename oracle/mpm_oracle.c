/*
 * mpm_oracle.c -- CPU restatement of DexDeform's differentiable MLS-MPM substep and its adjoint.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / reference legs of bench.py may load it.  The product path (dexdeform_b200/) never
 * links or calls anything in oracle/.
 *
 * Every function restates one reference function; the reference location is cited as
 * (file:line) relative to /root/reference/.  Buffers use the reference's own layouts
 * (AoS vec3 = 3 floats, mat3 = 9 floats row-major, quat = (w,x,y,z); mpm/types.py:25-75) so the
 * same host arrays can be fed to the reference library, to this oracle and to the CUDA product.
 *
 * Parity pin: the reference has no golden vectors for this path; the oracle is pinned
 * against outputs of the reference's own CUDA library (oracle/_ref/libmaniskill_mpm.so, built by
 * oracle/build_ref.sh from the unmodified sources) generated on a B200 by tests/golden/make_golden.py
 * and committed under tests/golden/.
 *
 * Arithmetic follows the reference expression by expression (float where it is float, double where
 * C++ promotion makes it double).  Results are not bit-identical to the GPU because nvcc contracts
 * a*b+c into FMA and uses its own expf/logf; tests state the tolerance.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off mpm_oracle.c -o libmpm_oracle.so -lm
 */
#include <math.h>
#include <stdio.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#define ATOMIC _Pragma("omp atomic")
#define PARFOR _Pragma("omp parallel for schedule(static)")
#else
#define ATOMIC
#define PARFOR
#endif

typedef struct { float x, y, z; } v3;
typedef struct { float m[3][3]; } m3;
typedef struct { float w, x, y, z; } q4;

#define NORM_EPS 1e-8       /* integrator.cu:11 */
#define SIG_CLIP_EPS 0.05   /* integrator.cu:12 */
#define NORM_EPS2 1e-30     /* shape.h:4 */

/* ------------------------------------------------------------------ vec3 (vec3.h:8-170) */
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vdivs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 vcross(v3 a, v3 b) {
  return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float vget(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static inline v3 vabs(v3 a) { return V(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
static inline v3 vmaxs(v3 a, float b) { return V(fmaxf(a.x, b), fmaxf(a.y, b), fmaxf(a.z, b)); }

/* integrator.cu:28-31 */
static inline float norm_eps(v3 a) { return sqrtf(vdot(a, a) + (float)NORM_EPS); } /* eps is a float parameter */
/* shape.h:7-14 */
static inline float length30(v3 a) { return (float)sqrt((double)vdot(a, a) + NORM_EPS2); }
static inline v3 normalized(v3 a) { return vdivs(a, length30(a)); }
/* shape.h:16-25 */
static inline v3 normalized_backward(v3 vec, v3 g) {
  float doted = (float)((double)vdot(vec, vec) + NORM_EPS2);
  float s = vdot(vdivs(vec, doted), g);
  v3 t = vsub(g, vscale(vec, s));
  float k = (float)(1. / (double)sqrtf(doted)); /* 1./sqrt(float) is double, narrowed by operator*(vec3,float) */
  return vscale(t, k);
}

/* ------------------------------------------------------------------ mat3 (mat3.h:8-186) */
static inline m3 mzero(void) { m3 r; memset(&r, 0, sizeof r); return r; }
static inline m3 mdiag(v3 d) { m3 r = mzero(); r.m[0][0] = d.x; r.m[1][1] = d.y; r.m[2][2] = d.z; return r; }
static inline m3 mident(float d) { return mdiag(V(d, d, d)); }
static inline m3 mT(m3 a) { m3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i]; return r; }
/* mat3.h:86-96: accumulates k = 0,1,2 in order starting from 0 */
static inline m3 mmul(m3 a, m3 b) {
  m3 r = mzero();
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) r.m[i][j] += a.m[i][k] * b.m[k][j];
  return r;
}
/* mat3.h:98-101: col0*b.x + col1*b.y + col2*b.z */
static inline v3 mvec(m3 a, v3 b) {
  return V(a.m[0][0] * b.x + a.m[0][1] * b.y + a.m[0][2] * b.z,
           a.m[1][0] * b.x + a.m[1][1] * b.y + a.m[1][2] * b.z,
           a.m[2][0] * b.x + a.m[2][1] * b.y + a.m[2][2] * b.z);
}
static inline m3 madd(m3 a, m3 b) { m3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
static inline m3 msub(m3 a, m3 b) { m3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] - b.m[i][j]; return r; }
static inline m3 mhad(m3 a, m3 b) { m3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] * b.m[i][j]; return r; }
static inline m3 mscale(m3 a, float s) { m3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] * s; return r; }
/* mat3.h:172-174: outer(a,b)[i][j] = a_i b_j */
static inline m3 mouter(v3 a, v3 b) {
  m3 r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = vget(a, i) * vget(b, j);
  return r;
}
static inline float msum(m3 a) { float s = 0.f; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s += a.m[i][j]; return s; }

/* ------------------------------------------------------------------ quat (quat.h:5-97) */
static inline q4 qinv(q4 q) { q4 r = {q.w, -q.x, -q.y, -q.z}; return r; }
/* quat.h:14-19 */
static inline v3 qrot(q4 q, v3 v) {
  v3 u = V(q.x, q.y, q.z);
  v3 uv = vcross(u, v);
  v3 uuv = vcross(u, uv);
  return vadd(v, vscale(vadd(vscale(uv, q.w), uuv), 2.f));
}
/* quat.h:89-97 */
static inline v3 spatial_transform(v3 p, q4 q, v3 pt) { return vadd(p, qrot(q, pt)); }
static inline v3 spatial_transform_inv(v3 p, q4 q, v3 pt) { return qrot(qinv(q), vsub(pt, p)); }
/* quat.h:25-47 */
static inline void qmul_backward(q4 q, v3 v, v3 g, q4 *gq, v3 *gv) {
  v3 u = V(q.x, q.y, q.z);
  v3 uv = vcross(u, v);
  gq->w += vdot(uv, g) * 2.f;
  v3 g_uv = vscale(g, 2.f * q.w);
  v3 g_uuv = vscale(g, 2.f);
  v3 g_u = vcross(uv, g_uuv);
  g_uv = vadd(g_uv, vcross(g_uuv, u));
  g_u = vadd(g_u, vcross(v, g_uv));
  *gv = vadd(*gv, vadd(g, vcross(g_uv, u)));
  gq->x += g_u.x; gq->y += g_u.y; gq->z += g_u.z;
}
/* quat.h:50-62 */
static inline void spatial_transform_backward(v3 p, q4 q, v3 pt, v3 g, v3 *gp, q4 *gq, v3 *gpt) {
  (void)p;
  *gp = vadd(*gp, g);
  qmul_backward(q, pt, g, gq, gpt);
}
/* quat.h:64-80 */
static inline void inv_spatial_transform_backward(v3 p, q4 q, v3 pt, v3 g, v3 *gp, q4 *gq, v3 *gpt) {
  q4 tq = {0.f, 0.f, 0.f, 0.f};
  v3 tp = V(0.f, 0.f, 0.f);
  qmul_backward(qinv(q), vsub(pt, p), g, &tq, &tp);
  q4 c = qinv(tq);
  gq->w += c.w; gq->x += c.x; gq->y += c.y; gq->z += c.z;
  *gp = vsub(*gp, tp);
  *gpt = vadd(*gpt, tp);
}

/* ------------------------------------------------------------------ shapes (shape.h:27-152) */
static inline int get_type(q4 tfsr) { return (int)floorf(tfsr.w + 0.1f); }
static inline void abs_backward(v3 gx, v3 *g) {
  if (gx.x < 0) g->x *= -1;
  if (gx.y < 0) g->y *= -1;
  if (gx.z < 0) g->z *= -1;
}
/* shape.h:37-61 */
static inline float shape_sdf(q4 tfsr, q4 args, v3 gx) {
  int type = get_type(tfsr);
  float sdf = 0.f;
  if (type == 0) {
    v3 q = vsub(vabs(gx), V(args.w, args.x, args.y));
    sdf = length30(vmaxs(q, 0.f)) + fminf(fmaxf(fmaxf(q.x, q.y), q.z), 0.f);
  } else if (type == 1) {
    v3 p2 = gx;
    float r = args.w, h = args.x;
    p2.y += h / 2;
    p2.y -= fminf(fmaxf(p2.y, 0.f), h);
    sdf = length30(p2) - r;
  }
  return sdf - tfsr.z;
}
/* shape.h:65-102 */
static inline v3 shape_grad(q4 tfsr, q4 args, v3 gx) {
  int type = get_type(tfsr);
  v3 grad = V(0.f, 0.f, 0.f);
  if (type == 0) {
    v3 q = vsub(vabs(gx), V(args.w, args.x, args.y));
    float inside = fmaxf(fmaxf(q.x, q.y), q.z);
    if (inside <= 0) {
      if (q.x == inside) grad.x += 1;
      if (q.y == inside) grad.y += 1;
      if (q.z == inside) grad.z += 1;
    } else {
      grad = normalized(vmaxs(q, 0.f));
    }
    abs_backward(gx, &grad);
  } else if (type == 1) {
    v3 p2 = gx;
    float h = args.x;
    p2.y += h / 2;
    p2.y -= fminf(fmaxf(p2.y, 0.f), h);
    return normalized(p2);
  }
  return grad;
}
/* shape.h:104-152.  NOTE the capsule branch back-propagates a zero vector: grad_in is initialised to 0
 * and never set from grad_out before normalized_backward (shape.h:144) -- reproduced bug-for-bug. */
static inline v3 shape_grad_backward(q4 tfsr, q4 args, v3 gx, v3 gout) {
  int type = get_type(tfsr);
  v3 gin = V(0.f, 0.f, 0.f);
  if (type == 0) {
    v3 q = vsub(vabs(gx), V(args.w, args.x, args.y));
    float inside = fmaxf(fmaxf(q.x, q.y), q.z);
    gin = gout;
    abs_backward(gx, &gin);
    if (inside <= 0) {
      gin = V(0.f, 0.f, 0.f);
    } else {
      v3 q2 = vmaxs(q, 0.f);
      gin = normalized_backward(q2, gin);
      if (q.x < 0) gin.x = 0;
      if (q.y < 0) gin.y = 0;
      if (q.z < 0) gin.z = 0;
      abs_backward(gx, &gin);
    }
  } else if (type == 1) {
    v3 p2 = gx;
    float h = args.x;
    p2.y += h / 2;
    float zero_y_grad = (p2.y >= 0.f && p2.y <= h);
    p2.y -= fminf(fmaxf(p2.y, 0.f), h);
    gin = normalized_backward(p2, gin);
    gin.y *= zero_y_grad;
    return gin;
  }
  return gin;
}

/* ------------------------------------------------------------------ 3x3 SVD in double (svd.h:16-413)
 * Same algorithm as the reference (ericjang/svd3, after McAdams et al.): Jacobi eigen-analysis of A^T A with
 * approximate Givens rotations accumulated in a quaternion (8 sweeps x 3 conjugations), column sort with
 * sign-preserving swaps, then Givens QR of A V.  Kept algorithmically identical so that U and V (which are
 * not unique for repeated singular values) agree with the reference's. */
#define SVD_GAMMA ((double)5.82842712474619f)  /* svd.h:18 (float literal promoted) */
#define SVD_CSTAR ((double)0.9238795325112867f)
#define SVD_SSTAR ((double)0.3826834323650897f)
#define SVD_EPS 1e-6

/* svd.h:109-124 */
static void approx_givens(double a11, double a12, double a22, double *ch, double *sh) {
  double c = 2 * (a11 - a22), s = a12;
  int b = SVD_GAMMA * s * s < c * c;
  double w = 1.0 / sqrt(c * c + s * s);
  *ch = b ? w * c : SVD_CSTAR;
  *sh = b ? w * s : SVD_SSTAR;
}
/* svd.h:126-188.  S holds (s11,s21,s22,s31,s32,s33); q = (x,y,z,w). */
static void jacobi_conjugation(int x, int y, int z, double *S, double *q) {
  double ch, sh;
  approx_givens(S[0], S[1], S[2], &ch, &sh);
  double scale = ch * ch + sh * sh;
  double a = (ch * ch - sh * sh) / scale;
  double b = (2 * sh * ch) / scale;
  double s11 = S[0], s21 = S[1], s22 = S[2], s31 = S[3], s32 = S[4], s33 = S[5];
  double n11 = a * (a * s11 + b * s21) + b * (a * s21 + b * s22);
  double n21 = a * (-b * s11 + a * s21) + b * (-b * s21 + a * s22);
  double n22 = -b * (-b * s11 + a * s21) + a * (-b * s21 + a * s22);
  double n31 = a * s31 + b * s32;
  double n32 = -b * s31 + a * s32;
  double n33 = s33;
  double tmp[3] = {q[0] * sh, q[1] * sh, q[2] * sh};
  sh *= q[3];
  q[0] *= ch; q[1] *= ch; q[2] *= ch; q[3] *= ch;
  q[z] += sh;
  q[3] -= tmp[z];
  q[x] += tmp[y];
  q[y] -= tmp[x];
  /* cyclic re-labelling for the next (p,q) pair, svd.h:174-187 */
  S[0] = n22; S[1] = n32; S[2] = n33; S[3] = n21; S[4] = n31; S[5] = n11;
}
static inline void cond_swap(int c, double *X, double *Y) { double Z = *X; *X = c ? *Y : *X; *Y = c ? Z : *Y; }
static inline void cond_neg_swap(int c, double *X, double *Y) { double Z = -*X; *X = c ? *Y : *X; *Y = c ? Z : *Y; }
/* svd.h:293-310 */
static void qr_givens(double a1, double a2, double *ch, double *sh) {
  double rho0 = a1 * a1 + a2 * a2;
  double rho = rho0 / sqrt(rho0); /* accurateSqrt, svd.h:23 */
  *sh = rho > SVD_EPS ? a2 : 0;
  *ch = fabs(a1) + fmax(rho, SVD_EPS);
  int b = a1 < 0;
  cond_swap(b, sh, ch);
  double w = 1.0 / sqrt((*ch) * (*ch) + (*sh) * (*sh));
  *ch *= w; *sh *= w;
}
/* svd.h:346-413 */
static void svd3(m3 A, m3 *Uo, v3 *sig, m3 *Vo) {
  double a[3][3], v[3][3], B[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = A.m[i][j];
  /* A^T A, lower triangle (svd.h:61-80, 363-365) */
  double ata[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) ata[i][j] = a[0][i] * a[0][j] + a[1][i] * a[1][j] + a[2][i] * a[2][j];
  double S[6] = {ata[0][0], ata[1][0], ata[1][1], ata[2][0], ata[2][1], ata[2][2]};
  double q[4] = {0, 0, 0, 1};
  for (int it = 0; it < 8; ++it) { /* svd.h:193-210 */
    jacobi_conjugation(0, 1, 2, S, q);
    jacobi_conjugation(1, 2, 0, S, q);
    jacobi_conjugation(2, 0, 1, S, q);
  }
  { /* quatToMat3, svd.h:82-107 (no normalisation: V is scaled by |q|^2 exactly as in the reference) */
    double w = q[3], x = q[0], y = q[1], z = q[2];
    double qxx = x * x, qyy = y * y, qzz = z * z, qxz = x * z, qxy = x * y, qyz = y * z, qwx = w * x, qwy = w * y, qwz = w * z;
    v[0][0] = 1 - 2 * (qyy + qzz); v[0][1] = 2 * (qxy - qwz); v[0][2] = 2 * (qxz + qwy);
    v[1][0] = 2 * (qxy + qwz); v[1][1] = 1 - 2 * (qxx + qzz); v[1][2] = 2 * (qyz - qwx);
    v[2][0] = 2 * (qxz - qwy); v[2][1] = 2 * (qyz + qwx); v[2][2] = 1 - 2 * (qxx + qyy);
  }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) B[i][j] = a[i][0] * v[0][j] + a[i][1] * v[1][j] + a[i][2] * v[2][j];
  { /* sortSingularValues, svd.h:212-246 */
    double rho1 = B[0][0] * B[0][0] + B[1][0] * B[1][0] + B[2][0] * B[2][0];
    double rho2 = B[0][1] * B[0][1] + B[1][1] * B[1][1] + B[2][1] * B[2][1];
    double rho3 = B[0][2] * B[0][2] + B[1][2] * B[1][2] + B[2][2] * B[2][2];
    int c = rho1 < rho2;
    for (int i = 0; i < 3; ++i) { cond_neg_swap(c, &B[i][0], &B[i][1]); cond_neg_swap(c, &v[i][0], &v[i][1]); }
    cond_swap(c, &rho1, &rho2);
    c = rho1 < rho3;
    for (int i = 0; i < 3; ++i) { cond_neg_swap(c, &B[i][0], &B[i][2]); cond_neg_swap(c, &v[i][0], &v[i][2]); }
    cond_swap(c, &rho1, &rho3);
    c = rho2 < rho3;
    for (int i = 0; i < 3; ++i) { cond_neg_swap(c, &B[i][1], &B[i][2]); cond_neg_swap(c, &v[i][1], &v[i][2]); }
  }
  /* QRDecomposition, svd.h:248-344 */
  double ch1, sh1, ch2, sh2, ch3, sh3, ca, cb, r[3][3], b2[3][3];
  qr_givens(B[0][0], B[1][0], &ch1, &sh1);
  ca = 1 - 2 * sh1 * sh1; cb = 2 * ch1 * sh1;
  for (int j = 0; j < 3; ++j) { r[0][j] = ca * B[0][j] + cb * B[1][j]; r[1][j] = -cb * B[0][j] + ca * B[1][j]; r[2][j] = B[2][j]; }
  qr_givens(r[0][0], r[2][0], &ch2, &sh2);
  ca = 1 - 2 * sh2 * sh2; cb = 2 * ch2 * sh2;
  for (int j = 0; j < 3; ++j) { b2[0][j] = ca * r[0][j] + cb * r[2][j]; b2[1][j] = r[1][j]; b2[2][j] = -cb * r[0][j] + ca * r[2][j]; }
  qr_givens(b2[1][1], b2[2][1], &ch3, &sh3);
  ca = 1 - 2 * sh3 * sh3; cb = 2 * ch3 * sh3;
  for (int j = 0; j < 3; ++j) { r[0][j] = b2[0][j]; r[1][j] = ca * b2[1][j] + cb * b2[2][j]; r[2][j] = -cb * b2[1][j] + ca * b2[2][j]; }
  double sh12 = sh1 * sh1, sh22 = sh2 * sh2, sh32 = sh3 * sh3, u[3][3];
  u[0][0] = (-1 + 2 * sh12) * (-1 + 2 * sh22);
  u[0][1] = 4 * ch2 * ch3 * (-1 + 2 * sh12) * sh2 * sh3 + 2 * ch1 * sh1 * (-1 + 2 * sh32);
  u[0][2] = 4 * ch1 * ch3 * sh1 * sh3 - 2 * ch2 * (-1 + 2 * sh12) * sh2 * (-1 + 2 * sh32);
  u[1][0] = 2 * ch1 * sh1 * (1 - 2 * sh22);
  u[1][1] = -8 * ch1 * ch2 * ch3 * sh1 * sh2 * sh3 + (-1 + 2 * sh12) * (-1 + 2 * sh32);
  u[1][2] = -2 * ch3 * sh3 + 4 * sh1 * (ch3 * sh1 * sh3 + ch1 * ch2 * sh2 * (-1 + 2 * sh32));
  u[2][0] = 2 * ch2 * sh2;
  u[2][1] = 2 * ch3 * (1 - 2 * sh22) * sh3;
  u[2][2] = (-1 + 2 * sh22) * (-1 + 2 * sh32);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Uo->m[i][j] = (float)u[i][j]; Vo->m[i][j] = (float)v[i][j]; }
  sig->x = (float)r[0][0]; sig->y = (float)r[1][1]; sig->z = (float)r[2][2];
}

void orc_svd3(const float *A, float *U, float *sig, float *Vm, int n) {
  for (int p = 0; p < n; ++p) svd3(((const m3 *)A)[p], (m3 *)U + p, (v3 *)sig + p, (m3 *)Vm + p);
}

/* ------------------------------------------------------------------ helpers shared by the particle kernels */
typedef struct { int x, y, z; } i3;
static inline int grid_index(int x, int y, int z, i3 d) { return (x * d.y + y) * d.z + z; } /* vec3.h:201-209 */

typedef struct { i3 base; v3 fx; float w[3][3]; float w1[3][3]; } stencil;
/* integrator.cu:348-352 (weights), :493 (their derivatives); w[axis][i] */
static inline stencil make_stencil(v3 x, float inv_dx) {
  stencil s;
  v3 t = V(x.x * inv_dx - 0.5f, x.y * inv_dx - 0.5f, x.z * inv_dx - 0.5f);
  s.base.x = (int)floorf(t.x); s.base.y = (int)floorf(t.y); s.base.z = (int)floorf(t.z);
  s.fx = V(x.x * inv_dx - (float)s.base.x, x.y * inv_dx - (float)s.base.y, x.z * inv_dx - (float)s.base.z);
  for (int a = 0; a < 3; ++a) {
    float f = vget(s.fx, a);
    s.w[a][0] = 0.5f * ((1.5f - f) * (1.5f - f));
    s.w[a][1] = 0.75f - (f - 1.f) * (f - 1.f);
    s.w[a][2] = 0.5f * ((f - 0.5f) * (f - 0.5f));
    s.w1[a][0] = -inv_dx * (1.5f - f);
    s.w1[a][1] = inv_dx * ((-2.f) * f + 2.0f);
    s.w1[a][2] = -inv_dx * (f * (-1.f) + 0.5f);
  }
  return s;
}
/* integrator.cu:33-40 */
static inline v3 dw(const stencil *s, int i, int j, int k) {
  return V(s->w1[0][i] * s->w[1][j] * s->w[2][k], s->w[0][i] * s->w1[1][j] * s->w[2][k], s->w[0][i] * s->w[1][j] * s->w1[2][k]);
}

/* integrator.cu:42-67 */
static float von_mises(m3 F, m3 U, v3 s, m3 Vm, float yield, float mu, m3 *outF) {
  v3 sn = vmaxs(s, (float)SIG_CLIP_EPS);
  v3 eps = V(logf(sn.x), logf(sn.y), logf(sn.z));
  float mean = (eps.x + eps.y + eps.z) / 3.f;
  v3 eh = V(eps.x - mean, eps.y - mean, eps.z - mean);
  float ehn = norm_eps(eh);
  float dg = ehn - yield / (2 * mu);
  if (dg > 0.f) {
    eps = vsub(eps, vscale(eh, dg / ehn));
    v3 e = V(expf(eps.x), expf(eps.y), expf(eps.z));
    *outF = mmul(mmul(U, mdiag(e)), mT(Vm));
    return e.x * e.y * e.z;
  }
  *outF = F;
  return s.x * s.y * s.z;
}

/* ------------------------------------------------------------------ forward kernels */
/* integrator.cu:84-100 */
void orc_compute_svd(const float *F, const float *C, float *newF, float *U, float *Vm, float *sig, float dt, int n) {
  PARFOR
  for (int p = 0; p < n; ++p) {
    m3 f = mmul(madd(mident(1.f), mscale(((const m3 *)C)[p], dt)), ((const m3 *)F)[p]);
    m3 u, v; v3 s;
    svd3(f, &u, &s, &v);
    ((m3 *)newF)[p] = f; ((m3 *)U)[p] = u; ((m3 *)Vm)[p] = v; ((v3 *)sig)[p] = s;
  }
}

/* integrator.cu:313-394 */
void orc_p2g(const float *px, const float *pv, const float *pm, const float *pvol, const float *pF, const float *pU,
             const float *psig, const float *pV, const float *pC, const float *mly, const int *grid_lower, const int *gdim,
             float dx, float inv_dx, float dt, float *outF, float *grid_mv, float *grid_m, int n) {
  i3 gd = {gdim[0], gdim[1], gdim[2]};
  PARFOR
  for (int p = 0; p < n; ++p) {
    v3 x = ((const v3 *)px)[p];
    x = vsub(x, vscale(V((float)grid_lower[0], (float)grid_lower[1], (float)grid_lower[2]), dx));
    m3 U = ((const m3 *)pU)[p], Vm = ((const m3 *)pV)[p], Ft = ((const m3 *)pF)[p];
    v3 sigma = ((const v3 *)psig)[p];
    stencil st = make_stencil(x, inv_dx);
    float mu = mly[3 * p], lam = mly[3 * p + 1], yield = mly[3 * p + 2];
    m3 nF;
    float J = von_mises(Ft, U, sigma, Vm, yield, mu, &nF);
    ((m3 *)outF)[p] = nF;
    m3 r = mmul(U, mT(Vm));
    m3 stress = madd(mscale(mmul(msub(nF, r), mT(nF)), 2.f * mu), mident(lam * J * (J - 1)));
    stress = mscale(stress, -dt * pvol[p] * 4.f * inv_dx * inv_dx);
    m3 affine = madd(stress, mscale(((const m3 *)pC)[p], pm[p]));
    v3 mv = vscale(((const v3 *)pv)[p], pm[p]);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) {
      float weight = st.w[0][i] * st.w[1][j] * st.w[2][k];
      v3 dpos = vscale(vsub(V((float)i, (float)j, (float)k), st.fx), dx);
      int idx = grid_index(st.base.x + i, st.base.y + j, st.base.z + k, gd);
      v3 c = vscale(vadd(mv, mvec(affine, dpos)), weight);
      float mw = pm[p] * weight;
      ATOMIC grid_m[idx] += mw;
      ATOMIC grid_mv[3 * idx] += c.x;
      ATOMIC grid_mv[3 * idx + 1] += c.y;
      ATOMIC grid_mv[3 * idx + 2] += c.z;
    }
  }
}

/* the contact test of integrator.cu:707-710 / :913-916; influence compared in double because of the 0.1 / 1. literals */
static inline int contact_active(float dist, float softness, float *influence) {
  float e = expf(-dist * softness);
  *influence = (float)fmin((double)e, 1.);
  return (softness > 0 && (double)(*influence) > 0.1) || dist <= 0;
}

/* integrator.cu:647-777 */
void orc_grid_op_v2(const float *grid_m, const float *grid_v_in, float *grid_body_v_in, const int *grid_lower, const float *gravity,
                    const float *body_pos, const float *body_rot, const float *next_pos, const float *next_rot, const float *tfsr_,
                    const float *args_, float dx, float inv_dx, float dt, float ground_friction, float *out_v, const int *gdim, int nb) {
  (void)inv_dx;
  i3 gd = {gdim[0], gdim[1], gdim[2]};
  int dim = gd.x * gd.y * gd.z;
  PARFOR
  for (int tid = 0; tid < dim; ++tid) {
    if (!(grid_m[tid] > 1e-12)) continue;
    int gx_ = tid / gd.z / gd.y, gy_ = (tid / gd.z) % gd.y, gz_ = tid % gd.z;
    float m = grid_m[tid];
    v3 mv = ((const v3 *)grid_v_in)[tid];
    v3 v = vscale(mv, 1.f / m);
    v = vadd(v, vscale(((const v3 *)gravity)[0], dt));
    v3 lower = vscale(V((float)grid_lower[0], (float)grid_lower[1], (float)grid_lower[2]), dx);
    v3 gx = vadd(lower, vscale(V((float)gx_, (float)gy_, (float)gz_), dx));
    for (int b = 0; b < nb; ++b) {
      ((v3 *)grid_body_v_in)[(size_t)tid * (nb + 1) + b] = v;
      v3 bx = ((const v3 *)body_pos)[b];
      q4 bq = ((const q4 *)body_rot)[b];
      q4 tfsr = ((const q4 *)tfsr_)[b], sargs = ((const q4 *)args_)[b];
      float friction = tfsr.x, softness = tfsr.y;
      v3 gxb = spatial_transform_inv(bx, bq, gx);
      float dist = shape_sdf(tfsr, sargs, gxb);
      float infl;
      if (contact_active(dist, softness, &infl)) {
        v3 nrm = qrot(bq, normalized(shape_grad(tfsr, sargs, gxb)));
        v3 bv = vdivs(vsub(spatial_transform(((const v3 *)next_pos)[b], ((const q4 *)next_rot)[b], gxb), gx), dt);
        v3 rel = vsub(v, bv);
        float nc = vdot(rel, nrm);
        v3 vt = vsub(rel, vscale(nrm, fminf(nc, 0.f)));
        if (nc < 0. && (double)vdot(vt, vt) > 1e-30) {
          float vtn = length30(vt);
          vt = vscale(vscale(vt, 1.f / vtn), fmaxf(0.f, vtn + nc * friction));
        }
        v = vadd(vadd(bv, vscale(rel, 1 - infl)), vscale(vt, infl));
      }
    }
    ((v3 *)grid_body_v_in)[(size_t)tid * (nb + 1) + nb] = v;
    const int bound = 3;
    if (gx_ < bound && v.x < 0) v.x = 0;
    if (gx_ > gd.x - bound && v.x > 0) v.x = 0;
    if (gy_ < bound && v.y < 0) {
      if (ground_friction > 0.f) {
        if (ground_friction < 99.f) {
          float lin = v.y;
          v3 vit = V(v.x, 0.f, v.z);
          float lit = norm_eps(vit);
          v = vscale(vit, fmaxf((float)(1. + (double)(ground_friction * lin / lit)), 0.f));
        } else {
          v = V(0.f, 0.f, 0.f);
        }
      }
      v.y = 0;
    }
    if (gy_ > gd.y - bound && v.y > 0) v.y = 0;
    if (gz_ < bound && v.z < 0) v.z = 0;
    if (gz_ > gd.z - bound && v.z > 0) v.z = 0;
    ((v3 *)out_v)[tid] = v;
  }
}

/* integrator.cu:1059-1109 */
void orc_g2p(const float *px, const float *grid_v, const int *grid_lower, float dx, float inv_dx, float dt, const int *gdim,
             float *out_v, float ground_height, float *out_C, float *out_x, int n) {
  i3 gd = {gdim[0], gdim[1], gdim[2]};
  PARFOR
  for (int p = 0; p < n; ++p) {
    v3 lower = vscale(V((float)grid_lower[0], (float)grid_lower[1], (float)grid_lower[2]), dx);
    v3 x = vsub(((const v3 *)px)[p], lower);
    stencil st = make_stencil(x, inv_dx);
    v3 nv = V(0.f, 0.f, 0.f);
    m3 nC = mzero();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) {
      float weight = st.w[0][i] * st.w[1][j] * st.w[2][k];
      v3 dpos = vsub(V((float)i, (float)j, (float)k), st.fx);
      v3 v = ((const v3 *)grid_v)[grid_index(st.base.x + i, st.base.y + j, st.base.z + k, gd)];
      nv = vadd(nv, vscale(v, weight));
      nC = madd(nC, mscale(mouter(v, dpos), weight * inv_dx * 4.f));
    }
    v3 hi = V(((float)gd.x - 3.f) * dx, ((float)gd.y - 3.f) * dx, ((float)gd.z - 3.f) * dx);
    float lo = ground_height * dx;
    v3 t = vadd(x, vscale(nv, dt));
    v3 nx = V(fmaxf(fminf(t.x, hi.x), lo), fmaxf(fminf(t.y, hi.y), lo), fmaxf(fminf(t.z, hi.z), lo));
    ((v3 *)out_x)[p] = vadd(nx, lower);
    ((v3 *)out_v)[p] = nv;
    ((m3 *)out_C)[p] = nC;
  }
}

/* ------------------------------------------------------------------ adjoint kernels */
/* integrator.cu:1527-1614 */
void orc_g2p_grad(const float *px, const float *grid_v, const int *grid_lower, float dx, float inv_dx, float dt, const int *gdim,
                  const float *out_v, float ground_height, const float *out_C, const float *out_x, int n, float *x_grad,
                  float *grid_v_grad, const float *out_v_grad, const float *out_C_grad, const float *out_x_grad) {
  (void)out_C; (void)out_x;
  i3 gd = {gdim[0], gdim[1], gdim[2]};
  PARFOR
  for (int p = 0; p < n; ++p) {
    v3 lower = vscale(V((float)grid_lower[0], (float)grid_lower[1], (float)grid_lower[2]), dx);
    v3 x = vsub(((const v3 *)px)[p], lower);
    v3 gx = ((const v3 *)out_x_grad)[p];
    v3 gnv = ((const v3 *)out_v_grad)[p];
    m3 gnC = ((const m3 *)out_C_grad)[p];
    v3 nx = vadd(x, vscale(((const v3 *)out_v)[p], dt));
    v3 hi = V(((float)gd.x - 3.f) * dx, ((float)gd.y - 3.f) * dx, ((float)gd.z - 3.f) * dx);
    float lo = ground_height * dx;
    if (nx.x > hi.x || nx.x < lo) gx.x = 0;
    if (nx.y > hi.y || nx.y < lo) gx.y = 0;
    if (nx.z > hi.z || nx.z < lo) gx.z = 0;
    gnv = vadd(gnv, vscale(gx, dt));
    stencil st = make_stencil(x, inv_dx);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) {
      float weight = st.w[0][i] * st.w[1][j] * st.w[2][k];
      v3 dpos = vsub(V((float)i, (float)j, (float)k), st.fx);
      int tid = grid_index(st.base.x + i, st.base.y + j, st.base.z + k, gd);
      v3 v = ((const v3 *)grid_v)[tid];
      float xx = (float)((double)(weight * inv_dx) * 4.);
      v3 ggv = vadd(vscale(gnv, weight), vscale(mvec(gnC, dpos), xx));
      ATOMIC grid_v_grad[3 * tid] += ggv.x;
      ATOMIC grid_v_grad[3 * tid + 1] += ggv.y;
      ATOMIC grid_v_grad[3 * tid + 2] += ggv.z;
      gx = vadd(gx, vscale(vscale(mvec(mT(gnC), v), -inv_dx), xx));
      float gw = vdot(gnv, v) + (inv_dx * 4.f) * msum(mhad(mouter(v, dpos), gnC));
      gx = vadd(gx, vscale(dw(&st, i, j, k), gw));
    }
    v3 *o = (v3 *)x_grad + p;
    *o = vadd(*o, gx);
  }
}

/* integrator.cu:779-1057 */
void orc_grid_op_v2_grad(const float *grid_m, const float *grid_v_in, const float *grid_body_v_in, const int *grid_lower,
                         const float *gravity, const float *body_pos, const float *body_rot, const float *next_pos,
                         const float *next_rot, const float *tfsr_, const float *args_, float *grid_m_grad, float *grid_v_in_grad,
                         float *pos_grad, float *rot_grad, float *next_pos_grad, float *next_rot_grad, float dx, float inv_dx,
                         float dt, float ground_friction, const float *out_v, const float *out_v_grad, const int *gdim, int nb) {
  (void)gravity; (void)inv_dx; (void)out_v;
  i3 gd = {gdim[0], gdim[1], gdim[2]};
  int dim = gd.x * gd.y * gd.z;
  PARFOR
  for (int tid = 0; tid < dim; ++tid) {
    if (!(grid_m[tid] > 1e-12)) continue;
    int gx_ = tid / gd.z / gd.y, gy_ = (tid / gd.z) % gd.y, gz_ = tid % gd.z;
    v3 gv = ((const v3 *)out_v_grad)[tid];
    float m = grid_m[tid];
    v3 mv = ((const v3 *)grid_v_in)[tid];
    v3 vv = ((const v3 *)grid_body_v_in)[(size_t)tid * (nb + 1) + nb];
    v3 vin = vv;
    const int bound = 3;
    if (gx_ > gd.x - bound && vv.x > 0) vin.x = 0;
    if (gx_ < bound && vv.x < 0) vin.x = 0;
    float lin = 0.f, lit = 1.f;
    v3 vit = V(0.f, 0.f, 0.f);
    int hit_ground = gy_ < bound && vin.y < 0;
    if (hit_ground) {
      lin = vin.y;
      vit = V(vin.x, 0.f, vin.z);
      lit = norm_eps(vit);
      float flag = (float)(1. + (double)(ground_friction * lin / lit));
      vin = vscale(vit, fmaxf(flag, 0.f));
    }
    if (gz_ > gd.z - bound && vin.z > 0) gv.z = 0;
    if (gz_ < bound && vin.z < 0) gv.z = 0;
    if (gy_ > gd.y - bound && vin.y > 0) gv.y = 0;
    if (hit_ground) {
      gv.y = 0;
      float flag = (float)(1. + (double)(ground_friction * lin / lit));
      if (flag >= 0.) {
        v3 g_vit = vscale(gv, flag);
        float g_lin = ground_friction / lit * vdot(vit, gv);
        float g_lit = -ground_friction * lin / lit / lit * vdot(vit, gv);
        g_vit = vadd(g_vit, vscale(vdivs(vit, lit), g_lit));
        gv = V(g_vit.x, g_lin, g_vit.z);
      } else {
        gv = V(0.f, 0.f, 0.f);
      }
    }
    if (gx_ > gd.x - bound && vv.x > 0) gv.x = 0;
    if (gx_ < bound && vv.x < 0) gv.x = 0;

    v3 lower = vscale(V((float)grid_lower[0], (float)grid_lower[1], (float)grid_lower[2]), dx);
    v3 gx = vadd(lower, vscale(V((float)gx_, (float)gy_, (float)gz_), dx));
    for (int b = nb - 1; b >= 0; --b) {
      v3 v = ((const v3 *)grid_body_v_in)[(size_t)tid * (nb + 1) + b];
      v3 bx = ((const v3 *)body_pos)[b];
      q4 bq = ((const q4 *)body_rot)[b];
      q4 tfsr = ((const q4 *)tfsr_)[b], sargs = ((const q4 *)args_)[b];
      float friction = tfsr.x, softness = tfsr.y;
      v3 gxb = spatial_transform_inv(bx, bq, gx);
      float dist = shape_sdf(tfsr, sargs, gxb);
      float infl;
      if (!contact_active(dist, softness, &infl)) continue;
      v3 g_gxb = V(0.f, 0.f, 0.f), g_bx = V(0.f, 0.f, 0.f);
      q4 g_bq = {0.f, 0.f, 0.f, 0.f};
      float g_infl = 0.f;
      v3 un = shape_grad(tfsr, sargs, gxb);
      v3 rn = normalized(un);
      v3 nrm = qrot(bq, rn);
      v3 npos = ((const v3 *)next_pos)[b];
      q4 nrot = ((const q4 *)next_rot)[b];
      v3 bv = vdivs(vsub(spatial_transform(npos, nrot, gxb), gx), dt);
      v3 rel = vsub(v, bv);
      float nc = vdot(rel, nrm);
      v3 vt_in = vsub(rel, vscale(nrm, fminf(nc, 0.f)));
      int has_fric = nc < 0 && (double)vdot(vt_in, vt_in) > 1e-30;
      v3 vt = vt_in;
      float vtn = length30(vt_in);
      if (has_fric) vt = vscale(vscale(vt_in, 1.f / vtn), fmaxf(vtn + nc * friction, 0.f));
      float g_nc = 0.f;
      v3 g_bv = gv, g_rel = vscale(gv, 1 - infl), g_vt = vscale(gv, infl);
      g_infl += vdot(vsub(vt, rel), gv);
      if (has_fric) {
        float bf = vtn + nc * friction;
        if (bf > 0.) {
          g_nc += vdot(vt_in, g_vt) * friction / vtn;
          float g_vtn = -nc * g_nc / vtn;
          /* integrator.cu:969: grad * (1./norm) in double, then * bf */
          float inv = (float)(1. / (double)vtn);
          v3 t1 = vscale(g_vt, inv);
          g_vt = vadd(vscale(t1, bf), vdivs(vscale(vt_in, g_vtn), vtn));
        } else {
          g_vt = V(0.f, 0.f, 0.f);
        }
      }
      v3 g_n = V(0.f, 0.f, 0.f);
      g_rel = vadd(g_rel, g_vt);
      if (nc < 0.) {
        g_nc += -vdot(nrm, g_vt);
        g_n = vadd(g_n, vscale(g_vt, -nc));
      }
      g_rel = vadd(g_rel, vscale(nrm, g_nc));
      g_n = vadd(g_n, vscale(rel, g_nc));
      gv = g_rel;
      g_bv = vsub(g_bv, g_rel);
      v3 g_np = V(0.f, 0.f, 0.f);
      q4 g_nq = {0.f, 0.f, 0.f, 0.f};
      spatial_transform_backward(npos, nrot, gxb, vscale(g_bv, 1.f / dt), &g_np, &g_nq, &g_gxb);
      ATOMIC next_pos_grad[3 * b] += g_np.x;
      ATOMIC next_pos_grad[3 * b + 1] += g_np.y;
      ATOMIC next_pos_grad[3 * b + 2] += g_np.z;
      ATOMIC next_rot_grad[4 * b] += g_nq.w;
      ATOMIC next_rot_grad[4 * b + 1] += g_nq.x;
      ATOMIC next_rot_grad[4 * b + 2] += g_nq.y;
      ATOMIC next_rot_grad[4 * b + 3] += g_nq.z;
      v3 g_rn = V(0.f, 0.f, 0.f);
      qmul_backward(bq, rn, g_n, &g_bq, &g_rn);
      g_gxb = vadd(g_gxb, shape_grad_backward(tfsr, sargs, gxb, normalized_backward(un, g_rn)));
      float expdist = expf(-dist * softness);
      if (expdist <= 1) {
        float g_dist = -softness * expdist * g_infl;
        g_gxb = vadd(g_gxb, vscale(un, g_dist));
      }
      v3 g_tmp = V(0.f, 0.f, 0.f);
      inv_spatial_transform_backward(bx, bq, gx, g_gxb, &g_bx, &g_bq, &g_tmp);
      ATOMIC pos_grad[3 * b] += g_bx.x;
      ATOMIC pos_grad[3 * b + 1] += g_bx.y;
      ATOMIC pos_grad[3 * b + 2] += g_bx.z;
      ATOMIC rot_grad[4 * b] += g_bq.w;
      ATOMIC rot_grad[4 * b + 1] += g_bq.x;
      ATOMIC rot_grad[4 * b + 2] += g_bq.y;
      ATOMIC rot_grad[4 * b + 3] += g_bq.z;
    }
    /* integrator.cu:1054-1055: (1. / m) is double, narrowed to float by operator*(float,vec3) */
    float im = (float)(1. / (double)m);
    v3 *o = (v3 *)grid_v_in_grad + tid;
    *o = vadd(*o, vscale(gv, im));
    grid_m_grad[tid] += (-1.f / m / m) * vdot(mv, gv);
  }
}

/* integrator.cu:396-627 */
void orc_p2g_grad(const float *px, const float *pv, const float *pm, const float *pvol, const float *pF, const float *pU,
                  const float *psig, const float *pV, const float *pC, const float *mly, const int *grid_lower, const int *gdim,
                  float dx, float inv_dx, float dt, const float *outF, const float *grid_mv, const float *grid_m, float *x_grad,
                  float *v_grad, float *F_grad, float *C_grad, float *U_grad, float *sig_grad, float *V_grad,
                  const float *outF_grad, const float *grid_v_grad, const float *grid_m_grad, int n) {
  (void)outF; (void)grid_mv; (void)grid_m;
  i3 gd = {gdim[0], gdim[1], gdim[2]};
  PARFOR
  for (int p = 0; p < n; ++p) {
    m3 U = ((const m3 *)pU)[p], Vm = ((const m3 *)pV)[p];
    v3 sigma = ((const v3 *)psig)[p];
    float mu = mly[3 * p], lam = mly[3 * p + 1], yield = mly[3 * p + 2];
    float J;
    m3 nF;
    v3 sn = vmaxs(sigma, (float)SIG_CLIP_EPS);
    v3 eps = V(logf(sn.x), logf(sn.y), logf(sn.z));
    float mean = (eps.x + eps.y + eps.z) / 3.f;
    v3 eh = V(eps.x - mean, eps.y - mean, eps.z - mean);
    float ehn = norm_eps(eh);
    float dg = ehn - yield / (2 * mu);
    v3 ee = V(0.f, 0.f, 0.f);
    if (dg > 0.f) {
      v3 t = vsub(eps, vscale(eh, dg / ehn));
      ee = V(expf(t.x), expf(t.y), expf(t.z));
      nF = mmul(mmul(U, mdiag(ee)), mT(Vm));
      J = ee.x * ee.y * ee.z;
    } else {
      nF = ((const m3 *)pF)[p];
      J = sigma.x * sigma.y * sigma.z;
    }
    m3 r = mmul(U, mT(Vm));
    float gss = -dt * inv_dx * pvol[p] * 4.f * inv_dx;
    m3 stress = madd(mscale(mmul(msub(nF, r), mT(nF)), 2.f * mu), mident(lam * J * (J - 1)));
    m3 affine = madd(mscale(stress, gss), mscale(((const m3 *)pC)[p], pm[p]));
    float m_p = pm[p];
    v3 v_p = ((const v3 *)pv)[p];
    v3 x = ((const v3 *)px)[p];
    x = vsub(x, vscale(V((float)grid_lower[0], (float)grid_lower[1], (float)grid_lower[2]), dx));
    stencil st = make_stencil(x, inv_dx);
    m3 g_stress = mzero(), g_C = mzero();
    v3 g_x = V(0.f, 0.f, 0.f), g_v = V(0.f, 0.f, 0.f);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) {
      int idx = grid_index(st.base.x + i, st.base.y + j, st.base.z + k, gd);
      float N = st.w[0][i] * st.w[1][j] * st.w[2][k];
      v3 dpos = vscale(vsub(V((float)i, (float)j, (float)k), st.fx), dx);
      v3 ogv = ((const v3 *)grid_v_grad)[idx];
      v3 gN = dw(&st, i, j, k);
      m3 tmp = mouter(ogv, dpos);
      g_stress = madd(g_stress, mscale(tmp, N * gss));
      g_C = madd(g_C, mscale(tmp, N * m_p));
      float gm = grid_m_grad[idx];
      g_v = vadd(g_v, vscale(ogv, N * m_p));
      g_x = vadd(g_x, vscale(gN, gm * m_p));
      g_x = vadd(g_x, vscale(gN, vdot(v_p, ogv) * m_p));
      /* grad_dpos = -1 (int) * N -> float */
      g_x = vadd(g_x, vadd(vscale(mvec(mT(affine), ogv), -1 * N), vscale(gN, vdot(mvec(affine, dpos), ogv))));
    }
    { v3 *o = (v3 *)x_grad + p; *o = vadd(*o, g_x); }
    { v3 *o = (v3 *)v_grad + p; *o = vadd(*o, g_v); }
    { m3 *o = (m3 *)C_grad + p; *o = madd(*o, g_C); }
    m3 g_r = mscale(mmul(g_stress, nF), -2.f * mu);
    m3 g_U = mmul(g_r, Vm);
    m3 g_V = mmul(mT(g_r), U);
    m3 g_nF = madd(((const m3 *)outF_grad)[p], mscale(madd(mmul(mT(g_stress), msub(nF, r)), mmul(g_stress, nF)), 2.f * mu));
    float g_J = ((2 * J - 1) * lam) * (g_stress.m[0][0] + g_stress.m[1][1] + g_stress.m[2][2]);
    v3 g_sig = V(0.f, 0.f, 0.f);
    m3 g_F = mzero();
    if (dg > 0.f) {
      m3 E = mdiag(ee);
      g_U = madd(g_U, mmul(mmul(g_nF, Vm), E));
      g_V = madd(g_V, mmul(mmul(mT(g_nF), U), E));
      m3 t = mmul(mmul(mT(U), g_nF), Vm);
      v3 Fpart = V(t.m[0][0], t.m[1][1], t.m[2][2]);
      v3 Jpart = V(g_J * ee.y * ee.z, g_J * ee.x * ee.z, g_J * ee.x * ee.y);
      v3 g_eps = vmul(ee, vadd(Jpart, Fpart));
      v3 g_eh = vscale(g_eps, -dg / ehn);
      double g_ehn = (double)(-vdot(vdivs(eh, ehn), g_eps) * (yield / (2 * mu)) / ehn);
      v3 ehu = vdivs(eh, ehn);
      /* vec3 * double -> the double is converted to float at the call (vec3.h:47) */
      g_eh = vadd(g_eh, vscale(ehu, (float)g_ehn));
      /* sum/3. in double, converted to float by operator-(vec3,float) */
      float mean_g = (float)((double)(g_eh.x + g_eh.y + g_eh.z) / 3.);
      g_eps = vadd(g_eps, V(g_eh.x - mean_g, g_eh.y - mean_g, g_eh.z - mean_g));
      if (sigma.x >= SIG_CLIP_EPS) g_sig.x += g_eps.x / sigma.x;
      if (sigma.y >= SIG_CLIP_EPS) g_sig.y += g_eps.y / sigma.y;
      if (sigma.z >= SIG_CLIP_EPS) g_sig.z += g_eps.z / sigma.z;
    } else {
      g_sig = vadd(g_sig, V(g_J * sigma.y * sigma.z, g_J * sigma.x * sigma.z, g_J * sigma.x * sigma.y));
      g_F = madd(g_F, g_nF);
    }
    { m3 *o = (m3 *)U_grad + p; *o = madd(*o, g_U); }
    { m3 *o = (m3 *)V_grad + p; *o = madd(*o, g_V); }
    { v3 *o = (v3 *)sig_grad + p; *o = vadd(*o, g_sig); }
    { m3 *o = (m3 *)F_grad + p; *o = madd(*o, g_F); }
  }
}

/* integrator.cu:102-108 */
static inline float clamp_eps(float a) {
  if (a >= 0.) return fmaxf(a, 1e-6f);
  return fminf(a, -1e-6f);
}
/* integrator.cu:110-186 */
void orc_compute_svd_grad(const float *F, const float *C, const float *pU, const float *pV, const float *psig, float *newF_grad,
                          const float *U_grad, const float *V_grad, const float *sig_grad, float *F_grad, float *C_grad, float dt,
                          int n) {
  PARFOR
  for (int p = 0; p < n; ++p) {
    m3 u = ((const m3 *)pU)[p], v = ((const m3 *)pV)[p], gu = ((const m3 *)U_grad)[p], gv = ((const m3 *)V_grad)[p];
    v3 sigma = ((const v3 *)psig)[p];
    m3 sg = mdiag(sigma), gs = mdiag(((const v3 *)sig_grad)[p]);
    m3 vt = mT(v), ut = mT(u);
    m3 sigma_term = mmul(mmul(u, gs), vt);
    double s0 = sigma.x, s1 = sigma.y, s2 = sigma.z;
    s0 = s0 * s0; s1 = s1 * s1; s2 = s2 * s2;
    m3 FF;
    /* the double differences are narrowed to float by clamp(float a, ...) (integrator.cu:102) */
    FF.m[0][0] = 0.f; FF.m[0][1] = (float)(1.0 / clamp_eps((float)(s1 - s0))); FF.m[0][2] = (float)(1.0 / clamp_eps((float)(s2 - s0)));
    FF.m[1][0] = (float)(1.0 / clamp_eps((float)(s0 - s1))); FF.m[1][1] = 0.f; FF.m[1][2] = (float)(1.0 / clamp_eps((float)(s2 - s1)));
    FF.m[2][0] = (float)(1.0 / clamp_eps((float)(s0 - s2))); FF.m[2][1] = (float)(1.0 / clamp_eps((float)(s1 - s2))); FF.m[2][2] = 0.f;
    m3 u_term = mmul(mmul(u, mmul(mhad(FF, msub(mmul(ut, gu), mmul(mT(gu), u))), sg)), vt);
    m3 v_term = mmul(u, mmul(sg, mmul(mhad(FF, msub(mmul(vt, gv), mmul(mT(gv), v))), vt)));
    m3 G = madd(madd(madd(((const m3 *)newF_grad)[p], u_term), sigma_term), v_term);
    ((m3 *)newF_grad)[p] = G;
    { m3 *o = (m3 *)C_grad + p; *o = madd(*o, mscale(mmul(G, mT(((const m3 *)F)[p])), dt)); }
    { m3 *o = (m3 *)F_grad + p; *o = madd(*o, mmul(mT(madd(mident(1.f), mscale(((const m3 *)C)[p], dt))), G)); }
  }
}

/* integrator.cu:188-237 */
void orc_compute_dist(const float *px, const float *body_pos, const float *body_rot, const float *tfsr_, const float *args_,
                      float *dist, int nb, float *x_grad, float *pos_grad, float *rot_grad, const float *dist_grad,
                      int compute_grad, int n) {
  PARFOR
  for (int p = 0; p < n; ++p) {
    for (int b = 0; b < nb; ++b) {
      v3 bx = ((const v3 *)body_pos)[b];
      q4 bq = ((const q4 *)body_rot)[b];
      q4 tfsr = ((const q4 *)tfsr_)[b], sargs = ((const q4 *)args_)[b];
      v3 xp = ((const v3 *)px)[p];
      v3 gxb = spatial_transform_inv(bx, bq, xp);
      if (!compute_grad) {
        dist[(size_t)p * nb + b] = shape_sdf(tfsr, sargs, gxb);
      } else {
        float gd_ = dist_grad[(size_t)p * nb + b];
        v3 g_gxb = vscale(shape_grad(tfsr, sargs, gxb), gd_);
        v3 g_bx = V(0.f, 0.f, 0.f), g_x = V(0.f, 0.f, 0.f);
        q4 g_bq = {0.f, 0.f, 0.f, 0.f};
        inv_spatial_transform_backward(bx, bq, xp, g_gxb, &g_bx, &g_bq, &g_x);
        v3 *o = (v3 *)x_grad + p;
        *o = vadd(*o, g_x);
        ATOMIC pos_grad[3 * b] += g_bx.x;
        ATOMIC pos_grad[3 * b + 1] += g_bx.y;
        ATOMIC pos_grad[3 * b + 2] += g_bx.z;
        ATOMIC rot_grad[4 * b] += g_bq.w;
        ATOMIC rot_grad[4 * b + 1] += g_bq.x;
        ATOMIC rot_grad[4 * b + 2] += g_bq.y;
        ATOMIC rot_grad[4 * b + 3] += g_bq.z;
      }
    }
  }
}

/* integrator.cu:239-310 */
void orc_particle2mass(const float *px, const float *pm, const int *grid_lower, const int *gdim, float dx, float inv_dx,
                       float *grid_m, const float *grid_m_grad, float *x_grad, const int *ids, int id, int compute_grad, int n) {
  i3 gd = {gdim[0], gdim[1], gdim[2]};
  PARFOR
  for (int p = 0; p < n; ++p) {
    if (id != -1 && ids[p] != id) continue;
    v3 x = ((const v3 *)px)[p];
    x = vsub(x, vscale(V((float)grid_lower[0], (float)grid_lower[1], (float)grid_lower[2]), dx));
    stencil st = make_stencil(x, inv_dx);
    if (compute_grad) {
      v3 g = V(0.f, 0.f, 0.f);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) {
        int idx = grid_index(st.base.x + i, st.base.y + j, st.base.z + k, gd);
        g = vadd(g, vscale(dw(&st, i, j, k), grid_m_grad[idx] * pm[p]));
      }
      v3 *o = (v3 *)x_grad + p;
      *o = vadd(*o, g);
    } else {
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) {
        int idx = grid_index(st.base.x + i, st.base.y + j, st.base.z + k, gd);
        float c = pm[p] * (st.w[0][i] * st.w[1][j] * st.w[2][k]);
        ATOMIC grid_m[idx] += c;
      }
    }
  }
}

/* ------------------------------------------------------------------ leaf wrappers for golden tests */
void orc_von_mises(const float *F, const float *U, const float *sig, const float *Vm, const float *yield_mu, float *outF, float *J, int n) {
  for (int p = 0; p < n; ++p)
    J[p] = von_mises(((const m3 *)F)[p], ((const m3 *)U)[p], ((const v3 *)sig)[p], ((const m3 *)Vm)[p], yield_mu[2 * p], yield_mu[2 * p + 1], (m3 *)outF + p);
}
void orc_shape(const float *tfsr, const float *args, const float *gx, float *sdf, float *grad, const float *gout, float *gin, int n) {
  for (int p = 0; p < n; ++p) {
    q4 t = ((const q4 *)tfsr)[p], a = ((const q4 *)args)[p];
    v3 x = ((const v3 *)gx)[p];
    sdf[p] = shape_sdf(t, a, x);
    ((v3 *)grad)[p] = shape_grad(t, a, x);
    ((v3 *)gin)[p] = shape_grad_backward(t, a, x, ((const v3 *)gout)[p]);
  }
}
void orc_quat(const float *pos, const float *quat, const float *pt, const float *g, float *fwd, float *inv, float *gp_f, float *gq_f,
              float *gpt_f, float *gp_i, float *gq_i, float *gpt_i, int n) {
  for (int p = 0; p < n; ++p) {
    v3 P = ((const v3 *)pos)[p], X = ((const v3 *)pt)[p], G = ((const v3 *)g)[p];
    q4 Q = ((const q4 *)quat)[p];
    ((v3 *)fwd)[p] = spatial_transform(P, Q, X);
    ((v3 *)inv)[p] = spatial_transform_inv(P, Q, X);
    v3 a = V(0, 0, 0), c = V(0, 0, 0); q4 b = {0, 0, 0, 0};
    spatial_transform_backward(P, Q, X, G, &a, &b, &c);
    ((v3 *)gp_f)[p] = a; ((q4 *)gq_f)[p] = b; ((v3 *)gpt_f)[p] = c;
    a = V(0, 0, 0); c = V(0, 0, 0); b.w = b.x = b.y = b.z = 0;
    inv_spatial_transform_backward(P, Q, X, G, &a, &b, &c);
    ((v3 *)gp_i)[p] = a; ((q4 *)gq_i)[p] = b; ((v3 *)gpt_i)[p] = c;
  }
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
