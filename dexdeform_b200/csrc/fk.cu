// fk.cu -- Shadow-hand joint-velocity control + forward kinematics on the device (replaces HandSimulator.JointVel_Fk /
// hand_forward_kinematics / rigid_body_motion_hand, mpm/hand.py:20-65, 347-428) and its reverse mode (what torch autograd does
// for the reference): k_hand_fk / k_hand_fk_grad.
//
// One thread per (substep t, environment e, hand h): ramps the wrist pose, advances and clamps the 24 joint angles, walks
// the kinematic chains (alternating constant transforms and hinge joints) and writes position + quaternion of every
// collision primitive straight into the engine's pose table for state f+1+t.  The reference does this with ~100 small
// torch kernels per env step followed by a device->host->device copy per substep (mpm/simulator.py:553-559).
#include "mpm_math.cuh"
#include "../../include/dexdeform_mpm.h"
#include <cstdio>
#include <string>
#include <vector>

using namespace dd;

extern "C" int dd_sim_pose_table(dd_sim *sim, float **pos, float **rot, int *slots, int *n_envs, int *n_bodies);
extern "C" int dd_sim_pose_grad_table(dd_sim *sim, float **gpos, float **grot);
int dd_set_error(const char *msg);  // engine.cu

struct dd_hand {
  int nh, n_ops, n_mats, n_geoms;
  int *op_kind, *op_index, *op_reset, *op_g0, *op_g1, *geom_order;
  float *mats, *joint_pos, *joint_axis, *geom_local, *q_lo, *q_hi, *scale;
  int *action_map;
};

namespace {

struct Fr { M3 R; V3 p; };  // rigid frame
DD_DEV Fr fmul(const Fr &a, const Fr &b) { Fr r; r.R = mul(a.R, b.R); r.p = mul(a.R, b.p) + a.p; return r; }
DD_DEV Fr load_fr(const float *m) {  // row-major 4x4
  Fr f;
  f.R = m3(m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]);
  f.p = v3(m[3], m[7], m[11]);
  return f;
}
DD_DEV void store_fr(float *m, const Fr &f) {
  m[0] = f.R.a00; m[1] = f.R.a01; m[2] = f.R.a02; m[3] = f.p.x; m[4] = f.R.a10; m[5] = f.R.a11; m[6] = f.R.a12; m[7] = f.p.y;
  m[8] = f.R.a20; m[9] = f.R.a21; m[10] = f.R.a22; m[11] = f.p.z; m[12] = 0.f; m[13] = 0.f; m[14] = 0.f; m[15] = 1.f;
}
// pytorch3d.transforms.quaternion_to_matrix (w,x,y,z), not assuming unit norm
DD_DEV M3 quat_to_mat(float r, float i, float j, float k) {
  float s = 2.f / (r * r + i * i + j * j + k * k);
  return m3(1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r), s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r),
            s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j));
}
// pytorch3d.transforms.matrix_to_quaternion: four candidates, the one with the largest |component| wins.  Written with
// scalars only: local arrays here shared stack slots with the caller's joint-angle array (ptxas 12.9) and corrupted it.
DD_DEV float4 mat_to_quat(const M3 &m) {
  float t0 = 1.f + m.a00 + m.a11 + m.a22, t1 = 1.f + m.a00 - m.a11 - m.a22, t2 = 1.f - m.a00 + m.a11 - m.a22, t3 = 1.f - m.a00 - m.a11 + m.a22;
  float a0 = t0 > 0.f ? sqrtf(t0) : 0.f, a1 = t1 > 0.f ? sqrtf(t1) : 0.f, a2 = t2 > 0.f ? sqrtf(t2) : 0.f, a3 = t3 > 0.f ? sqrtf(t3) : 0.f;
  int best = 0;
  float ab = a0;
  if (a1 > ab) { best = 1; ab = a1; }
  if (a2 > ab) { best = 2; ab = a2; }
  if (a3 > ab) { best = 3; ab = a3; }
  float inv = 1.f / (2.f * fmaxf(ab, 0.1f));
  float4 q;
  if (best == 0) q = make_float4(a0 * a0, m.a21 - m.a12, m.a02 - m.a20, m.a10 - m.a01);
  else if (best == 1) q = make_float4(m.a21 - m.a12, a1 * a1, m.a10 + m.a01, m.a02 + m.a20);
  else if (best == 2) q = make_float4(m.a02 - m.a20, m.a10 + m.a01, a2 * a2, m.a12 + m.a21);
  else q = make_float4(m.a10 - m.a01, m.a20 + m.a02, m.a21 + m.a12, a3 * a3);
  return make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
}
// pytorch3d axis_angle_to_matrix (via the quaternion, with its small-angle series)
DD_DEV M3 axis_angle_to_mat(V3 aa) {
  float ang = sqrtf(dot(aa, aa)), half = 0.5f * ang;
  float s = ang < 1e-6f ? 0.5f - ang * ang / 48.f : sinf(half) / ang;
  return quat_to_mat(cosf(half), aa.x * s, aa.y * s, aa.z * s);
}

__global__ void k_hand_fk(dd_hand H, int S, int E, int nb, const float *__restrict__ base_pose, const float *__restrict__ joint_rot,
                          const float *__restrict__ action, float4 *__restrict__ pos_out, float4 *__restrict__ rot_out,
                          float *__restrict__ next_base, float *__restrict__ next_q, int has_base_action) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= S * E * H.nh) return;
  int h = tid % H.nh, e = (tid / H.nh) % E, t = tid / (H.nh * E);
  const float *act = action + ((size_t)e * H.nh + h) * 26;
  // ---- wrist: linear ramp of translation and axis-angle (hand.py:20-65)
  Fr base = load_fr(base_pose + ((size_t)e * H.nh + h) * 16);
  if (has_base_action) {
    float ramp = (float)(t + 1) / (float)S;
    V3 tr = v3(act[20] * H.scale[20], act[21] * H.scale[21], act[22] * H.scale[22]) * ramp;
    V3 rv = v3(act[23] * H.scale[23], act[24] * H.scale[24], act[25] * H.scale[25]) * ramp;
    float4 qb = mat_to_quat(base.R);
    float q0[4] = {qb.x, qb.y, qb.z, qb.w};
    float w = sqrtf(dot(rv, rv) + 1e-16f), sw = sinf(0.5f * w) / fminf(fmaxf(w, 1e-7f), 1e9f);
    float d0 = cosf(0.5f * w), d1 = rv.x * sw, d2 = rv.y * sw, d3 = rv.z * sw;
    float o0 = q0[0] * d0 - q0[1] * d1 - q0[2] * d2 - q0[3] * d3, o1 = q0[0] * d1 + q0[1] * d0 - q0[2] * d3 + q0[3] * d2;
    float o2 = q0[0] * d2 + q0[1] * d3 + q0[2] * d0 - q0[3] * d1, o3 = q0[0] * d3 - q0[1] * d2 + q0[2] * d1 + q0[3] * d0;
    float n = rsqrtf(o0 * o0 + o1 * o1 + o2 * o2 + o3 * o3);
    base.R = quat_to_mat(o0 * n, o1 * n, o2 * n, o3 * n);
    base.p = base.p + tr;
  }
  // ---- joints: q_t = clamp(q + (t+1) * clamp(a, -1, 1) * scale) (hand.py:412-416)
  float q[24];
  for (int j = 0; j < 24; ++j) {
    int a = H.action_map[j];
    float da = fminf(fmaxf(act[a], -1.f), 1.f) * H.scale[a];
    q[j] = fminf(fmaxf(joint_rot[((size_t)e * H.nh + h) * 24 + j] + da * (float)(t + 1), H.q_lo[j]), H.q_hi[j]);
  }
  if (t == S - 1) {
    if (next_base) store_fr(next_base + ((size_t)e * H.nh + h) * 16, base);
    if (next_q) for (int j = 0; j < 24; ++j) next_q[((size_t)e * H.nh + h) * 24 + j] = q[j];
  }
  // ---- chains (hand.py:366-376) and primitives (hand.py:377-381)
  Fr cur = base;
  size_t out0 = ((size_t)t * E + e) * nb + (size_t)h * H.n_geoms;
  for (int k = 0; k < H.n_ops; ++k) {
    if (H.op_reset[k]) cur = base;
    if (H.op_kind[k] == 0) {
      cur = fmul(cur, load_fr(H.mats + ((size_t)h * H.n_mats + H.op_index[k]) * 16));
    } else {
      int j = H.op_index[k];
      Fr T;
      const float *ax = H.joint_axis + ((size_t)h * 24 + j) * 3, *jp = H.joint_pos + ((size_t)h * 24 + j) * 3;
      T.R = axis_angle_to_mat(v3(ax[0] * q[j], ax[1] * q[j], ax[2] * q[j]));
      T.p = v3(jp[0], jp[1], jp[2]);
      cur = fmul(cur, T);
      for (int gi = H.op_g0[k]; gi < H.op_g1[k]; ++gi) {  // primitives carried by this joint (first visit only)
        int g = H.geom_order[gi];
        Fr G = fmul(cur, load_fr(H.geom_local + ((size_t)h * H.n_geoms + g) * 16));
        pos_out[out0 + g] = make_float4(G.p.x, G.p.y, G.p.z, 0.f);
        rot_out[out0 + g] = mat_to_quat(G.R);
      }
    }
  }
}

// ---- reverse mode ------------------------------------------------------------------------------------------------------
// Adjoint of k_hand_fk: one thread per (substep t, environment e, hand h) recomputes its chain frames, walks the ops backwards
// and adds its share of dL/d(action), dL/d(base_pose), dL/d(joint_rot) with atomics (S threads per hand).  It differentiates
// exactly what the forward kernel / the torch mirror evaluate (pytorch3d's matrix_to_quaternion branch selection and 0.1 floor,
// the clamps of actions and joint angles with torch's sub-gradient: 1 inside the closed interval).
constexpr int kMaxOps = 80;
DD_DEV M3 outer3(V3 a, V3 b) { return m3(a.x * b.x, a.x * b.y, a.x * b.z, a.y * b.x, a.y * b.y, a.y * b.z, a.z * b.x, a.z * b.y, a.z * b.z); }
// r = a (x) b:  ga += (gr.R b.R^T + gr.p (x) b.p, gr.p)
DD_DEV void fmul_adj_left(const Fr &b, const Fr &gr, Fr &ga) { ga.R = ga.R + mul_nt(gr.R, b.R) + outer3(gr.p, b.p); ga.p = ga.p + gr.p; }
DD_DEV M3 mat_to_quat_adj(const M3 &m, float4 gq) {
  float t0 = 1.f + m.a00 + m.a11 + m.a22, t1 = 1.f + m.a00 - m.a11 - m.a22, t2 = 1.f - m.a00 + m.a11 - m.a22, t3 = 1.f - m.a00 - m.a11 + m.a22;
  float a0 = t0 > 0.f ? sqrtf(t0) : 0.f, a1 = t1 > 0.f ? sqrtf(t1) : 0.f, a2 = t2 > 0.f ? sqrtf(t2) : 0.f, a3 = t3 > 0.f ? sqrtf(t3) : 0.f;
  int best = 0;
  float ab = a0;
  if (a1 > ab) { best = 1; ab = a1; }
  if (a2 > ab) { best = 2; ab = a2; }
  if (a3 > ab) { best = 3; ab = a3; }
  float den = fmaxf(ab, 0.1f), inv = 1.f / (2.f * den);
  float c0, c1, c2, c3;  // the selected candidate before the division
  if (best == 0) { c0 = a0 * a0; c1 = m.a21 - m.a12; c2 = m.a02 - m.a20; c3 = m.a10 - m.a01; }
  else if (best == 1) { c0 = m.a21 - m.a12; c1 = a1 * a1; c2 = m.a10 + m.a01; c3 = m.a02 + m.a20; }
  else if (best == 2) { c0 = m.a02 - m.a20; c1 = m.a10 + m.a01; c2 = a2 * a2; c3 = m.a12 + m.a21; }
  else { c0 = m.a10 - m.a01; c1 = m.a20 + m.a02; c2 = m.a21 + m.a12; c3 = a3 * a3; }
  float g0 = gq.x * inv, g1 = gq.y * inv, g2 = gq.z * inv, g3 = gq.w * inv;  // dL/dc
  float gbest = best == 0 ? g0 : best == 1 ? g1 : best == 2 ? g2 : g3;
  float g_ab = 2.f * ab * gbest;                                               // numerator entry q_abs^2
  if (ab > 0.1f) g_ab -= (gq.x * c0 + gq.y * c1 + gq.z * c2 + gq.w * c3) / (2.f * den * den);  // denominator 2 max(q_abs, 0.1)
  float g_t = ab > 0.f ? g_ab / (2.f * ab) : 0.f;                              // q_abs = sqrt(t) on t > 0
  M3 g = mzero();
  if (best == 0) {
    g.a00 = g_t; g.a11 = g_t; g.a22 = g_t;
    g.a21 += g1; g.a12 -= g1; g.a02 += g2; g.a20 -= g2; g.a10 += g3; g.a01 -= g3;
  } else if (best == 1) {
    g.a00 = g_t; g.a11 = -g_t; g.a22 = -g_t;
    g.a21 += g0; g.a12 -= g0; g.a10 += g2; g.a01 += g2; g.a02 += g3; g.a20 += g3;
  } else if (best == 2) {
    g.a00 = -g_t; g.a11 = g_t; g.a22 = -g_t;
    g.a02 += g0; g.a20 -= g0; g.a10 += g1; g.a01 += g1; g.a12 += g3; g.a21 += g3;
  } else {
    g.a00 = -g_t; g.a11 = -g_t; g.a22 = g_t;
    g.a10 += g0; g.a01 -= g0; g.a20 += g1; g.a02 += g1; g.a21 += g2; g.a12 += g2;
  }
  return g;
}
// adjoint of quat_to_mat(r, i, j, k) (no unit-norm assumption): returns dL/d(r, i, j, k)
DD_DEV float4 quat_to_mat_adj(float r, float i, float j, float k, const M3 &g) {
  float n = r * r + i * i + j * j + k * k, s = 2.f / n;
  float gs = g.a00 * -(j * j + k * k) + g.a01 * (i * j - k * r) + g.a02 * (i * k + j * r) + g.a10 * (i * j + k * r) + g.a11 * -(i * i + k * k) +
             g.a12 * (j * k - i * r) + g.a20 * (i * k - j * r) + g.a21 * (j * k + i * r) + g.a22 * -(i * i + j * j);
  float gr = s * (-k * g.a01 + j * g.a02 + k * g.a10 - i * g.a12 - j * g.a20 + i * g.a21);
  float gi = s * (j * (g.a01 + g.a10) + k * (g.a02 + g.a20) - 2.f * i * (g.a11 + g.a22) + r * (g.a21 - g.a12));
  float gj = s * (-2.f * j * (g.a00 + g.a22) + i * (g.a01 + g.a10) + r * (g.a02 - g.a20) + k * (g.a12 + g.a21));
  float gk = s * (-2.f * k * (g.a00 + g.a11) + r * (g.a10 - g.a01) + i * (g.a02 + g.a20) + j * (g.a12 + g.a21));
  float c = -s * s * gs;  // through s = 2 / n
  return make_float4(gr + c * r, gi + c * i, gj + c * j, gk + c * k);
}

__global__ void k_hand_fk_grad(dd_hand H, int S, int E, int nb, const float *__restrict__ base_pose, const float *__restrict__ joint_rot,
                               const float *__restrict__ action, const float4 *__restrict__ gpos, const float4 *__restrict__ grot,
                               const float *__restrict__ g_next_base, const float *__restrict__ g_next_q, float *__restrict__ g_base,
                               float *__restrict__ g_q, float *__restrict__ g_action, int has_base_action) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= S * E * H.nh) return;
  int h = tid % H.nh, e = (tid / H.nh) % E, t = tid / (H.nh * E);
  size_t eh = (size_t)e * H.nh + h;
  const float *act = action + eh * 26;
  const float ramp = (float)(t + 1) / (float)S;
  // ---- forward replay: wrist
  const Fr base0 = load_fr(base_pose + eh * 16);
  Fr base = base0;
  float q0w = 1.f, q0x = 0.f, q0y = 0.f, q0z = 0.f, d0 = 1.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, o0 = 1.f, o1 = 0.f, o2 = 0.f, o3 = 0.f, on = 1.f, w = 0.f, cw = 1.f;
  V3 rv = vzero();
  if (has_base_action) {
    V3 tr = v3(act[20] * H.scale[20], act[21] * H.scale[21], act[22] * H.scale[22]) * ramp;
    rv = v3(act[23] * H.scale[23], act[24] * H.scale[24], act[25] * H.scale[25]) * ramp;
    float4 qb = mat_to_quat(base0.R);
    q0w = qb.x; q0x = qb.y; q0y = qb.z; q0z = qb.w;
    w = sqrtf(dot(rv, rv) + 1e-16f);
    cw = fminf(fmaxf(w, 1e-7f), 1e9f);
    float sw = sinf(0.5f * w) / cw;
    d0 = cosf(0.5f * w); d1 = rv.x * sw; d2 = rv.y * sw; d3 = rv.z * sw;
    o0 = q0w * d0 - q0x * d1 - q0y * d2 - q0z * d3; o1 = q0w * d1 + q0x * d0 - q0y * d3 + q0z * d2;
    o2 = q0w * d2 + q0x * d3 + q0y * d0 - q0z * d1; o3 = q0w * d3 - q0x * d2 + q0y * d1 + q0z * d0;
    on = rsqrtf(o0 * o0 + o1 * o1 + o2 * o2 + o3 * o3);
    base.R = quat_to_mat(o0 * on, o1 * on, o2 * on, o3 * on);
    base.p = base0.p + tr;
  }
  // ---- forward replay: joints and chain frames
  float q[24], gq[24];
  for (int j = 0; j < 24; ++j) {
    int a = H.action_map[j];
    float da = fminf(fmaxf(act[a], -1.f), 1.f) * H.scale[a];
    q[j] = fminf(fmaxf(joint_rot[eh * 24 + j] + da * (float)(t + 1), H.q_lo[j]), H.q_hi[j]);
    gq[j] = (t == S - 1 && g_next_q) ? g_next_q[eh * 24 + j] : 0.f;
  }
  Fr fr[kMaxOps];
  {
    Fr cur = base;
    for (int k = 0; k < H.n_ops; ++k) {
      if (H.op_reset[k]) cur = base;
      if (H.op_kind[k] == 0) {
        cur = fmul(cur, load_fr(H.mats + ((size_t)h * H.n_mats + H.op_index[k]) * 16));
      } else {
        int j = H.op_index[k];
        const float *ax = H.joint_axis + ((size_t)h * 24 + j) * 3, *jp = H.joint_pos + ((size_t)h * 24 + j) * 3;
        Fr T;
        T.R = axis_angle_to_mat(v3(ax[0] * q[j], ax[1] * q[j], ax[2] * q[j]));
        T.p = v3(jp[0], jp[1], jp[2]);
        cur = fmul(cur, T);
      }
      fr[k] = cur;
    }
  }
  // ---- backward over the ops
  Fr gbase_t, gcur;
  gbase_t.R = mzero(); gbase_t.p = vzero(); gcur.R = mzero(); gcur.p = vzero();
  if (t == S - 1 && g_next_base) gbase_t = load_fr(g_next_base + eh * 16);
  size_t out0 = ((size_t)t * E + e) * nb + (size_t)h * H.n_geoms;
  for (int k = H.n_ops - 1; k >= 0; --k) {
    const Fr cur = fr[k];
    const Fr prev = H.op_reset[k] ? base : fr[k - 1];
    Fr X;
    int j = -1;
    if (H.op_kind[k] == 0) {
      X = load_fr(H.mats + ((size_t)h * H.n_mats + H.op_index[k]) * 16);
    } else {
      j = H.op_index[k];
      const float *ax = H.joint_axis + ((size_t)h * 24 + j) * 3, *jp = H.joint_pos + ((size_t)h * 24 + j) * 3;
      X.R = axis_angle_to_mat(v3(ax[0] * q[j], ax[1] * q[j], ax[2] * q[j]));
      X.p = v3(jp[0], jp[1], jp[2]);
      for (int gi = H.op_g0[k]; gi < H.op_g1[k]; ++gi) {  // primitives carried by this joint
        int g = H.geom_order[gi];
        Fr L = load_fr(H.geom_local + ((size_t)h * H.n_geoms + g) * 16), G = fmul(cur, L), gG;
        float4 gp = gpos[out0 + g], gr = grot[out0 + g];
        gG.p = v3(gp.x, gp.y, gp.z);
        gG.R = mat_to_quat_adj(G.R, gr);
        fmul_adj_left(L, gG, gcur);
      }
    }
    Fr gprev;
    gprev.R = mzero(); gprev.p = vzero();
    fmul_adj_left(X, gcur, gprev);
    if (j >= 0) {  // dL/dq_j = < prev.R^T gcur.R , [a]x X.R >   (X.R = exp(q_j [a]x))
      const float *ax = H.joint_axis + ((size_t)h * 24 + j) * 3;
      M3 gX = mul_tn(prev.R, gcur.R), R = X.R;
      float a0 = ax[0], a1 = ax[1], a2 = ax[2];
      M3 D = m3(-a2 * R.a10 + a1 * R.a20, -a2 * R.a11 + a1 * R.a21, -a2 * R.a12 + a1 * R.a22,
                a2 * R.a00 - a0 * R.a20, a2 * R.a01 - a0 * R.a21, a2 * R.a02 - a0 * R.a22,
                -a1 * R.a00 + a0 * R.a10, -a1 * R.a01 + a0 * R.a11, -a1 * R.a02 + a0 * R.a12);
      gq[j] += ddot(gX, D);
    }
    if (H.op_reset[k]) { gbase_t.R = gbase_t.R + gprev.R; gbase_t.p = gbase_t.p + gprev.p; gcur.R = mzero(); gcur.p = vzero(); }
    else gcur = gprev;
  }
  // ---- joints: q_t = clamp(q0 + (t+1) clamp(a) scale)
  for (int j = 0; j < 24; ++j) {
    int a = H.action_map[j];
    float da = fminf(fmaxf(act[a], -1.f), 1.f) * H.scale[a];
    float raw = joint_rot[eh * 24 + j] + da * (float)(t + 1);
    float g = (raw >= H.q_lo[j] && raw <= H.q_hi[j]) ? gq[j] : 0.f;
    if (g != 0.f) {
      atomicAdd(g_q + eh * 24 + j, g);
      if (act[a] >= -1.f && act[a] <= 1.f) atomicAdd(g_action + eh * 26 + a, g * (float)(t + 1) * H.scale[a]);
    }
  }
  // ---- wrist
  float *gb = g_base + eh * 16;
  if (!has_base_action) {
    const float r9[9] = {gbase_t.R.a00, gbase_t.R.a01, gbase_t.R.a02, gbase_t.R.a10, gbase_t.R.a11, gbase_t.R.a12, gbase_t.R.a20, gbase_t.R.a21, gbase_t.R.a22};
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) atomicAdd(gb + 4 * r + c, r9[3 * r + c]);
    atomicAdd(gb + 3, gbase_t.p.x); atomicAdd(gb + 7, gbase_t.p.y); atomicAdd(gb + 11, gbase_t.p.z);
    return;
  }
  atomicAdd(gb + 3, gbase_t.p.x); atomicAdd(gb + 7, gbase_t.p.y); atomicAdd(gb + 11, gbase_t.p.z);
  atomicAdd(g_action + eh * 26 + 20, gbase_t.p.x * H.scale[20] * ramp);
  atomicAdd(g_action + eh * 26 + 21, gbase_t.p.y * H.scale[21] * ramp);
  atomicAdd(g_action + eh * 26 + 22, gbase_t.p.z * H.scale[22] * ramp);
  // base_t.R = quat_to_mat(o / |o|), o = qb (x) dq
  float4 g_on = quat_to_mat_adj(o0 * on, o1 * on, o2 * on, o3 * on, gbase_t.R);
  float proj = (g_on.x * o0 + g_on.y * o1 + g_on.z * o2 + g_on.w * o3) * on * on;  // (on_hat . g) with on_hat = o * on
  float go0 = on * (g_on.x - o0 * proj), go1 = on * (g_on.y - o1 * proj), go2 = on * (g_on.z - o2 * proj), go3 = on * (g_on.w - o3 * proj);
  float4 g_qb = make_float4(go0 * d0 + go1 * d1 + go2 * d2 + go3 * d3, -go0 * d1 + go1 * d0 + go2 * d3 - go3 * d2,
                            -go0 * d2 - go1 * d3 + go2 * d0 + go3 * d1, -go0 * d3 + go1 * d2 - go2 * d1 + go3 * d0);
  float gd0 = go0 * q0w + go1 * q0x + go2 * q0y + go3 * q0z, gd1 = -go0 * q0x + go1 * q0w - go2 * q0z + go3 * q0y;
  float gd2 = -go0 * q0y + go1 * q0z + go2 * q0w - go3 * q0x, gd3 = -go0 * q0z - go1 * q0y + go2 * q0x + go3 * q0w;
  {
    float sh = sinf(0.5f * w), ch = cosf(0.5f * w), sw = sh / cw;
    float g_sw = rv.x * gd1 + rv.y * gd2 + rv.z * gd3;
    float dsw_dw = 0.5f * ch / cw - ((w >= 1e-7f && w <= 1e9f) ? sh / (cw * cw) : 0.f);
    float g_w = -0.5f * sh * gd0 + g_sw * dsw_dw;
    V3 g_rv = v3(gd1, gd2, gd3) * sw + rv * (g_w / w);
    atomicAdd(g_action + eh * 26 + 23, g_rv.x * H.scale[23] * ramp);
    atomicAdd(g_action + eh * 26 + 24, g_rv.y * H.scale[24] * ramp);
    atomicAdd(g_action + eh * 26 + 25, g_rv.z * H.scale[25] * ramp);
  }
  M3 gR0 = mat_to_quat_adj(base0.R, g_qb);
  const float r9[9] = {gR0.a00, gR0.a01, gR0.a02, gR0.a10, gR0.a11, gR0.a12, gR0.a20, gR0.a21, gR0.a22};
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) atomicAdd(gb + 4 * r + c, r9[3 * r + c]);
}

template <class T>
T *to_device(const std::vector<T> &v) {
  T *d = nullptr;
  cudaMalloc((void **)&d, sizeof(T) * (v.empty() ? 1 : v.size()));
  if (!v.empty()) cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
  return d;
}

}  // namespace

extern "C" {

int dd_hand_create(int n_hands, int n_ops, const int *op_kind, const int *op_index, const int *op_reset, int n_mats, const float *mats,
                   const float *joint_pos, const float *joint_axis, int n_geoms, const int *geom_joint, const float *geom_local,
                   const float *q_lower, const float *q_upper, const int *action_map, const float *action_scale, dd_hand **out) {
  if (!out || n_hands < 1 || n_ops < 1 || n_geoms < 1) return dd_set_error("dd_hand_create: bad arguments");
  dd_hand *H = new dd_hand();
  H->nh = n_hands; H->n_ops = n_ops; H->n_mats = n_mats; H->n_geoms = n_geoms;
  // attach every primitive to the first op that produces its joint's pose
  std::vector<int> g0(n_ops, 0), g1(n_ops, 0), order;
  std::vector<char> seen(24, 0);
  for (int k = 0; k < n_ops; ++k) {
    g0[k] = g1[k] = (int)order.size();
    if (op_kind[k] == 1 && !seen[op_index[k]]) {
      seen[op_index[k]] = 1;
      for (int g = 0; g < n_geoms; ++g)
        if (geom_joint[g] == op_index[k]) order.push_back(g);
      g1[k] = (int)order.size();
    }
  }
  if ((int)order.size() != n_geoms) { delete H; return dd_set_error("dd_hand_create: a primitive is attached to a joint that no chain reaches"); }
  H->op_kind = to_device(std::vector<int>(op_kind, op_kind + n_ops));
  H->op_index = to_device(std::vector<int>(op_index, op_index + n_ops));
  H->op_reset = to_device(std::vector<int>(op_reset, op_reset + n_ops));
  H->op_g0 = to_device(g0); H->op_g1 = to_device(g1); H->geom_order = to_device(order);
  H->mats = to_device(std::vector<float>(mats, mats + (size_t)n_hands * n_mats * 16));
  H->joint_pos = to_device(std::vector<float>(joint_pos, joint_pos + (size_t)n_hands * 24 * 3));
  H->joint_axis = to_device(std::vector<float>(joint_axis, joint_axis + (size_t)n_hands * 24 * 3));
  H->geom_local = to_device(std::vector<float>(geom_local, geom_local + (size_t)n_hands * n_geoms * 16));
  H->q_lo = to_device(std::vector<float>(q_lower, q_lower + 24));
  H->q_hi = to_device(std::vector<float>(q_upper, q_upper + 24));
  H->scale = to_device(std::vector<float>(action_scale, action_scale + 26));
  H->action_map = to_device(std::vector<int>(action_map, action_map + 24));
  *out = H;
  return 0;
}

void dd_hand_destroy(dd_hand *H) {
  if (!H) return;
  void *p[] = {H->op_kind, H->op_index, H->op_reset, H->op_g0, H->op_g1, H->geom_order, H->mats, H->joint_pos, H->joint_axis, H->geom_local,
               H->q_lo, H->q_hi, H->scale, H->action_map};
  for (void *x : p) cudaFree(x);
  delete H;
}

// base_pose (E, nh, 4, 4), joint_rot (E, nh, 24), action (E, nh, 26): DEVICE pointers.  Writes the poses of states f+1 .. f+S into
// the simulator and the end-of-step kinematic state into next_base / next_q (device, may alias the inputs, may be NULL).
int dd_hand_fk(dd_hand *H, dd_sim *sim, int f, int S, const float *base_pose, const float *joint_rot, const float *action, float *next_base,
               float *next_q, int has_base_action, cudaStream_t stream) {
  if (!H || !sim || !base_pose || !joint_rot || !action) return dd_set_error("dd_hand_fk: null argument");
  float *pos = nullptr, *rot = nullptr;
  int slots = 0, E = 0, nb = 0;
  if (dd_sim_pose_table(sim, &pos, &rot, &slots, &E, &nb)) return 1;
  if (nb != H->nh * H->n_geoms) return dd_set_error("dd_hand_fk: simulator body count does not match the hand tables");
  if (f < 0 || S < 1 || f + S >= slots) return dd_set_error("dd_hand_fk: substep range exceeds max_steps");
  int n = S * E * H->nh;
  size_t off = (size_t)(f + 1) * E * nb;
  k_hand_fk<<<(n + 63) / 64, 64, 0, stream>>>(*H, S, E, nb, base_pose, joint_rot, action, reinterpret_cast<float4 *>(pos) + off,
                                              reinterpret_cast<float4 *>(rot) + off, next_base, next_q, has_base_action);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return dd_set_error((std::string("dd_hand_fk: ") + cudaGetErrorString(e)).c_str());
  return 0;
}

// Reverse mode of dd_hand_fk for the same inputs: reads the pose gradients of states f+1 .. f+S from the simulator and ADDS
// dL/d(base_pose) (E, nh, 4, 4: rotation block and translation column), dL/d(joint_rot) (E, nh, 24) and dL/d(action) (E, nh, 26)
// to g_base / g_q / g_action (device, caller-zeroed).  g_next_base / g_next_q: gradients flowing into the end-of-step kinematic
// state from later env steps (device, may be NULL).
int dd_hand_fk_grad(dd_hand *H, dd_sim *sim, int f, int S, const float *base_pose, const float *joint_rot, const float *action,
                    const float *g_next_base, const float *g_next_q, float *g_base, float *g_q, float *g_action, int has_base_action,
                    cudaStream_t stream) {
  if (!H || !sim || !base_pose || !joint_rot || !action || !g_base || !g_q || !g_action) return dd_set_error("dd_hand_fk_grad: null argument");
  if (H->n_ops > kMaxOps) return dd_set_error("dd_hand_fk_grad: kinematic description has too many ops");
  float *pos = nullptr, *rot = nullptr, *gpos = nullptr, *grot = nullptr;
  int slots = 0, E = 0, nb = 0;
  if (dd_sim_pose_table(sim, &pos, &rot, &slots, &E, &nb) || dd_sim_pose_grad_table(sim, &gpos, &grot)) return 1;
  if (nb != H->nh * H->n_geoms) return dd_set_error("dd_hand_fk_grad: simulator body count does not match the hand tables");
  if (f < 0 || S < 1 || f + S >= slots) return dd_set_error("dd_hand_fk_grad: substep range exceeds max_steps");
  int n = S * E * H->nh;
  size_t off = (size_t)(f + 1) * E * nb;
  k_hand_fk_grad<<<(n + 63) / 64, 64, 0, stream>>>(*H, S, E, nb, base_pose, joint_rot, action, reinterpret_cast<const float4 *>(gpos) + off,
                                                   reinterpret_cast<const float4 *>(grot) + off, g_next_base, g_next_q, g_base, g_q, g_action, has_base_action);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return dd_set_error((std::string("dd_hand_fk_grad: ") + cudaGetErrorString(e)).c_str());
  return 0;
}

}  // extern "C"
