// fk.cu -- Shadow-hand joint-velocity control + forward kinematics on the device (replaces HandSimulator.JointVel_Fk /
// hand_forward_kinematics / rigid_body_motion_hand, mpm/hand.py:20-65, 347-428, for the forward pass).
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
#ifdef DD_FK_DEBUG
  if (tid == 0) printf("base R %f %f %f | %f %f %f | %f %f %f p %f %f %f q0 %f q1 %f nops %d nmats %d ngeoms %d\n", base.R.a00, base.R.a01, base.R.a02, base.R.a10, base.R.a11, base.R.a12, base.R.a20, base.R.a21, base.R.a22, base.p.x, base.p.y, base.p.z, q[0], q[1], H.n_ops, H.n_mats, H.n_geoms);
#endif
  for (int k = 0; k < H.n_ops; ++k) {
#ifdef DD_FK_DEBUG
    if (tid == 0 && k < 6) printf("op %d kind %d idx %d reset %d g0 %d g1 %d cur.p %f %f %f R00 %f R11 %f\n", k, H.op_kind[k], H.op_index[k], H.op_reset[k], H.op_g0[k], H.op_g1[k], cur.p.x, cur.p.y, cur.p.z, cur.R.a00, cur.R.a11);
#endif
    if (H.op_reset[k]) cur = base;
    if (H.op_kind[k] == 0) {
      cur = fmul(cur, load_fr(H.mats + ((size_t)h * H.n_mats + H.op_index[k]) * 16));
    } else {
      int j = H.op_index[k];
      Fr T;
      const float *ax = H.joint_axis + ((size_t)h * 24 + j) * 3, *jp = H.joint_pos + ((size_t)h * 24 + j) * 3;
      T.R = axis_angle_to_mat(v3(ax[0] * q[j], ax[1] * q[j], ax[2] * q[j]));
      T.p = v3(jp[0], jp[1], jp[2]);
#ifdef DD_FK_DEBUG
      if (tid == 0 && k < 4) printf("  joint %d q %f ax %f %f %f jp %f %f %f T.R %f %f %f / %f %f %f\n", j, q[j], ax[0], ax[1], ax[2], jp[0], jp[1], jp[2], T.R.a00, T.R.a01, T.R.a02, T.R.a10, T.R.a11, T.R.a12);
#endif
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

}  // extern "C"
