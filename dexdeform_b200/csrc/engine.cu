// engine.cu -- fused, batched MLS-MPM substep + adjoint for sm_100a (ABI-2, dd_* symbols in include/dexdeform_mpm.h).
//
// What differs from the reference pipeline (mpm/simulator.py:561-585, integrator.cu):
//   * particle state lives in float4 SoA planes, one checkpoint slot per substep (x,v,C | F), E environments batched
//   * compute_svd + p2g are one kernel; F~ and the matrices U, V never touch HBM (the reference writes/reads 120 B/particle).
//     What the adjoint needs of them is checkpointed compactly per substep: two quaternions, sigma and the affine matrix of the
//     scatter (64 B), so the backward pass runs no SVD, no QR and no stress evaluation
//   * tiled kernels: a warp owns a chunk of one 4^3-cell brick and a private 8^3-node shared-memory tile; particles are stored
//     in 32-wide rows laid out so that the lanes of a row sit in different cells AND different shared-memory bank groups
//     (no atomics, no bank conflicts); persistent launches hand chunks out through a ticket counter, largest first;
//     the inputs of the next row are staged with cp.async
//   * a grid node is one float4 (mv.xyz, m), scattered with one red.global.add.v4.f32 instead of four scalar atomics
//   * the per-body input velocities (grid_body_v_in, (nb+1)*12 B per node) are not stored: the adjoint re-derives them
//     from a contact bit-mask (contacts are velocity independent) by replaying the few contacting bodies
//   * p2g_grad + compute_svd_grad are one kernel; gradients of state t are written once (ping-pong grad slots)
//   * S substeps (or their reverse) are captured into one CUDA graph; poses for all substeps are device resident
#define DD_FLOAT_LENGTH 1
#include "mpm_math.cuh"
#include "../../include/dexdeform_mpm.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <functional>
#include <map>
#include <string>
#include <tuple>
#include <vector>

using namespace dd;

namespace {

thread_local std::string g_last_error;
int fail(const std::string &msg) { g_last_error = msg; return 1; }
#define DD_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(e_));        \
  } while (0)

constexpr int kT = 256;
// occupancy knobs (min resident blocks per SM) -- tuned with ncu, see profiles/
#ifndef DD_LB_P2G_TILE
#define DD_LB_P2G_TILE 4
#endif
#ifndef DD_LB_G2PG_TILE
#define DD_LB_G2PG_TILE 3
#endif
#ifndef DD_LB_G2PG_SCATTER
#define DD_LB_G2PG_SCATTER 6
#endif
#ifndef DD_LB_G2P_TILE
#define DD_LB_G2P_TILE 5
#endif
#ifndef DD_LB_P2GG_TILE
#define DD_LB_P2GG_TILE 3
#endif
#ifndef DD_LB_P2G_GRAD
#define DD_LB_P2G_GRAD 2
#endif
#ifndef DD_LB_GRID_GRAD
#define DD_LB_GRID_GRAD 3  // blocks of k_grid_grad_b per SM (80 registers; 4 blocks = 64 registers spill 96 B in the contact adjoint: 10k scene 49 -> 47 us, 64 x 10k 64 -> 58)
#endif
constexpr int kPlaneFloats = 45;  // 16 (x,v,C + pad) + 9 (F) + 4 (quaternion of V) + 16 (constitutive checkpoint: affine, sigma, quaternion of U)

struct KP {            // kernel parameters shared by all kernels
  int E, N, EN, nb, G, gx, gy, gz;
  float dx, inv_dx, dt, gf, gh;
  float g0, g1, g2;    // gravity (already scaled as the caller wants it)
};

struct SegView {       // device-resident description of one particle ordering (a "segment" of the rollout, see dd_sim)
  const int4 *chunks;  // (brick, start, count, last-row lane mask), largest first
  int *cnt;            // [0] chunks  [1] active bricks  [2] occupied bricks
  int *active;         // active brick list; grows when a particle's stencil reaches a new brick (activate_bricks)
  int *flags;          // per brick: 0 inactive, 1 active, 2 being activated
};

// ---- particle planes ----------------------------------------------------------------------------------------
// slot layout (floats): [0,4EN) P0=(x.x,x.y,x.z,v.x)  [4EN,8EN) P1=(v.y,v.z,C00,C01)  [8EN,12EN) P2=(C02,C10,C11,C12)
//                       [12EN,16EN) P3=(C20,C21,C22,0) [16EN,20EN) F0=(F00..F10) [20EN,24EN) F1=(F11..F21) [24EN,25EN) F2=F22
//                       [25EN,29EN) Q=(qx,qy,qz,qw): quaternion of V of the SVD that produced this slot's F (warm start / replay)
//                       [29EN,45EN) constitutive checkpoint of the substep that produced this slot, for the adjoint:
//                                   A0=(a00,a01,a02,a10) A1=(a11,a12,a20,a21) A2=(a22,sig.x,sig.y,sig.z) A3=quaternion of U; a = affine matrix of p2g
struct XVC { V3 x, v; M3 C; };
DD_DEV const float4 *plane4(const float *slot, int EN, int k) { return reinterpret_cast<const float4 *>(slot + (size_t)4 * k * EN); }
DD_DEV float4 *plane4(float *slot, int EN, int k) { return reinterpret_cast<float4 *>(slot + (size_t)4 * k * EN); }
DD_DEV float4 ldg_stream(const float4 *p) {  // streaming 128-bit load that does not pollute L1
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
DD_DEV XVC load_xvc(const float *slot, int EN, int p) {
  float4 a = ldg_stream(plane4(slot, EN, 0) + p), b = ldg_stream(plane4(slot, EN, 1) + p), c = ldg_stream(plane4(slot, EN, 2) + p),
         d = ldg_stream(plane4(slot, EN, 3) + p);
  XVC r;
  r.x = v3(a.x, a.y, a.z);
  r.v = v3(a.w, b.x, b.y);
  r.C = m3(b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z);
  return r;
}
DD_DEV void store_xvc(float *slot, int EN, int p, V3 x, V3 v, const M3 &C) {
  plane4(slot, EN, 0)[p] = make_float4(x.x, x.y, x.z, v.x);
  plane4(slot, EN, 1)[p] = make_float4(v.y, v.z, C.a00, C.a01);
  plane4(slot, EN, 2)[p] = make_float4(C.a02, C.a10, C.a11, C.a12);
  plane4(slot, EN, 3)[p] = make_float4(C.a20, C.a21, C.a22, 0.f);
}
DD_DEV M3 load_F(const float *slot, int EN, int p) {
  float4 a = ldg_stream(plane4(slot, EN, 4) + p), b = ldg_stream(plane4(slot, EN, 5) + p);
  float c = __ldg(slot + (size_t)24 * EN + p);
  return m3(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c);
}
DD_DEV void store_F(float *slot, int EN, int p, const M3 &F) {
  plane4(slot, EN, 4)[p] = make_float4(F.a00, F.a01, F.a02, F.a10);
  plane4(slot, EN, 5)[p] = make_float4(F.a11, F.a12, F.a20, F.a21);
  slot[(size_t)24 * EN + p] = F.a22;
}
DD_DEV float4 load_q(const float *slot, int EN, int p) { return ldg_stream(reinterpret_cast<const float4 *>(slot + (size_t)25 * EN) + p); }
DD_DEV void store_q(float *slot, int EN, int p, float4 q) { reinterpret_cast<float4 *>(slot + (size_t)25 * EN)[p] = q; }
DD_DEV const float4 *aux4(const float *slot, int EN, int k) { return reinterpret_cast<const float4 *>(slot + (size_t)(29 + 4 * k) * EN); }
DD_DEV float4 *aux4(float *slot, int EN, int k) { return reinterpret_cast<float4 *>(slot + (size_t)(29 + 4 * k) * EN); }
// shared-memory access that the compiler may neither reorder nor merge (the tile updates of one warp rely on program order).
// Volatile asm statements keep their mutual order; plain accesses to the tiles (fill, flush) are fenced by __syncwarp().
#ifdef DD_TILE_MEMCLOBBER
#define DD_TILE_CLOBBER : "memory"
#else
#define DD_TILE_CLOBBER
#endif
DD_DEV unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// (only the lanes that will store load: a lane that shares its cell with a lower lane of the row, or shadows an idle slot, would
// otherwise read a node another lane is updating in the same step -- harmless, its value is discarded, but a reported hazard)
DD_DEV float4 lds_v4_if(unsigned a, bool pred) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %5, 0; @q ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4]; }" : "+f"(r.x), "+f"(r.y), "+f"(r.z), "+f"(r.w) : "r"(a), "r"((int)pred) DD_TILE_CLOBBER);
  return r;
}
// Same, into a register quadruple the caller carries from update to update: lanes whose predicate is false keep whatever they
// had (finite garbage that is never stored), so no zero-initialisation is issued per update (three instructions of sixteen).
DD_DEV void lds_v4_into(float4 &r, unsigned a, bool pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %5, 0; @q ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4]; }" : "+f"(r.x), "+f"(r.y), "+f"(r.z), "+f"(r.w) : "r"(a), "r"((int)pred) DD_TILE_CLOBBER);
}
DD_DEV void sts_v4_if(unsigned a, float4 v, bool pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %5, 0; @q st.volatile.shared.v4.f32 [%0], {%1,%2,%3,%4}; }" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)pred) DD_TILE_CLOBBER);
#ifndef DD_TILE_NO_SYNCWARP
  // Orders the read-modify-write of one stencil offset before the next one for ALL lanes in the sense of the CUDA memory model
  // (lanes of a row alias each other's nodes at different offsets).  Without it the code relies on what the hardware does for a
  // converged warp executing straight-line code (shared-memory accesses of one warp are performed in program order); measured
  // cost of the barrier: none (config D 316 vs 317 us per substep pair), racecheck: profiles/r02_sanitizer.md.
  __syncwarp();
#endif
}
// Tile slot of node (tx,ty,tz) in an 8^3 tile.  A 128-bit shared-memory access is served one quarter-warp (8 lanes) at a
// time and is conflict-free when those 8 lanes hit 8 different 16-byte bank groups; the group of a node is the low three
// bits of its slot, (tz + 4 ty + 2 tx) mod 8.  The map is ADDITIVE, so moving every lane by the same stencil offset
// (i,j,k) rotates all groups by the same amount: lanes whose home cells lie in different groups stay conflict-free for
// all 27 nodes.  The particle order (k_sort_keys / k_interleave) puts group g's particles into lanes g, g+8, g+16, g+24.
DD_DEV int tile_group(int tx, int ty, int tz) { return (tz + 4 * ty + 2 * tx) & 7; }
DD_DEV int tile_slot(int tx, int ty, int tz) { return tx << 6 | ty << 3 | tile_group(tx, ty, tz); }
DD_DEV void red_add_v4(float4 *addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// programmatic dependent launch: let the next kernel on the stream be scheduled / wait for the previous one to complete
DD_DEV void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
DD_DEV void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
DD_DEV int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
// a + n b for a stencil offset n in {0,1,2} known after unrolling: nothing, one add or one fma per component (a plain a + b * 0.f
// costs an fma per component, which IEEE rules keep the compiler from dropping)
DD_DEV V3 step_n(V3 a, V3 b, int n) { return n == 0 ? a : n == 1 ? a + b : v3(fmaf(b.x, 2.f, a.x), fmaf(b.y, 2.f, a.y), fmaf(b.z, 2.f, a.z)); }

// stencil with the base clamped into the grid (identical to the reference for every in-domain particle)
DD_DEV Stencil make_stencil_safe(V3 x, const KP &kp) {
  Stencil s = make_stencil(x, kp.inv_dx);
  int bx = clampi(s.bx, 0, kp.gx - 3), by = clampi(s.by, 0, kp.gy - 3), bz = clampi(s.bz, 0, kp.gz - 3);
  if (bx != s.bx || by != s.by || bz != s.bz) {  // out-of-domain input: keep memory safe, weights follow the clamped base
    s.bx = bx; s.by = by; s.bz = bz;
  }
  return s;
}

// Constitutive update shared by forward and adjoint: F~ = (I + dt C) F, SVD, return mapping, affine matrix.
struct Constit {
  M3 Ft, U, Vm, nF, r, affine;
  V3 sigma;
  Plastic pl;
  float J, scale;
};
// q: warm-start quaternion in, converged quaternion out (svd_mode 1); max_sweeps = 0 replays a stored factorisation.
// qu: quaternion of U (svd_mode 1 only)
template <int SVD>
DD_DEV void constitutive(const XVC &s, const M3 &F, float4 m0, float yield, const KP &kp, Constit &c, float4 &q, int max_sweeps, float4 *qu = nullptr) {
  c.Ft = mul(mdiag(1.f) + s.C * kp.dt, F);
  if (SVD == 0) svd3_f64(c.Ft, c.U, c.sigma, c.Vm);
  else svd3_warm<false>(c.Ft, q.x, q.y, q.z, q.w, c.U, c.sigma, c.Vm, max_sweeps, reinterpret_cast<float *>(qu));
  c.J = von_mises<SVD == 1>(c.Ft, c.U, c.sigma, c.Vm, yield, m0.z, c.nF, c.pl);
  c.r = mul_nt(c.U, c.Vm);
  c.scale = -kp.dt * m0.y * 4.f * kp.inv_dx * kp.inv_dx;
  c.affine = c.scale * fixed_corotated(c.nF, c.r, c.J, m0.z, m0.w) + m0.x * s.C;
}
// what the forward pass leaves for the adjoint in the next slot (svd_mode 1): no SVD, QR or stress evaluation in the backward pass
DD_DEV void store_constit(float *nxt, int EN, int p, const Constit &c, float4 qu) {
  aux4(nxt, EN, 0)[p] = make_float4(c.affine.a00, c.affine.a01, c.affine.a02, c.affine.a10);
  aux4(nxt, EN, 1)[p] = make_float4(c.affine.a11, c.affine.a12, c.affine.a20, c.affine.a21);
  aux4(nxt, EN, 2)[p] = make_float4(c.affine.a22, c.sigma.x, c.sigma.y, c.sigma.z);
  aux4(nxt, EN, 3)[p] = qu;
}

// ---- forward kernels -------------------------------------------------------------------------------------------
// compute_svd + p2g (integrator.cu:84-100, 313-394) fused
template <int SVD, bool WRITE_F>
__global__ void __launch_bounds__(kT) k_p2g(KP kp, const float *__restrict__ cur, float *__restrict__ nxt, const float4 *__restrict__ mat0,
                                            const float *__restrict__ yield, float4 *__restrict__ grid) {
  int p = blockIdx.x * kT + threadIdx.x;
  if (p >= kp.EN) return;
  XVC s = load_xvc(cur, kp.EN, p);
  M3 F = load_F(cur, kp.EN, p);
  float4 m0 = __ldg(mat0 + p);
  Constit c;
  float4 q = load_q(cur, kp.EN, p), qu = make_float4(0.f, 0.f, 0.f, 1.f);
  constitutive<SVD>(s, F, m0, __ldg(yield + p), kp, c, q, 6, &qu);
  if (WRITE_F) { store_F(nxt, kp.EN, p, c.nF); store_q(nxt, kp.EN, p, q); store_constit(nxt, kp.EN, p, c, qu); }
  Stencil st = make_stencil_safe(s.x, kp);
  float m = m0.x;
  V3 mv = m * s.v;
  // affine * dpos is separable: dpos = (offset - fx) * dx
  V3 c0 = v3(c.affine.a00, c.affine.a10, c.affine.a20), c1 = v3(c.affine.a01, c.affine.a11, c.affine.a21), c2 = v3(c.affine.a02, c.affine.a12, c.affine.a22);
  float4 *g = grid + (size_t)(p / kp.N) * kp.G;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float wi = pick(st.w0, st.w1, st.w2, i, 0);
    V3 ai = mv + c0 * (((float)i - st.fx.x) * kp.dx);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float wij = wi * pick(st.w0, st.w1, st.w2, j, 1);
      V3 aij = ai + c1 * (((float)j - st.fx.y) * kp.dx);
      int row = ((st.bx + i) * kp.gy + st.by + j) * kp.gz + st.bz;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float w = wij * pick(st.w0, st.w1, st.w2, k, 2);
        V3 a = (aij + c2 * (((float)k - st.fx.z) * kp.dx)) * w;
        red_add_v4(g + row + k, a.x, a.y, a.z, m * w);
      }
    }
  }
}

struct BodyTables {  // device pointers; poses are per (slot, env, body), shapes shared by all envs
  const float4 *pos, *rot, *npos, *nrot;  // already offset to the slot; index env*nb + b
  const float4 *tfsr, *args;
  const float *cull;                      // conservative activation radius per body
};
DD_DEV Q4 q4f(float4 t) { Q4 q; q.w = t.x; q.x = t.y; q.y = t.z; q.z = t.w; return q; }
DD_DEV V3 v3f(float4 t) { return v3(t.x, t.y, t.z); }

struct Hit {
  V3 gxb, un, rn, nrm, bv, rel, vt_in, vt;
  float dist, infl, nc, vtn;
  bool has_fric;
};
// geometric part of the contact test (velocity independent), integrator.cu:705-710
DD_DEV bool contact_geom(V3 gx, V3 bx, Q4 bq, Q4 tfsr, Q4 sargs, float cull, Hit &h) {
  V3 d = gx - bx;
  if (dot(d, d) > cull * cull) return false;  // cannot be within the influence band
  h.gxb = qrot(qconj(bq), d);
  h.dist = shape_sdf(tfsr, sargs, h.gxb);
  return contact_active(h.dist, tfsr.y, h.infl);
}
// velocity part, integrator.cu:714-729
DD_DEV V3 contact_apply(V3 gx, V3 v, Q4 bq, V3 npos, Q4 nrot, Q4 tfsr, Q4 sargs, float dt, Hit &h) {
  h.un = shape_grad(tfsr, sargs, h.gxb);
  h.rn = normalized(h.un);
  h.nrm = qrot(bq, h.rn);
  h.bv = (xform(npos, nrot, h.gxb) - gx) / dt;
  h.rel = v - h.bv;
  h.nc = dot(h.rel, h.nrm);
  h.vt_in = h.rel - fminf(h.nc, 0.f) * h.nrm;
  h.has_fric = h.nc < 0.f && dot(h.vt_in, h.vt_in) > 1e-30f;
  h.vtn = length30(h.vt_in);
  h.vt = h.vt_in;
  if (h.has_fric) h.vt = h.vt_in * (1.f / h.vtn) * fmaxf(0.f, h.vtn + h.nc * tfsr.x);
  return h.bv + h.rel * (1 - h.infl) + h.vt * h.infl;
}
// bodies among `cand` whose influence band contains the node (the contact test of integrator.cu:705-710 for each of them)
DD_DEV unsigned long long contact_mask(V3 gx, int env, int nb, unsigned long long cand, const BodyTables &bt) {
  unsigned long long mask = 0ull;
  for (unsigned long long c = cand; c; c &= c - 1ull) {
    int b = __ffsll((long long)c) - 1, pb = env * nb + b;
    Hit h;
    if (contact_geom(gx, v3f(bt.pos[pb]), q4f(bt.rot[pb]), q4f(bt.tfsr[b]), q4f(bt.args[b]), bt.cull[b], h)) mask |= 1ull << b;
  }
  return mask;
}
DD_DEV V3 apply_bc(V3 v, int gx_, int gy_, int gz_, const KP &kp) {  // integrator.cu:734-774
  const int bound = 3;
  if (gx_ < bound && v.x < 0) v.x = 0;
  if (gx_ > kp.gx - bound && v.x > 0) v.x = 0;
  if (gy_ < bound && v.y < 0) {
    if (kp.gf > 0.f) {
      if (kp.gf < 99.f) {
        float lin = v.y;
        V3 vit = v3(v.x, 0.f, v.z);
        float lit = sqrtf(dot(vit, vit) + 1e-8f);
        v = vit * fmaxf(1.f + kp.gf * lin / lit, 0.f);
      } else {
        v = vzero();
      }
    }
    v.y = 0;
  }
  if (gy_ > kp.gy - bound && v.y > 0) v.y = 0;
  if (gz_ < bound && v.z < 0) v.z = 0;
  if (gz_ > kp.gz - bound && v.z > 0) v.z = 0;
  return v;
}

// grid_op_v2 (integrator.cu:647-777) on float4 nodes; dense over E*G nodes
__global__ void __launch_bounds__(kT) k_grid(KP kp, const float4 *__restrict__ grid, float4 *__restrict__ grid_v, BodyTables bt) {
  int node = blockIdx.x * kT + threadIdx.x;
  if (node >= kp.E * kp.G) return;
  float4 mm = grid[node];
  if (!(mm.w > 1e-12)) {
    grid_v[node] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  int env = node / kp.G, cell = node - env * kp.G;
  int gx_ = cell / kp.gz / kp.gy, gy_ = (cell / kp.gz) % kp.gy, gz_ = cell % kp.gz;
  V3 v = v3(mm.x, mm.y, mm.z) * (1.f / mm.w) + kp.dt * v3(kp.g0, kp.g1, kp.g2);
  V3 gx = v3((float)gx_, (float)gy_, (float)gz_) * kp.dx;
  for (int b = 0; b < kp.nb; ++b) {
    int pb = env * kp.nb + b;
    Hit h;
    Q4 bq = q4f(bt.rot[pb]), tfsr = q4f(bt.tfsr[b]), sargs = q4f(bt.args[b]);
    if (contact_geom(gx, v3f(bt.pos[pb]), bq, tfsr, sargs, bt.cull[b], h))
      v = contact_apply(gx, v, bq, v3f(bt.npos[pb]), q4f(bt.nrot[pb]), tfsr, sargs, kp.dt, h);
  }
  v = apply_bc(v, gx_, gy_, gz_, kp);
  grid_v[node] = make_float4(v.x, v.y, v.z, 0.f);
}

// APIC gather of g2p in separable form: per (i,j) row A = sum_k wz_k v_k and B = sum_k k wz_k v_k, then
//   v' = sum w_ij A,   sum N k v = sum w_ij B,   sum N i v = sum (i w_ij) A,   sum N j v = sum (j w_ij) A
// and C' = (4/dx) (those moments - v' (x) fx): ~9 instead of 16 floating-point instructions per node.
template <class Fetch>
DD_DEV void g2p_gather(Fetch fetch, const float (&wx)[3], const float (&wy)[3], const float (&wz)[3], V3 fx, float s4, V3 &nv, M3 &nC) {
  // components (x, y) of every sum travel as one packed pair (FFMA2 with the weight broadcast), z as a scalar
  float2 nvp = pk(0.f, 0.f), Cxp = nvp, Cyp = nvp, Czp = nvp;
  float nvz = 0.f, Cxz = 0.f, Cyz = 0.f, Czz = 0.f;
  const float kz1 = wz[1], kz2 = 2.f * wz[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float wij = wx[i] * wy[j];
      float4 t0 = fetch(i, j, 0), t1 = fetch(i, j, 1), t2 = fetch(i, j, 2);
      float2 Ap = fma2(wz[2], pk(t2.x, t2.y), fma2(wz[1], pk(t1.x, t1.y), mul2(wz[0], pk(t0.x, t0.y))));
      float Az = fmaf(wz[2], t2.z, fmaf(wz[1], t1.z, wz[0] * t0.z));
      float2 Bp = fma2(kz2, pk(t2.x, t2.y), mul2(kz1, pk(t1.x, t1.y)));
      float Bz = fmaf(kz2, t2.z, kz1 * t1.z);
      nvp = fma2(wij, Ap, nvp); nvz = fmaf(wij, Az, nvz);
      Czp = fma2(wij, Bp, Czp); Czz = fmaf(wij, Bz, Czz);
      if (i > 0) { float wi = wij * (float)i; Cxp = fma2(wi, Ap, Cxp); Cxz = fmaf(wi, Az, Cxz); }
      if (j > 0) { float wj = wij * (float)j; Cyp = fma2(wj, Ap, Cyp); Cyz = fmaf(wj, Az, Cyz); }
    }
  }
  nv = v3(nvp.x, nvp.y, nvz);
  V3 Cx = v3(Cxp.x, Cxp.y, Cxz), Cy = v3(Cyp.x, Cyp.y, Cyz), Cz = v3(Czp.x, Czp.y, Czz);
  V3 c0 = (Cx - nv * fx.x) * s4, c1 = (Cy - nv * fx.y) * s4, c2 = (Cz - nv * fx.z) * s4;
  nC = m3(c0.x, c1.x, c2.x, c0.y, c1.y, c2.y, c0.z, c1.z, c2.z);
}

// g2p (integrator.cu:1059-1109).  v' = sum w v_n ; C' = 4/dx * sum (w v_n) (x) (offset - fx)
__global__ void __launch_bounds__(kT) k_g2p(KP kp, const int *__restrict__ spos, const float *__restrict__ cur, float *__restrict__ nxt,
                                            const float4 *__restrict__ grid_v) {
  int p = blockIdx.x * kT + threadIdx.x;
  if (p >= kp.EN) return;
  if (spos) p = __ldg(spos + p);  // neighbouring threads take particles of the same / adjacent cells: the 27-node gathers hit L1
  float4 a = ldg_stream(plane4(cur, kp.EN, 0) + p);
  V3 x = v3(a.x, a.y, a.z);
  Stencil st = make_stencil_safe(x, kp);
  const float4 *g = grid_v + (size_t)(p / kp.N) * kp.G + (st.bx * kp.gy + st.by) * kp.gz + st.bz;
  float wx[3] = {st.w0.x, st.w1.x, st.w2.x}, wy[3] = {st.w0.y, st.w1.y, st.w2.y}, wz[3] = {st.w0.z, st.w1.z, st.w2.z};
  V3 nv;
  M3 nC;
  g2p_gather([&](int i, int j, int k) { return __ldg(g + (i * kp.gy + j) * kp.gz + k); }, wx, wy, wz, st.fx, kp.inv_dx * 4.f, nv, nC);
  V3 hi = v3(((float)kp.gx - 3.f) * kp.dx, ((float)kp.gy - 3.f) * kp.dx, ((float)kp.gz - 3.f) * kp.dx);
  float lo = kp.gh * kp.dx;
  V3 t = x + nv * kp.dt;
  store_xvc(nxt, kp.EN, p, v3(fmaxf(fminf(t.x, hi.x), lo), fmaxf(fminf(t.y, hi.y), lo), fmaxf(fminf(t.z, hi.z), lo)), nv, nC);
}

// ---- adjoint kernels -------------------------------------------------------------------------------------------
// g2p_grad (integrator.cu:1527-1614).  gin = gradients of state t+1, gout = gradients of state t (partial gx written)
__global__ void __launch_bounds__(kT) k_g2p_grad(KP kp, const float *__restrict__ cur, const float *__restrict__ nxt,
                                                 const float4 *__restrict__ grid_v, const float *__restrict__ gin, float *__restrict__ gout,
                                                 float4 *__restrict__ ggrid_v) {
  int p = blockIdx.x * kT + threadIdx.x;
  if (p >= kp.EN) return;
  float4 a = ldg_stream(plane4(cur, kp.EN, 0) + p);
  V3 x = v3(a.x, a.y, a.z);
  float4 n0 = ldg_stream(plane4(nxt, kp.EN, 0) + p), n1 = ldg_stream(plane4(nxt, kp.EN, 1) + p);
  V3 nvel = v3(n0.w, n1.x, n1.y);
  XVC g = load_xvc(gin, kp.EN, p);  // (gx', gv', gC')
  V3 gx = g.x, gnv = g.v;
  V3 nx = x + nvel * kp.dt;
  V3 hi = v3(((float)kp.gx - 3.f) * kp.dx, ((float)kp.gy - 3.f) * kp.dx, ((float)kp.gz - 3.f) * kp.dx);
  float lo = kp.gh * kp.dx;
  if (nx.x > hi.x || nx.x < lo) gx.x = 0;
  if (nx.y > hi.y || nx.y < lo) gx.y = 0;
  if (nx.z > hi.z || nx.z < lo) gx.z = 0;
  gnv += gx * kp.dt;
  Stencil st = make_stencil_safe(x, kp);
  V3 d0, d1, d2;
  stencil_dw(st, kp.inv_dx, d0, d1, d2);
  size_t goff = (size_t)(p / kp.N) * kp.G;
  float s4 = kp.inv_dx * 4.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      int row = ((st.bx + i) * kp.gy + st.by + j) * kp.gz + st.bz;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float wx = pick(st.w0, st.w1, st.w2, i, 0), wy = pick(st.w0, st.w1, st.w2, j, 1), wz = pick(st.w0, st.w1, st.w2, k, 2);
        float w = wx * wy * wz;
        V3 dpos = v3((float)i, (float)j, (float)k) - st.fx;
        float4 t = __ldg(grid_v + goff + row + k);
        V3 v = v3(t.x, t.y, t.z);
        float xx = w * s4;
        V3 cd = mul(g.C, dpos);
        V3 ggv = w * gnv + cd * xx;
        red_add_v4(ggrid_v + goff + row + k, ggv.x, ggv.y, ggv.z, 0.f);
        gx += (-kp.inv_dx * xx) * mul_t(g.C, v);
        float gw = dot(gnv, v) + s4 * dot(v, cd);
        gx += v3(pick(d0, d1, d2, i, 0) * wy * wz, wx * pick(d0, d1, d2, j, 1) * wz, wx * wy * pick(d0, d1, d2, k, 2)) * gw;
      }
    }
  }
  plane4(gout, kp.EN, 0)[p] = make_float4(gx.x, gx.y, gx.z, 0.f);
}

// grid_op_v2_grad (integrator.cu:779-1057) without the stored per-body velocities
// cand: bodies that can touch this node's brick at all (superset of the contact mask); zero_gv / zero_m: clear the node's
// entry of ggrid_v / grid after it has been consumed, so the next substep finds zeros without a separate memset pass
// mm_ck: where this node's (mv, m) of the substep is stored when it does not come from the dense grid (brick checkpoints), else null
DD_DEV void grid_grad_body(const KP &kp, int node, bool inr, int env_, int cx, int cy, int cz, float4 *__restrict__ grid,
                           float4 *__restrict__ ggrid_v, float4 *__restrict__ ggrid, const BodyTables &bt, float4 *gpos, float4 *grot,
                           float4 *gnpos, float4 *gnrot, unsigned long long cand, bool zero_gv, bool zero_m, const float4 *mm_ck = nullptr, bool use_ck = false) {
  float4 mm = !inr ? make_float4(0.f, 0.f, 0.f, 0.f) : !use_ck ? grid[node] : mm_ck ? *mm_ck : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 gvn = inr ? ggrid_v[node] : make_float4(0.f, 0.f, 0.f, 0.f);  // both node loads in flight together (cold HBM in the adjoint sweep)
  if (inr && zero_m) grid[node] = make_float4(0.f, 0.f, 0.f, 0.f);
  bool live = inr && mm.w > 1e-12;
  // a whole warp of empty nodes leaves early (the common case)
  if (!__any_sync(0xffffffffu, live)) {
    if (inr) {
      ggrid[node] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (zero_gv) ggrid_v[node] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  int env = 0, gx_ = 0, gy_ = 0, gz_ = 0;
  V3 gv = vzero(), mv = vzero(), gx = vzero(), v0 = vzero();
  unsigned long long mask = 0ull;
#ifndef DD_STAGE_STACK
#define DD_STAGE_STACK 12  // (6 -> 12: fewer forward replays where nodes touch many bodies; 10k-particle scene 57 -> 49 us, config D unchanged)
#endif
  constexpr int kStageStack = DD_STAGE_STACK;
#ifdef DD_GRID_UNION_FWD
  V3 stage_in[kStageStack];
#else
  Hit kept[kStageStack];  // the first contact stages of this node as the forward replay evaluated them (local memory; only contact nodes touch it)
#endif
  if (live) {
    env = env_; gx_ = cx; gy_ = cy; gz_ = cz;
    mv = v3(mm.x, mm.y, mm.z);
    v0 = mv * (1.f / mm.w) + kp.dt * v3(kp.g0, kp.g1, kp.g2);
    gx = v3((float)gx_, (float)gy_, (float)gz_) * kp.dx;
    // forward replay: contact mask + velocity after all bodies
    V3 v = v0;
    int nc = 0;
#ifdef DD_GRID_UNION_FWD
    for (unsigned long long c = cand; c; c &= c - 1ull) {
      int b = __ffsll((long long)c) - 1, pb = env * kp.nb + b;
      Hit h;
      Q4 bq = q4f(bt.rot[pb]), tfsr = q4f(bt.tfsr[b]), sargs = q4f(bt.args[b]);
      if (contact_geom(gx, v3f(bt.pos[pb]), bq, tfsr, sargs, bt.cull[b], h)) {
        mask |= 1ull << b;
        if (nc < kStageStack) stage_in[nc] = v;  // input velocity of this contact stage, for the reverse sweep
        ++nc;
        v = contact_apply(gx, v, bq, v3f(bt.npos[pb]), q4f(bt.nrot[pb]), tfsr, sargs, kp.dt, h);
      }
    }
#else
    mask = contact_mask(gx, env, kp.nb, cand, bt);  // two passes: the mask over the brick's candidates (geometry only, every lane in step), then the node's own stages in index order
    for (unsigned long long m = mask; m; m &= m - 1ull) {
      int b = __ffsll((long long)m) - 1, pb = env * kp.nb + b;
      Hit h;
      Q4 bq = q4f(bt.rot[pb]), tfsr = q4f(bt.tfsr[b]), sargs = q4f(bt.args[b]);
      contact_geom(gx, v3f(bt.pos[pb]), bq, tfsr, sargs, bt.cull[b], h);
      v = contact_apply(gx, v, bq, v3f(bt.npos[pb]), q4f(bt.nrot[pb]), tfsr, sargs, kp.dt, h);
      if (nc < kStageStack) kept[nc] = h;  // the whole stage record: the reverse sweep reads it back instead of re-deriving geometry and velocity stage
      ++nc;
    }
#endif
    V3 vv = v;
    float4 t = gvn;
    if (zero_gv) ggrid_v[node] = make_float4(0.f, 0.f, 0.f, 0.f);
    gv = v3(t.x, t.y, t.z);
    // boundary-condition adjoint (integrator.cu:829-893)
    V3 vin = vv;
    const int bound = 3;
    if (gx_ > kp.gx - bound && vv.x > 0) vin.x = 0;
    if (gx_ < bound && vv.x < 0) vin.x = 0;
    float lin = 0.f, lit = 1.f;
    V3 vit = vzero();
    bool hit_ground = gy_ < bound && vin.y < 0;
    if (hit_ground) {
      lin = vin.y;
      vit = v3(vin.x, 0.f, vin.z);
      lit = sqrtf(dot(vit, vit) + 1e-8f);
      float flag = 1.f + kp.gf * lin / lit;
      vin = vit * fmaxf(flag, 0.f);
    }
    if (gz_ > kp.gz - bound && vin.z > 0) gv.z = 0;
    if (gz_ < bound && vin.z < 0) gv.z = 0;
    if (gy_ > kp.gy - bound && vin.y > 0) gv.y = 0;
    if (hit_ground) {
      gv.y = 0;
      float flag = 1.f + kp.gf * lin / lit;
      if (flag >= 0.f) {
        V3 g_vit = flag * gv;
        float g_lin = kp.gf / lit * dot(vit, gv);
        float g_lit = -kp.gf * lin / lit / lit * dot(vit, gv);
        g_vit += g_lit * (vit / lit);
        gv = v3(g_vit.x, g_lin, g_vit.z);
      } else {
        gv = vzero();
      }
    }
    if (gx_ > kp.gx - bound && vv.x > 0) gv.x = 0;
    if (gx_ < bound && vv.x < 0) gv.x = 0;
  }
  // Bodies in reverse.  Every lane walks ITS OWN contact list (highest body first), so one trip of the loop runs the stage adjoint
  // of up to 32 different (node, body) pairs: the trip count is the longest list of a node (1-3), not the size of the union of the
  // warp's bodies (up to nb where a hand closes over a few nodes).  The pose gradients of a trip are then reduced once per distinct
  // body of that trip; all lanes take part in those warp reductions.
  const int lane_ = threadIdx.x & 31;
  const int wenv = __shfl_sync(0xffffffffu, env_, 0);  // a warp never straddles two environments (G is a multiple of 64)
#ifdef DD_UNION_LOOP  // (round-2 formulation, kept for A/B timing: one trip per body of the warp's union)
  unsigned long long wmask = mask;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmask |= __shfl_xor_sync(0xffffffffu, wmask, o);
  while (wmask) {
    const int b = 63 - __clzll((long long)wmask);
    wmask &= ~(1ull << b);
    const bool has = live && (mask >> b & 1ull);
#else
  unsigned long long rem = live ? mask : 0ull;
  while (__any_sync(0xffffffffu, rem != 0ull)) {
    const bool has = rem != 0ull;
    const int b = has ? 63 - __clzll((long long)rem) : -1;
    if (has) rem &= ~(1ull << b);
#endif
    V3 g_bx = vzero(), g_np = vzero();
    Q4 g_bq, g_nq;
    g_bq.w = g_bq.x = g_bq.y = g_bq.z = 0.f;
    g_nq = g_bq;
    if (has) {
      // input velocity of stage b: kept from the forward replay for the first kStageStack contacts of this node, otherwise
      // re-derived by replaying the contacting bodies before it
      V3 v = v0;
      unsigned long long lower = mask & ((1ull << b) - 1ull);
      int ci = __popcll(lower);
      int pb = env * kp.nb + b;
      V3 bx = v3f(bt.pos[pb]);
      Q4 bq = q4f(bt.rot[pb]), nrot = q4f(bt.nrot[pb]), tfsr = q4f(bt.tfsr[b]), sargs = q4f(bt.args[b]);
      Hit h;
#ifdef DD_GRID_UNION_FWD
      if (ci < kStageStack) { v = stage_in[ci]; lower = 0ull; }
#else
      if (ci < kStageStack) h = kept[ci];
      else
#endif
      {
        while (lower) {
          int c = __ffsll((long long)lower) - 1;
          lower &= lower - 1ull;
          int pc = env * kp.nb + c;
          Hit hc;
          Q4 cq = q4f(bt.rot[pc]), ct = q4f(bt.tfsr[c]), ca = q4f(bt.args[c]);
          contact_geom(gx, v3f(bt.pos[pc]), cq, ct, ca, bt.cull[c], hc);
          v = contact_apply(gx, v, cq, v3f(bt.npos[pc]), q4f(bt.nrot[pc]), ct, ca, kp.dt, hc);
        }
        contact_geom(gx, bx, bq, tfsr, sargs, bt.cull[b], h);
        contact_apply(gx, v, bq, v3f(bt.npos[pb]), nrot, tfsr, sargs, kp.dt, h);
      }
      float friction = tfsr.x, softness = tfsr.y;
      float g_nc = 0.f;
      V3 g_bv = gv, g_rel = gv * (1 - h.infl), g_vt = gv * h.infl;
      float g_infl = dot(h.vt - h.rel, gv);
      if (h.has_fric) {
        float bf = h.vtn + h.nc * friction;
        if (bf > 0.f) {
          g_nc += dot(h.vt_in, g_vt) * friction / h.vtn;
          float g_vtn = -h.nc * g_nc / h.vtn;
          g_vt = g_vt * (1.f / h.vtn) * bf + g_vtn * h.vt_in / h.vtn;
        } else {
          g_vt = vzero();
        }
      }
      V3 g_n = vzero();
      g_rel += g_vt;
      if (h.nc < 0.f) {
        g_nc += -dot(h.nrm, g_vt);
        g_n += (-h.nc) * g_vt;
      }
      g_rel += h.nrm * g_nc;
      g_n += h.rel * g_nc;
      gv = g_rel;
      g_bv = g_bv - g_rel;
      V3 g_gxb = vzero();
      xform_adj(nrot, h.gxb, g_bv * (1.f / kp.dt), g_np, g_nq, g_gxb);
      V3 g_rn = vzero();
      qrot_adj(bq, h.rn, g_n, g_bq, g_rn);
      g_gxb += shape_grad_adj(tfsr, sargs, h.gxb, normalized_adj(h.un, g_rn));
      float expdist = expf(-h.dist * softness);
      if (expdist <= 1) g_gxb += h.un * (-softness * expdist * g_infl);
      V3 g_tmp = vzero();
      xform_inv_adj(bx, bq, gx, g_gxb, g_bx, g_bq, g_tmp);
    }
    // Warp totals of the 14 pose-gradient components with a transposing butterfly: at every stage a lane hands half of its values
    // to its partner and keeps the other half, so 16 shuffles (8 + 4 + 2 + 1 + 1) replace 14 x 5; afterwards the even lane 2 i holds
    // the warp total of component i and adds it with ONE atomic (14 lanes in parallel instead of 14 atomics issued by lane 0).
    unsigned todo = __ballot_sync(0xffffffffu, has);
    while (todo) {
      const int bs = __shfl_sync(0xffffffffu, b, __ffs(todo) - 1);
      const bool mem = has && b == bs;
      todo &= ~__ballot_sync(0xffffffffu, mem);
      float r[16] = {g_np.x, g_np.y, g_np.z, g_nq.w, g_nq.x, g_nq.y, g_nq.z, g_bx.x, g_bx.y, g_bx.z, g_bq.w, g_bq.x, g_bq.y, g_bq.z, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 14; ++c) r[c] = mem ? r[c] : 0.f;
#pragma unroll
      for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
        const bool hi = (lane_ & off) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          float send = hi ? r[i] : r[i + half], keep = hi ? r[i + half] : r[i];
          r[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      r[0] += __shfl_xor_sync(0xffffffffu, r[0], 1);
#ifndef DD_NO_POSE_ATOMICS
      {
        const int comp = lane_ >> 1, pb = wenv * kp.nb + bs;
        if (!(lane_ & 1) && comp < 14) {
          float *dst = comp < 3 ? &gnpos[pb].x + comp : comp < 7 ? &gnrot[pb].x + (comp - 3) : comp < 10 ? &gpos[pb].x + (comp - 7) : &grot[pb].x + (comp - 10);
          atomicAdd(dst, r[0]);
        }
      }
#else
      if (r[0] == 12345.f) gnpos[0].x = r[0] + (float)(wenv + bs);
#endif
    }
  }
  if (inr) {
    if (live) {
      V3 o = gv * (1.f / mm.w);
      ggrid[node] = make_float4(o.x, o.y, o.z, (-1.f / mm.w / mm.w) * dot(mv, gv));
    } else {
      ggrid[node] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (zero_gv) ggrid_v[node] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
__global__ void __launch_bounds__(kT) k_grid_grad(KP kp, float4 *__restrict__ grid, float4 *__restrict__ ggrid_v,
                                                  float4 *__restrict__ ggrid, BodyTables bt, float4 *gpos, float4 *grot, float4 *gnpos,
                                                  float4 *gnrot) {
  int node = blockIdx.x * kT + threadIdx.x;
  bool inr = node < kp.E * kp.G;
  int env = inr ? node / kp.G : 0, cell = node - env * kp.G;
  grid_grad_body(kp, node, inr, env, cell / kp.gz / kp.gy, (cell / kp.gz) % kp.gy, cell % kp.gz, grid, ggrid_v, ggrid, bt, gpos, grot, gnpos, gnrot,
                 kp.nb >= 64 ? ~0ull : (1ull << kp.nb) - 1ull, false, false);
}

// p2g_grad + compute_svd_grad (integrator.cu:396-627, 110-186) fused; writes the complete gradient of state t.
// The gather over the 27 nodes is regrouped so that every node costs ~30 FP instructions:
//   T  = sum N (g_mv (x) dpos)          -> dL/dstress = scale T, dL/dC += m T
//   Sv = sum N g_mv                     -> dL/dv = m Sv, and the -N A^T g_mv term of dL/dx is -A^T Sv
//   dL/dx += sum gradN (m g_m + g_mv . (m v + A dpos))
constexpr int kTileN = 512;          // 8^3 nodes
constexpr int kTileWarps = 4;        // chunks per thread block (launch-bound hint; the launch picks the real number)
constexpr int kStageP2G = 9;         // staged float4 slots per lane: p2g_tile
constexpr int kStageG2PG = 7;        // g2p_grad_tile
constexpr int kStageP2GG = 16;       // p2g_grad_tile (slots listed at P2ggStager)
constexpr int kQueueF4 = 8;          // deferred-lane queue: 32 ints per warp (DeferQueue)
constexpr size_t kSmemP2G = (kTileN + kStageP2G * 32 + kQueueF4) * sizeof(float4);        // bytes per warp
constexpr size_t kSmemG2PG = (2 * kTileN + kStageG2PG * 32 + kQueueF4) * sizeof(float4);
constexpr size_t kSmemG2P = (kTileN + 32) * sizeof(float4);  // tile + one staged position per lane
constexpr size_t kSmemP2GG = (kTileN + kStageP2GG * 32) * sizeof(float4);                 // fp32-SVD variant; the fp64 variant does not stage
// a chunk is stored as R rows of 32 particles (the last row holds the remaining `last`); row j starts at start + 32 j
struct ChunkGeom { int env, ox, oy, oz, start, cnt, R, last, bx, by, bz; unsigned lastmask; };
// lanes (columns) of the short last row are given by a bit mask chosen at sort time (the columns whose bank group has particles
// left over after R - 1 full rows), not by a prefix: its particles are stored compacted in lane order
DD_DEV bool lane_on(const ChunkGeom &c, int j, int lane) { return j < c.R - 1 || (c.lastmask >> lane & 1u); }
DD_DEV int row_pos(const ChunkGeom &c, int j, int lane) {  // storage position of (row j, lane), or the chunk's first particle for an idle lane
  if (j < c.R - 1) return c.start + 32 * j + lane;
  return (c.lastmask >> lane & 1u) ? c.start + 32 * j + __popc(c.lastmask & ((1u << lane) - 1u)) : c.start;
}
DD_DEV ChunkGeom chunk_geom(int4 ch, const KP &kp) {
  ChunkGeom c;
  int nby = kp.gy >> 2, nbz = kp.gz >> 2, NB = (kp.gx >> 2) * nby * nbz;
  c.env = ch.x / NB;
  int b = ch.x - c.env * NB;
  c.bx = b / (nby * nbz); c.by = (b / nbz) % nby; c.bz = b % nbz;
  c.ox = c.bx * 4 - 1; c.oy = c.by * 4 - 1; c.oz = c.bz * 4 - 1;
  c.start = ch.y; c.cnt = ch.z;
  c.R = (c.cnt + 31) >> 5;
  c.last = c.cnt - 32 * (c.R - 1);
  c.lastmask = (unsigned)ch.w;
  return c;
}
// Which of the 27 bricks around the chunk's home brick are active (bit (dx+1)*9 + (dy+1)*3 + (dz+1)); whole warp calls.
DD_DEV unsigned chunk_active_mask(const int *active_flag, const ChunkGeom &cg, const KP &kp, int lane) {
  int nbx = kp.gx >> 2, nby = kp.gy >> 2, nbz = kp.gz >> 2;
  bool on = false;
  if (lane < 27) {
    int x = cg.bx + lane / 9 - 1, y = cg.by + (lane / 3) % 3 - 1, z = cg.bz + lane % 3 - 1;
    if ((unsigned)x < (unsigned)nbx && (unsigned)y < (unsigned)nby && (unsigned)z < (unsigned)nbz)
      on = *(const volatile int *)(active_flag + (size_t)cg.env * nbx * nby * nbz + (x * nby + y) * nbz + z) == 1;
  }
  return __ballot_sync(0xffffffffu, on);
}
// Asynchronous row staging for the tiled kernels.  A warp runs few rounds concurrently with few other warps, so a round that
// starts with loads from HBM stalls for a full memory round trip.  Instead every lane copies the float4s of ITS particle of
// the NEXT round into a per-warp staging area with cp.async (no registers, completes in the background of the current round)
// and picks them up at the start of that round.  A lane only ever touches its own staging slots.
DD_DEV void cp_async16(float4 *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
DD_DEV void cp_async4(float *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
DD_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
DD_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// Fill of the 8^3 tile from a dense grid (swizzled slots) with asynchronous copies: no registers, all 16 copies of a lane in
// flight at once; out-of-grid nodes are zero-filled by a copy of source size 0.  The caller waits for the group
// (cp_async_wait_all + __syncwarp) before the first read.
DD_DEV void cp_async16_zfill(float4 *smem_dst, const void *gsrc, unsigned src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
DD_DEV void fill_tile_async(float4 *tile, const float4 *__restrict__ grid_env, const KP &kp, int ox, int oy, int oz, int lane, float4 *zero_too = nullptr) {
#pragma unroll
  for (int n0 = 0; n0 < kTileN; n0 += 32) {
    int n = n0 + lane;
    int txx = n >> 6, tyy = (n >> 3) & 7, tzz = n & 7;
    int nx = ox + txx, ny = oy + tyy, nz = oz + tzz;
    bool ok = (unsigned)nx < (unsigned)kp.gx && (unsigned)ny < (unsigned)kp.gy && (unsigned)nz < (unsigned)kp.gz;
    int slot = tile_slot(txx, tyy, tzz);
    cp_async16_zfill(tile + slot, ok ? grid_env + (nx * kp.gy + ny) * kp.gz + nz : grid_env, ok ? 16u : 0u);
    if (zero_too) zero_too[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  cp_async_commit();
}
// Persistent chunk loop of the tiled kernels: every warp pulls chunk indices from a ticket counter until the list is
// drained, so a launch never ends with a nearly empty last wave.  sched[0] = next ticket, sched[1] = warps that have
// finished; the last warp to finish clears both for the next launch (launches are serialised on one stream).
DD_DEV int next_chunk(int *sched, int lane) {
  int c = 0;
  if (lane == 0) c = atomicAdd(sched, 1);
  return __shfl_sync(0xffffffffu, c, 0);
}
// A launch with no more chunks than warps needs no tickets at all: warp w takes list entry w.  (Small scenes: the atomic round
// trip in front of a warp's only chunk was 4-10 % of the stall samples; 10k particles: 83 -> 95 M particle-substeps/s.  With
// more chunks than warps the ticket for the first chunk stays: whichever warp is up first takes the largest chunk.)
#define DD_CHUNK_LOOP(ci) \
  const int gw_ = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)); \
  const bool few_ = nchunks <= (int)(gridDim.x * (blockDim.x >> 5)); \
  for (int ci = few_ ? gw_ : next_chunk(sched, lane); ci < nchunks; ci = few_ ? nchunks : next_chunk(sched, lane))
DD_DEV void chunks_done(int *sched, int lane) {
  if (lane == 0 && atomicAdd(sched + 1, 1) == (int)(gridDim.x * (blockDim.x >> 5)) - 1) { sched[0] = 0; sched[1] = 0; }
}
// The active region starts as the set of bricks some particle's stencil (+1 node of slack) touched at the last sort and GROWS
// while a segment runs: a particle whose stencil reaches a brick outside it activates that brick (k_p2g_tile is the first
// kernel of a substep to see a new position).  Activation = claim the brick's flag (0 -> 2), zero its nodes in this substep's
// scatter target (stale data of an earlier rollout may sit there), append it to the active list the grid kernels walk, publish
// (flag = 1).  The list only grows, so every later kernel of the segment -- forward or adjoint -- covers the brick.
// (tx,ty,tz) = stencil base in tile coordinates; returns true if this lane's stencil needs a brick the chunk's mask lacks.
DD_DEV bool stencil_misses(unsigned amask, int tx, int ty, int tz) {
  bool near = (unsigned)(tx + 3) <= 9u && (unsigned)(ty + 3) <= 9u && (unsigned)(tz + 3) <= 9u;  // inside the 3x3x3 brick neighbourhood
  if (!near) return true;
  if (amask == 0x7ffffffu) return false;
  int ax = (tx + 3) >> 2, bx = (tx + 5) >> 2, ay = (ty + 3) >> 2, by = (ty + 5) >> 2, az = (tz + 3) >> 2, bz = (tz + 5) >> 2;
  unsigned need = 0u;
  need |= 1u << (ax * 9 + ay * 3 + az); need |= 1u << (ax * 9 + ay * 3 + bz); need |= 1u << (ax * 9 + by * 3 + az); need |= 1u << (ax * 9 + by * 3 + bz);
  need |= 1u << (bx * 9 + ay * 3 + az); need |= 1u << (bx * 9 + ay * 3 + bz); need |= 1u << (bx * 9 + by * 3 + az); need |= 1u << (bx * 9 + by * 3 + bz);
  return (need & ~amask) != 0u;
}
// whole warp calls; `miss` marks the lanes whose stencil (base node sbx,sby,sbz, already clamped into the grid) needs activation
DD_DEV void activate_bricks(const KP &kp, int env, bool miss, int sbx, int sby, int sbz, int *flags, int *active, int *cnt, float4 *grid_env, int lane) {
  const unsigned full = 0xffffffffu;
  int nby = kp.gy >> 2, nbz = kp.gz >> 2, NB = (kp.gx >> 2) * nby * nbz;
  unsigned todo = __ballot_sync(full, miss);
  while (todo) {
    int src = __ffs(todo) - 1;
    todo &= todo - 1;
    int sx = __shfl_sync(full, sbx, src), sy = __shfl_sync(full, sby, src), sz = __shfl_sync(full, sbz, src);
    for (int c = 0; c < 8; ++c) {
      int X = (sx + ((c & 1) ? 2 : 0)) >> 2, Y = (sy + ((c & 2) ? 2 : 0)) >> 2, Z = (sz + ((c & 4) ? 2 : 0)) >> 2;
      int b = env * NB + (X * nby + Y) * nbz + Z;
      int state = 1;
      if (lane == 0) {
        state = *(volatile int *)(flags + b);
        if (state != 1) state = atomicCAS(flags + b, 0, 2);
      }
      state = __shfl_sync(full, state, 0);
      if (state == 0) {  // ours to activate
        for (int n = lane; n < 64; n += 32)
          grid_env[((X * 4 + (n >> 4)) * kp.gy + Y * 4 + ((n >> 2) & 3)) * kp.gz + Z * 4 + (n & 3)] = make_float4(0.f, 0.f, 0.f, 0.f);
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          active[atomicAdd(cnt + 1, 1)] = b;
          __threadfence();
          atomicExch(flags + b, 1);
        }
      } else if (state == 2) {  // another warp is zeroing it right now
        if (lane == 0) while (atomicAdd(flags + b, 0) == 2) {}
      }
      __syncwarp();
    }
  }
}

// TILE = true: node adjoints are read from a swizzled shared-memory tile at (tx,ty,tz); otherwise from the dense grid
// Staged inputs (tiled kernel, svd_mode 1): `stg` points at this lane's first staging slot (stride 32 float4); slots as in
// kP2ggSlot*.  hook.phase1_done() / hook.phase2_done() are called as soon as the respective slots have been read, so that the
// caller can start the asynchronous copies of the next round into them.
struct NoHook { DD_DEV void phase1_done() const {} DD_DEV void phase2_done() const {} };
// staging slots: 0 x|v  1 v|C  2 material  3-5 affine, sigma | 6,7 C  8,9 F  10 qU  11 qV  12,13 gF  14 partial gx  15 (F22, yield, gF22, -)
template <int SVD, bool TILE, class Hook = NoHook>
DD_DEV void p2g_grad_particle(const KP &kp, int p, const float *__restrict__ cur, const float *__restrict__ nxt, const float4 *__restrict__ mat0,
                              const float *__restrict__ yield, const float4 *__restrict__ ggrid, const float4 *tile, int ox, int oy, int oz,
                              const float *__restrict__ gin, float *__restrict__ gout, const float4 *stg = nullptr, const Hook &hook = Hook()) {
  // svd_mode 1: the gather needs only x, v, the mass and the affine matrix the forward pass left in the next slot; everything
  // else (F, C, the SVD factors, the incoming F gradient) is loaded after the gather so that it does not sit in registers
  XVC s;
  M3 F;
  float4 m0 = stg ? stg[2 * 32] : __ldg(mat0 + p);
  float yl = 0.f;
  Constit c;
  float4 aux2 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = aux2;
  if (SVD == 1) {
    float4 a = stg ? stg[0] : ldg_stream(plane4(cur, kp.EN, 0) + p);
    p1 = stg ? stg[32] : ldg_stream(plane4(cur, kp.EN, 1) + p);
    s.x = v3(a.x, a.y, a.z);
    s.v = v3(a.w, p1.x, p1.y);
    float4 a0 = stg ? stg[3 * 32] : ldg_stream(aux4(nxt, kp.EN, 0) + p), a1 = stg ? stg[4 * 32] : ldg_stream(aux4(nxt, kp.EN, 1) + p);
    aux2 = stg ? stg[5 * 32] : ldg_stream(aux4(nxt, kp.EN, 2) + p);
    c.affine = m3(a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, aux2.x);
    hook.phase1_done();
  } else {
    s = load_xvc(cur, kp.EN, p);
    F = load_F(cur, kp.EN, p);
    yl = __ldg(yield + p);
    float4 q = load_q(nxt, kp.EN, p);
    constitutive<SVD>(s, F, m0, yl, kp, c, q, 0);
  }
  float mu = m0.z, lam = m0.w, m_p = m0.x;
  Stencil st = make_stencil_safe(s.x, kp);
  V3 d0, d1, d2;
  stencil_dw(st, kp.inv_dx, d0, d1, d2);
  float wx[3] = {st.w0.x, st.w1.x, st.w2.x}, wy[3] = {st.w0.y, st.w1.y, st.w2.y}, wz[3] = {st.w0.z, st.w1.z, st.w2.z};
  float ex[3] = {d0.x, d1.x, d2.x}, ey[3] = {d0.y, d1.y, d2.y}, ez[3] = {d0.z, d1.z, d2.z};
  V3 c0 = v3(c.affine.a00, c.affine.a10, c.affine.a20) * kp.dx, c1 = v3(c.affine.a01, c.affine.a11, c.affine.a21) * kp.dx,
     c2 = v3(c.affine.a02, c.affine.a12, c.affine.a22) * kp.dx;
  V3 base = m_p * s.v - (c0 * st.fx.x + c1 * st.fx.y + c2 * st.fx.z);
  const float4 *gg = ggrid + (size_t)(p / kp.N) * kp.G + (st.bx * kp.gy + st.by) * kp.gz + st.bz;
  int tx = st.bx - ox, ty = st.by - oy, tz = st.bz - oz;
  bool in_tile = TILE && (unsigned)tx <= 5u && (unsigned)ty <= 5u && (unsigned)tz <= 5u;
  // Separable form of the 27-node gather.  With t = (g_mv, g_m) the node adjoints and val_n = base + c0 i + c1 j + c2 k:
  //   per (i,j) row:  A = sum_k wz_k g_mv,  B = sum_k k wz_k g_mv,  M = sum_k wz_k g_m   (and E, EB, EM with dwz instead of wz)
  //   sum_k wz_k (m g_m + g_mv . val) = m M + A . val_ij + B . c2          -> the x and y components of dL/dx
  //   sum_k dwz_k ( ... )             = m EM + E . val_ij + EB . c2        -> the z component
  //   Sv = sum w_ij A,  sum N k g_mv = sum w_ij B,  sum N i g_mv = sum (i w_ij) A,  sum N j g_mv = sum (j w_ij) A
  // T = sum N g_mv (x) (offset - fx) dx follows from those four sums after the loop (~24 instead of ~33 instructions a node).
  // Every sum comes in two flavours, with the z weights wz_k and with their derivatives ez_k: the two are computed together as
  // packed pairs (FFMA2, the node value broadcast to both halves), .x = weight flavour, .y = derivative flavour.
  V3 Sv = vzero(), g_x = vzero(), Tx = vzero(), Ty = vzero(), Tz = vzero();
  const float2 WE0 = pk(wz[0], ez[0]), WE1 = pk(wz[1], ez[1]), WE2 = pk(wz[2], ez[2]), KE2 = pk(2.f * wz[2], 2.f * ez[2]);
  auto row = [&](int i, int j, float wxi, float wyj, float exi, float eyj, float4 t0, float4 t1, float4 t2) {
    float wij = wxi * wyj, a1 = exi * wyj, a2 = wxi * eyj;
    V3 vij = step_n(step_n(base, c0, i), c1, j);
    float2 AEx = fma2(t2.x, WE2, fma2(t1.x, WE1, mul2(t0.x, WE0))), AEy = fma2(t2.y, WE2, fma2(t1.y, WE1, mul2(t0.y, WE0)));   // (A, E)
    float2 AEz = fma2(t2.z, WE2, fma2(t1.z, WE1, mul2(t0.z, WE0))), AEm = fma2(t2.w, WE2, fma2(t1.w, WE1, mul2(t0.w, WE0)));   // .. (M, EM)
    float2 BEx = fma2(t2.x, KE2, mul2(t1.x, WE1)), BEy = fma2(t2.y, KE2, mul2(t1.y, WE1)), BEz = fma2(t2.z, KE2, mul2(t1.z, WE1));  // (B, EB)
    float2 SS = fma2(c2.z, BEz, fma2(c2.y, BEy, fma2(c2.x, BEx, fma2(vij.z, AEz, fma2(vij.y, AEy, fma2(vij.x, AEx, mul2(m_p, AEm)))))));  // (S, SE)
    g_x.x = fmaf(a1, SS.x, g_x.x); g_x.y = fmaf(a2, SS.x, g_x.y); g_x.z = fmaf(wij, SS.y, g_x.z);
    Sv.x = fmaf(wij, AEx.x, Sv.x); Sv.y = fmaf(wij, AEy.x, Sv.y); Sv.z = fmaf(wij, AEz.x, Sv.z);
    Tz.x = fmaf(wij, BEx.x, Tz.x); Tz.y = fmaf(wij, BEy.x, Tz.y); Tz.z = fmaf(wij, BEz.x, Tz.z);
    if (i > 0) { float wi = wij * (float)i; Tx.x = fmaf(wi, AEx.x, Tx.x); Tx.y = fmaf(wi, AEy.x, Tx.y); Tx.z = fmaf(wi, AEz.x, Tx.z); }
    if (j > 0) { float wj = wij * (float)j; Ty.x = fmaf(wj, AEx.x, Ty.x); Ty.y = fmaf(wj, AEy.x, Ty.y); Ty.z = fmaf(wj, AEz.x, Ty.z); }
  };
  if (in_tile) {
    const float4 *trow = tile + (tx << 6 | ty << 3);
    int g0 = tz + 4 * ty + 2 * tx;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float4 *r_ = trow + (i << 6 | j << 3);
        int g = g0 + 2 * i + 4 * j;
        row(i, j, wx[i], wy[j], ex[i], ey[j], r_[g & 7], r_[(g + 1) & 7], r_[(g + 2) & 7]);
      }
  } else if (TILE) {  // left the tile since the last sort (rare): rolled loop over the dense grid, kept small on purpose
#pragma unroll 1
    for (int i = 0; i < 3; ++i)
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        const float4 *r_ = gg + (i * kp.gy + j) * kp.gz;
        row(i, j, pick(st.w0, st.w1, st.w2, i, 0), pick(st.w0, st.w1, st.w2, j, 1), pick(d0, d1, d2, i, 0), pick(d0, d1, d2, j, 1), __ldg(r_), __ldg(r_ + 1), __ldg(r_ + 2));
      }
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float4 *r_ = gg + (i * kp.gy + j) * kp.gz;
        row(i, j, wx[i], wy[j], ex[i], ey[j], __ldg(r_), __ldg(r_ + 1), __ldg(r_ + 2));
      }
  }
  M3 gF_next;
  float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
  if (SVD == 1) {  // second half of the loads, then the factors from their checkpoints (no SVD, no QR, no stress evaluation)
    float4 cc, d, f0, f1, qu, q, g0, g1, sc;
    if (stg) {
      cc = stg[6 * 32]; d = stg[7 * 32]; f0 = stg[8 * 32]; f1 = stg[9 * 32]; qu = stg[10 * 32]; q = stg[11 * 32]; g0 = stg[12 * 32]; g1 = stg[13 * 32];
      part = stg[14 * 32]; sc = stg[15 * 32];
    } else {
      cc = ldg_stream(plane4(cur, kp.EN, 2) + p); d = ldg_stream(plane4(cur, kp.EN, 3) + p);
      f0 = ldg_stream(plane4(cur, kp.EN, 4) + p); f1 = ldg_stream(plane4(cur, kp.EN, 5) + p);
      qu = ldg_stream(aux4(nxt, kp.EN, 3) + p); q = load_q(nxt, kp.EN, p);
      g0 = ldg_stream(plane4(gin, kp.EN, 4) + p); g1 = ldg_stream(plane4(gin, kp.EN, 5) + p);
      sc = make_float4(__ldg(cur + (size_t)24 * kp.EN + p), __ldg(yield + p), __ldg(gin + (size_t)24 * kp.EN + p), 0.f);
      part = plane4(gout, kp.EN, 0)[p];  // partial dL/dx written by the g2p adjoint
    }
    hook.phase2_done();
    s.C = m3(p1.z, p1.w, cc.x, cc.y, cc.z, cc.w, d.x, d.y, d.z);
    F = m3(f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, sc.x);
    yl = sc.y;
    gF_next = m3(g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, sc.z);
    c.Ft = mul(mdiag(1.f) + s.C * kp.dt, F);
    c.sigma = v3(aux2.y, aux2.z, aux2.w);
    c.U = quat_to_m3(qu.x, qu.y, qu.z, qu.w);
    c.Vm = quat_to_m3(q.x, q.y, q.z, q.w);
    c.J = von_mises<true>(c.Ft, c.U, c.sigma, c.Vm, yl, m0.z, c.nF, c.pl);
    c.r = mul_nt(c.U, c.Vm);
    c.scale = -kp.dt * m0.y * 4.f * kp.inv_dx * kp.inv_dx;
  }
  V3 t0 = (Tx - Sv * st.fx.x) * kp.dx, t1 = (Ty - Sv * st.fx.y) * kp.dx, t2 = (Tz - Sv * st.fx.z) * kp.dx;
  M3 T = m3(t0.x, t1.x, t2.x, t0.y, t1.y, t2.y, t0.z, t1.z, t2.z);
  g_x -= mul_t(c.affine, Sv);
  M3 g_stress = c.scale * T, g_C = m_p * T;
  V3 g_v = m_p * Sv;
  if (SVD != 1) {
    gF_next = load_F(gin, kp.EN, p);
    part = plane4(gout, kp.EN, 0)[p];
  }
  g_x += v3(part.x, part.y, part.z);
  // Adjoint of stress -> (F_new, R = U V^T, J) and of the return map, then through the SVD (integrator.cu:541-620, 131-159),
  // regrouped: with W = U^T g_R V and Y = U^T g_Fnew V every U/V gradient the reference materialises is
  //   U^T gU = W + Y E,   V^T gV = W^T + Y^T E      (E = diag(exp eps), plastic branch only)
  // and the three SVD-adjoint terms share one U ( . ) V^T sandwich.
  M3 gsF = mul(g_stress, c.nF);
  M3 g_nF = gF_next + (2.f * mu) * (mul_tn(g_stress, c.nF - c.r) + gsF);
  M3 W = mul_tn(c.U, mul((-2.f * mu) * gsF, c.Vm));
  float g_J = ((2 * c.J - 1) * lam) * trace(g_stress);
  V3 g_sig = vzero();
  M3 G = mzero();  // dL/dF~ accumulated directly
  M3 A = W, B = transpose(W);  // U^T gU and V^T gV
  if (c.pl.plastic) {
    const Plastic &pl = c.pl;
    M3 Y = mul_tn(c.U, mul(g_nF, c.Vm));
    A += mul_diag(Y, pl.ee);
    B += mul_diag(transpose(Y), pl.ee);
    V3 Fpart = diag(Y);
    V3 Jpart = v3(g_J * pl.ee.y * pl.ee.z, g_J * pl.ee.x * pl.ee.z, g_J * pl.ee.x * pl.ee.y);
    V3 g_eps = pl.ee * (Jpart + Fpart);
    V3 g_eh = (-pl.dg / pl.ehn) * g_eps;
    float g_ehn = -dot(pl.eh / pl.ehn, g_eps) * (yl / (2 * mu)) / pl.ehn;
    g_eh += (pl.eh / pl.ehn) * g_ehn;
    float mean_g = (g_eh.x + g_eh.y + g_eh.z) / 3.f;
    g_eps += v3(g_eh.x - mean_g, g_eh.y - mean_g, g_eh.z - mean_g);
    if (c.sigma.x >= 0.05) g_sig.x += g_eps.x / c.sigma.x;
    if (c.sigma.y >= 0.05) g_sig.y += g_eps.y / c.sigma.y;
    if (c.sigma.z >= 0.05) g_sig.z += g_eps.z / c.sigma.z;
  } else {
    g_sig += v3(g_J * c.sigma.y * c.sigma.z, g_J * c.sigma.x * c.sigma.z, g_J * c.sigma.x * c.sigma.y);
    G = g_nF;
  }
  {
    V3 sg = c.sigma;
    float d10 = (sg.y - sg.x) * (sg.y + sg.x), d20 = (sg.z - sg.x) * (sg.z + sg.x), d21 = (sg.z - sg.y) * (sg.z + sg.y);
    float k10 = __fdividef(1.f, clamp_eps(d10)), k20 = __fdividef(1.f, clamp_eps(d20)), k21 = __fdividef(1.f, clamp_eps(d21));
    float k01 = __fdividef(1.f, clamp_eps(-d10)), k02 = __fdividef(1.f, clamp_eps(-d20)), k12 = __fdividef(1.f, clamp_eps(-d21));
    M3 K = m3(0.f, k10, k20, k01, 0.f, k21, k02, k12, 0.f);
    // inner = (K o (A - A^T)) Sigma + Sigma (K o (B - B^T)) + diag(g_sigma)
    M3 inner = mul_diag(hadamard(K, A - transpose(A)), sg) + diag_mul(sg, hadamard(K, B - transpose(B))) + mdiag(g_sig);
    G += mul_nt(mul(c.U, inner), c.Vm);
  }
  g_C += kp.dt * mul_nt(G, F);
  M3 g_F = mul_tn(mdiag(1.f) + kp.dt * s.C, G);
  store_xvc(gout, kp.EN, p, g_x, g_v, g_C);
  store_F(gout, kp.EN, p, g_F);
}
template <int SVD>
__global__ void __launch_bounds__(kT, DD_LB_P2G_GRAD) k_p2g_grad(KP kp, const int *__restrict__ spos, const float *__restrict__ cur, const float *__restrict__ nxt,
                                                 const float4 *__restrict__ mat0, const float *__restrict__ yield, const float4 *__restrict__ ggrid,
                                                 const float *__restrict__ gin, float *__restrict__ gout) {
  int p = blockIdx.x * kT + threadIdx.x;
  if (p >= kp.EN) return;
  (void)spos;  // measured: strided state loads cost this kernel more than the gather locality gains
  p2g_grad_particle<SVD, false>(kp, p, cur, nxt, mat0, yield, ggrid, nullptr, 0, 0, 0, gin, gout);
}

// ================================================================================================ tiled kernels
// Particles are stored brick by brick (4x4x4 cells), a brick's particles cell-sorted and split into chunks.  One warp
// owns one chunk and a private 8x8x8-node shared-memory tile around the brick (1 node of slack on every side for drift
// since the last sort).  Inside a chunk lane l walks the sorted ranks [l*R, (l+1)*R), so the 32 particles of a round
// sit in 32 different cells and their read-modify-writes on the tile never collide; the storage order is the
// round-major transpose of that assignment, which keeps every global load coalesced.  Collisions that do occur
// (dense cells, drift) are detected with match.any: the second lane of a cell is absorbed by its owner lane (pairs, see
// k_p2g_tile), further lanes take the deferred queue below.  The tile is flushed with one vector reduction per touched node
// instead of one per (particle, node).  The slack for drift is one node on the high side of every axis (tile origin =
// brick origin - 1: a fresh stencil base sits at tile coordinate 0..4, the tile takes 0..5).

// Lanes that cannot use the tile in their row -- the THIRD and further lanes of a cell in a row (the second one is absorbed by
// the cell's owner lane, see "pairs" in k_p2g_tile), or a lane whose stencil has left the tile since the last sort -- are not
// served inside the row loop: their storage positions go to a small per-warp queue, and whenever 32 are waiting (and at the end
// of the kernel) every lane takes one, reads its inputs back and sends its 27 contributions straight to the grid.  Measured
// (-DDD_COUNT_DEFER): 0.5 % of the particles at config D, 3-4 % in the dense 10k-particle scenes (before the pairs: 3.3 % / 16 %).
#ifdef DD_COUNT_DEFER
__device__ unsigned long long g_defer_count[2];  // (diagnostic build) deferred lanes, rows with a deferred lane
#endif
struct DeferQueue {
  int *slots;  // 32 ints of shared memory, private to the warp
  int n;       // warp-uniform
  template <class Flush>
  DD_DEV void push(bool defer, int p, int lane, Flush flush) {
    unsigned dm = __ballot_sync(0xffffffffu, defer);
    if (dm == 0u) return;
#ifdef DD_DROP_DEFERRED
    return;  // (diagnostic build, WRONG results: what the direct path of the deferred lanes costs)
#endif
    int k = __popc(dm);
#ifdef DD_COUNT_DEFER
    if (lane == 0) { atomicAdd(&g_defer_count[0], (unsigned long long)k); atomicAdd(&g_defer_count[1], 1ull); }
#endif
    if (n + k > 32) flush();
    if (defer) slots[n + __popc(dm & ((1u << lane) - 1u))] = p;
    n += k;
  }
};
// scatter of one particle straight to the grid from what the forward pass stored: x, v of state t, the affine matrix in slot t+1
__device__ __noinline__ void p2g_direct(const KP &kp, int p, const float *cur, const float *nxt, const float4 *__restrict__ mat0, float4 *__restrict__ grid) {
  float4 a = __ldcg(plane4(cur, kp.EN, 0) + p), b = __ldcg(plane4(cur, kp.EN, 1) + p), m0 = __ldg(mat0 + p);
  float4 a0 = __ldcg(aux4(nxt, kp.EN, 0) + p), a1 = __ldcg(aux4(nxt, kp.EN, 1) + p), a2 = __ldcg(aux4(nxt, kp.EN, 2) + p);
  V3 x = v3(a.x, a.y, a.z), v = v3(a.w, b.x, b.y);
  Stencil st = make_stencil_safe(x, kp);
  float m = m0.x;
  V3 c0 = v3(a0.x, a0.w, a1.z) * kp.dx, c1 = v3(a0.y, a1.x, a1.w) * kp.dx, c2 = v3(a0.z, a1.y, a2.x) * kp.dx;
  V3 base = m * v - (c0 * st.fx.x + c1 * st.fx.y + c2 * st.fx.z);
  float wx[3] = {st.w0.x, st.w1.x, st.w2.x}, wy[3] = {st.w0.y, st.w1.y, st.w2.y}, wz[3] = {st.w0.z, st.w1.z, st.w2.z};
  float4 *g = grid + (size_t)(p / kp.N) * kp.G + (st.bx * kp.gy + st.by) * kp.gz + st.bz;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    V3 vi = step_n(base, c0, i);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      V3 vij = step_n(vi, c1, j);
      float wij = wx[i] * wy[j];
      float4 *r_ = g + (i * kp.gy + j) * kp.gz;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        V3 val = step_n(vij, c2, k);
        float w = wij * wz[k];
        red_add_v4(r_ + k, val.x * w, val.y * w, val.z * w, m * w);
      }
    }
  }
}

template <int SVD, bool WRITE_F>
__global__ void __launch_bounds__(32 * kTileWarps, DD_LB_P2G_TILE) k_p2g_tile(KP kp, SegView sg, const float *__restrict__ cur,
                                                                 float *__restrict__ nxt, const float4 *__restrict__ mat0,
                                                                 const float *__restrict__ yield, float4 *__restrict__ grid, int *sched) {
  extern __shared__ float4 dd_smem[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 *tile = dd_smem + warp * (kTileN + kStageP2G * 32 + kQueueF4), *stage = tile + kTileN + lane;
  DeferQueue dq{reinterpret_cast<int *>(tile + kTileN + kStageP2G * 32), 0};
  unsigned tbase = smem_u32(tile);
  pdl_launch_dependents();
  for (int n = lane; n < kTileN; n += 32) tile[n] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  pdl_wait();  // nothing produced by earlier kernels (tickets, counts, particles) is touched before this point
  const int nchunks = sg.cnt[0];
  const int4 *__restrict__ chunks = sg.chunks;
  auto stage_row = [&](int p) {  // x,v,C | 8 of F | quaternion | material: 8 float4; then F22 and the yield stress
#pragma unroll
    for (int k = 0; k < 6; ++k) cp_async16(stage + 32 * k, plane4(cur, kp.EN, k) + p);
    cp_async16(stage + 192, reinterpret_cast<const float4 *>(cur + (size_t)25 * kp.EN) + p);
    cp_async16(stage + 224, mat0 + p);
    cp_async4(reinterpret_cast<float *>(stage + 256), cur + (size_t)24 * kp.EN + p);
    cp_async4(reinterpret_cast<float *>(stage + 256) + 1, yield + p);
    cp_async_commit();
  };
  auto flush_queue = [&]() {
    __syncwarp();  // the affine matrices the queued particles' lanes stored are visible to the whole warp
    if (lane < dq.n) p2g_direct(kp, dq.slots[lane], cur, nxt, mat0, grid);
    __syncwarp();
    dq.n = 0;
  };
  // (no look-ahead across chunks: a warp that holds a chunk in reserve lengthens the tail of the launch -- measured)
  DD_CHUNK_LOOP(ci) {
  ChunkGeom cg = chunk_geom(chunks[ci], kp);
  stage_row(row_pos(cg, 0, lane));
  unsigned amask = chunk_active_mask(sg.flags, cg, kp, lane);
  float4 *g = grid + (size_t)cg.env * kp.G;
  for (int j = 0; j < cg.R; ++j) {
    bool act = lane_on(cg, j, lane);
    int p = row_pos(cg, j, lane);  // idle lanes shadow a valid particle, contribute nothing
    cp_async_wait_all();
    float4 r0 = stage[0], r1 = stage[32], r2 = stage[64], r3 = stage[96], r4 = stage[128], r5 = stage[160], q = stage[192], m0 = stage[224], r8 = stage[256], qu = make_float4(0.f, 0.f, 0.f, 1.f);
    if (j + 1 < cg.R) stage_row(row_pos(cg, j + 1, lane));
    XVC s;
    s.x = v3(r0.x, r0.y, r0.z);
    s.v = v3(r0.w, r1.x, r1.y);
    s.C = m3(r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z);
    M3 F = m3(r4.x, r4.y, r4.z, r4.w, r5.x, r5.y, r5.z, r5.w, r8.x);
    Constit c;
    constitutive<SVD>(s, F, m0, r8.y, kp, c, q, 6, &qu);
    if (WRITE_F && act) { store_F(nxt, kp.EN, p, c.nF); store_q(nxt, kp.EN, p, q); store_constit(nxt, kp.EN, p, c, qu); }
    Stencil st = make_stencil_safe(s.x, kp);
    float m = m0.x;
    V3 c0 = v3(c.affine.a00, c.affine.a10, c.affine.a20) * kp.dx, c1 = v3(c.affine.a01, c.affine.a11, c.affine.a21) * kp.dx,
       c2 = v3(c.affine.a02, c.affine.a12, c.affine.a22) * kp.dx;
    V3 base = m * s.v - (c0 * st.fx.x + c1 * st.fx.y + c2 * st.fx.z);  // value at node offset (0,0,0); +c_a per step along axis a
    float wx[3] = {st.w0.x, st.w1.x, st.w2.x}, wy[3] = {st.w0.y, st.w1.y, st.w2.y}, wz[3] = {st.w0.z, st.w1.z, st.w2.z};
    int tx = st.bx - cg.ox, ty = st.by - cg.oy, tz = st.bz - cg.oz;
    {  // (rare) the stencil reaches a brick outside the active region: activate it before anything lands there
      bool miss = act && stencil_misses(amask, tx, ty, tz);
      if (__any_sync(0xffffffffu, miss)) {
        activate_bricks(kp, cg.env, miss, st.bx, st.by, st.bz, sg.flags, sg.active, sg.cnt, g, lane);
        amask = chunk_active_mask(sg.flags, cg, kp, lane);
      }
    }
    bool in_tile = act && (unsigned)tx <= 5u && (unsigned)ty <= 5u && (unsigned)tz <= 5u;
    unsigned key = in_tile ? (unsigned)(tx << 6 | ty << 3 | tz) : 0x1000u + lane;
    unsigned peers = __match_any_sync(0xffffffffu, key);
    const unsigned lower = peers & ((1u << lane) - 1u);
    bool mine = in_tile && lower == 0u;  // lowest lane of its cell in this row
    // Two lanes of a row in the same cell (dense scenes: fewer occupied cells in a brick than lanes; 3 % of the particles at
    // config D, 12-16 % at 50 particles per cell): the SECOND lane of a cell does not go to the deferred queue -- its owner lane
    // fetches its scatter parameters (22 shuffles, only in rows that have such a pair) and adds both contributions in one
    // read-modify-write.  Third and further lanes of a cell, and lanes outside the tile, are deferred as before.
    bool second = in_tile && __popc(lower) == 1;
    const unsigned higher = lane < 31 ? peers & ~((2u << lane) - 1u) : 0u;
    const bool has_partner = mine && higher != 0u;
    if (!in_tile) { tx = ty = tz = 0; }
#ifdef DD_NO_PAIRS
    second = false;
#endif
    {
      // byte address of the node at stencil offset (i, jj, k): row base + (i << 6 | jj << 3) * 16 (an immediate after unrolling)
      // + the bank group rotated by 2 i + 4 jj + k; the eight rotations are formed once per row
      const unsigned rowb = tbase + 16u * (unsigned)(tx << 6 | ty << 3);
      const int g0 = tz + 4 * ty + 2 * tx;
      unsigned rot[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) rot[r] = rowb + (((unsigned)(g0 + r) & 7u) << 4);
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#ifndef DD_NO_PAIRS
      if (__any_sync(0xffffffffu, has_partner)) {
        const int src = has_partner ? __ffs(higher) - 1 : lane;
        auto get = [&](float v) { return __shfl_sync(0xffffffffu, v, src); };
        V3 bB = v3(get(base.x), get(base.y), get(base.z)), c0B = v3(get(c0.x), get(c0.y), get(c0.z)), c1B = v3(get(c1.x), get(c1.y), get(c1.z)),
           c2B = v3(get(c2.x), get(c2.y), get(c2.z));
        float mB = get(m);
        float wxB[3] = {get(wx[0]), get(wx[1]), get(wx[2])}, wyB[3] = {get(wy[0]), get(wy[1]), get(wy[2])}, wzB[3] = {get(wz[0]), get(wz[1]), get(wz[2])};
        if (!has_partner) { wxB[0] = wxB[1] = wxB[2] = 0.f; }  // no partner: the second contribution is zero
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          V3 vi = step_n(base, c0, i), viB = step_n(bB, c0B, i);
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) {
            V3 vij = step_n(vi, c1, jj), vijB = step_n(viB, c1B, jj);
            float wij = wx[i] * wy[jj], wijB = wxB[i] * wyB[jj];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              V3 val = step_n(vij, c2, k), valB = step_n(vijB, c2B, k);
              float w = wij * wz[k], wB = wijB * wzB[k];
              unsigned a = rot[(2 * i + 4 * jj + k) & 7] + 16u * (unsigned)(i << 6 | jj << 3);
              lds_v4_into(t, a, mine);
              t.x = fmaf(val.x, w, t.x); t.y = fmaf(val.y, w, t.y); t.z = fmaf(val.z, w, t.z); t.w = fmaf(m, w, t.w);
              t.x = fmaf(valB.x, wB, t.x); t.y = fmaf(valB.y, wB, t.y); t.z = fmaf(valB.z, wB, t.z); t.w = fmaf(mB, wB, t.w);
              sts_v4_if(a, t, mine);
            }
          }
        }
      } else
#endif
      {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          V3 vi = step_n(base, c0, i);
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) {
            V3 vij = step_n(vi, c1, jj);
            float wij = wx[i] * wy[jj];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              V3 val = step_n(vij, c2, k);
              float w = wij * wz[k];
              unsigned a = rot[(2 * i + 4 * jj + k) & 7] + 16u * (unsigned)(i << 6 | jj << 3);
              lds_v4_into(t, a, mine);
              t.x = fmaf(val.x, w, t.x); t.y = fmaf(val.y, w, t.y); t.z = fmaf(val.z, w, t.z); t.w = fmaf(m, w, t.w);
              sts_v4_if(a, t, mine);
            }
          }
        }
      }
    }
    dq.push(act && !mine && !second, p, lane, flush_queue);  // third lane of a cell in this row, or has left the tile
  }
  __syncwarp();
  for (int n = lane; n < kTileN; n += 32) {  // flush, and leave the tile zeroed for the next chunk
    int txx = n >> 6, tyy = (n >> 3) & 7, tzz = n & 7;
    float4 *tp = tile + tile_slot(txx, tyy, tzz);
    float4 t = *tp;
    int nx = cg.ox + txx, ny = cg.oy + tyy, nz = cg.oz + tzz;
    if (t.w != 0.f || t.x != 0.f || t.y != 0.f || t.z != 0.f) {
      *tp = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((unsigned)nx < (unsigned)kp.gx && (unsigned)ny < (unsigned)kp.gy && (unsigned)nz < (unsigned)kp.gz)
        red_add_v4(g + (nx * kp.gy + ny) * kp.gz + nz, t.x, t.y, t.z, t.w);
    }
  }
  __syncwarp();
  }
  flush_queue();
  chunks_done(sched, lane);
}

// g2p_grad on tiles (integrator.cu:1527-1614): grid velocities are gathered from a tile copy, their adjoint is scattered
// into a second tile.  With h_n = gv' + (4/dx) gC' (offset_n - fx) (affine in the offset, so evaluated incrementally):
//   d/d v_n  = w_n h_n ;  dL/dx = -(4/dx^2) gC'^T (sum w_n v_n) + sum gradN_n (v_n . h_n)
struct G2pgIn { V3 x, gx, gnv, nvel; M3 gC; };
DD_DEV G2pgIn g2pg_inputs(const KP &kp, float4 a, float4 n0, float4 n1, float4 g0, float4 g1, float4 g2, float4 g3) {
  G2pgIn r;
  r.x = v3(a.x, a.y, a.z);
  r.nvel = v3(n0.w, n1.x, n1.y);
  r.gx = v3(g0.x, g0.y, g0.z);
  r.gnv = v3(g0.w, g1.x, g1.y);
  r.gC = m3(g1.z, g1.w, g2.x, g2.y, g2.z, g2.w, g3.x, g3.y, g3.z);
  V3 hi = v3(((float)kp.gx - 3.f) * kp.dx, ((float)kp.gy - 3.f) * kp.dx, ((float)kp.gz - 3.f) * kp.dx);
  float lo = kp.gh * kp.dx;
  V3 nx = r.x + r.nvel * kp.dt;
  if (nx.x > hi.x || nx.x < lo) r.gx.x = 0;
  if (nx.y > hi.y || nx.y < lo) r.gx.y = 0;
  if (nx.z > hi.z || nx.z < lo) r.gx.z = 0;
  r.gnv += r.gx * kp.dt;
  return r;
}
// scatter half of one queued particle straight to the grid (see DeferQueue)
__device__ __noinline__ void g2pg_direct(const KP &kp, int p, const float *__restrict__ cur, const float *__restrict__ nxt, const float *__restrict__ gin, float4 *__restrict__ ggrid_v) {
  G2pgIn in = g2pg_inputs(kp, ldg_stream(plane4(cur, kp.EN, 0) + p), ldg_stream(plane4(nxt, kp.EN, 0) + p), ldg_stream(plane4(nxt, kp.EN, 1) + p),
                          ldg_stream(plane4(gin, kp.EN, 0) + p), ldg_stream(plane4(gin, kp.EN, 1) + p), ldg_stream(plane4(gin, kp.EN, 2) + p), ldg_stream(plane4(gin, kp.EN, 3) + p));
  Stencil st = make_stencil_safe(in.x, kp);
  float s4 = kp.inv_dx * 4.f;
  float wx[3] = {st.w0.x, st.w1.x, st.w2.x}, wy[3] = {st.w0.y, st.w1.y, st.w2.y}, wz[3] = {st.w0.z, st.w1.z, st.w2.z};
  V3 H0 = v3(in.gC.a00, in.gC.a10, in.gC.a20) * s4, H1 = v3(in.gC.a01, in.gC.a11, in.gC.a21) * s4, H2 = v3(in.gC.a02, in.gC.a12, in.gC.a22) * s4;
  V3 h0 = in.gnv - (H0 * st.fx.x + H1 * st.fx.y + H2 * st.fx.z);
  float4 *g = ggrid_v + (size_t)(p / kp.N) * kp.G + (st.bx * kp.gy + st.by) * kp.gz + st.bz;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    V3 hi_ = step_n(h0, H0, i);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      V3 hij = step_n(hi_, H1, j);
      float wij = wx[i] * wy[j];
      float4 *r_ = g + (i * kp.gy + j) * kp.gz;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        V3 h = step_n(hij, H2, k);
        float w = wij * wz[k];
        red_add_v4(r_ + k, w * h.x, w * h.y, w * h.z, 0.f);
      }
    }
  }
}
// (Measured and rejected, round 2: a block of two warps per chunk sharing the read-only velocity tile, each warp with its own
// adjoint tile and every second row -- 14 instead of 11 warps per SM, but two flushes per chunk: 88 -> 92 us at config D.)
__global__ void __launch_bounds__(32 * kTileWarps, DD_LB_G2PG_TILE) k_g2p_grad_tile(KP kp, SegView sg, const float *__restrict__ cur,
                                                                      const float *__restrict__ nxt, const float4 *__restrict__ grid_v,
                                                                      const float *__restrict__ gin, float *__restrict__ gout,
                                                                      float4 *__restrict__ ggrid_v, int *sched) {
  extern __shared__ float4 dd_smem[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kStage = kStageG2PG;  // staged float4s per particle: x | next (x,v) | incoming (gx, gv, gC)
  float4 *tv = dd_smem + warp * (2 * kTileN + kStage * 32 + kQueueF4), *tg = tv + kTileN, *stage = tg + kTileN + lane;
  DeferQueue dq{reinterpret_cast<int *>(tg + kTileN + kStage * 32), 0};
  unsigned gbase = smem_u32(tg);
  pdl_launch_dependents();
  pdl_wait();
  const int nchunks = sg.cnt[0];
  const int4 *__restrict__ chunks = sg.chunks;
  const float s4 = kp.inv_dx * 4.f;
  auto stage_row = [&](int p) {
    cp_async16(stage, plane4(cur, kp.EN, 0) + p);
    cp_async16(stage + 32, plane4(nxt, kp.EN, 0) + p);
    cp_async16(stage + 64, plane4(nxt, kp.EN, 1) + p);
#pragma unroll
    for (int k = 0; k < 4; ++k) cp_async16(stage + 96 + 32 * k, plane4(gin, kp.EN, k) + p);
    cp_async_commit();
  };
  auto flush_queue = [&]() {
    if (lane < dq.n) g2pg_direct(kp, dq.slots[lane], cur, nxt, gin, ggrid_v);
    __syncwarp();
    dq.n = 0;
  };
  // (no look-ahead across chunks: a warp that holds a chunk in reserve lengthens the tail of the launch -- measured)
  DD_CHUNK_LOOP(ci) {
  ChunkGeom cg = chunk_geom(chunks[ci], kp);
  stage_row(row_pos(cg, 0, lane));
  size_t goff = (size_t)cg.env * kp.G;
  fill_tile_async(tv, grid_v + goff, kp, cg.ox, cg.oy, cg.oz, lane, tg);
  cp_async_wait_all();
  __syncwarp();
  for (int j = 0; j < cg.R; ++j) {
    bool act = lane_on(cg, j, lane);
    int p = row_pos(cg, j, lane);
    cp_async_wait_all();
    G2pgIn in = g2pg_inputs(kp, stage[0], stage[32], stage[64], stage[96], stage[128], stage[160], stage[192]);
    if (j + 1 < cg.R) stage_row(row_pos(cg, j + 1, lane));
    Stencil st = make_stencil_safe(in.x, kp);
    V3 d0, d1, d2;
    stencil_dw(st, kp.inv_dx, d0, d1, d2);
    float wx[3] = {st.w0.x, st.w1.x, st.w2.x}, wy[3] = {st.w0.y, st.w1.y, st.w2.y}, wz[3] = {st.w0.z, st.w1.z, st.w2.z};
    float ex[3] = {d0.x, d1.x, d2.x}, ey[3] = {d0.y, d1.y, d2.y}, ez[3] = {d0.z, d1.z, d2.z};
    V3 H0 = v3(in.gC.a00, in.gC.a10, in.gC.a20) * s4, H1 = v3(in.gC.a01, in.gC.a11, in.gC.a21) * s4, H2 = v3(in.gC.a02, in.gC.a12, in.gC.a22) * s4;
    V3 h0 = in.gnv - (H0 * st.fx.x + H1 * st.fx.y + H2 * st.fx.z);
    int tx = st.bx - cg.ox, ty = st.by - cg.oy, tz = st.bz - cg.oz;
    bool in_tile = act && (unsigned)tx <= 5u && (unsigned)ty <= 5u && (unsigned)tz <= 5u;  // (every brick this stencil needs was activated by the forward pass)
    unsigned key = in_tile ? (unsigned)(tx << 6 | ty << 3 | tz) : 0x1000u + lane;
    unsigned peers = __match_any_sync(0xffffffffu, key);
    const unsigned lower = peers & ((1u << lane) - 1u);
    bool mine = in_tile && lower == 0u;  // lowest lane of its cell in this row: scatters into the tile
    bool second = in_tile && __popc(lower) == 1;  // second lane of its cell: its owner adds its contribution (see k_p2g_tile)
    const unsigned higher = lane < 31 ? peers & ~((2u << lane) - 1u) : 0u;
    const bool has_partner = mine && higher != 0u;
#ifdef DD_NO_PAIRS
    second = false;
#endif
    if (!in_tile) { tx = ty = tz = 0; }
    V3 gxs = vzero();
    const float4 *tvrow = tv + (tx << 6 | ty << 3);
    unsigned growb = gbase + 16u * (unsigned)(tx << 6 | ty << 3);
    int g0 = tz + 4 * ty + 2 * tx;
    // the gather half (read-only tile, every lane) and the scatter of the lanes that own their cell in this row
#ifndef DD_NO_PAIRS
    if (__any_sync(0xffffffffu, has_partner)) {
      const int src = has_partner ? __ffs(higher) - 1 : lane;
      auto get = [&](float v) { return __shfl_sync(0xffffffffu, v, src); };
      V3 h0B = v3(get(h0.x), get(h0.y), get(h0.z)), H0B = v3(get(H0.x), get(H0.y), get(H0.z)), H1B = v3(get(H1.x), get(H1.y), get(H1.z)),
         H2B = v3(get(H2.x), get(H2.y), get(H2.z));
      float wxB[3] = {get(wx[0]), get(wx[1]), get(wx[2])}, wyB[3] = {get(wy[0]), get(wy[1]), get(wy[2])}, wzB[3] = {get(wz[0]), get(wz[1]), get(wz[2])};
      if (!has_partner) { wxB[0] = wxB[1] = wxB[2] = 0.f; }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        V3 hi_ = step_n(h0, H0, i), hiB = step_n(h0B, H0B, i);
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
          V3 hij = step_n(hi_, H1, jj), hijB = step_n(hiB, H1B, jj);
          float wij = wx[i] * wy[jj], a1 = ex[i] * wy[jj], a2 = wx[i] * ey[jj], wijB = wxB[i] * wyB[jj];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            V3 h = step_n(hij, H2, k), hB = step_n(hijB, H2B, k);
            float w = wij * wz[k], wB = wijB * wzB[k];
            int so = (i << 6 | jj << 3) + ((g0 + 2 * i + 4 * jj + k) & 7);
            float4 t = tvrow[so];
            unsigned ga = growb + 16u * (unsigned)so;
            float4 o = lds_v4_if(ga, mine);
            o.x = fmaf(w, h.x, o.x); o.y = fmaf(w, h.y, o.y); o.z = fmaf(w, h.z, o.z);
            o.x = fmaf(wB, hB.x, o.x); o.y = fmaf(wB, hB.y, o.y); o.z = fmaf(wB, hB.z, o.z);
            sts_v4_if(ga, o, mine);
            float qn = t.x * h.x + t.y * h.y + t.z * h.z;
            float tt = wz[k] * qn, uu = ez[k] * qn;
            gxs.x = fmaf(a1, tt, gxs.x); gxs.y = fmaf(a2, tt, gxs.y); gxs.z = fmaf(wij, uu, gxs.z);
          }
        }
      }
    } else
#endif
    {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      V3 hi_ = step_n(h0, H0, i);
#pragma unroll
      for (int jj = 0; jj < 3; ++jj) {
        V3 hij = step_n(hi_, H1, jj);
        float wij = wx[i] * wy[jj], a1 = ex[i] * wy[jj], a2 = wx[i] * ey[jj];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          V3 h = step_n(hij, H2, k);
          float w = wij * wz[k];
          int so = (i << 6 | jj << 3) + ((g0 + 2 * i + 4 * jj + k) & 7);
          float4 t = tvrow[so];
          unsigned ga = growb + 16u * (unsigned)so;
          float4 o = lds_v4_if(ga, mine);  // (carrying one register quadruple across the updates, as k_p2g_tile does, only adds moves here)
          o.x = fmaf(w, h.x, o.x); o.y = fmaf(w, h.y, o.y); o.z = fmaf(w, h.z, o.z);
          sts_v4_if(ga, o, mine);
          float qn = t.x * h.x + t.y * h.y + t.z * h.z;
          float tt = wz[k] * qn, uu = ez[k] * qn;
          gxs.x = fmaf(a1, tt, gxs.x); gxs.y = fmaf(a2, tt, gxs.y); gxs.z = fmaf(wij, uu, gxs.z);
        }
      }
    }
    }
    if (!in_tile) gxs = vzero();
    if (act && !in_tile) {  // left the tile since the last sort (rare): gather from the grid, rolled
#pragma unroll 1
      for (int i = 0; i < 3; ++i)
#pragma unroll 1
        for (int jj = 0; jj < 3; ++jj)
#pragma unroll 1
          for (int k = 0; k < 3; ++k) {
            float wxi = pick(st.w0, st.w1, st.w2, i, 0), wyj = pick(st.w0, st.w1, st.w2, jj, 1), wzk = pick(st.w0, st.w1, st.w2, k, 2);
            V3 h = h0 + H0 * (float)i + H1 * (float)jj + H2 * (float)k;
            float4 t = __ldg(grid_v + goff + ((st.bx + i) * kp.gy + st.by + jj) * kp.gz + st.bz + k);
            float qn = t.x * h.x + t.y * h.y + t.z * h.z;
            gxs += v3(pick(d0, d1, d2, i, 0) * wyj * wzk, wxi * pick(d0, d1, d2, jj, 1) * wzk, wxi * wyj * pick(d0, d1, d2, k, 2)) * qn;
          }
    }
    dq.push(act && !mine && !second, p, lane, flush_queue);
    V3 gx = in.gx + gxs - (kp.inv_dx * s4) * mul_t(in.gC, in.nvel);  // sum_n w_n v_n is the velocity g2p stored in the next slot
    if (act) plane4(gout, kp.EN, 0)[p] = make_float4(gx.x, gx.y, gx.z, 0.f);
  }
  __syncwarp();
  for (int n = lane; n < kTileN; n += 32) {
    int txx = n >> 6, tyy = (n >> 3) & 7, tzz = n & 7;
    float4 t = tg[tile_slot(txx, tyy, tzz)];
    int nx = cg.ox + txx, ny = cg.oy + tyy, nz = cg.oz + tzz;
    if ((t.x != 0.f || t.y != 0.f || t.z != 0.f) && (unsigned)nx < (unsigned)kp.gx && (unsigned)ny < (unsigned)kp.gy && (unsigned)nz < (unsigned)kp.gz)
      red_add_v4(ggrid_v + goff + (nx * kp.gy + ny) * kp.gz + nz, t.x, t.y, t.z, 0.f);
  }
  __syncwarp();
  }
  flush_queue();
  chunks_done(sched, lane);
}

// p2g_grad on tiles
// asynchronous copies of one particle's inputs into its staging slots, in the two groups the adjoint consumes them in
struct P2ggStager {
  const KP &kp;
  const float *cur, *nxt, *gin, *yield;
  const float4 *mat0;
  float *gout;
  float4 *stg;
  int p_next;  // particle of the next round, or -1
  DD_DEV void stage1(int p) const {
    cp_async16(stg, plane4(cur, kp.EN, 0) + p); cp_async16(stg + 32, plane4(cur, kp.EN, 1) + p); cp_async16(stg + 2 * 32, mat0 + p);
    cp_async16(stg + 3 * 32, aux4(nxt, kp.EN, 0) + p); cp_async16(stg + 4 * 32, aux4(nxt, kp.EN, 1) + p); cp_async16(stg + 5 * 32, aux4(nxt, kp.EN, 2) + p);
    cp_async_commit();
  }
  DD_DEV void stage2(int p) const {
    cp_async16(stg + 6 * 32, plane4(cur, kp.EN, 2) + p); cp_async16(stg + 7 * 32, plane4(cur, kp.EN, 3) + p);
    cp_async16(stg + 8 * 32, plane4(cur, kp.EN, 4) + p); cp_async16(stg + 9 * 32, plane4(cur, kp.EN, 5) + p);
    cp_async16(stg + 10 * 32, aux4(nxt, kp.EN, 3) + p); cp_async16(stg + 11 * 32, reinterpret_cast<const float4 *>(nxt + (size_t)25 * kp.EN) + p);
    cp_async16(stg + 12 * 32, plane4(gin, kp.EN, 4) + p); cp_async16(stg + 13 * 32, plane4(gin, kp.EN, 5) + p);
    cp_async16(stg + 14 * 32, plane4(gout, kp.EN, 0) + p);
    float *sc = reinterpret_cast<float *>(stg + 15 * 32);
    cp_async4(sc, cur + (size_t)24 * kp.EN + p); cp_async4(sc + 1, yield + p); cp_async4(sc + 2, gin + (size_t)24 * kp.EN + p);
    cp_async_commit();
  }
  DD_DEV void phase1_done() const { if (p_next >= 0) stage1(p_next); }
  DD_DEV void phase2_done() const { if (p_next >= 0) stage2(p_next); }
};

template <int SVD>
__global__ void __launch_bounds__(32 * kTileWarps, DD_LB_P2GG_TILE) k_p2g_grad_tile(KP kp, SegView sg, const float *__restrict__ cur,
                                                                      const float *__restrict__ nxt, const float4 *__restrict__ mat0,
                                                                      const float *__restrict__ yield, const float4 *__restrict__ ggrid,
                                                                      const float *__restrict__ gin, float *__restrict__ gout, int *sched) {
  extern __shared__ float4 dd_smem[];
  constexpr bool STAGED = SVD == 1;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 *tile = dd_smem + warp * (kTileN + (STAGED ? kStageP2GG * 32 : 0));
  P2ggStager stager{kp, cur, nxt, gin, yield, mat0, gout, tile + kTileN + lane, -1};
  pdl_launch_dependents();
  pdl_wait();
  const int nchunks = sg.cnt[0];
  const int4 *__restrict__ chunks = sg.chunks;
  DD_CHUNK_LOOP(ci) {
    ChunkGeom cg = chunk_geom(chunks[ci], kp);
    if (STAGED) { int p0 = row_pos(cg, 0, lane); stager.stage1(p0); stager.stage2(p0); }
    fill_tile_async(tile, ggrid + (size_t)cg.env * kp.G, kp, cg.ox, cg.oy, cg.oz, lane);
    cp_async_wait_all();
    __syncwarp();
    for (int j = 0; j < cg.R; ++j) {
      if (STAGED) {
        // every lane takes part (idle lanes of a short last row shadow the chunk's first particle and write nothing)
        bool act = lane_on(cg, j, lane);
        stager.p_next = j + 1 < cg.R ? row_pos(cg, j + 1, lane) : -1;
        cp_async_wait_all();
        if (act) p2g_grad_particle<SVD, true, P2ggStager>(kp, row_pos(cg, j, lane), cur, nxt, mat0, yield, ggrid, tile, cg.ox, cg.oy, cg.oz, gin, gout, stager.stg, stager);
        else { stager.phase1_done(); stager.phase2_done(); }
      } else {
        if (!lane_on(cg, j, lane)) continue;
        p2g_grad_particle<SVD, true>(kp, row_pos(cg, j, lane), cur, nxt, mat0, yield, ggrid, tile, cg.ox, cg.oy, cg.oz, gin, gout);
      }
    }
    __syncwarp();
  }
  chunks_done(sched, lane);
}

// g2p on tiles: the 27 node velocities come from a shared-memory copy of the brick's neighbourhood
__global__ void __launch_bounds__(32 * kTileWarps, DD_LB_G2P_TILE) k_g2p_tile(KP kp, SegView sg, const float *__restrict__ cur,
                                                                 float *__restrict__ nxt, const float4 *__restrict__ grid_v, int *sched) {
  extern __shared__ float4 dd_smem[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 *tile = dd_smem + warp * (kTileN + 32), *xst = tile + kTileN + lane;
  pdl_launch_dependents();
  pdl_wait();
  const int nchunks = sg.cnt[0];
  const int4 *__restrict__ chunks = sg.chunks;
  V3 hi = v3(((float)kp.gx - 3.f) * kp.dx, ((float)kp.gy - 3.f) * kp.dx, ((float)kp.gz - 3.f) * kp.dx);
  float lo = kp.gh * kp.dx;
  DD_CHUNK_LOOP(ci) {
  ChunkGeom cg = chunk_geom(chunks[ci], kp);
  const float4 *genv = grid_v + (size_t)cg.env * kp.G;
#ifndef DD_G2P_NO_STAGE_X
  cp_async16(xst, plane4(cur, kp.EN, 0) + row_pos(cg, 0, lane));  // (joins the group of the tile fill)
#endif
  fill_tile_async(tile, genv, kp, cg.ox, cg.oy, cg.oz, lane);
  cp_async_wait_all();
  __syncwarp();
  for (int j = 0; j < cg.R; ++j) {
#ifndef DD_G2P_NO_STAGE_X
    // the position of the lane's particle of the NEXT row is copied into shared memory while this row is gathered (the plain
    // load at the top of the row was 27 % of the kernel's stall samples); idle lanes of a short last row shadow a valid particle
    cp_async_wait_all();
    float4 a = *xst;
    if (j + 1 < cg.R) { cp_async16(xst, plane4(cur, kp.EN, 0) + row_pos(cg, j + 1, lane)); cp_async_commit(); }
    if (!lane_on(cg, j, lane)) continue;
    int p = row_pos(cg, j, lane);
#else
    if (!lane_on(cg, j, lane)) continue;
    int p = row_pos(cg, j, lane);
    float4 a = ldg_stream(plane4(cur, kp.EN, 0) + p);
#endif
    V3 x = v3(a.x, a.y, a.z);
    Stencil st = make_stencil_safe(x, kp);
    float wx[3] = {st.w0.x, st.w1.x, st.w2.x}, wy[3] = {st.w0.y, st.w1.y, st.w2.y}, wz[3] = {st.w0.z, st.w1.z, st.w2.z};
    int tx = st.bx - cg.ox, ty = st.by - cg.oy, tz = st.bz - cg.oz;
    bool in_tile = (unsigned)tx <= 5u && (unsigned)ty <= 5u && (unsigned)tz <= 5u;
    const float4 *g = genv + (st.bx * kp.gy + st.by) * kp.gz + st.bz;
    V3 nv;
    M3 nC;
    if (in_tile) {
      const float4 *trow = tile + (tx << 6 | ty << 3);
      int g0 = tz + 4 * ty + 2 * tx;
      g2p_gather([&](int i, int jj, int k) { return trow[(i << 6 | jj << 3) + ((g0 + 2 * i + 4 * jj + k) & 7)]; }, wx, wy, wz, st.fx, kp.inv_dx * 4.f, nv, nC);
    } else {
      g2p_gather([&](int i, int jj, int k) { return __ldg(g + (i * kp.gy + jj) * kp.gz + k); }, wx, wy, wz, st.fx, kp.inv_dx * 4.f, nv, nC);
    }
    V3 t = x + nv * kp.dt;
    store_xvc(nxt, kp.EN, p, v3(fmaxf(fminf(t.x, hi.x), lo), fmaxf(fminf(t.y, hi.y), lo), fmaxf(fminf(t.z, hi.z), lo)), nv, nC);
  }
  __syncwarp();
  }
  chunks_done(sched, lane);
}

// ---- grid kernels restricted to the active bricks (3x3x3-brick neighbourhood of every occupied brick) -----------------
DD_DEV int brick_node(int brick, int local, const KP &kp, int &env, int &gx_, int &gy_, int &gz_) {
  int nby = kp.gy >> 2, nbz = kp.gz >> 2, NB = (kp.gx >> 2) * nby * nbz;
  env = brick / NB;
  int b = brick - env * NB;
  gx_ = (b / (nby * nbz)) * 4 + (local >> 4); gy_ = ((b / nbz) % nby) * 4 + ((local >> 2) & 3); gz_ = (b % nbz) * 4 + (local & 3);
  return env * kp.G + (gx_ * kp.gy + gy_) * kp.gz + gz_;
}
// (the active-brick count lives in device memory, cnt[1]: launches use a fixed upper-bound grid and stride over the list)
__global__ void __launch_bounds__(kT) k_zero_bricks(KP kp, const int *__restrict__ cnt, const int *__restrict__ active, float4 *a, float4 *b) {
  pdl_launch_dependents();
  pdl_wait();
  int total = cnt[1] * 64;
  for (int t = blockIdx.x * kT + threadIdx.x; t < total; t += gridDim.x * kT) {
    int env, x, y, z;
    int node = brick_node(active[t >> 6], t & 63, kp, env, x, y, z);
    a[node] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b) b[node] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// Four bricks (64 nodes each) per block.  Each 64-thread group first copies its environment's body poses into shared
// memory and builds a 64-bit mask of the bodies whose activation sphere reaches the brick's bounding box at all; the
// per-node loop then only visits those (typically 0-3 of the 19 primitives) in index order.
struct GridSm {
  float4 pos[4][64], rot[4][64], npos[4][64], nrot[4][64], tfsr[64], args[64];
  float cull[64];
  unsigned long long cand[4];
};
DD_DEV void stage_shapes(GridSm &sm, const KP &kp, const BodyTables &bt) {
  if ((int)threadIdx.x < kp.nb) { sm.tfsr[threadIdx.x] = bt.tfsr[threadIdx.x]; sm.args[threadIdx.x] = bt.args[threadIdx.x]; sm.cull[threadIdx.x] = bt.cull[threadIdx.x]; }
}
DD_DEV unsigned long long stage_bodies(GridSm &sm, const KP &kp, const BodyTables &bt, bool valid, int brick, int grp, int g, BodyTables &view) {
  if (threadIdx.x < 4) sm.cand[threadIdx.x] = 0ull;
  __syncthreads();
  int nby = kp.gy >> 2, nbz = kp.gz >> 2, NB = (kp.gx >> 2) * nby * nbz;
  int env = valid ? brick / NB : 0, bb = valid ? brick - env * NB : 0;
  if (valid && g < kp.nb) {
    int pb = env * kp.nb + g;
    float4 p = bt.pos[pb];
    sm.pos[grp][g] = p; sm.rot[grp][g] = bt.rot[pb]; sm.npos[grp][g] = bt.npos[pb]; sm.nrot[grp][g] = bt.nrot[pb];
    // Exact signed distances are 1-Lipschitz: if the distance at the centre of the brick's node box [lo, lo + 3 dx]^3 exceeds
    // the activation band by more than the half diagonal of the box, no node of the brick can touch this body.
    float lx = (float)((bb / (nby * nbz)) * 4) * kp.dx, ly = (float)(((bb / nbz) % nby) * 4) * kp.dx, lz = (float)((bb % nbz) * 4) * kp.dx, h = 1.5f * kp.dx;
    Q4 tf = q4f(bt.tfsr[g]);
    float dist = shape_sdf(tf, q4f(bt.args[g]), qrot(qconj(q4f(sm.rot[grp][g])), v3(lx + h, ly + h, lz + h) - v3(p.x, p.y, p.z)));
    float band = tf.y > 0.f ? 2.5f / tf.y : 0.f;
    if (dist <= band + 2.5981f * kp.dx * 1.001f + 1e-6f) atomicOr(&sm.cand[grp], 1ull << g);
  }
  __syncthreads();
  // body index inside the kernels is env*nb + b: bias the shared-memory pointers so the same expression lands on [b]
  view.pos = sm.pos[grp] - env * kp.nb; view.rot = sm.rot[grp] - env * kp.nb; view.npos = sm.npos[grp] - env * kp.nb; view.nrot = sm.nrot[grp] - env * kp.nb;
  view.tfsr = sm.tfsr; view.args = sm.args; view.cull = sm.cull;
  return sm.cand[grp];
}

// zero_next: the (distinct) scatter target of the NEXT substep, cleared here so that no separate zeroing pass is needed;
// zero_self: clear this substep's (m, mv) after use (forward-only mode with a single grid buffer)
// Brick checkpoints (BrickCk): when one dense grid pair per substep does not fit in HBM (many environments, long rollouts), the
// forward pass keeps (mv, m) and v_out of the ACTIVE bricks only, 64 nodes per brick, indexed by the brick's position in the
// segment's active list (the list only grows inside a segment, so positions are stable); `cap` bricks per substep.  The adjoint
// reads (mv, m) from there and writes v_out back into the dense working grid one substep ahead (see k_grid_grad_b).
struct BrickCk {
  float4 *m, *v;   // this substep's bricks (already offset), or null
  int *n;          // bricks stored for this substep
  int cap;
  int *status;     // bit 0: a substep had more active bricks than `cap`
};
__global__ void __launch_bounds__(kT, 4) k_grid_b(KP kp, const int *__restrict__ cnt, const int *__restrict__ active, float4 *__restrict__ grid,
                                               float4 *__restrict__ grid_v, BodyTables bt, float4 *__restrict__ zero_next, int zero_self, BrickCk ck) {
  __shared__ GridSm sm;
  pdl_launch_dependents();
  stage_shapes(sm, kp, bt);  // (shape tables are never written by a kernel)
  pdl_wait();
  const int nactive = cnt[1], grp = threadIdx.x >> 6, g = threadIdx.x & 63;
  if (ck.m && blockIdx.x == 0 && threadIdx.x == 0) {
    *ck.n = min(nactive, ck.cap);
    if (nactive > ck.cap) atomicOr(ck.status, 1);
  }
  for (int blk = blockIdx.x; blk * 4 < nactive; blk += gridDim.x) {
    if (blk != (int)blockIdx.x) __syncthreads();  // the previous iteration is done with the staged poses
    int t = blk * kT + threadIdx.x;
    bool valid = t < nactive * 64;
    int brick = valid ? active[t >> 6] : 0;
    BodyTables view;
    unsigned long long cand = stage_bodies(sm, kp, bt, valid, brick, grp, g, view);
    if (!valid) continue;
    int env, gx_, gy_, gz_;
    int node = brick_node(brick, g, kp, env, gx_, gy_, gz_);
    float4 mm = grid[node];
    if (zero_next) zero_next[node] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (zero_self) grid[node] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 *ckv = nullptr;
    if (ck.m && (t >> 6) < ck.cap) { ck.m[t] = mm; ckv = ck.v + t; }
    if (!(mm.w > 1e-12)) {
      grid_v[node] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ckv) *ckv = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    V3 v = v3(mm.x, mm.y, mm.z) * (1.f / mm.w) + kp.dt * v3(kp.g0, kp.g1, kp.g2);
    V3 gx = v3((float)gx_, (float)gy_, (float)gz_) * kp.dx;
    // (two passes -- contact mask first, then the node's own stages, as in grid_grad_body -- measured: no change, 21 us at the 10k-particle scene)
    for (unsigned long long c = cand; c; c &= c - 1ull) {
      int b = __ffsll((long long)c) - 1, pb = env * kp.nb + b;
      Hit h;
      Q4 bq = q4f(view.rot[pb]), tfsr = q4f(view.tfsr[b]), sargs = q4f(view.args[b]);
      if (contact_geom(gx, v3f(view.pos[pb]), bq, tfsr, sargs, view.cull[b], h))
        v = contact_apply(gx, v, bq, v3f(view.npos[pb]), q4f(view.nrot[pb]), tfsr, sargs, kp.dt, h);
    }
    v = apply_bc(v, gx_, gy_, gz_, kp);
    grid_v[node] = make_float4(v.x, v.y, v.z, 0.f);
    if (ckv) *ckv = make_float4(v.x, v.y, v.z, 0.f);
  }
}
// v_out of a substep from its brick checkpoint back into the dense working grid (first adjoint substep of a range or of a segment;
// in between k_grid_grad_b does it for the substep below on its way)
__global__ void __launch_bounds__(kT) k_restore_bricks(KP kp, const int *__restrict__ cnt, const int *__restrict__ active, const float4 *__restrict__ ck_v,
                                                       const int *__restrict__ ck_n, float4 *__restrict__ grid_v) {
  pdl_launch_dependents();
  pdl_wait();
  const int total = cnt[1] * 64, stored = *ck_n * 64;
  for (int t = blockIdx.x * kT + threadIdx.x; t < total; t += gridDim.x * kT) {
    int env, x, y, z;
    grid_v[brick_node(active[t >> 6], t & 63, kp, env, x, y, z)] = t < stored ? ck_v[t] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(kT, DD_LB_GRID_GRAD) k_grid_grad_b(KP kp, const int *__restrict__ cnt, const int *__restrict__ active, float4 *__restrict__ grid,
                                                    float4 *__restrict__ ggrid_v, float4 *__restrict__ ggrid, BodyTables bt, float4 *gpos,
                                                    float4 *grot, float4 *gnpos, float4 *gnrot, int zero_m, BrickCk ck, BrickCk below, float4 *__restrict__ grid_v) {
  // ck: this substep's brick checkpoint ((mv, m) is read from it instead of `grid`); below: the checkpoint of the substep the
  // adjoint visits next, whose v_out this kernel puts back into the dense `grid_v` (nobody reads grid_v between the gather of
  // this substep, which has completed, and the gather of the next one)
  __shared__ GridSm sm;
  pdl_launch_dependents();
  stage_shapes(sm, kp, bt);  // (shape tables are never written by a kernel)
  pdl_wait();
  const int nactive = cnt[1], grp = threadIdx.x >> 6, g = threadIdx.x & 63;
  const int stored = ck.m ? *ck.n * 64 : 0, stored_below = below.v ? *below.n * 64 : 0;
  for (int blk = blockIdx.x; blk * 4 < nactive; blk += gridDim.x) {
    if (blk != (int)blockIdx.x) __syncthreads();
    int t = blk * kT + threadIdx.x;
    bool inr = t < nactive * 64;
    int brick = inr ? active[t >> 6] : 0;
    BodyTables view;
    unsigned long long cand = stage_bodies(sm, kp, bt, inr, brick, grp, g, view);
    int env = 0, x = 0, y = 0, z = 0, node = 0;
    if (inr) node = brick_node(brick, g, kp, env, x, y, z);
    if (inr && below.v) grid_v[node] = t < stored_below ? below.v[t] : make_float4(0.f, 0.f, 0.f, 0.f);
    grid_grad_body(kp, node, inr, env, x, y, z, grid, ggrid_v, ggrid, view, gpos, grot, gnpos, gnrot, cand, true, zero_m != 0, t < stored ? ck.m + t : nullptr, ck.m != nullptr);
  }
}

// ---- layout conversion (original AoS order <-> sorted planes) --------------------------------------------------------
__global__ void k_pack(int EN, const int *__restrict__ perm, const float *__restrict__ x, const float *__restrict__ v, const float *__restrict__ F,
                       const float *__restrict__ C, float *slot) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= EN) return;
  int s = perm[i];
  if (x && v && C) {
    store_xvc(slot, EN, i, ld_v3(x, s), ld_v3(v, s), ld_m3(C, s));
  }
  if (F) {
    store_F(slot, EN, i, ld_m3(F, s));
    store_q(slot, EN, i, make_float4(0.f, 0.f, 0.f, 1.f));  // cold start for the first SVD from this state
  }
}
__global__ void k_unpack(int EN, const int *__restrict__ perm, const float *__restrict__ slot, float *x, float *v, float *F, float *C) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= EN) return;
  int s = perm[i];
  XVC a = load_xvc(slot, EN, i);
  if (x) st_v3(x, s, a.x);
  if (v) st_v3(v, s, a.v);
  if (C) st_m3(C, s, a.C);
  if (F) st_m3(F, s, load_F(slot, EN, i));
}
// slot += packed(values) for whichever of gx, gv, gF, gC is given
__global__ void k_add_grad(int EN, const int *__restrict__ perm, const float *__restrict__ gx, const float *__restrict__ gv,
                           const float *__restrict__ gF, const float *__restrict__ gC, float *slot) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= EN) return;
  int s = perm[i];
  XVC a = load_xvc(slot, EN, i);
  if (gx) a.x += ld_v3(gx, s);
  if (gv) a.v += ld_v3(gv, s);
  if (gC) a.C += ld_m3(gC, s);
  store_xvc(slot, EN, i, a.x, a.v, a.C);
  if (gF) store_F(slot, EN, i, load_F(slot, EN, i) + ld_m3(gF, s));
}
__global__ void k_pack_mat(int EN, const int *__restrict__ perm, const float *__restrict__ mass, const float *__restrict__ vol,
                           const float *__restrict__ mly, float4 *mat0, float *yield) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= EN) return;
  int s = perm[i];
  mat0[i] = make_float4(mass[s], vol[s], mly[3 * s], mly[3 * s + 1]);
  yield[i] = mly[3 * s + 2];
}
// sort key of a particle: environment-major, then 4x4x4-cell brick (x-major like the grid), then cell inside the brick
// x_aos: positions in the caller's order (dd_sim_set_state), or x_plane: plane 0 of a checkpoint slot (device re-sort)
__global__ void k_sort_keys(KP kp, const float *__restrict__ x_aos, const float4 *__restrict__ x_plane, unsigned *__restrict__ keys, int *__restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kp.EN) return;
  V3 x;
  if (x_aos) x = ld_v3(x_aos, i);
  else { float4 a = x_plane[i]; x = v3(a.x, a.y, a.z); }
  int cx = clampi((int)floorf(x.x * kp.inv_dx - 0.5f), 0, kp.gx - 1), cy = clampi((int)floorf(x.y * kp.inv_dx - 0.5f), 0, kp.gy - 1),
      cz = clampi((int)floorf(x.z * kp.inv_dx - 0.5f), 0, kp.gz - 1);
  int nby = (kp.gy + 3) >> 2, nbz = (kp.gz + 3) >> 2;
  unsigned brick = ((cx >> 2) * nby + (cy >> 2)) * nbz + (cz >> 2);
  // within the brick: bank group of the home cell (tile coordinates = local + 1) first, then the 8 cells of that group
  int lx = cx & 3, ly = cy & 3, lz = cz & 3;
  unsigned cell = (unsigned)tile_group(lx + 1, ly + 1, lz + 1) << 3 | (unsigned)(lx << 1 | ly >> 1);
  keys[i] = (unsigned)(i / kp.N) * (unsigned)kp.G + (brick << 6 | cell);
  idx[i] = i;
}

// On the sorted keys: flags the first particle of every brick, and -- once per occupied cell, the active region depends on the
// cell only -- marks the bricks reached by that cell's stencil [base, base + 2] padded by one node on each side.
__global__ void k_mark_heads(KP kp, const unsigned *__restrict__ keys, char *__restrict__ flags, int *__restrict__ active_flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kp.EN) return;
  unsigned key = keys[i], prev = i ? keys[i - 1] : ~key;
  flags[i] = (i == 0) || (key >> 6) != (prev >> 6);
  if (key == prev) return;
  int nbx = kp.gx >> 2, nby = kp.gy >> 2, nbz = kp.gz >> 2, NB = nbx * nby * nbz;
  int env = (int)(key / (unsigned)kp.G), within = (int)(key - (unsigned)env * (unsigned)kp.G), brick = within >> 6, code = within & 63;
  // invert the in-brick code of k_sort_keys: group g = (lz + 4 ly + 2 lx + 7) mod 8, sub = lx << 1 | ly >> 1
  int g = code >> 3, lx = (code >> 1) & 3, t = (g - 2 * lx - 7) & 7, lz = t & 3, ly = (code & 1) << 1 | t >> 2;
  int cx = (brick / (nby * nbz)) * 4 + lx, cy = ((brick / nbz) % nby) * 4 + ly, cz = (brick % nbz) * 4 + lz;
  int b0x = clampi(cx, 0, kp.gx - 3), b0y = clampi(cy, 0, kp.gy - 3), b0z = clampi(cz, 0, kp.gz - 3);
  int x0 = max(b0x - 1, 0) >> 2, x1 = min((b0x + 3) >> 2, nbx - 1), y0 = max(b0y - 1, 0) >> 2, y1 = min((b0y + 3) >> 2, nby - 1),
      z0 = max(b0z - 1, 0) >> 2, z1 = min((b0z + 3) >> 2, nbz - 1);
  for (int x_ = x0; x_ <= x1; ++x_)
    for (int y_ = y0; y_ <= y1; ++y_)
      for (int z_ = z0; z_ <= z1; ++z_) {
        int *fl = active_flag + (size_t)env * NB + (x_ * nby + y_) * nbz + z_;
        if (!*fl) *fl = 1;
      }
}
// one thread per occupied brick: split its particles into `nsub` chunks.
// Chunk c takes the cell-sorted ranks r = c (mod nsub) of the brick -- a thinned copy of the whole brick, so that the
// lanes of a round still sit in different cells -- and owns the storage range after chunks 0..c-1.
// (launched over an upper bound of threads; the number of occupied bricks is counters[2], written by the stream compaction)
__global__ void k_make_chunks(KP kp, const int *__restrict__ head_pos, const unsigned *__restrict__ keys, int chunk_max,
                              int4 *__restrict__ chunks, int4 *__restrict__ chunk_src, int *__restrict__ counters) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int nbricks = counters[2];
  if (k >= nbricks) return;
  int start = head_pos[k], end = k + 1 < nbricks ? head_pos[k + 1] : kp.EN, cnt = end - start;
  int brick = (int)(keys[start] >> 6);
  int nsub = (cnt + chunk_max - 1) / chunk_max;
  int c0 = atomicAdd(&counters[0], nsub);
  int s = start;
  for (int c = 0; c < nsub; ++c) {
    int n_c = (cnt - c + nsub - 1) / nsub;
    chunks[c0 + c] = make_int4(brick, s, n_c, 0);
    chunk_src[c0 + c] = make_int4(start, cnt, c, nsub);
    s += n_c;
  }
}
// Storage order: every chunk is an R-row x 32-column table stored row-major; a warp processes one row (round) at a time,
// lane l = column l.  Columns l with l mod 8 = g are filled, in cell order, with the chunk's particles whose home cell
// lies in bank group g (see tile_slot), so the eight lanes of a quarter-warp sit in eight different groups and the tile
// accesses of a round are free of bank conflicts; the four columns of a group take consecutive quarters of the group's
// cell-sorted list, so the lanes of a round also sit in different cells (no read-modify-write collisions).  Groups are
// not equally populated: what does not fit into a group's own columns spills into the free slots of the others (a few
// percent of the particles, costing at most one extra wavefront where they sit).  One warp per chunk.
__global__ void k_interleave(KP kp, const int *__restrict__ counters, int4 *__restrict__ chunks, const int4 *__restrict__ chunk_src, const unsigned *__restrict__ keys_sorted,
                             const int *__restrict__ perm_in, int *__restrict__ perm_out, int *__restrict__ spos) {
  int ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (ci >= counters[0]) return;
  const unsigned full = 0xffffffffu;
  ChunkGeom cg = chunk_geom(chunks[ci], kp);
  int4 src = chunk_src[ci];  // (brick start, brick count, c, nsub): item r of the chunk is sorted rank src.x + src.z + r * src.w
  int cnt = cg.cnt, R = cg.R;
  // first item of every group (lanes 0..7; lane 8 holds cnt)
  int gs = cnt;
  if (lane < 8) {
    int lo = 0, hi = cnt;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if ((int)((keys_sorted[src.x + src.z + mid * src.w] >> 3) & 7u) >= lane) hi = mid; else lo = mid + 1;
    }
    gs = lo;
  }
  int g_cnt = __shfl_down_sync(full, gs, 1) - gs;                        // valid on lanes 0..7
  // The short last row: its `last` slots go to the columns whose group still has particles after R - 1 full rows (one more per
  // quarter-warp, up to four per group), so that the idle lanes of that row absorb the imbalance between the groups instead of
  // spills into foreign columns; what is left goes to the lowest free columns.
  int need = __shfl_sync(full, g_cnt, lane & 7) - 4 * (R - 1);           // of this column's group
  int want = need > (lane >> 3) ? 1 : 0, pre = want;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(full, pre, o); if (lane >= o) pre += t; }
  bool bit = want && pre - want < cg.last;
  int given = __popc(__ballot_sync(full, bit)), nb2 = bit ? 0 : 1, pre2 = nb2;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(full, pre2, o); if (lane >= o) pre2 += t; }
  bit = bit || (nb2 && pre2 - nb2 < cg.last - given);
  unsigned lastmask = __ballot_sync(full, bit);
  if (lane == 0) chunks[ci].w = (int)lastmask;
  int cap = R - 1 + (bit ? 1 : 0);                                       // capacity of column `lane`
  int cap1 = __shfl_sync(full, cap, (lane & 7) + 8), cap2 = __shfl_sync(full, cap, (lane & 7) + 16), cap3 = __shfl_sync(full, cap, (lane & 7) + 24);
  int cap0 = __shfl_sync(full, cap, lane & 7);
  int capg = cap0 + cap1 + cap2 + cap3;                                  // capacity of group (lane & 7)
  int over = lane < 8 ? max(g_cnt - capg, 0) : 0, spill_base = over;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) { int t = __shfl_up_sync(full, spill_base, o); if (lane >= o) spill_base += t; }
  spill_base -= over;                                                    // exclusive prefix over groups (lanes 0..7)
  int gc = __shfl_sync(full, g_cnt, lane & 7);                           // population of this column's group
  int before = (lane >> 3) == 0 ? 0 : (lane >> 3) == 1 ? cap0 : (lane >> 3) == 2 ? cap0 + cap1 : cap0 + cap1 + cap2;
  int used = min(max(gc - before, 0), cap), fr = cap - used, free_base = fr;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(full, free_base, o); if (lane >= o) free_base += t; }
  free_base -= fr;                                                       // exclusive prefix of free slots over columns
  for (int r0 = 0; r0 < cnt; r0 += 32) {
    int r = r0 + lane;
    bool live = r < cnt;
    int from = src.x + src.z + (live ? r : 0) * src.w;
    int g = (int)((keys_sorted[from] >> 3) & 7u);
    int c0 = __shfl_sync(full, cap0, g), c1 = __shfl_sync(full, cap1, g), c2 = __shfl_sync(full, cap2, g);
    int sb = __shfl_sync(full, spill_base, g), ov = __shfl_sync(full, over, g);
    // The first `ov` items of the group spill.  They belong to the group's first cell, whose other particles sit at the
    // top of column g, while free slots are at the bottom of other columns: a spilled particle does not share its round
    // with a particle of its own cell.
    int idx = r - __shfl_sync(full, gs, g) - ov;
    int col, row;
    bool spilled = idx < 0;
    if (idx < c0) { col = g; row = idx; }
    else if (idx < c0 + c1) { col = g + 8; row = idx - c0; }
    else if (idx < c0 + c1 + c2) { col = g + 16; row = idx - c0 - c1; }
    else { col = g + 24; row = idx - c0 - c1 - c2; }
    int sidx = sb + idx + ov;
    for (int l = 0; l < 32; ++l) {  // spilled items: the sidx-th free slot, columns in order
      int fb = __shfl_sync(full, free_base, l), ff = __shfl_sync(full, fr, l), uu = __shfl_sync(full, used, l);
      if (spilled && sidx >= fb && sidx < fb + ff) { col = l; row = uu + sidx - fb; }
    }
    if (live) {
      int pos = cg.start + 32 * row + (row < R - 1 ? col : __popc(lastmask & ((1u << col) - 1u)));
      perm_out[pos] = perm_in[from];
      spos[from] = pos;  // sorted rank -> storage position, for the flat gather kernels
    }
  }
}

// largest chunks first: the persistent tiled kernels hand chunks out in list order, so the launch ends on the small ones
// (the sort runs over the whole capacity of the list: entries past the live count get size 0 and sort to the end)
__global__ void k_chunk_keys(int cap, const int *__restrict__ counters, const int4 *__restrict__ chunks, int *__restrict__ keys, int *__restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  keys[i] = i < counters[0] ? chunks[i].z : 0;
  idx[i] = i;
}
__global__ void k_chunk_permute(int cap, const int *__restrict__ counters, const int *__restrict__ order, const int4 *__restrict__ chunks, const int4 *__restrict__ chunk_src,
                                int4 *__restrict__ chunks_out, int4 *__restrict__ chunk_src_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap || i >= counters[0]) return;
  chunks_out[i] = chunks[order[i]];
  chunk_src_out[i] = chunk_src[order[i]];
}
__global__ void k_pad4(int n, const float *__restrict__ src, int w, float4 *dst) {  // (n, w<=4) floats -> float4
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < w; ++k) t[k] = src[(size_t)i * w + k];
  dst[i] = make_float4(t[0], t[1], t[2], t[3]);
}
__global__ void k_unpad4(int n, const float4 *__restrict__ src, int w, float *dst, int accumulate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 t = src[i];
  float a[4] = {t.x, t.y, t.z, t.w};
  for (int k = 0; k < w; ++k) dst[(size_t)i * w + k] = accumulate ? dst[(size_t)i * w + k] + a[k] : a[k];
}
__global__ void k_add4(int n, const float *__restrict__ src, int w, float4 *dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < w; ++k) t[k] = src[(size_t)i * w + k];
  float4 d = dst[i];
  dst[i] = make_float4(d.x + t[0], d.y + t[1], d.z + t[2], d.w + t[3]);
}

// ---- device-side re-sort at a segment boundary ---------------------------------------------------------------------------
// dst[i] = src[from[i]] for the state planes (x, v, C | F | quaternion of V): the head copy of a boundary slot in the new order
__global__ void k_gather_planes(int EN, const int *__restrict__ from, const float *__restrict__ src, float *__restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= EN) return;
  int j = from[i];
#pragma unroll
  for (int k = 0; k < 6; ++k) plane4(dst, EN, k)[i] = plane4(src, EN, k)[j];
  dst[(size_t)24 * EN + i] = src[(size_t)24 * EN + j];
  store_q(dst, EN, i, reinterpret_cast<const float4 *>(src + (size_t)25 * EN)[j]);
}
// gradient planes of a boundary state from the new order back into the previous one: dst[from[i]] = src[i]
__global__ void k_scatter_grad(int EN, const int *__restrict__ from, const float *__restrict__ src, float *__restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= EN) return;
  int j = from[i];
#pragma unroll
  for (int k = 0; k < 6; ++k) plane4(dst, EN, k)[j] = plane4(src, EN, k)[i];
  dst[(size_t)24 * EN + j] = src[(size_t)24 * EN + i];
}
__global__ void k_compose_perm(int EN, const int *__restrict__ from, const int *__restrict__ perm_prev, int *__restrict__ perm_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < EN) perm_out[i] = perm_prev[from[i]];
}
__global__ void k_copy_poses(int n, const float4 *__restrict__ ps, const float4 *__restrict__ rs, float4 *__restrict__ pd, float4 *__restrict__ rd) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { pd[i] = ps[i]; rd[i] = rs[i]; }
}

// compute_dist (integrator.cu:188-237) on the engine layout, fused with the observation GradModel.get_obs assembles from it
// (mpm/torch_wrapper.py:46-66): row `perm[p]` of `out` (ORIGINAL particle order, `ld` floats per row) receives
// [x | v |] dist_0 .. dist_{nb-1}; WITH_XV = false writes the distances only (ld = nb: dd_sim_compute_dist).
template <bool WITH_XV>
__global__ void __launch_bounds__(kT) k_obs(KP kp, const int *__restrict__ perm, const float *__restrict__ slot, BodyTables bt, float *__restrict__ out, int ld) {
  // Rows of up to 32 floats go through shared memory: a lane computes the row of ITS particle, then the warp writes the 32 rows one
  // after the other, lane l the l-th float -- contiguous 4 ld-byte segments instead of ld scalar stores per lane into 32 different
  // rows of the caller's particle order (1M particles x 25 floats: 164 -> 75 us).
  __shared__ float rows[kT / 32][32][33];
  int p = blockIdx.x * kT + threadIdx.x;
  const bool live = p < kp.EN;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool staged = ld <= 32;
  float *srow = rows[w][lane];
  int target = 0;
  if (live) {
    int env = p / kp.N;
    float4 a = plane4(slot, kp.EN, 0)[p];
    V3 xp = v3(a.x, a.y, a.z);
    target = perm[p];
    float *o = staged ? srow : out + (size_t)target * ld;
    if (WITH_XV) {
      float4 b = plane4(slot, kp.EN, 1)[p];
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y;
      o += 6;
    }
    for (int b = 0; b < kp.nb; ++b) {
      int pb = env * kp.nb + b;
      o[b] = shape_sdf(q4f(bt.tfsr[b]), q4f(bt.args[b]), xform_inv(v3f(bt.pos[pb]), q4f(bt.rot[pb]), xp));
    }
  }
  if (!staged) return;
  __syncwarp();
  const unsigned livemask = __ballot_sync(0xffffffffu, live);
  for (int r = 0; r < 32; ++r) {
    int trg = __shfl_sync(0xffffffffu, target, r);
    if ((livemask >> r & 1u) && lane < ld) out[(size_t)trg * ld + lane] = rows[w][r][lane];
  }
}
// Adjoint of k_obs: row perm[p] of `g` (ld floats) holds [gx | gv |] gdist.  gx, gv are added to the gradient slot; the distance
// gradients go back to the particle position and the body poses.  A body none of whose 32 distance gradients in the warp is
// non-zero is skipped (observation gradients of a rollout are mostly exact zeros: torch hands zeros_like for every past
// observation), the 7 pose-gradient components of a body are reduced with a transposing butterfly (9 shuffles) into per-block
// shared-memory sums, and a block issues one global atomic per touched (body, component) at its end.
constexpr int kMaxBodiesObs = 64;
template <bool WITH_XV>
__global__ void __launch_bounds__(kT) k_obs_grad(KP kp, const int *__restrict__ perm, const float *__restrict__ slot, BodyTables bt, const float *__restrict__ g, int ld,
                                                 float *gslot, float4 *gpos, float4 *grot) {
  __shared__ float acc[kMaxBodiesObs * 8];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int p = blockIdx.x * kT + threadIdx.x;
  const bool live = p < kp.EN;
  const int first = blockIdx.x * kT, last = min(first + kT, kp.EN) - 1;
  const bool block_one_env = first / kp.N == last / kp.N && kp.nb <= kMaxBodiesObs;  // sums of the whole block belong to one environment
  for (int i = threadIdx.x; i < kp.nb * 8 && block_one_env; i += kT) acc[i] = 0.f;
  __syncthreads();
  const int env = live ? p / kp.N : last / kp.N;
  const float *row = g + (size_t)(live ? perm[p] : 0) * ld;  // (staging the rows through shared memory as k_obs does: 87 -> 99 us at 1M particles)
  V3 xp = vzero(), g_x = vzero(), g_v = vzero();
  if (live) {
    float4 a = plane4(slot, kp.EN, 0)[p];
    xp = v3(a.x, a.y, a.z);
    if (WITH_XV) { g_x = v3(row[0], row[1], row[2]); g_v = v3(row[3], row[4], row[5]); row += 6; }
  }
  const int env0 = __shfl_sync(full, env, 0);
  const bool warp_one_env = __all_sync(full, env == env0);
  for (int b = 0; b < kp.nb; ++b) {
    float gd = live ? row[b] : 0.f;
    if (!__any_sync(full, gd != 0.f)) continue;
    int pb = env * kp.nb + b;
    V3 bx = v3f(bt.pos[pb]);
    Q4 bq = q4f(bt.rot[pb]), tfsr = q4f(bt.tfsr[b]), sargs = q4f(bt.args[b]);
    V3 g_bx = vzero();
    Q4 g_bq;
    g_bq.w = g_bq.x = g_bq.y = g_bq.z = 0.f;
    if (gd != 0.f) xform_inv_adj(bx, bq, xp, shape_grad(tfsr, sargs, xform_inv(bx, bq, xp)) * gd, g_bx, g_bq, g_x);
    float r[8] = {g_bx.x, g_bx.y, g_bx.z, g_bq.w, g_bq.x, g_bq.y, g_bq.z, 0.f};
    if (warp_one_env) {
#pragma unroll
      for (int half = 4, off = 16; half >= 1; half >>= 1, off >>= 1) {  // afterwards lanes 4 c .. 4 c + 3 hold partial sums of component c
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          float send = hi ? r[i] : r[i + half], keep = hi ? r[i + half] : r[i];
          r[i] = keep + __shfl_xor_sync(full, send, off);
        }
      }
      r[0] += __shfl_xor_sync(full, r[0], 2);
      r[0] += __shfl_xor_sync(full, r[0], 1);
      const int comp = lane >> 2;
      if ((lane & 3) == 0 && comp < 7 && r[0] != 0.f) {
        if (block_one_env) atomicAdd(&acc[b * 8 + comp], r[0]);
        else atomicAdd(comp < 3 ? &gpos[env0 * kp.nb + b].x + comp : &grot[env0 * kp.nb + b].x + (comp - 3), r[0]);
      }
    } else if (gd != 0.f) {  // a warp that straddles two environments (N not a multiple of 32)
      atomicAdd(&gpos[pb].x, r[0]); atomicAdd(&gpos[pb].y, r[1]); atomicAdd(&gpos[pb].z, r[2]);
      atomicAdd(&grot[pb].x, r[3]); atomicAdd(&grot[pb].y, r[4]); atomicAdd(&grot[pb].z, r[5]); atomicAdd(&grot[pb].w, r[6]);
    }
  }
  if (live) {
    float4 *o0 = plane4(gslot, kp.EN, 0) + p;
    float4 t = *o0;
    *o0 = make_float4(t.x + g_x.x, t.y + g_x.y, t.z + g_x.z, t.w + g_v.x);
    if (WITH_XV) {
      float4 *o1 = plane4(gslot, kp.EN, 1) + p;
      float4 u = *o1;
      *o1 = make_float4(u.x + g_v.y, u.y + g_v.z, u.z, u.w);
    }
  }
  __syncthreads();
  if (block_one_env) {
    const int e = first / kp.N;
    for (int i = threadIdx.x; i < kp.nb * 8; i += kT) {
      int b = i >> 3, comp = i & 7;
      float v = acc[i];
      if (comp < 7 && v != 0.f) atomicAdd(comp < 3 ? &gpos[e * kp.nb + b].x + comp : &grot[e * kp.nb + b].x + (comp - 3), v);
    }
  }
}

// particle2mass (integrator.cu:239-310) on the engine layout: density-grid observation of object `id` (-1 = all) and its
// adjoint into the gradient slot.  `ids` is in ORIGINAL particle order (may be null when id == -1).
__global__ void __launch_bounds__(kT) k_grid_mass(KP kp, const int *__restrict__ perm, const float *__restrict__ slot, const float4 *__restrict__ mat0,
                                                  const int *__restrict__ ids, int id, float *grid_m, const float *__restrict__ grid_m_grad,
                                                  float *gslot, int need_grad) {
  int p = blockIdx.x * kT + threadIdx.x;
  if (p >= kp.EN) return;
  if (id != -1 && ids[perm[p]] != id) return;
  float4 a = plane4(slot, kp.EN, 0)[p];
  Stencil st = make_stencil_safe(v3(a.x, a.y, a.z), kp);
  V3 d0, d1, d2;
  stencil_dw(st, kp.inv_dx, d0, d1, d2);
  float m = mat0[p].x;
  size_t goff = (size_t)(p / kp.N) * kp.G;
  V3 g = vzero();
#pragma unroll 1
  for (int i = 0; i < 3; ++i)
#pragma unroll 1
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float wx = pick(st.w0, st.w1, st.w2, i, 0), wy = pick(st.w0, st.w1, st.w2, j, 1), wz = pick(st.w0, st.w1, st.w2, k, 2);
        size_t node = goff + ((st.bx + i) * kp.gy + st.by + j) * kp.gz + st.bz + k;
        if (need_grad) g += v3(pick(d0, d1, d2, i, 0) * wy * wz, wx * pick(d0, d1, d2, j, 1) * wz, wx * wy * pick(d0, d1, d2, k, 2)) * (grid_m_grad[node] * m);
        else atomicAdd(&grid_m[node], m * (wx * wy * wz));
      }
  if (need_grad) {
    float4 *o = plane4(gslot, kp.EN, 0) + p;
    float4 t = *o;
    *o = make_float4(t.x + g.x, t.y + g.y, t.z + g.z, t.w);
  }
}

inline int nblk(long long n) { return (int)((n + kT - 1) / kT); }

}  // namespace

// ======================================================================================================= host side
// Storage model.  A rollout of max_steps substeps is cut into SEGMENTS of `resort_interval` substeps (one segment when 0).
// All states of a segment are stored in one particle order -- the cell order of the segment's first state -- described by
// device-resident tables (Segment).  The first state of segments 1.. is therefore stored twice: as the last slot of the previous
// segment (old order, "tail" copy) and, re-sorted on the device, as the first slot of its own segment ("head" copy); the
// adjoint permutes the gradient of that state back once per boundary.  Every count the kernels need (chunks, active bricks)
// lives in device memory, so building a segment never synchronises with the host and cached CUDA graphs stay valid across
// re-sorts.  Host-side bookkeeping only tracks which physical slots hold data written under the current ordering of their
// segment (epochs): reading anything else is an error instead of silently permuted data.
struct Segment {
  int *perm = nullptr;       // storage position -> particle index in the caller's order
  int *from_prev = nullptr;  // storage position -> storage position in the previous segment (valid when linked)
  float4 *mat0 = nullptr;    // (mass, vol, mu, lambda) in storage order
  float *yield = nullptr;
  int4 *chunks = nullptr;
  int *active = nullptr, *flags = nullptr, *cnt = nullptr;
  int epoch = 0;             // 0: never built
  bool linked = false;       // built by re-sorting the previous segment's last state: gradients may flow across the boundary
  SegView view() const { SegView v; v.chunks = chunks; v.cnt = cnt; v.active = active; v.flags = flags; return v; }
};

struct dd_sim {
  dd_sim_config cfg;
  KP kp;
  int slots = 0;             // logical states: max_steps + 1
  int L = 0, nseg = 1;       // substeps per segment (0: the whole rollout is one segment), number of segments
  int pslots = 0;            // physical slots = slots + nseg - 1
  size_t slot_floats = 0;    // kPlaneFloats * EN
  float *ckpt = nullptr;     // pslots * slot_floats
  float *grad[2] = {nullptr, nullptr};
  int grad_holds[2] = {-1, -1};   // logical state whose gradient the slot holds
  int grad_order[2] = {-1, -1};   // segment whose particle order it is stored in (-1: all zeros, any order)
  float *gtmp = nullptr;     // 25 * EN floats, gradient planes in flight at a segment boundary
  float4 *grid = nullptr, *grid_v = nullptr, *ggrid_v = nullptr, *ggrid = nullptr;
  float4 *pos = nullptr, *rot = nullptr, *gpos = nullptr, *grot = nullptr;  // slots * E * nb
  float4 *tfsr = nullptr, *args = nullptr;
  float *cull = nullptr;
  // scratch of the sort pipeline (shared by all segments; builds are serialised on the caller's stream)
  unsigned *keys = nullptr, *keys_alt = nullptr;
  int *iota = nullptr, *sorted_idx = nullptr, *stor_idx = nullptr, *perm_tmp = nullptr;
  void *cub_tmp = nullptr;
  size_t cub_bytes = 0;
  int4 *chunks_tmp = nullptr, *chunk_src_tmp = nullptr, *chunk_src = nullptr;
  int *chunk_sort = nullptr;  // 4 * chunk_cap ints: sizes, sorted sizes, indices, sorted indices
  void *csort_tmp = nullptr;
  size_t csort_bytes = 0;
  int chunk_cap = 0, chunk_max = 512;
  int *head_pos = nullptr, *spos = nullptr;
  char *head_flags = nullptr;
  int NBtot = 0, occ_cap = 0;
  int *counters = nullptr;   // [4],[5] chunk tickets of the persistent tile kernels
  void *sel_tmp = nullptr;
  size_t sel_bytes = 0;
  // per-segment tables, pooled
  std::vector<Segment> segs;
  std::vector<int> slot_epoch;  // per physical slot: epoch of its segment when it was written (-1: nothing)
  int epoch_counter = 0;
  int *perm_pool = nullptr, *from_pool = nullptr, *active_pool = nullptr, *flags_pool = nullptr, *cnt_pool = nullptr;
  float4 *mat0_pool = nullptr;
  float *yield_pool = nullptr;
  int4 *chunks_pool = nullptr;
  float4 *gridck = nullptr, *gridvck = nullptr;  // max_steps * E * G each when (dense) grid checkpoints are on
  bool grid_ckpt = false;
  float4 *brick_m = nullptr, *brick_v = nullptr;  // brick checkpoints: max_steps * brick_cap * 64 each (see BrickCk)
  int *brick_n = nullptr, *status = nullptr;
  int brick_cap = 0;
  bool brick_ckpt = false;
  BrickCk bck(int f) const {
    BrickCk c = {nullptr, nullptr, nullptr, 0, nullptr};
    if (brick_ckpt && f >= 0 && f < slots - 1) {
      size_t o = (size_t)f * brick_cap * 64;
      c.m = brick_m + o; c.v = brick_v + o; c.n = brick_n + f; c.cap = brick_cap; c.status = status;
    }
    return c;
  }
  float *mat_aos = nullptr;  // (mass | vol | mu_lam_yield) in the caller's order, re-packed for every new ordering
  bool have_material = false;
  float *stage = nullptr;    // 24 * EN floats (x|v|F|C in original AoS order) or E*N*nb for dist
  size_t stage_floats = 0;
  std::map<std::tuple<int, int, int, int>, cudaGraphExec_t> graphs;
  std::map<std::tuple<int, int, int, int>, long long> graph_launches;  // kernels per replay of each cached graph
  long long launches = 0;    // kernels launched (or replayed through graphs) since creation
  cudaStream_t side = nullptr;  // uploads that may overlap the sort in dd_sim_set_state
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // persistent launch geometry of the tiled kernels (resident blocks per SM x SMs) and which gather variants run
  int sms = 1;
  bool pdl = false;          // programmatic dependent launch between the hot kernels (DD_PDL=1; measured: no gain inside CUDA graphs, -2 % at config D)
  int pb_p2g = 0, pb_g2pg = 0, pb_g2p = 0, pb_p2gg = 0;
  bool g2p_tiled = false, p2gg_tiled = false;
  int w_p2g = 4, w_g2pg = 1, w_g2p = 4, w_p2gg = 4;  // warps per block (the warps of a block are independent; this only sets the shared-memory granularity)
  // upper-bound grids: the live counts are read on the device
  int tile_blocks(int per_device, int wpb) const { return std::max(1, std::min((chunk_cap + wpb - 1) / wpb, per_device)); }
  int brick_blocks() const { return std::max(1, std::min((NBtot + 3) / 4, sms * 32)); }  // (one resident wave striding over the list measured slower: E=64, grid_grad_b 70 -> 119 us)

  // ---- logical state f <-> segment and physical slot
  int seg_of(int f) const { return L > 0 ? std::min(f / L, nseg - 1) : 0; }          // segment in which state f is the INPUT of a substep
  bool is_boundary(int f) const { return L > 0 && f > 0 && f % L == 0 && f / L < nseg; }  // stored twice
  int phys_cur(int f) const { return f + seg_of(f); }                                 // head copy (or the only copy)
  int phys_tail(int f) const { return f >= 1 ? phys_cur(f - 1) + 1 : phys_cur(0); }   // the copy written by substep f-1
  int seg_of_phys(int p) const { return L > 0 ? std::min(p / (L + 1), nseg - 1) : 0; }
  bool valid(int p) const { int e = slot_epoch[p]; return e > 0 && e == segs[seg_of_phys(p)].epoch; }
  float *pslot(int p) const { return ckpt + (size_t)p * slot_floats; }
  float4 *G(int f) const { return grid_ckpt ? gridck + (size_t)f * kp.E * kp.G : grid; }
  float4 *GV(int f) const { return grid_ckpt ? gridvck + (size_t)f * kp.E * kp.G : grid_v; }
  BodyTables tables(int f) const {
    BodyTables bt;
    size_t o = (size_t)f * kp.E * kp.nb, o1 = (size_t)(f + 1 < slots ? f + 1 : f) * kp.E * kp.nb;
    bt.pos = pos + o; bt.rot = rot + o; bt.npos = pos + o1; bt.nrot = rot + o1;
    bt.tfsr = tfsr; bt.args = args; bt.cull = cull;
    return bt;
  }
};

namespace {

// Launch of a hot-path kernel.  With programmatic dependent launch (DD_PDL=1) the kernel may be scheduled while its
// predecessor on the stream drains: its blocks become resident as SM resources free up, run their prologue (zeroing the
// shared-memory tile, ...) and block in griddepcontrol.wait until the predecessor has completed and flushed.  Every block of
// every hot kernel executes the wait, so completion stays transitive along the chain of kernels.
template <class K, class... A>
void launch_hot(dd_sim *s, K kernel, int grid, int block, size_t smem, cudaStream_t st, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = s->pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, args...);
}

using Mark = std::function<void(const char *)>;
inline void mark(const Mark *m, const char *name) { if (m) (*m)(name); }
int fwd_launches(const dd_sim *s) { return 3; }
int bwd_launches(const dd_sim *s) { return s->cfg.tile_mode ? (s->grid_ckpt || s->brick_ckpt ? 3 : 5) : 5; }
constexpr int kBuildLaunches = 16;  // kernels of one segment build (sort, compaction, chunk tables, gather)

template <int SVD>
void enqueue_forward_substep(dd_sim *s, int f, cudaStream_t st, const Mark *mk = nullptr) {
  const KP &kp = s->kp;
  const Segment &sg = s->segs[s->seg_of(f)];
  float *cur = s->pslot(s->phys_cur(f)), *nxt = s->pslot(s->phys_cur(f) + 1);
  if (s->cfg.tile_mode) {
    // invariant: the scatter target of substep f is already zero on the active bricks (cleared by the previous grid
    // kernel, or by dd_sim_forward for the first substep of a range; bricks activated on the fly clear themselves)
    launch_hot(s, k_p2g_tile<SVD, true>, s->tile_blocks(s->pb_p2g, s->w_p2g), 32 * s->w_p2g, s->w_p2g * kSmemP2G, st, kp, sg.view(), cur, nxt, sg.mat0, sg.yield, s->G(f), s->counters + 4);
    mark(mk, "p2g_tile (svd+return map+scatter)");
    float4 *zn = (s->grid_ckpt && f + 1 < s->slots - 1) ? s->G(f + 1) : nullptr;
    launch_hot(s, k_grid_b, s->brick_blocks(), kT, 0, st, kp, sg.cnt, sg.active, s->G(f), s->GV(f), s->tables(f), zn, s->grid_ckpt ? 0 : 1, s->bck(f));
    mark(mk, "grid_b (grid update + contact)");
    if (s->g2p_tiled) launch_hot(s, k_g2p_tile, s->tile_blocks(s->pb_g2p, s->w_g2p), 32 * s->w_g2p, s->w_g2p * kSmemG2P, st, kp, sg.view(), cur, nxt, s->GV(f), s->counters + 4);
    else k_g2p<<<nblk(kp.EN), kT, 0, st>>>(kp, s->spos, cur, nxt, s->GV(f));
    mark(mk, "g2p");
  } else {
    cudaMemsetAsync(s->grid, 0, sizeof(float4) * (size_t)kp.E * kp.G, st);
    k_p2g<SVD, true><<<nblk(kp.EN), kT, 0, st>>>(kp, cur, nxt, sg.mat0, sg.yield, s->grid);
    k_grid<<<nblk((long long)kp.E * kp.G), kT, 0, st>>>(kp, s->grid, s->grid_v, s->tables(f));
    k_g2p<<<nblk(kp.EN), kT, 0, st>>>(kp, nullptr, cur, nxt, s->grid_v);
  }
  s->launches += fwd_launches(s);
}
template <int SVD>
void enqueue_backward_substep(dd_sim *s, int f, cudaStream_t st, const Mark *mk = nullptr, bool restore_here = true, bool restore_below = false) {
  const KP &kp = s->kp;
  const Segment &sg = s->segs[s->seg_of(f)];
  float *cur = s->pslot(s->phys_cur(f)), *nxt = s->pslot(s->phys_cur(f) + 1);
  float *gin = s->grad[(f + 1) & 1], *gout = s->grad[f & 1];
  size_t eg = (size_t)kp.E * kp.G, ep = (size_t)kp.E * kp.nb;
  float4 *gp = s->gpos + (size_t)f * ep, *gr = s->grot + (size_t)f * ep, *gnp = s->gpos + (size_t)(f + 1) * ep, *gnr = s->grot + (size_t)(f + 1) * ep;
  if (s->cfg.tile_mode) {
    // invariant: ggrid_v is zero on the active bricks (k_grid_grad_b clears what it consumes)
    BrickCk none = {nullptr, nullptr, nullptr, 0, nullptr};
    if (s->brick_ckpt) {  // v_out of this substep back into the dense working grid, unless the substep above has already done it
      if (restore_here) {
        launch_hot(s, k_restore_bricks, s->brick_blocks(), kT, 0, st, kp, sg.cnt, sg.active, s->bck(f).v, s->bck(f).n, s->grid_v);
        s->launches += 1;
      }
    } else if (!s->grid_ckpt) {  // no grid checkpoints at all: re-run scatter and grid update like the reference does
      launch_hot(s, k_p2g_tile<SVD, false>, s->tile_blocks(s->pb_p2g, s->w_p2g), 32 * s->w_p2g, s->w_p2g * kSmemP2G, st, kp, sg.view(), cur, nxt, sg.mat0, sg.yield, s->grid, s->counters + 4);
      launch_hot(s, k_grid_b, s->brick_blocks(), kT, 0, st, kp, sg.cnt, sg.active, s->grid, s->grid_v, s->tables(f), nullptr, 0, none);
    }
    launch_hot(s, k_g2p_grad_tile, s->tile_blocks(s->pb_g2pg, s->w_g2pg), 32 * s->w_g2pg, s->w_g2pg * kSmemG2PG, st, kp, sg.view(), cur, nxt, s->GV(f), gin, gout, s->ggrid_v, s->counters + 4);
    mark(mk, "g2p_grad_tile");
    launch_hot(s, k_grid_grad_b, s->brick_blocks(), kT, 0, st, kp, sg.cnt, sg.active, s->G(f), s->ggrid_v, s->ggrid, s->tables(f), gp, gr, gnp, gnr, s->grid_ckpt || s->brick_ckpt ? 0 : 1,
               s->bck(f), restore_below ? s->bck(f - 1) : none, s->grid_v);
    mark(mk, "grid_grad_b");
    if (s->p2gg_tiled) launch_hot(s, k_p2g_grad_tile<SVD>, s->tile_blocks(s->pb_p2gg, s->w_p2gg), 32 * s->w_p2gg, s->w_p2gg * (SVD == 1 ? kSmemP2GG : kSmemG2P), st, kp, sg.view(), cur, nxt, sg.mat0, sg.yield, s->ggrid, gin, gout, s->counters + 4);
    else k_p2g_grad<SVD><<<nblk(kp.EN), kT, 0, st>>>(kp, s->spos, cur, nxt, sg.mat0, sg.yield, s->ggrid, gin, gout);
    mark(mk, "p2g_grad (+svd adjoint)");
  } else {
    cudaMemsetAsync(s->grid, 0, sizeof(float4) * eg, st);
    cudaMemsetAsync(s->ggrid_v, 0, sizeof(float4) * eg, st);
    k_p2g<SVD, false><<<nblk(kp.EN), kT, 0, st>>>(kp, cur, nxt, sg.mat0, sg.yield, s->grid);
    k_grid<<<nblk((long long)eg), kT, 0, st>>>(kp, s->grid, s->grid_v, s->tables(f));
    k_g2p_grad<<<nblk(kp.EN), kT, 0, st>>>(kp, cur, nxt, s->grid_v, gin, gout, s->ggrid_v);
    k_grid_grad<<<nblk((long long)eg), kT, 0, st>>>(kp, s->grid, s->ggrid_v, s->ggrid, s->tables(f), gp, gr, gnp, gnr);
    k_p2g_grad<SVD><<<nblk(kp.EN), kT, 0, st>>>(kp, nullptr, cur, nxt, sg.mat0, sg.yield, s->ggrid, gin, gout);
  }
  s->launches += bwd_launches(s);
}

// Enqueue the construction of segment k's ordering from positions: `x_aos` (caller's order, dd_sim_set_state) or plane 0 of
// physical slot `src` stored in segment kprev's order (device re-sort).  Afterwards stor_idx[i] = source index of storage
// position i, seg.perm is final and, for a device re-sort, seg.from_prev = stor_idx and the state planes of `src` have been
// gathered into physical slot `dst`.  Nothing here waits for the device.
int enqueue_build(dd_sim *s, int k, const float *x_aos, int src, int kprev, int dst, cudaStream_t st) {
  const KP &kp = s->kp;
  Segment &sg = s->segs[k];
  const int EN = kp.EN;
  const int *order = s->iota;  // storage position -> source index
  if (s->cfg.sort_particles) {
    k_sort_keys<<<nblk(EN), kT, 0, st>>>(kp, x_aos, x_aos ? nullptr : reinterpret_cast<const float4 *>(s->pslot(src)), s->keys, s->iota);
    int bits = 1;
    while (bits < 32 && (1ull << bits) < (unsigned long long)kp.E * kp.G) ++bits;
    DD_CUDA(cub::DeviceRadixSort::SortPairs(s->cub_tmp, s->cub_bytes, s->keys, s->keys_alt, s->iota, s->sorted_idx, EN, 0, bits, st));
    order = s->sorted_idx;
    if (s->cfg.tile_mode) {
      DD_CUDA(cudaMemsetAsync(sg.flags, 0, sizeof(int) * s->NBtot, st));
      DD_CUDA(cudaMemsetAsync(sg.cnt, 0, sizeof(int) * 4, st));
      k_mark_heads<<<nblk(EN), kT, 0, st>>>(kp, s->keys_alt, s->head_flags, sg.flags);
      DD_CUDA(cub::DeviceSelect::Flagged(s->sel_tmp, s->sel_bytes, cub::CountingInputIterator<int>(0), s->head_flags, s->head_pos, sg.cnt + 2, EN, st));
      k_make_chunks<<<nblk(s->occ_cap), kT, 0, st>>>(kp, s->head_pos, s->keys_alt, s->chunk_max, s->chunks_tmp, s->chunk_src_tmp, sg.cnt);
      DD_CUDA(cub::DeviceSelect::Flagged(s->sel_tmp, s->sel_bytes, cub::CountingInputIterator<int>(0), sg.flags, sg.active, sg.cnt + 1, s->NBtot, st));
      int cap = s->chunk_cap;
      int *ck = s->chunk_sort, *cks = ck + cap, *ci = ck + 2 * cap, *cis = ck + 3 * cap;
      k_chunk_keys<<<nblk(cap), kT, 0, st>>>(cap, sg.cnt, s->chunks_tmp, ck, ci);
      DD_CUDA(cub::DeviceRadixSort::SortPairsDescending(s->csort_tmp, s->csort_bytes, ck, cks, ci, cis, cap, 0, 32, st));
      k_chunk_permute<<<nblk(cap), kT, 0, st>>>(cap, sg.cnt, cis, s->chunks_tmp, s->chunk_src_tmp, sg.chunks, s->chunk_src);
      k_interleave<<<nblk((long long)cap * 32), kT, 0, st>>>(kp, sg.cnt, sg.chunks, s->chunk_src, s->keys_alt, s->sorted_idx, s->stor_idx, s->spos);
      order = s->stor_idx;
    }
  }
  if (x_aos) {
    DD_CUDA(cudaMemcpyAsync(sg.perm, order, sizeof(int) * EN, cudaMemcpyDeviceToDevice, st));
  } else {
    k_compose_perm<<<nblk(EN), kT, 0, st>>>(EN, order, s->segs[kprev].perm, s->perm_tmp);  // (kprev may be k: compose out of place)
    DD_CUDA(cudaMemcpyAsync(sg.from_prev, order, sizeof(int) * EN, cudaMemcpyDeviceToDevice, st));
    DD_CUDA(cudaMemcpyAsync(sg.perm, s->perm_tmp, sizeof(int) * EN, cudaMemcpyDeviceToDevice, st));
    k_gather_planes<<<nblk(EN), kT, 0, st>>>(EN, sg.from_prev, s->pslot(src), s->pslot(dst));
  }
  if (s->have_material) {
    float *a = s->mat_aos;
    k_pack_mat<<<nblk(EN), kT, 0, st>>>(EN, sg.perm, a, a + EN, a + 2 * (size_t)EN, sg.mat0, sg.yield);
  }
  DD_CUDA(cudaGetLastError());
  s->launches += kBuildLaunches;
  return 0;
}
// host bookkeeping that goes with enqueue_build: a new ordering invalidates everything stored under the old one
void segment_rebuilt(dd_sim *s, int k, bool linked, int first_phys) {
  Segment &sg = s->segs[k];
  sg.epoch = ++s->epoch_counter;
  sg.linked = linked;
  s->slot_epoch[first_phys] = sg.epoch;
  for (int i = 0; i < 2; ++i)
    if (s->grad_order[i] == k) { s->grad_holds[i] = -1; s->grad_order[i] = -1; }
}
// gradient of boundary state (held in grad slot `gi`, order of segment k) back into the order of segment k-1
void enqueue_grad_permute(dd_sim *s, int k, int gi, cudaStream_t st) {
  const int EN = s->kp.EN;
  k_scatter_grad<<<nblk(EN), kT, 0, st>>>(EN, s->segs[k].from_prev, s->grad[gi], s->gtmp);
  cudaMemcpyAsync(s->grad[gi], s->gtmp, sizeof(float) * 25 * (size_t)EN, cudaMemcpyDeviceToDevice, st);
  s->launches += 2;
}

// run `body` either directly or as a cached CUDA graph keyed by (kind, f0, n, variant)
template <class Fn>
int run_graphed(dd_sim *s, int kind, int f0, int n, int variant, cudaStream_t st, Fn body) {
  if (!s->cfg.use_graphs) {
    body(st);
    DD_CUDA(cudaGetLastError());
    return 0;
  }
  auto key = std::make_tuple(kind, f0, n, variant);
  auto it = s->graphs.find(key);
  long long before = s->launches;
  if (it == s->graphs.end()) {
    cudaStream_t cap = st;
    cudaStream_t tmp = nullptr;
    if (cap == nullptr) {  // the legacy default stream cannot be captured
      DD_CUDA(cudaStreamCreateWithFlags(&tmp, cudaStreamNonBlocking));
      cap = tmp;
    }
    DD_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    body(cap);
    cudaGraph_t graph = nullptr;
    DD_CUDA(cudaStreamEndCapture(cap, &graph));
    cudaGraphExec_t exec = nullptr;
    DD_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    DD_CUDA(cudaGraphDestroy(graph));
    if (tmp) DD_CUDA(cudaStreamDestroy(tmp));
    it = s->graphs.emplace(key, exec).first;
    s->graph_launches[key] = s->launches - before;
  }
  DD_CUDA(cudaGraphLaunch(it->second, st));
  s->launches = before + s->graph_launches[key];
  return 0;
}

// results written to device memory are ordered by the stream: only host destinations make a getter wait
bool on_device(const void *p) {
  if (!p) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice;
}
int finish_readback(cudaStream_t st, std::initializer_list<const void *> outs) {
  for (const void *p : outs)
    if (!on_device(p)) { DD_CUDA(cudaStreamSynchronize(st)); break; }
  return 0;
}

int check_range(dd_sim *s, int f0, int n, const char *what) {
  if (!s) return fail(std::string(what) + ": null simulator");
  if (f0 < 0 || n < 0 || f0 + n >= s->slots) return fail(std::string(what) + ": substep range [" + std::to_string(f0) + ", " + std::to_string(f0 + n) + "] exceeds max_steps=" + std::to_string(s->slots - 1));
  return 0;
}
// which stored copy of state f to read: the head copy if it is current, else the copy the previous substep wrote
int pick_copy(dd_sim *s, int f, int *phys, int *seg, const char *what) {
  if (s->valid(s->phys_cur(f))) { *phys = s->phys_cur(f); *seg = s->seg_of(f); return 0; }
  if (f >= 1 && s->valid(s->phys_tail(f))) { *phys = s->phys_tail(f); *seg = s->seg_of(f - 1); return 0; }
  return fail(std::string(what) + ": state " + std::to_string(f) + " is not available (never computed, or stored under a particle order that a later "
              "dd_sim_set_state / re-sort replaced)");
}
// the copy of state f that is stored in the same order as its gradient slot; fixes the order of an all-zero slot
int grad_copy(dd_sim *s, int f, int *phys, int *seg, const char *what) {
  int gi = f & 1;
  if (s->grad_holds[gi] != f) return fail(std::string(what) + ": gradient slot does not hold state " + std::to_string(f) + " (call dd_sim_zero_grad or run the backward pass down to it first)");
  if (s->grad_order[gi] < 0) {  // zeros: take the order the backward pass will consume it in (the substep that produced the state)
    if (f >= 1 && s->valid(s->phys_tail(f))) s->grad_order[gi] = s->seg_of(f - 1);
    else s->grad_order[gi] = s->seg_of(f);
  }
  *seg = s->grad_order[gi];
  *phys = *seg == s->seg_of(f) ? s->phys_cur(f) : s->phys_tail(f);
  return 0;
}

}  // namespace

int dd_set_error(const char *msg) { return fail(msg); }  // shared with fk.cu

extern "C" {

const char *dd_last_error(void) { return g_last_error.c_str(); }

// device pointers of the pose table: float4 (x,y,z,0) / (w,x,y,z) per (slot, env, body)
int dd_sim_pose_table(dd_sim *s, float **pos, float **rot, int *slots, int *n_envs, int *n_bodies) {
  if (!s) return fail("dd_sim_pose_table: null simulator");
  if (pos) *pos = reinterpret_cast<float *>(s->pos);
  if (rot) *rot = reinterpret_cast<float *>(s->rot);
  if (slots) *slots = s->slots;
  if (n_envs) *n_envs = s->kp.E;
  if (n_bodies) *n_bodies = s->kp.nb;
  return 0;
}
int dd_sim_pose_grad_table(dd_sim *s, float **gpos, float **grot) {
  if (!s) return fail("dd_sim_pose_grad_table: null simulator");
  if (gpos) *gpos = reinterpret_cast<float *>(s->gpos);
  if (grot) *grot = reinterpret_cast<float *>(s->grot);
  return 0;
}

int dd_sim_create(const dd_sim_config *cfg, dd_sim **out) {
  if (!cfg || !out) return fail("dd_sim_create: null argument");
  if (cfg->n_envs < 1 || cfg->n_particles < 1 || cfg->max_steps < 1) return fail("dd_sim_create: n_envs, n_particles, max_steps must be >= 1");
  if (cfg->n_bodies < 0 || cfg->n_bodies > 64) return fail("dd_sim_create: n_bodies must be in [0, 64]");
  if (cfg->grid_x < 8 || cfg->grid_y < 8 || cfg->grid_z < 8) return fail("dd_sim_create: grid must be at least 8^3");
  if (((long long)cfg->grid_x * cfg->grid_y * cfg->grid_z) % 32 != 0) return fail("dd_sim_create: grid size must be a multiple of 32 nodes");
  if ((long long)cfg->n_envs * cfg->grid_x * cfg->grid_y * cfg->grid_z >= (1ll << 32)) return fail("dd_sim_create: n_envs * grid nodes must stay below 2^32 (32-bit sort keys)");
  if ((long long)cfg->n_envs * cfg->n_particles >= (1ll << 31)) return fail("dd_sim_create: n_envs * n_particles must stay below 2^31");
  if (cfg->resort_interval < 0) return fail("dd_sim_create: resort_interval must be >= 0");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) return fail("dd_sim_create: no CUDA device (dexdeform_b200 has no CPU fallback)");
  dd_sim *s = new dd_sim();
  s->cfg = *cfg;
  cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming);
  KP &kp = s->kp;
  kp.E = cfg->n_envs; kp.N = cfg->n_particles; kp.EN = kp.E * kp.N; kp.nb = cfg->n_bodies;
  kp.gx = cfg->grid_x; kp.gy = cfg->grid_y; kp.gz = cfg->grid_z; kp.G = kp.gx * kp.gy * kp.gz;
  kp.dx = cfg->dx; kp.inv_dx = 1.0f / cfg->dx; kp.dt = cfg->dt; kp.gf = cfg->ground_friction; kp.gh = cfg->ground_height;
  kp.g0 = cfg->gravity[0]; kp.g1 = cfg->gravity[1]; kp.g2 = cfg->gravity[2];
  s->slots = cfg->max_steps + 1;
  // segments: re-sorting needs the sorted, tiled layout
  s->L = (cfg->tile_mode && cfg->sort_particles && cfg->resort_interval < cfg->max_steps) ? cfg->resort_interval : 0;
  s->nseg = s->L > 0 ? (cfg->max_steps + s->L - 1) / s->L : 1;
  s->pslots = s->slots + s->nseg - 1;
  size_t ENp = ((size_t)kp.EN + 3) / 4 * 4;
  if (ENp != (size_t)kp.EN) { delete s; return fail("dd_sim_create: n_envs * n_particles must be a multiple of 4 (float4 planes)"); }
  s->slot_floats = kPlaneFloats * ENp;
  size_t eg = (size_t)kp.E * kp.G, ep = (size_t)kp.E * (kp.nb > 0 ? kp.nb : 1) * s->slots;
  const int nseg = s->nseg;
#define DD_ALLOC(ptr, bytes)                                                                                   \
  do {                                                                                                         \
    cudaError_t e_ = cudaMalloc((void **)&(ptr), (bytes));                                                     \
    if (e_ != cudaSuccess) { dd_sim_destroy(s); return fail(std::string("cudaMalloc " #ptr ": ") + cudaGetErrorString(e_)); } \
    cudaMemset((ptr), 0, (bytes));                                                                             \
  } while (0)
  DD_ALLOC(s->ckpt, sizeof(float) * s->slot_floats * s->pslots);
  DD_ALLOC(s->grad[0], sizeof(float) * 28 * ENp);  // 25 floats per particle: the gradient of (x, v, C | F)
  DD_ALLOC(s->grad[1], sizeof(float) * 28 * ENp);
  DD_ALLOC(s->gtmp, sizeof(float) * 25 * ENp);
  DD_ALLOC(s->grid, sizeof(float4) * eg);
  DD_ALLOC(s->grid_v, sizeof(float4) * eg);
  DD_ALLOC(s->ggrid_v, sizeof(float4) * eg);
  DD_ALLOC(s->ggrid, sizeof(float4) * eg);
  DD_ALLOC(s->pos, sizeof(float4) * ep);
  DD_ALLOC(s->rot, sizeof(float4) * ep);
  DD_ALLOC(s->gpos, sizeof(float4) * ep);
  DD_ALLOC(s->grot, sizeof(float4) * ep);
  DD_ALLOC(s->tfsr, sizeof(float4) * 64);
  DD_ALLOC(s->args, sizeof(float4) * 64);
  DD_ALLOC(s->cull, sizeof(float) * 64);
  DD_ALLOC(s->keys, sizeof(unsigned) * ENp);
  DD_ALLOC(s->keys_alt, sizeof(unsigned) * ENp);
  DD_ALLOC(s->iota, sizeof(int) * ENp);
  DD_ALLOC(s->sorted_idx, sizeof(int) * ENp);
  DD_ALLOC(s->stor_idx, sizeof(int) * ENp);
  DD_ALLOC(s->perm_tmp, sizeof(int) * ENp);
  DD_ALLOC(s->mat_aos, sizeof(float) * 5 * ENp);
  DD_ALLOC(s->perm_pool, sizeof(int) * ENp * nseg);
  DD_ALLOC(s->from_pool, sizeof(int) * ENp * nseg);
  DD_ALLOC(s->mat0_pool, sizeof(float4) * ENp * nseg);
  DD_ALLOC(s->yield_pool, sizeof(float) * ENp * nseg);
  DD_ALLOC(s->cnt_pool, sizeof(int) * 8 * nseg);
  DD_ALLOC(s->counters, sizeof(int) * 8);
  cub::DeviceRadixSort::SortPairs(nullptr, s->cub_bytes, s->keys, s->keys_alt, s->iota, s->sorted_idx, kp.EN);
  DD_ALLOC(s->cub_tmp, s->cub_bytes + 16);
  s->segs.resize(nseg);
  s->slot_epoch.assign(s->pslots, -1);
  if (cfg->tile_mode) {
    if (!cfg->sort_particles) { dd_sim_destroy(s); return fail("dd_sim_create: tile_mode requires sort_particles"); }
    if ((kp.gx | kp.gy | kp.gz) & 3) { dd_sim_destroy(s); return fail("dd_sim_create: tile_mode needs grid dimensions that are multiples of 4"); }
    s->NBtot = kp.E * (kp.gx >> 2) * (kp.gy >> 2) * (kp.gz >> 2);
    {
      const int cap = 227 * 1024;
      auto up_to = [&](size_t per_warp) { return (int)std::min<size_t>(8 * per_warp, (size_t)cap); };
      cudaFuncSetAttribute(k_g2p_grad_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, up_to(kSmemG2PG));
      cudaFuncSetAttribute(k_g2p_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, up_to(kSmemG2P));
      cudaFuncSetAttribute(k_p2g_tile<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, up_to(kSmemP2G));
      cudaFuncSetAttribute(k_p2g_tile<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, up_to(kSmemP2G));
      cudaFuncSetAttribute(k_p2g_tile<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, up_to(kSmemP2G));
      cudaFuncSetAttribute(k_p2g_tile<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, up_to(kSmemP2G));
      cudaFuncSetAttribute(k_p2g_grad_tile<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, up_to(kSmemG2P));
      cudaFuncSetAttribute(k_p2g_grad_tile<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, up_to(kSmemP2GG));
    }
    {
      int dev = 0, occ = 1;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&s->sms, cudaDevAttrMultiProcessorCount, dev);
      const int sms = s->sms;
      auto knob = [](const char *name, int dflt) { const char *e = getenv(name); int v = e ? atoi(e) : dflt; return v >= 1 && v <= 8 ? v : dflt; };
      s->w_p2g = knob("DD_WPB_P2G", s->w_p2g); s->w_g2pg = knob("DD_WPB_G2PG", s->w_g2pg); s->w_g2p = knob("DD_WPB_G2P", s->w_g2p); s->w_p2gg = knob("DD_WPB_P2GG", s->w_p2gg);
      auto per_device = [&](auto kernel, int wpb, size_t smem) { occ = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 32 * wpb, smem * wpb); return std::max(occ, 1) * sms; };
      const char *e0 = getenv("DD_PDL");
      s->pdl = e0 && atoi(e0) != 0;
      if (cfg->svd_mode == 0) s->pb_p2g = per_device(k_p2g_tile<0, true>, s->w_p2g, kSmemP2G); else s->pb_p2g = per_device(k_p2g_tile<1, true>, s->w_p2g, kSmemP2G);
      s->pb_p2gg = cfg->svd_mode == 0 ? per_device(k_p2g_grad_tile<0>, s->w_p2gg, kSmemG2P) : per_device(k_p2g_grad_tile<1>, s->w_p2gg, kSmemP2GG);
      s->pb_g2pg = per_device(k_g2p_grad_tile, s->w_g2pg, kSmemG2PG);
      s->pb_g2p = per_device(k_g2p_tile, s->w_g2p, kSmemG2P);
      const char *e1 = getenv("DD_G2P_TILE"), *e2 = getenv("DD_P2GG_TILE");
      s->g2p_tiled = !(e1 && atoi(e1) == 0);  // default: tiled gather
      s->p2gg_tiled = e2 ? atoi(e2) != 0 : cfg->svd_mode == 1;  // default: the staged tile kernel with the fp32 SVD, the flat kernel otherwise
    }
    // Particles per chunk (one warp, one tile).  Large chunks amortise the tile fill / flush (256 is the measured optimum at
    // config D); a batch that cannot give every resident warp about two chunks gets smaller ones, because a warp walks the
    // rows of its chunk one after the other and a nearly empty machine is then bound by that latency.
    const char *ecm = getenv("DD_CHUNK_MAX");  // (tuning knob like DD_WPB_*: overrides the automatic choice, not the caller's)
    if (cfg->chunk_max > 0) s->chunk_max = cfg->chunk_max;
    else if (ecm && atoi(ecm) >= 32) s->chunk_max = std::min(atoi(ecm), 512);
    else {
      int warps = std::max(1, s->pb_p2g * s->w_p2g), per_warp = kp.EN / warps;
      s->chunk_max = per_warp >= 256 ? 256 : std::min(256, std::max(64, (per_warp / 2 + 31) / 32 * 32));
      // A batch that makes more than two scheduling waves of 256-particle chunks on the scatter kernel takes 320-particle chunks
      // when that shortens the modelled launch, ceil(waves) x rows per chunk (chunks average ~0.94 of the maximum).  Measured on a
      // pass of workload E: 128 x 10k particles 205.6 -> 198.9 ms; 64 x 10k (1.1 waves) and config D (1.9) are left at 256, where 320 is
      // neutral / 1 % slower.
      auto launch_rows = [&](int c) { double w = kp.EN / (0.94 * c) / warps; return std::ceil(w) * c; };
      if (kp.EN / (0.94 * 256) / warps > 2.0 && launch_rows(320) < launch_rows(256)) s->chunk_max = 320;
    }
    s->occ_cap = std::min(kp.EN, s->NBtot);
    s->chunk_cap = kp.EN / s->chunk_max + s->occ_cap + 1;  // every occupied brick adds at most one partly filled chunk
    DD_ALLOC(s->chunks_tmp, sizeof(int4) * s->chunk_cap);
    DD_ALLOC(s->chunk_src_tmp, sizeof(int4) * s->chunk_cap);
    DD_ALLOC(s->chunk_src, sizeof(int4) * s->chunk_cap);
    DD_ALLOC(s->chunk_sort, sizeof(int) * 4 * s->chunk_cap);
    cub::DeviceRadixSort::SortPairsDescending(nullptr, s->csort_bytes, s->chunk_sort, s->chunk_sort, s->chunk_sort, s->chunk_sort, s->chunk_cap);
    DD_ALLOC(s->csort_tmp, s->csort_bytes + 16);
    DD_ALLOC(s->head_pos, sizeof(int) * (s->occ_cap + 1));
    DD_ALLOC(s->spos, sizeof(int) * ENp);
    DD_ALLOC(s->head_flags, ENp);
    DD_ALLOC(s->chunks_pool, sizeof(int4) * (size_t)s->chunk_cap * nseg);
    DD_ALLOC(s->active_pool, sizeof(int) * (size_t)s->NBtot * nseg);
    DD_ALLOC(s->flags_pool, sizeof(int) * (size_t)s->NBtot * nseg);
    size_t b1 = 0, b2 = 0;
    cub::DeviceSelect::Flagged(nullptr, b1, cub::CountingInputIterator<int>(0), s->head_flags, s->head_pos, s->cnt_pool, kp.EN);
    cub::DeviceSelect::Flagged(nullptr, b2, cub::CountingInputIterator<int>(0), s->flags_pool, s->active_pool, s->cnt_pool, s->NBtot);
    s->sel_bytes = std::max(b1, b2);
    DD_ALLOC(s->sel_tmp, s->sel_bytes + 16);
    // per-substep grid checkpoints (no scatter / grid-update replay in the backward pass) if they fit
    size_t need = sizeof(float4) * eg * 2 * (size_t)cfg->max_steps, fr = 0, tot = 0;
    cudaMemGetInfo(&fr, &tot);
    bool want = cfg->grid_ckpt != 0;
    DD_ALLOC(s->status, sizeof(int) * 4);
    if (want && cfg->grid_ckpt != 2 && need < (size_t)(0.6 * (double)fr)) {
      s->grid_ckpt = true;
      DD_ALLOC(s->gridck, sizeof(float4) * eg * (size_t)cfg->max_steps);
      DD_ALLOC(s->gridvck, sizeof(float4) * eg * (size_t)cfg->max_steps);
    } else if (want) {
      // brick checkpoints: a fixed number of bricks per substep out of a fifth of the free memory (2 KB per brick and substep)
      size_t per_brick = sizeof(float4) * 64 * 2 * (size_t)cfg->max_steps;
      size_t cap = std::min<size_t>((size_t)s->NBtot, (size_t)(0.2 * (double)fr) / per_brick);
      if (const char *e = getenv("DD_BRICK_CAP")) cap = std::min<size_t>(cap, (size_t)std::max(1, atoi(e)));  // (tests: force the overflow report)
      if (cap >= 8) {
        s->brick_ckpt = true;
        s->brick_cap = (int)cap;
        DD_ALLOC(s->brick_m, sizeof(float4) * 64 * cap * (size_t)cfg->max_steps);
        DD_ALLOC(s->brick_v, sizeof(float4) * 64 * cap * (size_t)cfg->max_steps);
        DD_ALLOC(s->brick_n, sizeof(int) * (size_t)cfg->max_steps);
      }
    }
  }
  for (int k = 0; k < nseg; ++k) {
    Segment &sg = s->segs[k];
    sg.perm = s->perm_pool + (size_t)k * ENp; sg.from_prev = s->from_pool + (size_t)k * ENp;
    sg.mat0 = s->mat0_pool + (size_t)k * ENp; sg.yield = s->yield_pool + (size_t)k * ENp;
    sg.cnt = s->cnt_pool + 8 * k;
    if (cfg->tile_mode) {
      sg.chunks = s->chunks_pool + (size_t)k * s->chunk_cap;
      sg.active = s->active_pool + (size_t)k * s->NBtot; sg.flags = s->flags_pool + (size_t)k * s->NBtot;
    }
  }
  s->stage_floats = std::max((size_t)kp.EN * (24 > kp.nb ? 24 : kp.nb), (size_t)7 * s->slots * kp.E * kp.nb);  // states / distances, or every pose of every slot
  DD_ALLOC(s->stage, sizeof(float) * s->stage_floats);
#undef DD_ALLOC
  std::vector<int> ident(kp.EN);
  for (int i = 0; i < kp.EN; ++i) ident[i] = i;
  cudaMemcpy(s->iota, ident.data(), sizeof(int) * kp.EN, cudaMemcpyHostToDevice);
  cudaMemcpy(s->segs[0].perm, ident.data(), sizeof(int) * kp.EN, cudaMemcpyHostToDevice);
  *out = s;
  return 0;
}

void dd_sim_destroy(dd_sim *s) {
  if (!s) return;
  for (auto &kv : s->graphs) cudaGraphExecDestroy(kv.second);
  if (s->side) cudaStreamDestroy(s->side);
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->ev_join) cudaEventDestroy(s->ev_join);
  void *ptrs[] = {s->ckpt, s->grad[0], s->grad[1], s->gtmp, s->grid, s->grid_v, s->ggrid_v, s->ggrid, s->pos, s->rot, s->gpos, s->grot,
                  s->tfsr, s->args, s->cull, s->stage, s->keys, s->keys_alt, s->iota, s->sorted_idx, s->stor_idx, s->perm_tmp, s->cub_tmp, s->mat_aos,
                  s->chunks_tmp, s->chunk_src_tmp, s->chunk_src, s->chunk_sort, s->csort_tmp, s->head_pos, s->spos, s->head_flags, s->counters, s->sel_tmp,
                  s->perm_pool, s->from_pool, s->mat0_pool, s->yield_pool, s->cnt_pool, s->chunks_pool, s->active_pool, s->flags_pool, s->gridck, s->gridvck, s->brick_m, s->brick_v, s->brick_n, s->status};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  delete s;
}

long long dd_sim_launch_count(dd_sim *s) { return s ? s->launches : 0; }
#ifdef DD_COUNT_DEFER
extern "C" void dd_debug_defer_counts(unsigned long long *out) { cudaDeviceSynchronize(); cudaMemcpyFromSymbol(out, g_defer_count, 16); }
#endif

int dd_sim_set_material(dd_sim *s, const float *mass, const float *vol, const float *mu_lam_yield, cudaStream_t st) {
  if (!s || !mass || !vol || !mu_lam_yield) return fail("dd_sim_set_material: null argument");
  int EN = s->kp.EN;
  float *a = s->mat_aos, *b = a + EN, *c = b + EN;
  s->have_material = true;
  DD_CUDA(cudaMemcpyAsync(a, mass, sizeof(float) * EN, cudaMemcpyDefault, st));
  DD_CUDA(cudaMemcpyAsync(b, vol, sizeof(float) * EN, cudaMemcpyDefault, st));
  DD_CUDA(cudaMemcpyAsync(c, mu_lam_yield, sizeof(float) * 3 * EN, cudaMemcpyDefault, st));
  for (auto &sg : s->segs)
    if (sg.epoch > 0 || &sg == &s->segs[0]) k_pack_mat<<<nblk(EN), kT, 0, st>>>(EN, sg.perm, a, b, c, sg.mat0, sg.yield);
  DD_CUDA(cudaGetLastError());
  return 0;
}

int dd_sim_set_bodies(dd_sim *s, const float *tfsr, const float *args) {
  if (!s) return fail("dd_sim_set_bodies: null simulator");
  int nb = s->kp.nb;
  if (nb == 0) return 0;
  if (!tfsr || !args) return fail("dd_sim_set_bodies: null argument");
  std::vector<float> t(4 * nb), a(4 * nb), cull(nb);
  DD_CUDA(cudaMemcpy(t.data(), tfsr, sizeof(float) * 4 * nb, cudaMemcpyDefault));
  DD_CUDA(cudaMemcpy(a.data(), args, sizeof(float) * 4 * nb, cudaMemcpyDefault));
  for (int b = 0; b < nb; ++b) {
    // a node can only be active if dist < ln(10)/softness (or <= 0); dist >= |p| - R_bound - round
    float type = t[4 * b], soft = t[4 * b + 2], round = t[4 * b + 3];
    float rb = ((int)floorf(type + 0.1f) == 0) ? sqrtf(a[4 * b] * a[4 * b] + a[4 * b + 1] * a[4 * b + 1] + a[4 * b + 2] * a[4 * b + 2])
                                               : fabsf(a[4 * b]) + fabsf(a[4 * b + 1]);
    cull[b] = (rb + fabsf(round) + (soft > 0 ? 2.5f / soft : 0.f)) * 1.001f + 1e-5f;
  }
  DD_CUDA(cudaMemcpy(s->tfsr, t.data(), sizeof(float) * 4 * nb, cudaMemcpyHostToDevice));
  DD_CUDA(cudaMemcpy(s->args, a.data(), sizeof(float) * 4 * nb, cudaMemcpyHostToDevice));
  DD_CUDA(cudaMemcpy(s->cull, cull.data(), sizeof(float) * nb, cudaMemcpyHostToDevice));
  return 0;
}

// State f from the caller's arrays (host or device).  Starts a new particle order for f's segment; nothing here waits for
// the device (the caller synchronises before releasing host buffers).
int dd_sim_set_state(dd_sim *s, int f, const float *x, const float *v, const float *F, const float *C, cudaStream_t st) {
  if (check_range(s, f, 0, "dd_sim_set_state")) return 1;
  if (!x || !v || !F || !C) return fail("dd_sim_set_state: x, v, F, C are all required");
  int EN = s->kp.EN;
  float *sx = s->stage, *sv = sx + 3 * (size_t)EN, *sF = sv + 3 * (size_t)EN, *sC = sF + 9 * (size_t)EN;
  // positions first: the sort needs nothing else, so v, F and C are copied on a side stream while the sort kernels run
  DD_CUDA(cudaMemcpyAsync(sx, x, sizeof(float) * 3 * EN, cudaMemcpyDefault, st));
  DD_CUDA(cudaEventRecord(s->ev_fork, st));
  DD_CUDA(cudaStreamWaitEvent(s->side, s->ev_fork, 0));  // the staging buffer is free again once earlier work on st is done
  DD_CUDA(cudaMemcpyAsync(sv, v, sizeof(float) * 3 * EN, cudaMemcpyDefault, s->side));
  DD_CUDA(cudaMemcpyAsync(sF, F, sizeof(float) * 9 * EN, cudaMemcpyDefault, s->side));
  DD_CUDA(cudaMemcpyAsync(sC, C, sizeof(float) * 9 * EN, cudaMemcpyDefault, s->side));
  DD_CUDA(cudaEventRecord(s->ev_join, s->side));
  int k = s->seg_of(f), p = s->phys_cur(f);
  if (s->cfg.sort_particles || s->segs[k].epoch == 0) {
    if (s->cfg.tile_mode) DD_CUDA(cudaMemsetAsync(s->counters + 4, 0, sizeof(int) * 2, st));  // chunk tickets: start clean even after an aborted launch
    if (enqueue_build(s, k, sx, -1, -1, p, st)) return 1;
    segment_rebuilt(s, k, false, p);
  } else {
    s->slot_epoch[p] = s->segs[k].epoch;
  }
  DD_CUDA(cudaStreamWaitEvent(st, s->ev_join, 0));
  k_pack<<<nblk(EN), kT, 0, st>>>(EN, s->segs[k].perm, sx, sv, sF, sC, s->pslot(p));
  DD_CUDA(cudaGetLastError());
  return 0;
}

// Rolling window (mpm/simulator.py:626-634): state f_src becomes state 0, re-sorted on the device, poses included.
int dd_sim_roll(dd_sim *s, int f_src, cudaStream_t st) {
  if (check_range(s, f_src, 0, "dd_sim_roll")) return 1;
  if (f_src == 0) return 0;
  int src = 0, kprev = 0;
  if (pick_copy(s, f_src, &src, &kprev, "dd_sim_roll")) return 1;
  if (s->cfg.sort_particles) {
    if (enqueue_build(s, 0, nullptr, src, kprev, 0, st)) return 1;
  } else {
    DD_CUDA(cudaMemcpyAsync(s->pslot(0), s->pslot(src), sizeof(float) * 29 * (size_t)s->kp.EN, cudaMemcpyDeviceToDevice, st));
  }
  segment_rebuilt(s, 0, false, 0);
  if (!s->cfg.sort_particles) s->slot_epoch[0] = s->segs[0].epoch;
  size_t ep = (size_t)s->kp.E * s->kp.nb;
  if (ep) k_copy_poses<<<nblk((long long)ep), kT, 0, st>>>((int)ep, s->pos + (size_t)f_src * ep, s->rot + (size_t)f_src * ep, s->pos, s->rot);
  DD_CUDA(cudaGetLastError());
  return 0;
}

int dd_sim_get_state(dd_sim *s, int f, float *x, float *v, float *F, float *C, cudaStream_t st) {
  if (check_range(s, f, 0, "dd_sim_get_state")) return 1;
  int p = 0, k = 0;
  if (pick_copy(s, f, &p, &k, "dd_sim_get_state")) return 1;
  int EN = s->kp.EN;
  float *sx = s->stage, *sv = sx + 3 * (size_t)EN, *sF = sv + 3 * (size_t)EN, *sC = sF + 9 * (size_t)EN;
  k_unpack<<<nblk(EN), kT, 0, st>>>(EN, s->segs[k].perm, s->pslot(p), x ? sx : nullptr, v ? sv : nullptr, F ? sF : nullptr, C ? sC : nullptr);
  DD_CUDA(cudaGetLastError());
  if (x) DD_CUDA(cudaMemcpyAsync(x, sx, sizeof(float) * 3 * EN, cudaMemcpyDefault, st));
  if (v) DD_CUDA(cudaMemcpyAsync(v, sv, sizeof(float) * 3 * EN, cudaMemcpyDefault, st));
  if (F) DD_CUDA(cudaMemcpyAsync(F, sF, sizeof(float) * 9 * EN, cudaMemcpyDefault, st));
  if (C) DD_CUDA(cudaMemcpyAsync(C, sC, sizeof(float) * 9 * EN, cudaMemcpyDefault, st));
  if (finish_readback(st, {x, v, F, C})) return 1;
  return 0;
}

int dd_sim_set_poses(dd_sim *s, int f0, int count, const float *pos, const float *rot, cudaStream_t st) {
  if (!s) return fail("dd_sim_set_poses: null simulator");
  if (s->kp.nb == 0 || count == 0) return 0;
  if (f0 < 0 || count < 0 || f0 + count > s->slots) return fail("dd_sim_set_poses: slot range out of bounds");
  if (!pos || !rot) return fail("dd_sim_set_poses: null argument");
  size_t n = (size_t)count * s->kp.E * s->kp.nb;
  if (7 * n > s->stage_floats) return fail("dd_sim_set_poses: too many poses for the staging buffer; upload in smaller chunks");
  float *sp = s->stage, *sr = sp + 3 * n;
  DD_CUDA(cudaMemcpyAsync(sp, pos, sizeof(float) * 3 * n, cudaMemcpyDefault, st));
  DD_CUDA(cudaMemcpyAsync(sr, rot, sizeof(float) * 4 * n, cudaMemcpyDefault, st));
  size_t o = (size_t)f0 * s->kp.E * s->kp.nb;
  k_pad4<<<nblk((long long)n), kT, 0, st>>>((int)n, sp, 3, s->pos + o);
  k_pad4<<<nblk((long long)n), kT, 0, st>>>((int)n, sr, 4, s->rot + o);
  DD_CUDA(cudaGetLastError());
  return 0;
}
int dd_sim_get_poses(dd_sim *s, int f0, int count, float *pos, float *rot, cudaStream_t st) {
  if (!s) return fail("dd_sim_get_poses: null simulator");
  if (s->kp.nb == 0 || count == 0) return 0;
  if (f0 < 0 || count < 0 || f0 + count > s->slots) return fail("dd_sim_get_poses: slot range out of bounds");
  size_t n = (size_t)count * s->kp.E * s->kp.nb, o = (size_t)f0 * s->kp.E * s->kp.nb;
  if (7 * n > s->stage_floats) return fail("dd_sim_get_poses: too many poses for the staging buffer; download in smaller chunks");
  float *sp = s->stage, *sr = sp + 3 * n;
  k_unpad4<<<nblk((long long)n), kT, 0, st>>>((int)n, s->pos + o, 3, sp, 0);
  k_unpad4<<<nblk((long long)n), kT, 0, st>>>((int)n, s->rot + o, 4, sr, 0);
  DD_CUDA(cudaGetLastError());
  if (pos) DD_CUDA(cudaMemcpyAsync(pos, sp, sizeof(float) * 3 * n, cudaMemcpyDefault, st));
  if (rot) DD_CUDA(cudaMemcpyAsync(rot, sr, sizeof(float) * 4 * n, cudaMemcpyDefault, st));
  if (finish_readback(st, {pos, rot})) return 1;
  return 0;
}

int dd_sim_forward(dd_sim *s, int f0, int n, cudaStream_t st) {
  if (check_range(s, f0, n, "dd_sim_forward")) return 1;
  if (n == 0) return 0;
  // ---- plan: the first state must exist; a range that starts on a segment boundary re-sorts there unless the head copy is current
  bool rebuild0 = false;
  if (s->is_boundary(f0) && !s->valid(s->phys_cur(f0))) {
    if (!s->valid(s->phys_tail(f0))) return fail("dd_sim_forward: state " + std::to_string(f0) + " is not available");
    rebuild0 = true;
  } else if (!s->valid(s->phys_cur(f0))) {
    return fail("dd_sim_forward: state " + std::to_string(f0) + " is not available (dd_sim_set_state it, or run the substeps before it)");
  }
  for (int f = f0; f < f0 + n; ++f) {
    int k = s->seg_of(f);
    if (s->is_boundary(f) && (f > f0 || rebuild0)) segment_rebuilt(s, k, true, s->phys_cur(f));
    s->slot_epoch[s->phys_cur(f) + 1] = s->segs[k].epoch;
    if (s->is_boundary(f + 1)) s->slot_epoch[s->phys_cur(f + 1)] = -1;  // the head copy of the next segment is out of date now
  }
  size_t ep = (size_t)s->kp.E * s->kp.nb;
  auto body = [&](cudaStream_t q) {
    // reference: states[f+i+1].clear_grad in the forward pass (mpm/simulator.py:570-571) -- pose gradients only here,
    // particle gradients are overwritten, not accumulated, by the adjoint kernels
    if (ep) {
      cudaMemsetAsync(s->gpos + (size_t)(f0 + 1) * ep, 0, sizeof(float4) * ep * n, q);
      cudaMemsetAsync(s->grot + (size_t)(f0 + 1) * ep, 0, sizeof(float4) * ep * n, q);
    }
    for (int f = f0; f < f0 + n; ++f) {
      bool resort = s->is_boundary(f) && (f > f0 || rebuild0);
      if (resort) enqueue_build(s, s->seg_of(f), nullptr, s->phys_tail(f), s->seg_of(f - 1), s->phys_cur(f), q);
      if (s->cfg.tile_mode && (f == f0 || resort)) {  // scatter target of the first substep under this active list
        const Segment &sg = s->segs[s->seg_of(f)];
        launch_hot(s, k_zero_bricks, s->brick_blocks(), kT, 0, q, s->kp, sg.cnt, sg.active, s->G(f), nullptr);
        s->launches += 1;
      }
      if (s->cfg.svd_mode == 0) enqueue_forward_substep<0>(s, f, q); else enqueue_forward_substep<1>(s, f, q);
    }
  };
  return run_graphed(s, 0, f0, n, rebuild0 ? 1 : 0, st, body);
}

int dd_sim_zero_grad(dd_sim *s, int f, cudaStream_t st) {
  if (check_range(s, f, 0, "dd_sim_zero_grad")) return 1;
  size_t ep = (size_t)s->kp.E * s->kp.nb;
  DD_CUDA(cudaMemsetAsync(s->grad[f & 1], 0, sizeof(float) * 25 * (size_t)s->kp.EN, st));
  s->grad_holds[f & 1] = f;
  s->grad_order[f & 1] = -1;
  if (ep) {
    DD_CUDA(cudaMemsetAsync(s->gpos + (size_t)f * ep, 0, sizeof(float4) * ep, st));
    DD_CUDA(cudaMemsetAsync(s->grot + (size_t)f * ep, 0, sizeof(float4) * ep, st));
  }
  return 0;
}
// pose gradients of state f only (GradModel.zero_grad clears states[0]'s gradients, mpm/torch_wrapper.py:26)
int dd_sim_zero_pose_grads(dd_sim *s, int f0, int count, cudaStream_t st) {
  if (!s) return fail("dd_sim_zero_pose_grads: null simulator");
  if (f0 < 0 || count < 0 || f0 + count > s->slots) return fail("dd_sim_zero_pose_grads: slot range out of bounds");
  size_t ep = (size_t)s->kp.E * s->kp.nb;
  if (ep && count) {
    DD_CUDA(cudaMemsetAsync(s->gpos + (size_t)f0 * ep, 0, sizeof(float4) * ep * count, st));
    DD_CUDA(cudaMemsetAsync(s->grot + (size_t)f0 * ep, 0, sizeof(float4) * ep * count, st));
  }
  return 0;
}

int dd_sim_add_state_grad(dd_sim *s, int f, const float *gx, const float *gv, const float *gF, const float *gC, cudaStream_t st) {
  if (check_range(s, f, 0, "dd_sim_add_state_grad")) return 1;
  int p = 0, k = 0;
  if (grad_copy(s, f, &p, &k, "dd_sim_add_state_grad")) return 1;
  int EN = s->kp.EN;
  float *sx = s->stage, *sv = sx + 3 * (size_t)EN, *sF = sv + 3 * (size_t)EN, *sC = sF + 9 * (size_t)EN;
  if (gx) DD_CUDA(cudaMemcpyAsync(sx, gx, sizeof(float) * 3 * EN, cudaMemcpyDefault, st));
  if (gv) DD_CUDA(cudaMemcpyAsync(sv, gv, sizeof(float) * 3 * EN, cudaMemcpyDefault, st));
  if (gF) DD_CUDA(cudaMemcpyAsync(sF, gF, sizeof(float) * 9 * EN, cudaMemcpyDefault, st));
  if (gC) DD_CUDA(cudaMemcpyAsync(sC, gC, sizeof(float) * 9 * EN, cudaMemcpyDefault, st));
  k_add_grad<<<nblk(EN), kT, 0, st>>>(EN, s->segs[k].perm, gx ? sx : nullptr, gv ? sv : nullptr, gF ? sF : nullptr, gC ? sC : nullptr, s->grad[f & 1]);
  DD_CUDA(cudaGetLastError());
  return 0;
}

int dd_sim_get_state_grad(dd_sim *s, int f, float *gx, float *gv, float *gF, float *gC, cudaStream_t st) {
  if (check_range(s, f, 0, "dd_sim_get_state_grad")) return 1;
  int p = 0, k = 0;
  if (grad_copy(s, f, &p, &k, "dd_sim_get_state_grad")) return 1;
  int EN = s->kp.EN;
  float *sx = s->stage, *sv = sx + 3 * (size_t)EN, *sF = sv + 3 * (size_t)EN, *sC = sF + 9 * (size_t)EN;
  k_unpack<<<nblk(EN), kT, 0, st>>>(EN, s->segs[k].perm, s->grad[f & 1], gx ? sx : nullptr, gv ? sv : nullptr, gF ? sF : nullptr, gC ? sC : nullptr);
  DD_CUDA(cudaGetLastError());
  if (gx) DD_CUDA(cudaMemcpyAsync(gx, sx, sizeof(float) * 3 * EN, cudaMemcpyDefault, st));
  if (gv) DD_CUDA(cudaMemcpyAsync(gv, sv, sizeof(float) * 3 * EN, cudaMemcpyDefault, st));
  if (gF) DD_CUDA(cudaMemcpyAsync(gF, sF, sizeof(float) * 9 * EN, cudaMemcpyDefault, st));
  if (gC) DD_CUDA(cudaMemcpyAsync(gC, sC, sizeof(float) * 9 * EN, cudaMemcpyDefault, st));
  if (finish_readback(st, {gx, gv, gF, gC})) return 1;
  return 0;
}

int dd_sim_backward(dd_sim *s, int f0, int n, cudaStream_t st) {
  if (check_range(s, f0, n, "dd_sim_backward")) return 1;
  if (n == 0) return 0;
  const int ft = f0 + n, gi = ft & 1;
  if (s->grad_holds[gi] != ft) return fail("dd_sim_backward: no gradient seeded for state " + std::to_string(ft) + " (dd_sim_zero_grad + dd_sim_add_state_grad)");
  // ---- plan: the seed must be (or be brought) in the order of the segment that produced state ft
  const int want = s->seg_of(ft - 1);
  bool permute_top = false;
  if (s->grad_order[gi] >= 0 && s->grad_order[gi] != want) {
    if (s->is_boundary(ft) && s->grad_order[gi] == s->seg_of(ft)) permute_top = true;
    else return fail("dd_sim_backward: the gradient of state " + std::to_string(ft) + " is stored in the particle order of another segment");
  }
  for (int f = f0; f < ft; ++f) {
    if (!s->valid(s->phys_cur(f)) || !s->valid(s->phys_cur(f) + 1))
      return fail("dd_sim_backward: the checkpoints of substep " + std::to_string(f) + " are not available (run dd_sim_forward over it first; dd_sim_set_state or a re-sort replaced them)");
    if (s->is_boundary(f + 1) && (f + 1 < ft || permute_top) && !s->segs[s->seg_of(f + 1)].linked)
      return fail("dd_sim_backward: cannot back-propagate across state " + std::to_string(f + 1) + ": it was set with dd_sim_set_state, not computed");
  }
  auto body = [&](cudaStream_t q) {
    for (int f = ft - 1; f >= f0; --f) {
      if (s->is_boundary(f + 1) && (f + 1 < ft || permute_top)) enqueue_grad_permute(s, s->seg_of(f + 1), (f + 1) & 1, q);
      // brick checkpoints: substep f restores v_out of substep f-1 on its way when both use the same active list
      bool here = f == ft - 1 || s->seg_of(f) != s->seg_of(f + 1), below = s->brick_ckpt && f > f0 && s->seg_of(f - 1) == s->seg_of(f);
      if (s->cfg.svd_mode == 0) enqueue_backward_substep<0>(s, f, q, nullptr, here, below); else enqueue_backward_substep<1>(s, f, q, nullptr, here, below);
    }
  };
  int rc = run_graphed(s, 1, f0, n, permute_top ? 1 : 0, st, body);
  if (rc) return rc;
  s->grad_holds[f0 & 1] = f0;
  s->grad_order[f0 & 1] = s->seg_of(f0);
  s->grad_holds[(f0 + 1) & 1] = f0 + 1;
  s->grad_order[(f0 + 1) & 1] = s->seg_of(f0);
  return 0;
}

int dd_sim_get_pose_grads(dd_sim *s, int f0, int count, float *gpos, float *grot, cudaStream_t st) {
  if (!s) return fail("dd_sim_get_pose_grads: null simulator");
  if (s->kp.nb == 0 || count == 0) return 0;
  if (f0 < 0 || count < 0 || f0 + count > s->slots) return fail("dd_sim_get_pose_grads: slot range out of bounds");
  size_t n = (size_t)count * s->kp.E * s->kp.nb, o = (size_t)f0 * s->kp.E * s->kp.nb;
  if (7 * n > s->stage_floats) return fail("dd_sim_get_pose_grads: too many poses for the staging buffer; download in smaller chunks");
  float *sp = s->stage, *sr = sp + 3 * n;
  k_unpad4<<<nblk((long long)n), kT, 0, st>>>((int)n, s->gpos + o, 3, sp, 0);
  k_unpad4<<<nblk((long long)n), kT, 0, st>>>((int)n, s->grot + o, 4, sr, 0);
  DD_CUDA(cudaGetLastError());
  if (gpos) DD_CUDA(cudaMemcpyAsync(gpos, sp, sizeof(float) * 3 * n, cudaMemcpyDefault, st));
  if (grot) DD_CUDA(cudaMemcpyAsync(grot, sr, sizeof(float) * 4 * n, cudaMemcpyDefault, st));
  if (finish_readback(st, {gpos, grot})) return 1;
  return 0;
}

int dd_sim_add_pose_grads(dd_sim *s, int f, const float *gpos, const float *grot, cudaStream_t st) {
  if (check_range(s, f, 0, "dd_sim_add_pose_grads")) return 1;
  if (s->kp.nb == 0) return 0;
  size_t n = (size_t)s->kp.E * s->kp.nb, o = (size_t)f * n;
  float *sp = s->stage, *sr = sp + 3 * n;
  if (gpos) {
    DD_CUDA(cudaMemcpyAsync(sp, gpos, sizeof(float) * 3 * n, cudaMemcpyDefault, st));
    k_add4<<<nblk((long long)n), kT, 0, st>>>((int)n, sp, 3, s->gpos + o);
  }
  if (grot) {
    DD_CUDA(cudaMemcpyAsync(sr, grot, sizeof(float) * 4 * n, cudaMemcpyDefault, st));
    k_add4<<<nblk((long long)n), kT, 0, st>>>((int)n, sr, 4, s->grot + o);
  }
  DD_CUDA(cudaGetLastError());
  return 0;
}

// (a device destination is written directly; anything else goes through the staging buffer)
int dd_sim_compute_dist(dd_sim *s, int f, float *dist, cudaStream_t st) {
  if (check_range(s, f, 0, "dd_sim_compute_dist")) return 1;
  if (s->kp.nb == 0) return 0;
  if (!dist) return fail("dd_sim_compute_dist: null output");
  int p = 0, k = 0;
  if (pick_copy(s, f, &p, &k, "dd_sim_compute_dist")) return 1;
  size_t n = (size_t)s->kp.EN * s->kp.nb;
  const bool direct = on_device(dist);
  if (!direct && n > s->stage_floats) return fail("dd_sim_compute_dist: too many bodies for the staging buffer; pass a device pointer");
  k_obs<false><<<nblk(s->kp.EN), kT, 0, st>>>(s->kp, s->segs[k].perm, s->pslot(p), s->tables(f), direct ? dist : s->stage, s->kp.nb);
  DD_CUDA(cudaGetLastError());
  if (!direct) {
    DD_CUDA(cudaMemcpyAsync(dist, s->stage, sizeof(float) * n, cudaMemcpyDefault, st));
    DD_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}

int dd_sim_compute_dist_grad(dd_sim *s, int f, const float *dist_grad, cudaStream_t st) {
  if (check_range(s, f, 0, "dd_sim_compute_dist_grad")) return 1;
  if (s->kp.nb == 0) return 0;
  if (!dist_grad) return fail("dd_sim_compute_dist_grad: null argument");
  int p = 0, k = 0;
  if (grad_copy(s, f, &p, &k, "dd_sim_compute_dist_grad")) return 1;
  if (!s->valid(p)) return fail("dd_sim_compute_dist_grad: state " + std::to_string(f) + " is not available");
  size_t n = (size_t)s->kp.EN * s->kp.nb, ep = (size_t)s->kp.E * s->kp.nb;
  const bool direct = on_device(dist_grad);
  if (!direct) {
    if (n > s->stage_floats) return fail("dd_sim_compute_dist_grad: too many bodies for the staging buffer; pass a device pointer");
    DD_CUDA(cudaMemcpyAsync(s->stage, dist_grad, sizeof(float) * n, cudaMemcpyDefault, st));
  }
  k_obs_grad<false><<<nblk(s->kp.EN), kT, 0, st>>>(s->kp, s->segs[k].perm, s->pslot(p), s->tables(f), direct ? dist_grad : s->stage, s->kp.nb, s->grad[f & 1],
                                                   s->gpos + (size_t)f * ep, s->grot + (size_t)f * ep);
  DD_CUDA(cudaGetLastError());
  return 0;
}

// The particle observation of GradModel.get_obs (mpm/torch_wrapper.py:46-66) in one kernel: obs (E, N, 6 + nb) = [x | v | dist]
// in the caller's particle order, written straight into DEVICE memory (no staging, no concatenation).
int dd_sim_get_obs(dd_sim *s, int f, float *obs, cudaStream_t st) {
  if (check_range(s, f, 0, "dd_sim_get_obs")) return 1;
  if (!obs || !on_device(obs)) return fail("dd_sim_get_obs: obs must be a device pointer (host callers: dd_sim_get_state + dd_sim_compute_dist)");
  int p = 0, k = 0;
  if (pick_copy(s, f, &p, &k, "dd_sim_get_obs")) return 1;
  k_obs<true><<<nblk(s->kp.EN), kT, 0, st>>>(s->kp, s->segs[k].perm, s->pslot(p), s->tables(f), obs, 6 + s->kp.nb);
  DD_CUDA(cudaGetLastError());
  return 0;
}
// Its adjoint (GradModel.set_obs_grad, mpm/torch_wrapper.py:68-105): gobs[..., :3] and [..., 3:6] are added to the gradients of x
// and v of state f, gobs[..., 6:] is back-propagated through the signed distances into x and the pose gradients of state f.
int dd_sim_add_obs_grad(dd_sim *s, int f, const float *gobs, cudaStream_t st) {
  if (check_range(s, f, 0, "dd_sim_add_obs_grad")) return 1;
  if (!gobs || !on_device(gobs)) return fail("dd_sim_add_obs_grad: gobs must be a device pointer (host callers: dd_sim_add_state_grad + dd_sim_compute_dist_grad)");
  int p = 0, k = 0;
  if (grad_copy(s, f, &p, &k, "dd_sim_add_obs_grad")) return 1;
  if (!s->valid(p)) return fail("dd_sim_add_obs_grad: state " + std::to_string(f) + " is not available");
  size_t ep = (size_t)s->kp.E * s->kp.nb;
  k_obs_grad<true><<<nblk(s->kp.EN), kT, 0, st>>>(s->kp, s->segs[k].perm, s->pslot(p), s->tables(f), gobs, 6 + s->kp.nb, s->grad[f & 1],
                                                  s->gpos + (size_t)f * ep, s->grot + (size_t)f * ep);
  DD_CUDA(cudaGetLastError());
  return 0;
}

// Per-kernel device times of one forward + one backward substep (tile mode), CUDA events on `st`, averaged over reps.
// names_out receives '\n'-separated kernel labels; returns the number of kernels in *n_out.
int dd_sim_profile_substep(dd_sim *s, int f, int reps, float *ms_out, char *names_out, int names_cap, int *n_out, cudaStream_t st) {
  if (check_range(s, f, 1, "dd_sim_profile_substep")) return 1;
  if (!s->cfg.tile_mode || !s->grid_ckpt) return fail("dd_sim_profile_substep: needs tile_mode with grid checkpoints");
  if (s->grad_holds[(f + 1) & 1] != f + 1) return fail("dd_sim_profile_substep: seed a gradient for state f+1 first");
  if (!s->valid(s->phys_cur(f))) return fail("dd_sim_profile_substep: state f is not available");
  std::vector<cudaEvent_t> ev;
  std::vector<std::string> names;
  std::vector<double> acc;
  const Segment &sg = s->segs[s->seg_of(f)];
  for (int r = 0; r < reps; ++r) {
    size_t k = 0;
    auto new_event = [&]() {
      if (k >= ev.size()) { cudaEvent_t e_; cudaEventCreate(&e_); ev.push_back(e_); }
      cudaEventRecord(ev[k++], st);
    };
    launch_hot(s, k_zero_bricks, s->brick_blocks(), kT, 0, st, s->kp, sg.cnt, sg.active, s->G(f), nullptr);  // untimed: scatter target of this substep
    new_event();
    Mark mk = [&](const char *name) { if (r == 0) names.push_back(name); new_event(); };
    if (s->cfg.svd_mode == 0) { enqueue_forward_substep<0>(s, f, st, &mk); enqueue_backward_substep<0>(s, f, st, &mk); }
    else { enqueue_forward_substep<1>(s, f, st, &mk); enqueue_backward_substep<1>(s, f, st, &mk); }
    DD_CUDA(cudaStreamSynchronize(st));
    if (r == 0) acc.assign(names.size(), 0.0);
    for (size_t i = 0; i + 1 < k; ++i) { float ms = 0.f; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); acc[i] += ms; }
  }
  for (auto e_ : ev) cudaEventDestroy(e_);
  std::string joined;
  for (size_t i = 0; i < names.size(); ++i) { ms_out[i] = (float)(acc[i] / reps); joined += names[i]; joined += '\n'; }
  if ((int)joined.size() + 1 > names_cap) return fail("dd_sim_profile_substep: names buffer too small");
  std::memcpy(names_out, joined.c_str(), joined.size() + 1);
  *n_out = (int)names.size();
  s->slot_epoch[s->phys_cur(f) + 1] = sg.epoch;
  s->grad_holds[f & 1] = f;
  s->grad_order[f & 1] = s->seg_of(f);
  return 0;
}

// density-grid observation (mpm/simulator.py:323-354).  ids: (E*N) object ids in original order or NULL; out: (E, gx, gy, gz)
int dd_sim_compute_grid_mass(dd_sim *s, int f, const int *ids, int id, float *out, cudaStream_t st) {
  if (check_range(s, f, 0, "dd_sim_compute_grid_mass")) return 1;
  if (!out) return fail("dd_sim_compute_grid_mass: null output");
  if (id != -1 && !ids) return fail("dd_sim_compute_grid_mass: object ids required when id != -1");
  int p = 0, k = 0;
  if (pick_copy(s, f, &p, &k, "dd_sim_compute_grid_mass")) return 1;
  size_t eg = (size_t)s->kp.E * s->kp.G;
  float *buf = reinterpret_cast<float *>(s->ggrid);  // scratch: E*G floats fit in the float4 adjoint grid
  int *dids = reinterpret_cast<int *>(s->keys);
  if (ids) DD_CUDA(cudaMemcpyAsync(dids, ids, sizeof(int) * s->kp.EN, cudaMemcpyDefault, st));
  DD_CUDA(cudaMemsetAsync(buf, 0, sizeof(float) * eg, st));
  k_grid_mass<<<nblk(s->kp.EN), kT, 0, st>>>(s->kp, s->segs[k].perm, s->pslot(p), s->segs[k].mat0, ids ? dids : nullptr, id, buf, nullptr, nullptr, 0);
  DD_CUDA(cudaGetLastError());
  DD_CUDA(cudaMemcpyAsync(out, buf, sizeof(float) * eg, cudaMemcpyDefault, st));
  DD_CUDA(cudaMemsetAsync(buf, 0, sizeof(float) * eg, st));  // ggrid must stay zero outside the active bricks
  if (finish_readback(st, {out})) return 1;
  return 0;
}
int dd_sim_compute_grid_mass_grad(dd_sim *s, int f, const int *ids, int id, const float *grid_m_grad, cudaStream_t st) {
  if (check_range(s, f, 0, "dd_sim_compute_grid_mass_grad")) return 1;
  if (!grid_m_grad) return fail("dd_sim_compute_grid_mass_grad: null argument");
  if (id != -1 && !ids) return fail("dd_sim_compute_grid_mass_grad: object ids required when id != -1");
  int p = 0, k = 0;
  if (grad_copy(s, f, &p, &k, "dd_sim_compute_grid_mass_grad")) return 1;
  if (!s->valid(p)) return fail("dd_sim_compute_grid_mass_grad: state " + std::to_string(f) + " is not available");
  size_t eg = (size_t)s->kp.E * s->kp.G;
  float *buf = reinterpret_cast<float *>(s->ggrid);
  int *dids = reinterpret_cast<int *>(s->keys);
  if (ids) DD_CUDA(cudaMemcpyAsync(dids, ids, sizeof(int) * s->kp.EN, cudaMemcpyDefault, st));
  DD_CUDA(cudaMemcpyAsync(buf, grid_m_grad, sizeof(float) * eg, cudaMemcpyDefault, st));
  k_grid_mass<<<nblk(s->kp.EN), kT, 0, st>>>(s->kp, s->segs[k].perm, s->pslot(p), s->segs[k].mat0, ids ? dids : nullptr, id, nullptr, buf, s->grad[f & 1], 1);
  DD_CUDA(cudaGetLastError());
  DD_CUDA(cudaMemsetAsync(buf, 0, sizeof(float) * eg, st));
  return 0;
}

// diagnostics: out = {segment of state f, chunks, active bricks, occupied bricks, epoch, linked} (synchronises)
int dd_sim_segment_info(dd_sim *s, int f, int *out, cudaStream_t st) {
  if (check_range(s, f, 0, "dd_sim_segment_info")) return 1;
  if (!out) return fail("dd_sim_segment_info: null output");
  int k = s->seg_of(f), host[4] = {0, 0, 0, 0};
  if (s->cfg.tile_mode) {
    DD_CUDA(cudaMemcpyAsync(host, s->segs[k].cnt, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
    DD_CUDA(cudaStreamSynchronize(st));
  }
  out[0] = k; out[1] = host[0]; out[2] = host[1]; out[3] = host[2]; out[4] = s->segs[k].epoch; out[5] = s->segs[k].linked ? 1 : 0;
  out[6] = s->nseg; out[7] = s->L;
  return 0;
}

int dd_sim_sync(dd_sim *s, cudaStream_t st) {
  if (!s) return fail("dd_sim_sync: null simulator");
  int status = 0;
  if (s->status) DD_CUDA(cudaMemcpyAsync(&status, s->status, sizeof(int), cudaMemcpyDeviceToHost, st));
  DD_CUDA(cudaStreamSynchronize(st));
  if (status & 1) {
    DD_CUDA(cudaMemsetAsync(s->status, 0, sizeof(int), st));
    return fail("dd_sim_sync: a substep had more active bricks than the brick checkpoints hold (" + std::to_string(s->brick_cap) +
                " per substep): gradients computed since the last sync are invalid; create the simulator with grid_ckpt = 0 (replay) or fewer max_steps");
  }
  return 0;
}

}  // extern "C"
