/*
 * dexdeform_mpm.h -- C ABI of libmaniskill_mpm.so as built by dexdeform_b200 (sm_100a).
 *
 * Part 1 ("ABI-1") is, symbol for symbol and argument for argument, the interface the reference binds with
 * ctypes in mpm/types.py:103-290 and implements in mpm/csrc/integrator.cu:1616-2125.  Dropping this library at
 * mpm/libmaniskill_mpm.so (mpm/types.py:12-19) replaces the reference's native backend without touching its Python.
 *   - all pointers are raw device pointers owned by the caller (mpm/types.py:294-312); kernels never allocate
 *   - vec3 = 3 floats, mat3 = 9 floats row-major, quat = (w,x,y,z), ivec3 = 3 ints (mpm/types.py:25-75)
 *   - `const int *grid_dim` is the reference's `ivec3 const&` (pointer to 3 ints)
 *   - outputs are accumulated into caller-zeroed buffers (mpm/simulator.py:72-85)
 *   - no return codes; CUDA errors are printed as "CUDA Error: ..." and execution continues (mpm/csrc/common.h:29-34)
 *
 * Part 2 ("ABI-2", dd_* symbols) is the fused, batched engine the new Python host classes drive; see DESIGN.md.
 */
#ifndef DEXDEFORM_MPM_H
#define DEXDEFORM_MPM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st *cudaStream_t;
#endif

/* ------------------------------------------------------------------------------------------------ ABI-1 */

/* integrator.cu:1619-1625 -- dynamic grid origin (never called by the reference's substep; kept for the ABI) */
void compute_grid_lower(void *particle_x, float dx, float inv_dx, void *grid_lower, int dim, cudaStream_t stream);

/* integrator.cu:1627-1638 -- newF = (I + dt C) F ; (U, sig, V) = svd(newF) */
void compute_svd(void *F, void *C, void *newF, void *U, void *V, void *sig, float dt, int dim, cudaStream_t stream);

/* integrator.cu:1640-1657 -- adjoint of compute_svd; adds into F_grad, C_grad, overwrites newF_grad with the total */
void compute_svd_grad(void *F, void *C, void *U, void *V, void *sig, void *newF_grad, void *U_grad, void *V_grad, void *sig_grad,
                      void *F_grad, void *C_grad, float dt, int dim, cudaStream_t stream);

/* integrator.cu:1659-1690 -- return mapping, stress, APIC scatter of mass and momentum; writes out_particle_F */
void p2g(void *particle_x, void *particle_v, void *particle_m, void *particle_vol, void *particle_F, void *particle_U, void *particle_sig,
         void *particle_V, void *particle_C, void *particle_mu_lam_yield, void *grid_lower, const int *grid_dim, float dx, float inv_dx,
         float dt, void *out_particle_F, void *out_grid_mv, void *out_grid_m, int dim, cudaStream_t stream);

/* integrator.cu:1692-1742 -- adjoint of p2g */
void p2g_grad(void *particle_x, void *particle_v, void *particle_m, void *particle_vol, void *particle_F, void *particle_U, void *particle_sig,
              void *particle_V, void *particle_C, void *particle_mu_lam_yield, void *grid_lower, const int *grid_dim, float dx, float inv_dx,
              float dt, void *out_particle_F, void *out_grid_mv, void *out_grid_m, void *particle_x_grad, void *particle_v_grad,
              void *particle_F_grad, void *particle_C_grad, void *particle_U_grad, void *particle_sig_grad, void *particle_V_grad,
              void *out_particle_F_grad, void *out_grid_v_grad, void *out_grid_m_grad, int dim, cudaStream_t stream);

/* integrator.cu:1754-1779 -- normalise, gravity, sequential SDF contact against n_bodies primitives, boundary conditions */
void grid_op_v2(void *grid_m, void *grid_v_in, void *grid_body_v_in, void *grid_lower, void *gravity, void *body_pos, void *body_rot,
                void *next_body_pos, void *next_body_rot, void *body_type_friction_softness_round, void *body_args, float dx, float inv_dx,
                float dt, float ground_friction, void *out_grid_v, const int *grid_dim, int n_bodies, cudaStream_t stream);

/* integrator.cu:1781-1819 -- adjoint of grid_op_v2 incl. pose gradients at t and t+1 */
void grid_op_v2_grad(void *grid_m, void *grid_v_in, void *grid_body_v_in, void *grid_lower, void *gravity, void *body_pos, void *body_rot,
                     void *next_body_pos, void *next_body_rot, void *body_type_friction_softness_round, void *body_args, void *grid_m_grad,
                     void *grid_v_in_grad, void *body_pos_grad, void *body_rot_grad, void *next_body_pos_grad, void *next_body_rot_grad,
                     float dx, float inv_dx, float dt, float ground_friction, void *out_grid_v, void *out_grid_v_grad, const int *grid_dim,
                     int n_bodies, cudaStream_t stream);

/* integrator.cu:1821-1835 -- APIC gather, advection, position clamp */
void g2p(void *particle_x, void *grid_v, void *grid_lower, float dx, float inv_dx, float dt, const int *grid_dim, void *out_particle_v,
         float ground_height, void *out_particle_C, void *out_particle_x, int dim, cudaStream_t stream);

/* integrator.cu:1837-1861 -- adjoint of g2p */
void g2p_grad(void *particle_x, void *grid_v, void *grid_lower, float dx, float inv_dx, float dt, const int *grid_dim, void *out_particle_v,
              float ground_height, void *out_particle_C, void *out_particle_x, int dim, void *particle_x_grad, void *grid_v_grad,
              void *out_particle_v_grad, void *out_particle_C_grad, void *out_particle_x_grad, cudaStream_t stream);

/* integrator.cu:1935-1960 -- particle-to-primitive signed distances (N, n_bodies) and their adjoint */
void compute_dist(void *particle_x, void *body_pos, void *body_rot, void *body_type_friction_softness_round, void *body_args, void *dist,
                  int n_bodies, void *particle_x_grad, void *body_pos_grad, void *body_rot_grad, void *dist_grad, int compute_grad, int dim,
                  cudaStream_t stream);

/* integrator.cu:1962-1984 -- density-grid observation of object `id` (-1 = all) and its adjoint */
void particle2mass(void *particle_x, void *particle_m, void *grid_lower, const int *grid_dim, float dx, float inv_dx, void *out_grid_m,
                   void *out_grid_m_grad, void *particle_x_grad, void *particle_ids, int id, int compute_grad, int dim, cudaStream_t stream);

/* integrator.cu:1990-2072 -- device memory and stream helpers used by mpm/types.py:294-400 */
void *cuda_alloc(size_t size);
void cuda_free(void *ptr);
void print_memory_info(void);
cudaStream_t cuda_stream_create(void);
void cuda_stream_destroy(cudaStream_t stream);
void cuda_stream_sync(cudaStream_t stream);
void cuda_upload(void *device_ptr, void *host_ptr, size_t size);
void cuda_download(void *host_ptr, void *device_ptr, size_t size);
void cuda_copy(void *dst, void *src, size_t size);
void cuda_copy2d(void *dst, size_t dpitch, void *src, size_t spitch, size_t width, size_t height);
void cuda_zero(void *ptr, size_t size);
void cuda_zero_async(void *ptr, size_t size, cudaStream_t stream);
void cuda_upload_async(void *device_ptr, void *host_ptr, size_t size, cudaStream_t stream);
void cuda_download_async(void *host_ptr, void *device_ptr, size_t size, cudaStream_t stream);
void cuda_copy_async(void *dst, void *src, size_t size, cudaStream_t stream);

/* integrator.cu:1863-1933, 2074-2124 -- renderer entry points.  Outside the hot path: exported as no-ops so that
 * mpm/types.py can still set their argtypes. */
typedef struct { void *array; unsigned long long texture; } dd_texture_resources;
#ifdef __cplusplus
void render(void *sdf_volume, void *box_min, void *box_max, void *color_volume, void *body_pos, void *body_rot, void *tfsr, void *body_args,
            void *camera_rot, void *camera_pos, void *camera_intrinsic, void *color_buffer, void *depth_buffer, float sdf_threshold,
            const int *grid_dim, int n_bodies, bool visualize_shape, const int *image_dim, int max_ray_depth, int spp, float ground_height,
            int seed, void *light_direction, cudaStream_t stream);
#endif
void particle_sdf(void *volume, void *particle_x, void *particle_color, void *bbox_min, void *bbox_max, int bake_size, const int *grid_dim,
                  float inv_dx, void *sdf, void *color, void *sdf_tmp, int n_particles, cudaStream_t stream);
dd_texture_resources create_volume(float *data, int x, int y, int z);
void destroy_volume(dd_texture_resources tex);


/* ------------------------------------------------------------------------------------------------ ABI-2
 * Fused, batched engine.  One dd_sim owns E independent environments of N particles each (same grid, same primitive
 * shapes, per-environment poses), a ring of per-substep state checkpoints in HBM (slot f = state after f substeps;
 * what mpm/simulator.py:206 keeps as states[f]) and two ping-pong gradient slots.  Every function returns 0 on
 * success and non-zero on failure with a message in dd_last_error().  Array arguments may be host or device pointers
 * (cudaMemcpyDefault); particle arrays are in the caller's ORIGINAL particle order with the reference's AoS layouts,
 * shaped (E, N, 3|9); pose arrays are (count, E, n_bodies, 3|4 wxyz).  Work is enqueued on `stream`.
 *
 * Replaces, for E environments at once: MPMSimulator.substep / substep_grad (mpm/simulator.py:561-585), set_pose
 * (:553-559), State.set_state/get_state (:118-134), get_dists (:294-321) and the gradient plumbing of
 * GradModel.set_obs_grad / diff_forward.backward (mpm/torch_wrapper.py:79-141). */
typedef struct dd_sim dd_sim;
typedef struct {
  int n_envs;          /* E */
  int n_particles;     /* N per environment; E*N must be a multiple of 4 */
  int n_bodies;        /* primitives per environment, <= 64 */
  int grid_x, grid_y, grid_z;
  int max_steps;       /* checkpoint slots - 1 */
  float dx, dt;
  float ground_friction, ground_height;   /* mpm/simulator.py:490-494 */
  float gravity[3];    /* as uploaded by the reference, i.e. cfg.gravity * 30 (mpm/simulator.py:385) */
  int svd_mode;        /* 0: reference-order fp64 Jacobi SVD, 1: fp32 in-register SVD (production) */
  int use_graphs;      /* capture forward/backward ranges into CUDA graphs, cached per (f0, n) */
  int sort_particles;  /* dd_sim_set_state re-orders particles by (env, 4^3-cell brick, cell); results are returned in the caller's order */
  int tile_mode;       /* 1: warp-private shared-memory tiles for the scatters + grid kernels on active bricks only; 0: global reductions on the dense grid */
  int grid_ckpt;       /* grids for the adjoint, so that it does not replay scatter + grid update: 0 none (replay); 1 one dense grid pair per
                          substep when that fits in HBM, else the ACTIVE bricks of every substep only ("brick checkpoints", a fixed
                          number of bricks per substep; dd_sim_sync reports a substep that needed more); 2 brick checkpoints always */
  int chunk_max;       /* particles per warp-chunk in tile mode (0 = choose from the problem size) */
  int resort_interval; /* tile mode: re-sort the particles on the device every this many substeps inside dd_sim_forward (0 = only in
                          dd_sim_set_state).  States at multiples of the interval are stored in both orders; the adjoint permutes
                          the gradient back.  mpm/simulator.py has no equivalent (its kernels are order independent). */
} dd_sim_config;

const char *dd_last_error(void);
int dd_sim_create(const dd_sim_config *cfg, dd_sim **out);
void dd_sim_destroy(dd_sim *sim);
long long dd_sim_launch_count(dd_sim *sim);   /* kernels launched (incl. graph replays) since creation */
int dd_sim_set_material(dd_sim *sim, const float *mass, const float *vol, const float *mu_lam_yield, cudaStream_t stream);
int dd_sim_set_bodies(dd_sim *sim, const float *tfsr, const float *args);   /* (nb,4) each; simulator.py:388-404 */
int dd_sim_set_state(dd_sim *sim, int f, const float *x, const float *v, const float *F, const float *C, cudaStream_t stream);
int dd_sim_get_state(dd_sim *sim, int f, float *x, float *v, float *F, float *C, cudaStream_t stream);   /* any may be NULL; synchronises */
int dd_sim_set_poses(dd_sim *sim, int f0, int count, const float *pos, const float *rot, cudaStream_t stream);
int dd_sim_get_poses(dd_sim *sim, int f0, int count, float *pos, float *rot, cudaStream_t stream);
/* rolling window of MPMSimulator.step (mpm/simulator.py:626-634): state f_src (particles + poses) becomes state 0, re-sorted on the device */
int dd_sim_roll(dd_sim *sim, int f_src, cudaStream_t stream);
int dd_sim_forward(dd_sim *sim, int f0, int n_substeps, cudaStream_t stream);    /* slots f0 -> f0+n */
int dd_sim_zero_grad(dd_sim *sim, int f, cudaStream_t stream);                   /* start a backward pass at state f */
int dd_sim_zero_pose_grads(dd_sim *sim, int f0, int count, cudaStream_t stream);  /* pose gradients of states f0 .. f0+count-1 only */
int dd_sim_add_state_grad(dd_sim *sim, int f, const float *gx, const float *gv, const float *gF, const float *gC, cudaStream_t stream);
int dd_sim_get_state_grad(dd_sim *sim, int f, float *gx, float *gv, float *gF, float *gC, cudaStream_t stream);
int dd_sim_backward(dd_sim *sim, int f0, int n_substeps, cudaStream_t stream);   /* substeps f0+n-1 ... f0 */
int dd_sim_get_pose_grads(dd_sim *sim, int f0, int count, float *gpos, float *grot, cudaStream_t stream);
int dd_sim_add_pose_grads(dd_sim *sim, int f, const float *gpos, const float *grot, cudaStream_t stream);
int dd_sim_compute_dist(dd_sim *sim, int f, float *dist, cudaStream_t stream);   /* (E, N, nb) */
int dd_sim_compute_dist_grad(dd_sim *sim, int f, const float *dist_grad, cudaStream_t stream);
/* GradModel.get_obs / set_obs_grad (mpm/torch_wrapper.py:46-105) without the host round trip and the concatenation: obs and gobs are
 * DEVICE arrays (E, N, 6 + nb) = [x | v | dist] in the caller's particle order.  dd_sim_add_obs_grad adds gobs[..., :6] to the gradient
 * of (x, v) of state f and back-propagates gobs[..., 6:] through the signed distances (into x and the pose gradients of state f). */
int dd_sim_get_obs(dd_sim *sim, int f, float *obs, cudaStream_t stream);
int dd_sim_add_obs_grad(dd_sim *sim, int f, const float *gobs, cudaStream_t stream);
/* density-grid observation of object `id` (-1: all particles) and its adjoint; ids (E*N ints, original order) may be NULL for id == -1 */
int dd_sim_compute_grid_mass(dd_sim *sim, int f, const int *ids, int id, float *out, cudaStream_t stream);
int dd_sim_compute_grid_mass_grad(dd_sim *sim, int f, const int *ids, int id, const float *grid_m_grad, cudaStream_t stream);
int dd_sim_sync(dd_sim *sim, cudaStream_t stream);
/* diagnostics (synchronises): out[8] = {segment of state f, chunks, active bricks, occupied bricks, epoch, linked, n segments, interval} */
int dd_sim_segment_info(dd_sim *sim, int f, int *out, cudaStream_t stream);
/* measurement aid: device time of every kernel of one forward + backward substep (CUDA events on `stream`) */
int dd_sim_profile_substep(dd_sim *sim, int f, int reps, float *ms_out, char *names_out, int names_cap, int *n_out, cudaStream_t stream);

/* device pointers of the engine's pose table: float4 (x,y,z,0) and (w,x,y,z) per (slot, env, body) */
int dd_sim_pose_table(dd_sim *sim, float **pos, float **rot, int *slots, int *n_envs, int *n_bodies);
int dd_sim_pose_grad_table(dd_sim *sim, float **gpos, float **grot);   /* same layout, gradients */

/* ---- Shadow-hand kinematics on the device (replaces HandSimulator.JointVel_Fk + hand_forward_kinematics, mpm/hand.py:347-428).
 * Tables as produced by dexdeform_b200.mujoco_parser.hand_tables; all pointer arguments of dd_hand_fk are DEVICE pointers. */
typedef struct dd_hand dd_hand;
int dd_hand_create(int n_hands, int n_ops, const int *op_kind, const int *op_index, const int *op_reset, int n_mats, const float *mats,
                   const float *joint_pos, const float *joint_axis, int n_geoms, const int *geom_joint, const float *geom_local,
                   const float *q_lower, const float *q_upper, const int *action_map, const float *action_scale, dd_hand **out);
void dd_hand_destroy(dd_hand *hand);
int dd_hand_fk(dd_hand *hand, dd_sim *sim, int f, int n_substeps, const float *base_pose, const float *joint_rot, const float *action,
               float *next_base, float *next_q, int has_base_action, cudaStream_t stream);
/* Reverse mode of dd_hand_fk (what torch autograd computes for the reference's JointVel_Fk): reads the pose gradients of states
 * f+1 .. f+n_substeps from the simulator and ADDS dL/d(base_pose) (E, nh, 4, 4), dL/d(joint_rot) (E, nh, 24), dL/d(action) (E, nh, 26)
 * to the caller-zeroed device buffers; g_next_base / g_next_q: gradients w.r.t. the end-of-step kinematic state (may be NULL). */
int dd_hand_fk_grad(dd_hand *hand, dd_sim *sim, int f, int n_substeps, const float *base_pose, const float *joint_rot, const float *action,
                    const float *g_next_base, const float *g_next_q, float *g_base, float *g_q, float *g_action, int has_base_action,
                    cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DEXDEFORM_MPM_H */
