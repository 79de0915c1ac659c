"""ctypes tables for the fused engine entry points (ABI-2, ``dd_*`` in include/dexdeform_mpm.h)."""
import ctypes
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

P = c_void_p
S = c_void_p  # cudaStream_t


class dd_sim_config(ctypes.Structure):
    _fields_ = [("n_envs", c_int), ("n_particles", c_int), ("n_bodies", c_int), ("grid_x", c_int), ("grid_y", c_int),
                ("grid_z", c_int), ("max_steps", c_int), ("dx", c_float), ("dt", c_float), ("ground_friction", c_float),
                ("ground_height", c_float), ("gravity", c_float * 3), ("svd_mode", c_int), ("use_graphs", c_int), ("sort_particles", c_int), ("tile_mode", c_int), ("grid_ckpt", c_int), ("chunk_max", c_int),
                ("resort_interval", c_int)]


ABI2 = {
    "dd_last_error": (c_char_p, []),
    "dd_sim_create": (c_int, [ctypes.POINTER(dd_sim_config), ctypes.POINTER(c_void_p)]),
    "dd_sim_destroy": (None, [P]),
    "dd_sim_launch_count": (c_longlong, [P]),
    "dd_sim_set_material": (c_int, [P, P, P, P, S]),
    "dd_sim_set_bodies": (c_int, [P, P, P]),
    "dd_sim_set_state": (c_int, [P, c_int, P, P, P, P, S]),
    "dd_sim_get_state": (c_int, [P, c_int, P, P, P, P, S]),
    "dd_sim_set_poses": (c_int, [P, c_int, c_int, P, P, S]),
    "dd_sim_get_poses": (c_int, [P, c_int, c_int, P, P, S]),
    "dd_sim_roll": (c_int, [P, c_int, S]),
    "dd_sim_forward": (c_int, [P, c_int, c_int, S]),
    "dd_sim_zero_grad": (c_int, [P, c_int, S]),
    "dd_sim_zero_pose_grads": (c_int, [P, c_int, c_int, S]),
    "dd_sim_add_state_grad": (c_int, [P, c_int, P, P, P, P, S]),
    "dd_sim_get_state_grad": (c_int, [P, c_int, P, P, P, P, S]),
    "dd_sim_backward": (c_int, [P, c_int, c_int, S]),
    "dd_sim_get_pose_grads": (c_int, [P, c_int, c_int, P, P, S]),
    "dd_sim_add_pose_grads": (c_int, [P, c_int, P, P, S]),
    "dd_sim_compute_dist": (c_int, [P, c_int, P, S]),
    "dd_sim_compute_dist_grad": (c_int, [P, c_int, P, S]),
    "dd_sim_get_obs": (c_int, [P, c_int, P, S]),
    "dd_sim_add_obs_grad": (c_int, [P, c_int, P, S]),
    "dd_sim_compute_grid_mass": (c_int, [P, c_int, P, c_int, P, S]),
    "dd_sim_compute_grid_mass_grad": (c_int, [P, c_int, P, c_int, P, S]),
    "dd_sim_sync": (c_int, [P, S]),
    "dd_sim_segment_info": (c_int, [P, c_int, P, S]),
    "dd_sim_pose_table": (c_int, [P, P, P, P, P, P]),
    "dd_sim_pose_grad_table": (c_int, [P, P, P]),
    "dd_hand_create": (c_int, [c_int, c_int, P, P, P, c_int, P, P, P, c_int, P, P, P, P, P, P, P]),
    "dd_hand_destroy": (None, [P]),
    "dd_hand_fk": (c_int, [P, P, c_int, c_int, P, P, P, P, P, c_int, S]),
    "dd_hand_fk_grad": (c_int, [P, P, c_int, c_int, P, P, P, P, P, P, P, P, c_int, S]),
    "dd_sim_profile_substep": (c_int, [P, c_int, c_int, P, P, c_int, P, S]),
}


def bind_abi2(library):
    for name, (res, args) in ABI2.items():
        fn = getattr(library, name)
        fn.restype = res
        fn.argtypes = args
    return library
