"""Drives any library that speaks the reference C ABI (product, reference GPU, reference CPU shim, OracleLib) through
the reference's own call sequences: ``substep`` / ``substep_grad`` of mpm/simulator.py:561-585 with buffers laid out as
``State`` / ``TempState`` (mpm/simulator.py:46-145)."""
import numpy as np

from dexdeform_b200.types import array, float32, ivec3, mat3, quat, vec3

STATE_FIELDS = ("x", "v", "F", "C")


class Abi1Sim:
    def __init__(self, library, scene, max_steps):
        L = self.lib = library
        s = self.scene = scene
        n, nb = s["n"], s["nb"]
        self.n, self.nb = n, nb
        gd = s["grid_dim"]
        self.grid_dim = ivec3(int(gd[0]), int(gd[1]), int(gd[2]))
        G = self.G = int(gd[0]) * int(gd[1]) * int(gd[2])
        A = lambda dt, ln: array(dtype=dt, length=ln, library=L)
        self.states = []
        for _ in range(max_steps + 1):
            st = dict(x=A(vec3, n), v=A(vec3, n), F=A(mat3, n), C=A(mat3, n), x_grad=A(vec3, n), v_grad=A(vec3, n),
                      F_grad=A(mat3, n), C_grad=A(mat3, n), body_pos=A(vec3, max(nb, 1)), body_rot=A(quat, max(nb, 1)),
                      body_pos_grad=A(vec3, max(nb, 1)), body_rot_grad=A(quat, max(nb, 1)))
            self.states.append(st)
        t = self.temp = dict(
            grid_m=A(float32, G), grid_v_in=A(vec3, G), grid_v_out=A(vec3, G), grid_m_grad=A(float32, G),
            grid_v_in_grad=A(vec3, G), grid_v_out_grad=A(vec3, G), grid_body_v_in=A(vec3, G * (nb + 1)),
            F=A(mat3, n), U=A(mat3, n), V=A(mat3, n), sig=A(vec3, n), F_grad=A(mat3, n), U_grad=A(mat3, n), V_grad=A(mat3, n),
            sig_grad=A(vec3, n))
        self.grid_lower = array(dtype=ivec3, length=1, library=L)
        self.gravity = A(vec3, 1)
        self.gravity.upload(s["gravity"])
        self.mass, self.vol, self.mly = A(float32, n), A(float32, n), A(vec3, n)
        self.mass.upload(s["mass"]); self.vol.upload(s["vol"]); self.mly.upload(s["mu_lam_yield"])
        self.tfsr, self.args = A(quat, max(nb, 1)), A(quat, max(nb, 1))
        if nb:
            self.tfsr.upload(s["tfsr"]); self.args.upload(s["args"])
        self.stream = L.cuda_stream_create()
        self.dx, self.inv_dx, self.dt = float(s["dx"]), float(s["inv_dx"]), float(s["dt"])
        self.gf, self.gh = float(s["ground_friction"]), float(s["ground_height"])
        st0 = self.states[0]
        for k in STATE_FIELDS:
            st0[k].upload(s[k])
        if nb:
            for f in range(min(max_steps, len(s["pos"]) - 1) + 1):
                self.states[f]["body_pos"].upload(s["pos"][f])
                self.states[f]["body_rot"].upload(s["rot"][f])

    # ---- kernels with the argument order of mpm/simulator.py:435-551
    def compute_svd(self, cur):
        t = self.temp
        self.lib.compute_svd(cur["F"].data_ptr, cur["C"].data_ptr, t["F"].data_ptr, t["U"].data_ptr, t["V"].data_ptr, t["sig"].data_ptr,
                             self.dt, self.n, self.stream)

    def p2g(self, cur, nxt):
        t = self.temp
        self.lib.p2g(cur["x"].data_ptr, cur["v"].data_ptr, self.mass.data_ptr, self.vol.data_ptr, t["F"].data_ptr, t["U"].data_ptr,
                     t["sig"].data_ptr, t["V"].data_ptr, cur["C"].data_ptr, self.mly.data_ptr, self.grid_lower.data_ptr, self.grid_dim,
                     self.dx, self.inv_dx, self.dt, nxt["F"].data_ptr, t["grid_v_in"].data_ptr, t["grid_m"].data_ptr, self.n, self.stream)

    def grid_op(self, cur, nxt):
        t = self.temp
        self.lib.grid_op_v2(t["grid_m"].data_ptr, t["grid_v_in"].data_ptr, t["grid_body_v_in"].data_ptr, self.grid_lower.data_ptr,
                            self.gravity.data_ptr, cur["body_pos"].data_ptr, cur["body_rot"].data_ptr, nxt["body_pos"].data_ptr,
                            nxt["body_rot"].data_ptr, self.tfsr.data_ptr, self.args.data_ptr, self.dx, self.inv_dx, self.dt, self.gf,
                            t["grid_v_out"].data_ptr, self.grid_dim, self.nb, self.stream)

    def g2p(self, cur, nxt):
        t = self.temp
        self.lib.g2p(cur["x"].data_ptr, t["grid_v_out"].data_ptr, self.grid_lower.data_ptr, self.dx, self.inv_dx, self.dt, self.grid_dim,
                     nxt["v"].data_ptr, self.gh, nxt["C"].data_ptr, nxt["x"].data_ptr, self.n, self.stream)

    def g2p_grad(self, cur, nxt):
        t = self.temp
        self.lib.g2p_grad(cur["x"].data_ptr, t["grid_v_out"].data_ptr, self.grid_lower.data_ptr, self.dx, self.inv_dx, self.dt,
                          self.grid_dim, nxt["v"].data_ptr, self.gh, nxt["C"].data_ptr, nxt["x"].data_ptr, self.n, cur["x_grad"].data_ptr,
                          t["grid_v_out_grad"].data_ptr, nxt["v_grad"].data_ptr, nxt["C_grad"].data_ptr, nxt["x_grad"].data_ptr, self.stream)

    def grid_op_grad(self, cur, nxt):
        t = self.temp
        self.lib.grid_op_v2_grad(t["grid_m"].data_ptr, t["grid_v_in"].data_ptr, t["grid_body_v_in"].data_ptr, self.grid_lower.data_ptr,
                                 self.gravity.data_ptr, cur["body_pos"].data_ptr, cur["body_rot"].data_ptr, nxt["body_pos"].data_ptr,
                                 nxt["body_rot"].data_ptr, self.tfsr.data_ptr, self.args.data_ptr, t["grid_m_grad"].data_ptr,
                                 t["grid_v_in_grad"].data_ptr, cur["body_pos_grad"].data_ptr, cur["body_rot_grad"].data_ptr,
                                 nxt["body_pos_grad"].data_ptr, nxt["body_rot_grad"].data_ptr, self.dx, self.inv_dx, self.dt, self.gf,
                                 t["grid_v_out"].data_ptr, t["grid_v_out_grad"].data_ptr, self.grid_dim, self.nb, self.stream)

    def p2g_grad(self, cur, nxt):
        t = self.temp
        self.lib.p2g_grad(cur["x"].data_ptr, cur["v"].data_ptr, self.mass.data_ptr, self.vol.data_ptr, t["F"].data_ptr, t["U"].data_ptr,
                          t["sig"].data_ptr, t["V"].data_ptr, cur["C"].data_ptr, self.mly.data_ptr, self.grid_lower.data_ptr, self.grid_dim,
                          self.dx, self.inv_dx, self.dt, nxt["F"].data_ptr, t["grid_v_in"].data_ptr, t["grid_m"].data_ptr,
                          cur["x_grad"].data_ptr, cur["v_grad"].data_ptr, t["F_grad"].data_ptr, cur["C_grad"].data_ptr, t["U_grad"].data_ptr,
                          t["sig_grad"].data_ptr, t["V_grad"].data_ptr, nxt["F_grad"].data_ptr, t["grid_v_in_grad"].data_ptr,
                          t["grid_m_grad"].data_ptr, self.n, self.stream)

    def compute_svd_grad(self, cur):
        t = self.temp
        self.lib.compute_svd_grad(cur["F"].data_ptr, cur["C"].data_ptr, t["U"].data_ptr, t["V"].data_ptr, t["sig"].data_ptr,
                                  t["F_grad"].data_ptr, t["U_grad"].data_ptr, t["V_grad"].data_ptr, t["sig_grad"].data_ptr,
                                  cur["F_grad"].data_ptr, cur["C_grad"].data_ptr, self.dt, self.n, self.stream)

    def compute_dist(self, st, dist, dist_grad, need_grad):
        self.lib.compute_dist(st["x"].data_ptr, st["body_pos"].data_ptr, st["body_rot"].data_ptr, self.tfsr.data_ptr, self.args.data_ptr,
                              dist.data_ptr, self.nb, st["x_grad"].data_ptr, st["body_pos_grad"].data_ptr, st["body_rot_grad"].data_ptr,
                              dist_grad.data_ptr, int(need_grad), self.n, self.stream)

    # ---- sequences of mpm/simulator.py:561-585
    def clear_temp(self):
        for k in ("grid_m", "grid_v_in", "grid_v_out"):
            self.temp[k].zero(self.stream)

    def clear_temp_grad(self):
        for k in ("grid_v_in_grad", "grid_v_out_grad", "grid_m_grad", "sig_grad", "F_grad", "U_grad", "V_grad"):
            self.temp[k].zero(self.stream)

    def substep(self, f):
        cur, nxt = self.states[f], self.states[f + 1]
        self.clear_temp()
        self.compute_svd(cur)
        self.p2g(cur, nxt)
        self.grid_op(cur, nxt)
        self.g2p(cur, nxt)

    def substep_grad(self, f):
        cur, nxt = self.states[f], self.states[f + 1]
        self.clear_temp()
        self.clear_temp_grad()
        self.compute_svd(cur)
        self.p2g(cur, nxt)
        self.grid_op(cur, nxt)
        self.g2p_grad(cur, nxt)
        self.grid_op_grad(cur, nxt)
        self.p2g_grad(cur, nxt)
        self.compute_svd_grad(cur)

    def sync(self):
        self.lib.cuda_stream_sync(self.stream)

    def get(self, f, *names):
        self.sync()
        return {k: self.states[f][k].download() for k in (names or STATE_FIELDS)}

    def get_temp(self, *names):
        self.sync()
        return {k: self.temp[k].download() for k in names}


def loss_seed(n, seed=1):
    """Deterministic dL/d(state) used to seed backward passes in parity tests."""
    rng = np.random.default_rng(seed)
    return dict(x_grad=rng.normal(size=(n, 3)).astype(np.float32), v_grad=rng.normal(size=(n, 3)).astype(np.float32) * 0.1,
                F_grad=rng.normal(size=(n, 9)).astype(np.float32) * 0.01, C_grad=rng.normal(size=(n, 9)).astype(np.float32) * 1e-4)
