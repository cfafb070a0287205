"""ctypes binding of the CPU oracle (oracle/liba1mpc_oracle.so).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (a1-qp-mpc-controller_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODE_EXACT, MODE_OSQP_DEFAULT, MODE_OSQP_TIGHT, MODE_BUILD_ONLY = 0, 1, 2, 3


class Config(C.Structure):
    """mirror of a1mpc_config (include/a1mpc.h)"""
    _fields_ = [("horizon", C.c_int), ("precision", C.c_int), ("dt", C.c_double),
                ("mu", C.c_double), ("fz_min", C.c_double), ("fz_max", C.c_double),
                ("mass", C.c_double), ("inertia", C.c_double * 9),
                ("q", C.c_double * 13), ("r", C.c_double * 12),
                ("max_iter", C.c_int), ("tol", C.c_double)]


class Inputs(C.Structure):
    _fields_ = [("x0", C.c_void_p), ("rot", C.c_void_p), ("foot", C.c_void_p), ("ref", C.c_void_p),
                ("contact", C.c_void_p), ("ld", C.c_size_t)]


def build(force=False):
    so = os.path.join(_HERE, "liba1mpc_oracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "a1mpc_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liba1mpc_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.oracle_time_reference_path.restype = C.c_double
    return _LIB


def make_config(horizon=10, dt=0.0025, mu=0.3, fz_max=180.0, mass=12.0,
                inertia=(0.0158533, 0, 0, 0, 0.0377999, 0, 0, 0, 0.0456542),
                q=(20, 10, 1, 0, 0, 420, .05, .05, .05, 30, 30, 10, 0), r=(1e-7,) * 12):
    c = Config()
    c.horizon, c.precision, c.dt, c.mu, c.fz_min, c.fz_max, c.mass = horizon, 64, dt, mu, 0.0, fz_max, mass
    c.inertia[:] = inertia
    c.q[:] = q
    c.r[:] = r
    c.max_iter, c.tol = 0, 0.0
    return c


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Batch:
    """host SoA batch: x0[12,B], rot[9,B], foot[12,B], ref[9,B], contact[B]"""

    def __init__(self, x0, rot, foot, ref, contact):
        self.x0 = np.ascontiguousarray(x0, dtype=np.float64)
        self.rot = np.ascontiguousarray(rot, dtype=np.float64)
        self.foot = np.ascontiguousarray(foot, dtype=np.float64)
        self.ref = np.ascontiguousarray(ref, dtype=np.float64)
        self.contact = np.ascontiguousarray(contact, dtype=np.uint32)
        self.B = self.x0.shape[1]
        assert self.x0.shape == (12, self.B) and self.rot.shape == (9, self.B)
        assert self.foot.shape == (12, self.B) and self.ref.shape == (9, self.B) and self.contact.shape == (self.B,)

    def c_inputs(self):
        s = Inputs()
        s.x0, s.rot, s.foot, s.ref, s.contact, s.ld = _ptr(self.x0), _ptr(self.rot), _ptr(self.foot), _ptr(self.ref), _ptr(self.contact), self.B
        return s

    def slice(self, lo, hi):
        return Batch(self.x0[:, lo:hi], self.rot[:, lo:hi], self.foot[:, lo:hi], self.ref[:, lo:hi], self.contact[lo:hi])


def build_qp(cfg, batch, b):
    N = cfg.horizon
    n, m = 12 * N, 20 * N
    H = np.zeros((n, n)); g = np.zeros(n); A = np.zeros((m, n)); lb = np.zeros(m); ub = np.zeros(m)
    inp = batch.c_inputs()
    lib().oracle_build_qp(C.byref(cfg), C.byref(inp), b, _ptr(H), _ptr(g), _ptr(A), _ptr(lb), _ptr(ub))
    return H, g, A, lb, ub


def rollout(cfg, batch, b):
    """mpc_states, mpc_states_d, A_d, B_d_list, A_qp, B_qp of QP b (the ConvexMpc / A1CtrlStates members compute_grf leaves)"""
    N = cfg.horizon
    o = dict(mpc_states=np.zeros(13), mpc_states_d=np.zeros(13 * N), A_d=np.zeros((13, 13)), B_d_list=np.zeros((13 * N, 12)),
             A_qp=np.zeros((13 * N, 13)), B_qp=np.zeros((13 * N, 12 * N)))
    inp = batch.c_inputs()
    lib().oracle_rollout(C.byref(cfg), C.byref(inp), b, *[_ptr(o[k]) for k in ("mpc_states", "mpc_states_d", "A_d", "B_d_list", "A_qp", "B_qp")])
    return o


def qp_mats(cfg, A_d, B_d_list, x0, x_d):
    n = 12 * cfg.horizon
    H = np.zeros((n, n)); g = np.zeros(n)
    A_d = np.ascontiguousarray(A_d, dtype=np.float64); B_d_list = np.ascontiguousarray(B_d_list, dtype=np.float64)
    x0 = np.ascontiguousarray(x0, dtype=np.float64); x_d = np.ascontiguousarray(x_d, dtype=np.float64)
    lib().oracle_qp_mats(C.byref(cfg), _ptr(A_d), _ptr(B_d_list), _ptr(x0), _ptr(x_d), _ptr(H), _ptr(g))
    return H, g


def compute_grf_batch(cfg, batch, mode=MODE_EXACT, eps=0.0, nthreads=1, want_u=False):
    """returns f_body [12,B], info [B,8] (, u_full [B,12N])"""
    B = batch.B
    f = np.zeros((12, B)); info = np.zeros((B, 8))
    u = np.zeros((B, 12 * cfg.horizon)) if want_u else None
    inp = batch.c_inputs()
    lib().oracle_compute_grf_batch(C.byref(cfg), B, C.byref(inp), mode, C.c_double(eps), nthreads, _ptr(f),
                                   _ptr(u) if want_u else None, _ptr(info))
    return (f, info, u) if want_u else (f, info)


def compute_grf_batch_ext(cfg, batch, sched=None, normals=None, mode=MODE_EXACT, nthreads=1, want_u=False):
    """BASELINE config 4 (extension): per-step contact schedule [N,B] and/or terrain normals [12,B]"""
    B = batch.B
    f = np.zeros((12, B)); info = np.zeros((B, 8))
    u = np.zeros((B, 12 * cfg.horizon)) if want_u else None
    sc = np.ascontiguousarray(sched, dtype=np.uint32) if sched is not None else None
    nm = np.ascontiguousarray(normals, dtype=np.float64) if normals is not None else None
    inp = batch.c_inputs()
    lib().oracle_compute_grf_batch_ext(C.byref(cfg), B, C.byref(inp), _ptr(sc) if sc is not None else None, _ptr(nm) if nm is not None else None,
                                       mode, nthreads, _ptr(f), _ptr(u) if want_u else None, _ptr(info))
    return (f, info, u) if want_u else (f, info)


def solve_dense(cfg, H, g, contact, mode=MODE_EXACT):
    n = 12 * cfg.horizon
    u = np.zeros(n); info = np.zeros(8)
    H = np.ascontiguousarray(H, dtype=np.float64); g = np.ascontiguousarray(g, dtype=np.float64)
    lib().oracle_solve_dense(C.byref(cfg), _ptr(H), _ptr(g), C.c_uint32(int(contact)), mode, _ptr(u), _ptr(info))
    return u, info


def grf_qp_single(root_acc, rot_z, rot, foot, contact, mode=MODE_EXACT):
    f = np.zeros(12); info = np.zeros(8)
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (root_acc, rot_z, rot, foot)]
    lib().oracle_grf_qp_single(_ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), C.c_uint32(int(contact)), mode, _ptr(f), _ptr(info))
    return f, info


def joint_torques(f_grf, f_kin, jac, contact, km_foot, torques_gravity, tau_prev=None):
    """one robot: f_grf[12], f_kin[12], jac[36] -> tau[12]"""
    tau = np.zeros(12) if tau_prev is None else np.array(tau_prev, dtype=np.float64)
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (f_grf, f_kin, jac, km_foot, torques_gravity)]
    lib().oracle_joint_torques(_ptr(a[0]), _ptr(a[1]), _ptr(a[2]), C.c_uint32(int(contact)), _ptr(a[3]), _ptr(a[4]), _ptr(tau))
    return tau


def update_plan(gp, movement_mode, gait_counter, gait_counter_speed, lin_vel, lin_vel_d, rot_z, rot, root_pos):
    """one robot; gp = a1mpc.GaitParams-like (fields counter_per_gait, ..., horizon)"""
    gc = np.array(gait_counter, dtype=np.float64)
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (gait_counter_speed, lin_vel, lin_vel_d, rot_z, rot, root_pos)]
    dfp = np.array(list(gp.default_foot_pos), dtype=np.float64)
    plan = C.c_uint32(); sched = np.zeros(gp.horizon, dtype=np.uint32)
    trel = np.zeros(12); tabs = np.zeros(12); tw = np.zeros(12)
    lib().oracle_update_plan(C.c_double(gp.counter_per_gait), C.c_double(gp.counter_per_swing), C.c_double(gp.control_dt), _ptr(dfp),
                             C.c_double(gp.foot_delta_x_limit), C.c_double(gp.foot_delta_y_limit), int(movement_mode), _ptr(gc), _ptr(a[0]),
                             _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), _ptr(a[4]), _ptr(a[5]), int(gp.horizon), C.byref(plan), _ptr(sched),
                             _ptr(trel), _ptr(tabs), _ptr(tw))
    return gc, int(plan.value), sched, trel, tabs, tw


def leg_kinematics(q, rho_opt, rho_fix):
    """one leg: q[3], rho_opt[3], rho_fix[5] -> p[3], J[3,3] (J[a,k] = d p_a / d q_k)"""
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (q, rho_opt, rho_fix)]
    p = np.zeros(3); J = np.zeros(9)
    lib().oracle_leg_kinematics(_ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(p), _ptr(J))
    return p, J.reshape(3, 3)


def ref_leg_kinematics(q, rho_opt, rho_fix):
    """the REFERENCE's own A1Kinematics::fk / jac, from oracle/_ref/libref_kin.so (`make -C oracle ref`; only where
    /root/reference exists).  Returns None when that build is absent."""
    so = os.path.join(_HERE, "_ref", "libref_kin.so")
    if not os.path.exists(so):
        return None
    global _REFKIN
    try:
        _REFKIN
    except NameError:
        _REFKIN = C.CDLL(so)
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (q, rho_opt, rho_fix)]
    p = np.zeros(3); J = np.zeros(9)
    _REFKIN.ref_leg_kinematics(_ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(p), _ptr(J))
    return p, J.reshape(3, 3)


def ekf_init(foot_pos_rel, rot):
    """one robot: foot_pos_rel[12] leg-major, rot[9] -> x[18], P[18,18]"""
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (foot_pos_rel, rot)]
    x = np.zeros(18); P = np.zeros(324)
    lib().oracle_ekf_init(_ptr(a[0]), _ptr(a[1]), _ptr(x), _ptr(P))
    return x, P.reshape(18, 18)


def ekf_update(x, P, dt, assume_flat_ground, movement_mode, imu_acc, imu_ang_vel, rot, foot_pos_rel, foot_vel_rel, foot_force):
    """one robot, returns new x[18], P[18,18], root_pos, root_lin_vel, estimated_contacts mask, rc"""
    x = np.array(x, dtype=np.float64); P = np.ascontiguousarray(np.array(P, dtype=np.float64).reshape(324))
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (imu_acc, imu_ang_vel, rot, foot_pos_rel, foot_vel_rel, foot_force)]
    pos = np.zeros(3); vel = np.zeros(3); ec = C.c_uint32()
    rc = lib().oracle_ekf_update(C.c_double(dt), int(assume_flat_ground), C.c_uint32(int(movement_mode)), *[_ptr(v) for v in a], _ptr(x), _ptr(P),
                                 _ptr(pos), _ptr(vel), C.byref(ec))
    return x, P.reshape(18, 18), pos, vel, int(ec.value), rc


def time_reference_path(cfg, batch, nthreads):
    f = np.zeros((12, batch.B))
    inp = batch.c_inputs()
    sec = lib().oracle_time_reference_path(C.byref(cfg), batch.B, C.byref(inp), nthreads, _ptr(f))
    return sec, f


def hardware_threads():
    return int(lib().oracle_hardware_threads())


def effective_cores():
    """host cores this process may actually use: min(scheduler affinity, cgroup CPU quota).  std::thread::hardware_concurrency()
    (hardware_threads) reports the machine's logical CPUs even inside a container limited to a fraction of them -- the 1-GPU
    lease of round 1 said 128 and delivered 11."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    quota = None
    try:
        with open("/sys/fs/cgroup/cpu.max") as fh:          # cgroup v2: "<quota> <period>" or "max <period>"
            q, per = fh.read().split()
            if q != "max":
                quota = float(q) / float(per)
    except Exception:
        try:
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as fq, open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as fp:
                q, per = float(fq.read()), float(fp.read())
                if q > 0:
                    quota = q / per
        except Exception:
            pass
    eff = n if quota is None else max(1, min(n, int(np.ceil(quota))))
    return eff, dict(affinity=n, cgroup_quota=quota, hardware_concurrency=hardware_threads())


_GEN = None


def gen_states(B, config_id=2, stream=0):
    """the benchmark's synthetic states from oracle/liba1mpc_gen.so (csrc/a1mpc_gen.cpp compiled host-only): identical to
    a1mpc.gen_states, without loading the CUDA library"""
    global _GEN
    if _GEN is None:
        so = os.path.join(_HERE, "liba1mpc_gen.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _GEN = C.CDLL(so)
    x0 = np.zeros((12, B)); rot = np.zeros((9, B)); foot = np.zeros((12, B)); ref = np.zeros((9, B)); contact = np.zeros(B, dtype=np.uint32)
    rc = _GEN.a1mpc_gen_states(int(config_id), C.c_uint64(int(stream)), int(B), _ptr(x0), _ptr(rot), _ptr(foot), _ptr(ref), _ptr(contact))
    assert rc == 0
    return dict(x0=x0, rot=rot, foot=foot, ref=ref, contact=contact)


def test_mpc_fixture():
    """The hand-built state of the reference's only standalone driver (test/test_mpc.cpp:15-91)."""
    cfg = make_config(horizon=10, mass=15.0, q=(1, 1, 1, 0, 0, 50, 0, 0, 1, 1, 1, 1, 0), r=(1e-6,) * 12)
    x0 = np.zeros((12, 1)); x0[5, 0] = 0.15
    rot = np.eye(3).reshape(9, 1)
    foot = np.array([[.17, .15, -.35], [.17, -.15, -.35], [-.17, .15, -.35], [-.17, -.15, -.35]]).reshape(12, 1)
    ref = np.zeros((9, 1)); ref[8, 0] = 0.15
    contact = np.array([0b0101], dtype=np.uint32)  # FL and RL (contacts[0], contacts[2])
    return cfg, Batch(x0, rot, foot, ref, contact)
