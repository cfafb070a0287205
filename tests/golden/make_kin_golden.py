"""Generates tests/golden/kinematics_v1.json from the REFERENCE's own A1Kinematics::fk / jac, compiled from
/root/reference/src/a1_cpp/src/legKinematics/A1Kinematics.cpp by `make -C oracle ref` (oracle/_ref/libref_kin.so; works only
in the authoring container, where the reference is mounted).  The vectors travel; the reference does not.
   python tests/golden/make_kin_golden.py"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "all", "ref"])
from oracle import oracle_py as O  # noqa: E402

rng = np.random.default_rng(20260923)
# rho_fix of the four A1 legs (GazeboA1ROS.cpp:76-97): leg_offset_x, leg_offset_y, motor_offset, upper/lower leg length
RHO_FIX = [[0.1805, 0.047, 0.0838, 0.21, 0.21], [0.1805, -0.047, -0.0838, 0.21, 0.21],
           [-0.1805, 0.047, 0.0838, 0.21, 0.21], [-0.1805, -0.047, -0.0838, 0.21, 0.21]]
cases = []
for i in range(96):
    leg = i % 4
    q = (np.array([0.0, 0.8, -1.6]) + rng.normal(0, 0.5, 3)) if i >= 8 else np.array([[0, 0, 0], [0.3, 0, 0], [0, 0.7, 0], [0, 0, -1.2]][i % 4], dtype=float)
    rho_opt = np.zeros(3) if i % 3 == 0 else rng.normal(0, 0.02, 3)     # the reference runs with rho_opt = 0 (GazeboA1ROS.cpp:95)
    res = O.ref_leg_kinematics(q, rho_opt, RHO_FIX[leg])
    assert res is not None, "oracle/_ref/libref_kin.so missing: run where /root/reference exists"
    p, J = res
    cases.append(dict(leg=leg, q=q.tolist(), rho_opt=rho_opt.tolist(), rho_fix=RHO_FIX[leg], p=p.tolist(), J=J.reshape(9).tolist()))
out = dict(source="reference A1Kinematics::fk / jac (legKinematics/A1Kinematics.cpp), built by oracle/Makefile target `ref`",
           layout="J row-major: J[3a+k] = d p_a / d q_k", cases=cases)
with open(os.path.join(ROOT, "tests", "golden", "kinematics_v1.json"), "w") as fh:
    json.dump(out, fh, indent=0)
print("wrote", len(cases), "cases")
