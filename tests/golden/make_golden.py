"""Generates tests/golden/golden_v1.json.

The reference holds NO golden vector for this path (test/test_mpc.cpp prints and returns 0) and cannot be
built offline (Eigen, OSQP, OsqpEigen, ROS absent), so these fixtures are produced by the committed oracle
(oracle/a1mpc_oracle.cpp) and are accepted into the file only if
  * the exact long-double solve carries a KKT certificate on the literal 12N-variable problem
    (stationarity <= 1e-12, primal/dual violation <= 1e-9), and
  * the independent OSQP-algorithm restatement run to eps 1e-11 agrees to <= 1e-5 N on the step-0 forces.
Run:  python tests/golden/make_golden.py      (CPU only, ~1 min)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "a1-qp-mpc-controller_b200")); sys.path.insert(0, ROOT)
import a1mpc
from oracle import oracle_py as O

WEIGHTS = {
    "gazebo": dict(mass=12.0, inertia=(0.0158533, 0, 0, 0, 0.0377999, 0, 0, 0, 0.0456542),
                   q=(20, 10, 1, 0, 0, 420, .05, .05, .05, 30, 30, 10, 0), r=(1e-7,) * 12),          # config/gazebo_a1_mpc.yaml
    "hardware": dict(mass=13.5, inertia=(0.0178533, 0, 0, 0, 0.0377999, 0, 0, 0, 0.0456542),
                     q=(150, 150, 50, 0, 0, 80, .2, .2, .2, .3, .3, .3, 0), r=(1e-2, 1e-2, 1e-3) * 4),  # config/hardware_a1_mpc.yaml
}


def main():
    cases = []
    for horizon, wname, cid, nqp in ((10, "gazebo", 2, 10), (10, "gazebo", 4, 6), (10, "hardware", 2, 4), (10, "hardware", 4, 4), (20, "gazebo", 2, 4)):
        st = a1mpc.gen_states(64, cid, stream=99)
        # make sure every stance-count class is present: force a few patterns
        pat = [0b1001, 0b0110, 0b1111, 0b0001, 0b0111, 0b1110, 0b0011, 0b1000, 0b1011, 0b0101]
        for i in range(nqp):
            st["contact"][i] = pat[i % len(pat)]
        if horizon == 20:
            st["contact"][:nqp] = [0b1001, 0b0110, 0b0001, 0b0111][:nqp]
        cfg = O.make_config(horizon=horizon, **WEIGHTS[wname])
        ob = O.Batch(st["x0"][:, :nqp], st["rot"][:, :nqp], st["foot"][:, :nqp], st["ref"][:, :nqp], st["contact"][:nqp])
        f, info = O.compute_grf_batch(cfg, ob, O.MODE_EXACT)
        ft, infot = O.compute_grf_batch(cfg, ob, O.MODE_OSQP_TIGHT)
        for b in range(nqp):
            assert info[b, 1] == 1, "exact solve not verified"
            assert info[b, 2] <= 1e-12 and info[b, 3] <= 1e-9 and info[b, 4] <= 1e-9, info[b]
            d = float(np.abs(f[:, b] - ft[:, b]).max())
            assert d <= 1e-5, "OSQP-tight disagrees by %.2e N (case %s N=%d b=%d)" % (d, wname, horizon, b)
            cases.append(dict(horizon=horizon, weights=wname, x0=st["x0"][:, b].tolist(), rot=st["rot"][:, b].tolist(),
                              foot=st["foot"][:, b].tolist(), ref=st["ref"][:, b].tolist(), contact=int(st["contact"][b]),
                              f_body=f[:, b].tolist(), osqp_tight_diff=d, kkt=[float(info[b, 2]), float(info[b, 3]), float(info[b, 4])]))
    # the reference's own (only) standalone driver state: test/test_mpc.cpp:15-91
    cfg, ob = O.test_mpc_fixture()
    f, info = O.compute_grf_batch(cfg, ob, O.MODE_EXACT)
    fixture = dict(f_body=f[:, 0].tolist(),
                   survey_known_answer=dict(FL=[0.0, -12.8370306335, 42.7901021118], RL=[0.0, -12.8370306335, 42.7901021118],
                                            source="SURVEY.md Appendix C (surveyor's independent numpy derivation)"))
    out = dict(version=1, weights=WEIGHTS, cases=cases, test_mpc_fixture=fixture)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote %d cases" % len(cases))


if __name__ == "__main__":
    main()
