// C entry points around the reference's own A1Kinematics (compiled from /root/reference by `make -C oracle ref`).
// TEST INFRASTRUCTURE: pins oracle_leg_kinematics and generates tests/golden/kinematics_v1.json.
#include "legKinematics/A1Kinematics.h"

extern "C" {
// p[3]; J[9] ROW-major (J[3a+k] = d p_a / d q_k); the reference fills its Matrix3d column by column
int ref_leg_kinematics(const double* q, const double* rho_opt, const double* rho_fix, double* p, double* J) {
  A1Kinematics kin;
  Eigen::Vector3d qv;
  Eigen::VectorXd ro(3), rf(5);
  for (int i = 0; i < 3; ++i) { qv.data()[i] = q[i]; ro.data()[i] = rho_opt[i]; }
  for (int i = 0; i < 5; ++i) rf.data()[i] = rho_fix[i];
  Eigen::Vector3d pv = kin.fk(qv, ro, rf);
  Eigen::Matrix3d Jm = kin.jac(qv, ro, rf);
  for (int a = 0; a < 3; ++a) p[a] = pv.data()[a];
  for (int a = 0; a < 3; ++a)
    for (int k = 0; k < 3; ++k) J[3 * a + k] = Jm.data()[3 * k + a];
  return 0;
}
}
