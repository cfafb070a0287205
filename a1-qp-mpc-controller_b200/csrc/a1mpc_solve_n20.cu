// a1mpc_solve_n20.cu -- instantiations of the fused kernel for horizon 20 (BASELINE config 3)
#define A1MPC_HORIZON 20
#include "a1mpc_solve_n10.cu"
