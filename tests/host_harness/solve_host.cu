// Test infrastructure (not part of the product library): the 2-D inner solver and the damped linear solve of
// generalized-icp_b200/csrc/solve.cuh compiled for the HOST, so that tests/test_host.py can check them against the
// reference's own inner problems without a GPU.  Built by the test with nvcc into a temporary directory.
#include "../../generalized-icp_b200/csrc/solve.cuh"

extern "C" int gicp_test_solve2d(const double* red, int max_it, double* out) {
    using DD = gicp::Dim<2>;
    double dR[2][2], dt[2], th = 0.0, f = 0.0;
    gicp::inner_solve_2d(red, red + DD::NH, red[DD::NQ], max_it, dR, dt, &th, &f);
    out[0] = dt[0]; out[1] = dt[1]; out[2] = th; out[3] = f;
    out[4] = dR[0][0]; out[5] = dR[0][1]; out[6] = dR[1][0]; out[7] = dR[1][1];
    return 0;
}

extern "C" int gicp_test_spd_solve6(const double* A36, double* b6) {
    double A[6][6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) A[i][j] = A36[i * 6 + j];
    return gicp::spd_solve<6>(A, b6) ? 0 : 1;
}
