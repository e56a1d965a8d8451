// Test infrastructure (not part of the product library): the covariance tail of K2 - closed-form symmetric eigen-solve
// and plane-to-plane regularisation, generalized-icp_b200/csrc/knn_cov.cuh + common.cuh - compiled for the HOST, so
// that tests/test_host.py can check it against numpy (gicp.py:11-16) without a GPU, including degenerate inputs.
#include "../../generalized-icp_b200/csrc/knn_cov.cuh"

extern "C" int gicp_test_regularised_cov(int dim, const double* S6, double lam_t, double lam_n, double* C) {
    if (dim == 2) gicp::regularised_cov<2>(S6, false, lam_t, lam_n, C);
    else gicp::regularised_cov<3>(S6, false, lam_t, lam_n, C);
    return 0;
}
