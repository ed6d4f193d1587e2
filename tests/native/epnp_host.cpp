// Host build of the DEVICE EPnP source (csrc/epnp.cuh, CUDA qualifiers defined away) so that the port can be checked
// bit for bit against oracle/pnp_oracle.c and live cv2 on a machine without a GPU (tests/test_oracle_pnp.py).
// Test infrastructure only.
#define EPNP_HOST_BUILD
#define __device__
#define __restrict__
#define __forceinline__ inline
#define __noinline__
#include "../../stereo-visual-slam_b200/csrc/epnp.cuh"

extern "C" void epnp_host(const float* xyz, const float* uv, const int* idx, const double* K, double* R, double* t,
                          double* rvec, double* R_back) {
    static EpnpWork w;
    epnp5_dev(w, xyz, uv, idx, K[0], K[4], K[2], K[5], R, t);
    rodrigues_to_vec_dev(R, rvec);
    rodrigues_to_mat_dev(rvec, R_back);
}
