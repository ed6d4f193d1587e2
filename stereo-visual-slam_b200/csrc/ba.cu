// placeholder until the BA kernels land
#include "common.cuh"
int vslam_ba_init(vslam_ctx* ctx) { (void)ctx; return VSLAM_OK; }
void vslam_ba_free(vslam_ctx* ctx) { (void)ctx; }
