// placeholder until the ORB kernels land
#include "common.cuh"
int vslam_orb_init(vslam_ctx* ctx) { (void)ctx; return VSLAM_OK; }
void vslam_orb_free(vslam_ctx* ctx) { (void)ctx; }
