// front_f1024.cu -- k_front instantiations for frame_size 1024 (see frontend_kernel.cuh)
#include "front_inst.cuh"

cudaError_t b2_launch_front_1024(int in, int mode, b2::FrontParams &p, int num_sms, long long task_bound,
                                cudaStream_t st) {
  return b2::launch_front_size<1024>(in, mode, p, num_sms, task_bound, st);
}

cudaError_t b2_launch_pair_1024(int in, b2::FrontParams &p, int num_sms, long long task_bound, cudaStream_t st) {
  return b2::launch_pair_size<1024>(in, p, num_sms, task_bound, st);
}

cudaError_t b2_launch_warp_1024(int in, b2::FrontParams &p, int num_sms, long long task_bound, cudaStream_t st) {
  return b2::launch_warp_size<1024>(in, p, num_sms, task_bound, st);
}
