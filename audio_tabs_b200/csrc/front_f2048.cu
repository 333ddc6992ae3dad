// front_f2048.cu -- k_front instantiations for frame_size 2048 (see frontend_kernel.cuh)
#include "front_inst.cuh"

cudaError_t b2_launch_front_2048(int in, int mode, b2::FrontParams &p, int num_sms, long long task_bound,
                                cudaStream_t st) {
  return b2::launch_front_size<2048>(in, mode, p, num_sms, task_bound, st);
}

cudaError_t b2_launch_pair_2048(int in, b2::FrontParams &p, int num_sms, long long task_bound, cudaStream_t st) {
  return b2::launch_pair_size<2048>(in, p, num_sms, task_bound, st);
}

cudaError_t b2_launch_warp_2048(int in, b2::FrontParams &p, int num_sms, long long task_bound, cudaStream_t st) {
  return b2::launch_warp_size<2048>(in, p, num_sms, task_bound, st);
}
