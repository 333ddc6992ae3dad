// front_multi.cu -- k_front_multi instantiations: all resolutions of one hop in one launch (see frontend_pair_kernel.cuh)
#include "front_inst.cuh"

cudaError_t b2_launch_multi(int in, b2::MultiParams &m, int num_sms, long long task_bound, cudaStream_t st) {
  switch (in) {
    case b2::IN_F32_MONO: return b2::launch_multi_one<b2::IN_F32_MONO>(m, num_sms, task_bound, st);
    case b2::IN_F32_STEREO: return b2::launch_multi_one<b2::IN_F32_STEREO>(m, num_sms, task_bound, st);
    case b2::IN_I16_MONO: return b2::launch_multi_one<b2::IN_I16_MONO>(m, num_sms, task_bound, st);
    case b2::IN_I16_STEREO: return b2::launch_multi_one<b2::IN_I16_STEREO>(m, num_sms, task_bound, st);
  }
  return cudaErrorInvalidValue;
}
