// VN-DGCNN shape encoder entry points (kernels land in a later milestone of this round).
#include "smb_common.cuh"
#include "smb_kernels.h"

extern "C" {

size_t smb_encoder_workspace_bytes(const smb_encoder_weights* w, int32_t n_clouds, int32_t n_points) {
  (void)w; (void)n_clouds; (void)n_points;
  return 0;
}

int smb_vn_dgcnn_encode(const smb_encoder_weights* w, const float* clouds, int32_t n_clouds, int32_t n_points, float* latent,
                        void* workspace, size_t workspace_bytes, void* stream) {
  (void)w; (void)clouds; (void)n_clouds; (void)n_points; (void)latent; (void)workspace; (void)workspace_bytes; (void)stream;
  smb::set_error_msg("smb_vn_dgcnn_encode: not built yet");
  return SMB_E_UNSUPPORTED;
}

}  // extern "C"
