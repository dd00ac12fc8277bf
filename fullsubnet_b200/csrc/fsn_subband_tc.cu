// tcgen05 sub-band LSTM stack (placeholder until the kernel lands).
#include "fsn_internal.cuh"
namespace fsn {
bool sb_tc_supported(const fsn_model_desc*) { return false; }
size_t sb_tc_packed_bytes(const fsn_model_desc*) { return 0; }
int sb_tc_pack(const fsn_model_desc*, const fsn_seq_weights*, void*, cudaStream_t) {
  set_error("FSN_PREC_F16_TC not built");
  return FSN_ERR_UNSUPPORTED;
}
int sb_tc_forward(const SbTcArgs&, cudaStream_t) {
  set_error("FSN_PREC_F16_TC not built");
  return FSN_ERR_UNSUPPORTED;
}
}  // namespace fsn
