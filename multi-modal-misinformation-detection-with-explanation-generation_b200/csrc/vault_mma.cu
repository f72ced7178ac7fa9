// placeholder until the tcgen05 path lands
#include "common.cuh"
int mmf_mma_vault_changed(mmf_handle*) { return MMF_OK; }
int mmf_mma_supported(const mmf_handle*, int64_t, int) { return 0; }
void mmf_mma_destroy(mmf_handle*) {}
int mmf_mma_search(mmf_handle* h, const float*, int64_t, int, double, float*, int64_t*, uint64_t*, float*, cudaStream_t) {
  return mmf_set_error(h, MMF_ERR_UNSUPPORTED, "tcgen05 path not built");
}
