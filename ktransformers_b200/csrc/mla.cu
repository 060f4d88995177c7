// placeholder translation unit: filled in by the MLA decode kernels (see mla.cu history)
#include "common.cuh"
extern "C" {
size_t ktb200_mla_workspace_bytes(int, int, int) { return 0; }
int ktb200_mla_decode(const ktb200_mla_params*, void*) { ktb::set_error("mla_decode: not built"); return KTB200_ESTATE; }
int ktb200_mla_kv_write(void*, int, const void*, const void*, const int*, const int*, int, void*) { ktb::set_error("mla_kv_write: not built"); return KTB200_ESTATE; }
}
