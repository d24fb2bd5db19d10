// placeholder until the tcgen05 engine lands (next commit)
#include "common.cuh"
using namespace ynet;
extern "C" {
int ynet_tc_supported(void) { return 0; }
int ynet_tc_pack_f32_to_c8(const float*, int32_t, int32_t, int32_t, int32_t, int64_t, void*, int32_t, void*) {
  set_error("tensor-core engine not built"); return YNET_E_UNSUPPORTED; }
int ynet_tc_unpack_c8_to_f32(const void*, int32_t, int32_t, int32_t, int32_t, int32_t, float*, void*) {
  set_error("tensor-core engine not built"); return YNET_E_UNSUPPORTED; }
int64_t ynet_tc_packed_weight_bytes(int32_t, int32_t, const int32_t*) { return 0; }
int ynet_tc_pack_weights(const float*, int32_t, int32_t, const int32_t*, const int32_t*, void*, void*) {
  set_error("tensor-core engine not built"); return YNET_E_UNSUPPORTED; }
int ynet_tc_conv3x3(const ynet_tc_src*, int32_t, int32_t, int32_t, int32_t, const void*, const float*, int32_t, int32_t,
                    void*, int32_t, void*) {
  set_error("tensor-core engine not built"); return YNET_E_UNSUPPORTED; }
}
