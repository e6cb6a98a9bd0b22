// placeholder — replaced by the real engine
#include "common.cuh"
extern "C" {
int afb_engine_create(const afb_model_desc*, afb_engine**) { afb::set_last_error("engine not built"); return AFB_ERR_UNSUPPORTED; }
void afb_engine_destroy(afb_engine*) {}
int afb_engine_bind(afb_engine*, const afb_weights*) { return AFB_ERR_UNSUPPORTED; }
int afb_engine_set_lora_scale(afb_engine*, float) { return AFB_ERR_UNSUPPORTED; }
size_t afb_engine_workspace_bytes(const afb_engine*, int32_t, int32_t, int32_t) { return 0; }
int afb_engine_reserve(afb_engine*, int32_t, int32_t, int32_t) { return AFB_ERR_UNSUPPORTED; }
int afb_engine_forward(afb_engine*, const afb_forward_args*, void*) { return AFB_ERR_UNSUPPORTED; }
int afb_engine_denoise(afb_engine*, const afb_denoise_args*, void*) { return AFB_ERR_UNSUPPORTED; }
}
