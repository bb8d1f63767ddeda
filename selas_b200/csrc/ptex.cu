// ptex.cu -- stub (replaced by the real loader)
#include "internal.h"
namespace selas { const void* ptexDevicePointer(SelasPtex*) { return nullptr; } }
using namespace selas;
extern "C" {
int selas_cuda_ptex_load(SelasContext* ctx, const char*, SelasPtex** out) { if (out) *out = nullptr; setError(ctx, "ptex: not built"); return SELAS_ERR_UNSUPPORTED; }
int selas_cuda_ptex_load_memory(SelasContext* ctx, const void*, size_t, SelasPtex** out) { if (out) *out = nullptr; setError(ctx, "ptex: not built"); return SELAS_ERR_UNSUPPORTED; }
int selas_cuda_ptex_destroy(SelasPtex*) { return SELAS_OK; }
int selas_cuda_ptex_eval(SelasPtex*, size_t, const int32_t*, const float*, const float*, float*) { return SELAS_ERR_UNSUPPORTED; }
int selas_cuda_ptex_eval_device(SelasPtex*, size_t, const int32_t*, const float*, const float*, float*, void*) { return SELAS_ERR_UNSUPPORTED; }
}
