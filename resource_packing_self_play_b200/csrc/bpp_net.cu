// bpp_net.cu — batched policy/value forward (placeholder translation unit; the real kernels land next).
#include <stdint.h>
#include "../../include/bpp_b200.h"
extern "C" int bpp_net_create(int, int, int, int, int, bpp_net** out) { if (out) *out = nullptr; return BPP_E_STATE; }
extern "C" int bpp_net_destroy(bpp_net*) { return BPP_OK; }
extern "C" int bpp_net_set_param(bpp_net*, const char*, const float*, int64_t) { return BPP_E_STATE; }
extern "C" int bpp_net_commit(bpp_net*, void*) { return BPP_E_STATE; }
extern "C" int bpp_net_forward(bpp_net*, int, const int32_t*, const uint32_t*, const int32_t*, const int32_t*, float*,
                               float*, void*) { return BPP_E_STATE; }
