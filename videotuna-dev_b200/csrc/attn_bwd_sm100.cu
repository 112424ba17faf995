// placeholder until the backward kernel lands
#include "attn_common.h"
namespace vt {
cudaError_t attn_bwd_set_debug_ptr(unsigned int*) { return cudaSuccess; }
}
