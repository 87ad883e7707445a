// K5: weight-streaming small-M (KV-cached decode) FFN kernels.  Under construction: until the kernel lands the
// entry points report "outside envelope" so the C-ABI uses the tiled tcgen05 kernel for every token count.
#include "l32_internal.cuh"

namespace l32 {

int ffn_decode_swiglu(const void*, const void*, const void*, void*, int, int, int, int, cudaStream_t) {
    return L32_ERR_BAD_SHAPE;
}
int ffn_decode_linear(const void*, const void*, void*, int, int, int, int, cudaStream_t) { return L32_ERR_BAD_SHAPE; }

}  // namespace l32
