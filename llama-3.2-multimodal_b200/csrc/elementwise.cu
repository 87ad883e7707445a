// Elementwise SwiGLU pieces that cannot ride in a GEMM epilogue:
//  * swiglu_bwd_elementwise: d_gate, d_up from an externally supplied d_act (the autograd boundary of
//    SwiGLUFunction sits between the activation and the down projection, reference
//    Tools/swiglu/FusedSwiglu.py:32-40; math of the never-launched reference kernel swiglu.cu:204-210).
//  * swiglu_act_elementwise: act = silu(gate) * up (unfused reference point for the fusion-saving benchmark).
// HBM-bound: 128-bit coalesced accesses, grid-stride, fp32 math.
#include "l32_internal.cuh"

namespace l32 {
namespace {

L32_DEVICE uint4 ldg_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
L32_DEVICE void stg_v4(void* p, const uint4& v) {
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const T* __restrict__ d_act, const T* __restrict__ gate,
                                                         const T* __restrict__ up, T* __restrict__ d_gate,
                                                         T* __restrict__ d_up, int64_t nvec) {
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < nvec; i += gridDim.x * 256ll) {
        const uint4 a = ldg_v4(d_act + i * 8), g = ldg_v4(gate + i * 8), u = ldg_v4(up + i * 8);
        const uint32_t av[4] = {a.x, a.y, a.z, a.w}, gv[4] = {g.x, g.y, g.z, g.w}, uv[4] = {u.x, u.y, u.z, u.w};
        uint32_t og[4], ou[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 da = Pack2<T>::unpack(av[j]), gg = Pack2<T>::unpack(gv[j]), uu = Pack2<T>::unpack(uv[j]);
            const float s0 = sigmoid_f32(gg.x), s1 = sigmoid_f32(gg.y);
            og[j] = Pack2<T>::pack(da.x * uu.x * (s0 * (1.0f + gg.x * (1.0f - s0))),
                                   da.y * uu.y * (s1 * (1.0f + gg.y * (1.0f - s1))));
            ou[j] = Pack2<T>::pack(da.x * gg.x * s0, da.y * gg.y * s1);
        }
        stg_v4(d_gate + i * 8, make_uint4(og[0], og[1], og[2], og[3]));
        stg_v4(d_up + i * 8, make_uint4(ou[0], ou[1], ou[2], ou[3]));
    }
}

template <typename T>
__global__ void __launch_bounds__(256) swiglu_act_kernel(const T* __restrict__ gate, const T* __restrict__ up,
                                                         T* __restrict__ act, int64_t nvec) {
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < nvec; i += gridDim.x * 256ll) {
        const uint4 g = ldg_v4(gate + i * 8), u = ldg_v4(up + i * 8);
        const uint32_t gv[4] = {g.x, g.y, g.z, g.w}, uv[4] = {u.x, u.y, u.z, u.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 gg = Pack2<T>::unpack(gv[j]), uu = Pack2<T>::unpack(uv[j]);
            o[j] = Pack2<T>::pack(silu_f32(gg.x) * uu.x, silu_f32(gg.y) * uu.y);
        }
        stg_v4(act + i * 8, make_uint4(o[0], o[1], o[2], o[3]));
    }
}

int ew_grid(int64_t nvec) {
    const int64_t want = (nvec + 255) / 256;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
    return static_cast<int>(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace

cudaError_t swiglu_bwd_elementwise(const void* d_act, const void* gate, const void* up, void* d_gate, void* d_up,
                                   int64_t n, int dtype, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const int64_t nvec = n / 8;
    if (dtype == L32_BF16)
        swiglu_bwd_kernel<__nv_bfloat16><<<ew_grid(nvec), 256, 0, s>>>(
            (const __nv_bfloat16*)d_act, (const __nv_bfloat16*)gate, (const __nv_bfloat16*)up, (__nv_bfloat16*)d_gate,
            (__nv_bfloat16*)d_up, nvec);
    else
        swiglu_bwd_kernel<__half><<<ew_grid(nvec), 256, 0, s>>>((const __half*)d_act, (const __half*)gate, (const __half*)up,
                                                                (__half*)d_gate, (__half*)d_up, nvec);
    count_launch();
    return cudaGetLastError();
}

cudaError_t swiglu_act_elementwise(const void* gate, const void* up, void* act, int64_t n, int dtype, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const int64_t nvec = n / 8;
    if (dtype == L32_BF16)
        swiglu_act_kernel<__nv_bfloat16><<<ew_grid(nvec), 256, 0, s>>>((const __nv_bfloat16*)gate, (const __nv_bfloat16*)up,
                                                                       (__nv_bfloat16*)act, nvec);
    else
        swiglu_act_kernel<__half><<<ew_grid(nvec), 256, 0, s>>>((const __half*)gate, (const __half*)up, (__half*)act, nvec);
    count_launch();
    return cudaGetLastError();
}

}  // namespace l32
