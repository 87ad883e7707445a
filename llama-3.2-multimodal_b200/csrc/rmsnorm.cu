// K1/K2: fused residual Add-RMSNorm forward and RMSNorm backward for sm_100a.
//
// Replaces the reference kernels rmsnorm_kernel_fused / rmsnorm_backward_kernel
// (reference Tools/rmsnorm/rmsnorm.cuh:13-108 and :110-154) behind the host entry points
// rmsnorm_forward / rmsnorm_backward (reference Tools/rmsnorm/rmsnorm.cu:7-61).
//
// Design (HBM-bound; algorithmic traffic 3*C*2 B per row fwd with residual, 3*C*2 B per row bwd):
//  * forward : one CTA per row, the whole row lives in registers (single read of x and residual),
//              128-bit coalesced loads/stores, fp32 add + sum of squares, warp-shuffle reduction plus
//              one shared-memory hop, rsqrtf; optional h = x + residual output (training) and
//              optional in-place residual update (raw reference ABI).
//  * backward: 512-thread CTAs, each row handled by a group of threads, grid-stride over rows,
//              d_weight accumulated in registers per thread and written once per CTA as an fp32
//              partial row; a second tiny kernel column-reduces the partials (no per-element atomics,
//              deterministic).
#include "l32_internal.cuh"

namespace l32 {

template <typename T>
L32_DEVICE void unpack8(const uint4& v, float (&f)[8]) {
    float2 a = Pack2<T>::unpack(v.x), b = Pack2<T>::unpack(v.y), c = Pack2<T>::unpack(v.z), d = Pack2<T>::unpack(v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
template <typename T>
L32_DEVICE uint4 pack8(const float (&f)[8]) {
    uint4 v;
    v.x = Pack2<T>::pack(f[0], f[1]); v.y = Pack2<T>::pack(f[2], f[3]);
    v.z = Pack2<T>::pack(f[4], f[5]); v.w = Pack2<T>::pack(f[6], f[7]);
    return v;
}
L32_DEVICE uint4 ld_stream_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
L32_DEVICE uint4 ld_v4(const void* p) {   // plain (coherent) load: used when the buffer is also written
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
L32_DEVICE void st_v4(void* p, const uint4& v) {
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
L32_DEVICE float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over a group of RT consecutive threads (RT multiple of 32). `slot` = per-group smem scratch
// of RT/32 floats. All threads of the group get the result.
template <int RT>
L32_DEVICE float group_sum(float v, float* slot, int tid_in_group, int bar_id) {
    v = warp_sum(v);
    if constexpr (RT == 32) {
        return v;
    } else {
        constexpr int NW = RT / 32;
        if ((tid_in_group & 31) == 0) slot[tid_in_group >> 5] = v;
        if constexpr (RT == 512) __syncthreads(); else named_bar_sync(bar_id, RT);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NW; ++i) s += slot[i];
        return s;
    }
}

// ------------------------------------------------------------------------------------------------
// Forward, fast path: C % 8 == 0, C <= RT * VPT * 8. One CTA (RT threads) per row.
// ------------------------------------------------------------------------------------------------
template <typename T, int RT, int VPT, bool kHasResidual, bool kWriteH>
__global__ void __launch_bounds__(RT) add_rmsnorm_fwd_kernel(
    const T* __restrict__ x, const T* residual, const T* __restrict__ weight, T* __restrict__ y,
    T* h_out, float* __restrict__ rms_out, int64_t rows, int C, float eps) {
    __shared__ float red[16];
    const int tid = threadIdx.x;
    const int nvec = C >> 3;
    const int64_t row = blockIdx.x;
    const size_t base = static_cast<size_t>(row) * C;

    uint4 xv[VPT], rv[VPT], wv[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = tid + i * RT;
        if (v < nvec) xv[i] = ld_stream_v4(x + base + (size_t)v * 8);
    }
    if constexpr (kHasResidual) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int v = tid + i * RT;
            // h_out may alias residual (raw reference ABI updates residual in place): coherent load.
            if (v < nvec) rv[i] = ld_v4(residual + base + (size_t)v * 8);
        }
    }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = tid + i * RT;
        if (v < nvec) wv[i] = __ldg(reinterpret_cast<const uint4*>(weight) + v);
    }

    float h[VPT][8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = tid + i * RT;
        if (v < nvec) {
            unpack8<T>(xv[i], h[i]);
            if constexpr (kHasResidual) {
                float r[8];
                unpack8<T>(rv[i], r);
#pragma unroll
                for (int j = 0; j < 8; ++j) h[i][j] += r[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) ss = fmaf(h[i][j], h[i][j], ss);
        }
    }
    if constexpr (kWriteH) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int v = tid + i * RT;
            if (v < nvec) st_v4(h_out + base + (size_t)v * 8, pack8<T>(h[i]));
        }
    }
    ss = group_sum<RT>(ss, red, tid, 1);
    const float var = ss / static_cast<float>(C) + eps;
    const float inv = rsqrtf(var);
    if (rms_out != nullptr && tid == 0) rms_out[row] = sqrtf(var);

#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = tid + i * RT;
        if (v < nvec) {
            float w[8], o[8];
            unpack8<T>(wv[i], w);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = h[i][j] * inv * w[j];
            st_v4(y + base + (size_t)v * 8, pack8<T>(o));
        }
    }
}

// Forward, generic path: any C, scalar accesses, two passes (second pass re-reads, L2-resident).
template <typename T>
__global__ void __launch_bounds__(256) add_rmsnorm_fwd_generic_kernel(
    const T* x, const T* residual, const T* weight, T* y, T* h_out, float* rms_out, int64_t rows, int C, float eps) {
    __shared__ float red[8];
    const int64_t row = blockIdx.x;
    const size_t base = static_cast<size_t>(row) * C;
    float ss = 0.f;
    for (int i = threadIdx.x; i < C; i += 256) {
        float h = static_cast<float>(x[base + i]);
        if (residual != nullptr) h += static_cast<float>(residual[base + i]);
        ss = fmaf(h, h, ss);
    }
    ss = group_sum<256>(ss, red, threadIdx.x, 1);
    const float var = ss / static_cast<float>(C) + eps;
    const float inv = rsqrtf(var);
    if (rms_out != nullptr && threadIdx.x == 0) rms_out[row] = sqrtf(var);
    for (int i = threadIdx.x; i < C; i += 256) {
        float h = static_cast<float>(x[base + i]);
        if (residual != nullptr) h += static_cast<float>(residual[base + i]);
        if (h_out != nullptr) h_out[base + i] = static_cast<T>(h);   // each element touched by one thread only
        y[base + i] = static_cast<T>(h * inv * static_cast<float>(weight[i]));
    }
}

// ------------------------------------------------------------------------------------------------
// Backward, fast path: C % 8 == 0, C <= RT * VPT * 8. CTA = 512 threads = (512/RT) row groups.
//   rstd = 1/rms; xhat = h*rstd; wdy = dy*w; c1 = mean(xhat*wdy)
//   dx = (wdy - xhat*c1) * rstd;   dw[c] = sum_rows dy*xhat
// (same algebra as reference rmsnorm.cuh:124-152 with inp := h, minus its extra 1e-6 and atomics)
// ------------------------------------------------------------------------------------------------
template <typename T, int RT, int VPT>
__global__ void __launch_bounds__(512) rmsnorm_bwd_kernel(
    const T* __restrict__ dy, const T* __restrict__ h, const T* __restrict__ weight, const float* __restrict__ rms,
    T* __restrict__ dx, float* __restrict__ dw_partial, int64_t rows, int C) {
    constexpr int G = 512 / RT;
    __shared__ float red[G][16];
    extern __shared__ float dw_smem[];   // [G-1][C] when G > 1
    const int tid = threadIdx.x;
    const int g = tid / RT;
    const int t = tid % RT;
    const int nvec = C >> 3;

    float w[VPT][8], dwacc[VPT][8];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = t + i * RT;
        if (v < nvec) {
            uint4 wv = __ldg(reinterpret_cast<const uint4*>(weight) + v);
            unpack8<T>(wv, w[i]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) dwacc[i][j] = 0.f;
    }

    const float invC = 1.0f / static_cast<float>(C);
    // Uniform trip count per CTA so the group barriers inside group_sum stay convergent.
    const int64_t rows_per_iter = static_cast<int64_t>(gridDim.x) * G;
    const int64_t iters = (rows + rows_per_iter - 1) / rows_per_iter;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t row = it * rows_per_iter + static_cast<int64_t>(blockIdx.x) * G + g;
        const bool active = row < rows;
        const size_t base = static_cast<size_t>(active ? row : 0) * C;
        uint4 gv[VPT], hv[VPT];
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int v = t + i * RT;
            if (active && v < nvec) {
                gv[i] = ld_stream_v4(dy + base + (size_t)v * 8);
                hv[i] = ld_stream_v4(h + base + (size_t)v * 8);
            }
        }
        const float rstd = active ? 1.0f / rms[row] : 0.f;
        float wdy[VPT][8], xh[VPT][8];
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int v = t + i * RT;
            if (active && v < nvec) {
                float gg[8], hh[8];
                unpack8<T>(gv[i], gg);
                unpack8<T>(hv[i], hh);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    xh[i][j] = hh[j] * rstd;
                    wdy[i][j] = gg[j] * w[i][j];
                    dot = fmaf(xh[i][j], wdy[i][j], dot);
                    dwacc[i][j] = fmaf(gg[j], xh[i][j], dwacc[i][j]);
                }
            }
        }
        dot = group_sum<RT>(dot, red[g], t, 1 + g);
        const float c1 = dot * invC;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int v = t + i * RT;
            if (active && v < nvec) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = (wdy[i][j] - xh[i][j] * c1) * rstd;
                st_v4(dx + base + (size_t)v * 8, pack8<T>(o));
            }
        }
        if constexpr (RT > 32) {
            // red[g] is rewritten next iteration: make sure every thread of the group has read it.
            if constexpr (RT == 512) __syncthreads(); else named_bar_sync(1 + g, RT);
        }
    }

    // Fold the G row groups of this CTA, then one fp32 partial row per CTA.
    if constexpr (G > 1) {
        if (g > 0) {
#pragma unroll
            for (int i = 0; i < VPT; ++i) {
                const int v = t + i * RT;
                if (v < nvec) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) dw_smem[(size_t)(g - 1) * C + v * 8 + j] = dwacc[i][j];
                }
            }
        }
        __syncthreads();
    }
    if (g == 0) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int v = t + i * RT;
            if (v < nvec) {
                float4 lo, hi;
                float s[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    s[j] = dwacc[i][j];
                    if constexpr (G > 1) {
                        for (int gg = 1; gg < G; ++gg) s[j] += dw_smem[(size_t)(gg - 1) * C + v * 8 + j];
                    }
                }
                lo = make_float4(s[0], s[1], s[2], s[3]);
                hi = make_float4(s[4], s[5], s[6], s[7]);
                float4* dst = reinterpret_cast<float4*>(dw_partial + (size_t)blockIdx.x * C + (size_t)v * 8);
                dst[0] = lo;
                dst[1] = hi;
            }
        }
    }
}

// Backward, generic path: one CTA (256 threads) per grid-stride row, scalar accesses,
// dw partial per CTA accumulated in global memory owned by that CTA (no atomics).
template <typename T>
__global__ void __launch_bounds__(256) rmsnorm_bwd_generic_kernel(
    const T* dy, const T* h, const T* weight, const float* rms, T* dx, float* dw_partial, int64_t rows, int C) {
    __shared__ float red[8];
    float* my_dw = dw_partial + (size_t)blockIdx.x * C;
    for (int i = threadIdx.x; i < C; i += 256) my_dw[i] = 0.f;
    const float invC = 1.0f / static_cast<float>(C);
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const size_t base = static_cast<size_t>(row) * C;
        const float rstd = 1.0f / rms[row];
        float dot = 0.f;
        for (int i = threadIdx.x; i < C; i += 256) {
            const float gg = static_cast<float>(dy[base + i]);
            const float xh = static_cast<float>(h[base + i]) * rstd;
            dot = fmaf(xh, gg * static_cast<float>(weight[i]), dot);
            my_dw[i] += gg * xh;   // index i is owned by exactly one thread of this CTA
        }
        dot = group_sum<256>(dot, red, threadIdx.x, 1);
        const float c1 = dot * invC;
        for (int i = threadIdx.x; i < C; i += 256) {
            const float gg = static_cast<float>(dy[base + i]);
            const float xh = static_cast<float>(h[base + i]) * rstd;
            dx[base + i] = static_cast<T>((gg * static_cast<float>(weight[i]) - xh * c1) * rstd);
        }
        __syncthreads();
    }
}

// dw[c] = sum_p partial[p][c], cast to T. One thread per column; consecutive threads read
// consecutive columns (coalesced).
template <typename T>
__global__ void __launch_bounds__(128) rmsnorm_dw_reduce_kernel(const float* __restrict__ partial, T* __restrict__ dw,
                                                                int nparts, int C) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= C) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int p = 0;
    for (; p + 4 <= nparts; p += 4) {
        s0 += partial[(size_t)(p + 0) * C + c];
        s1 += partial[(size_t)(p + 1) * C + c];
        s2 += partial[(size_t)(p + 2) * C + c];
        s3 += partial[(size_t)(p + 3) * C + c];
    }
    for (; p < nparts; ++p) s0 += partial[(size_t)p * C + c];
    dw[c] = static_cast<T>((s0 + s1) + (s2 + s3));
}

// ------------------------------------------------------------------------------------------------
// Host launchers
// ------------------------------------------------------------------------------------------------
template <typename T, int RT, int VPT>
static cudaError_t launch_fwd_fast(const T* x, const T* residual, const T* weight, T* y, T* h_out, float* rms,
                                   int64_t rows, int C, float eps, cudaStream_t s) {
    dim3 grid(static_cast<unsigned>(rows)), block(RT);
    if (residual != nullptr) {
        if (h_out != nullptr)
            add_rmsnorm_fwd_kernel<T, RT, VPT, true, true><<<grid, block, 0, s>>>(x, residual, weight, y, h_out, rms, rows, C, eps);
        else
            add_rmsnorm_fwd_kernel<T, RT, VPT, true, false><<<grid, block, 0, s>>>(x, residual, weight, y, h_out, rms, rows, C, eps);
    } else {
        if (h_out != nullptr)
            add_rmsnorm_fwd_kernel<T, RT, VPT, false, true><<<grid, block, 0, s>>>(x, residual, weight, y, h_out, rms, rows, C, eps);
        else
            add_rmsnorm_fwd_kernel<T, RT, VPT, false, false><<<grid, block, 0, s>>>(x, residual, weight, y, h_out, rms, rows, C, eps);
    }
    count_launch();
    return cudaGetLastError();
}

template <typename T>
static cudaError_t add_rmsnorm_fwd_t(const T* x, const T* residual, const T* weight, T* y, T* h_out, float* rms,
                                     int64_t rows, int C, float eps, cudaStream_t s) {
    if (rows == 0) return cudaSuccess;
    const bool aligned = (C % 8 == 0) && is_aligned16(x) && is_aligned16(residual) && is_aligned16(weight) &&
                         is_aligned16(y) && is_aligned16(h_out);
    if (aligned && C <= 16384) {
        if (C <= 32 * 4 * 8) return launch_fwd_fast<T, 32, 4>(x, residual, weight, y, h_out, rms, rows, C, eps, s);
        if (C <= 64 * 4 * 8) return launch_fwd_fast<T, 64, 4>(x, residual, weight, y, h_out, rms, rows, C, eps, s);
        if (C <= 128 * 4 * 8) return launch_fwd_fast<T, 128, 4>(x, residual, weight, y, h_out, rms, rows, C, eps, s);
        if (C <= 256 * 4 * 8) return launch_fwd_fast<T, 256, 4>(x, residual, weight, y, h_out, rms, rows, C, eps, s);
        return launch_fwd_fast<T, 512, 4>(x, residual, weight, y, h_out, rms, rows, C, eps, s);
    }
    add_rmsnorm_fwd_generic_kernel<T><<<static_cast<unsigned>(rows), 256, 0, s>>>(x, residual, weight, y, h_out, rms, rows, C, eps);
    count_launch();
    return cudaGetLastError();
}

static int bwd_grid(int64_t rows, int rows_per_cta) {
    const int64_t want = (rows + rows_per_cta - 1) / rows_per_cta;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 2;
    return static_cast<int>(want < cap ? (want < 1 ? 1 : want) : cap);
}

template <typename T, int RT, int VPT>
static cudaError_t launch_bwd_fast(const T* dy, const T* h, const T* weight, const float* rms, T* dx, float* partial,
                                   int64_t rows, int C, int grid, cudaStream_t s) {
    constexpr int G = 512 / RT;
    const size_t smem = (G > 1) ? static_cast<size_t>(G - 1) * C * sizeof(float) : 0;
    auto* k = rmsnorm_bwd_kernel<T, RT, VPT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    k<<<grid, 512, smem, s>>>(dy, h, weight, rms, dx, partial, rows, C);
    count_launch();
    return cudaGetLastError();
}

// workspace layout: [grid][C] fp32, grid <= 2 * num_sms
size_t rmsnorm_bwd_workspace_bytes(int64_t rows, int C) {
    (void)rows;
    return static_cast<size_t>(num_sms()) * 2 * static_cast<size_t>(C) * sizeof(float);
}

template <typename T>
static cudaError_t rmsnorm_bwd_t(const T* dy, const T* h, const T* weight, const float* rms, T* dx, T* dw,
                                 float* workspace, int64_t rows, int C, cudaStream_t s) {
    if (rows == 0) {
        if (dw != nullptr) return cudaMemsetAsync(dw, 0, sizeof(T) * C, s);
        return cudaSuccess;
    }
    const bool aligned = (C % 8 == 0) && is_aligned16(dy) && is_aligned16(h) && is_aligned16(weight) && is_aligned16(dx);
    int grid;
    cudaError_t e;
    if (aligned && C <= 8192) {
        if (C <= 32 * 2 * 8) { grid = bwd_grid(rows, 16); e = launch_bwd_fast<T, 32, 2>(dy, h, weight, rms, dx, workspace, rows, C, grid, s); }
        else if (C <= 64 * 2 * 8) { grid = bwd_grid(rows, 8); e = launch_bwd_fast<T, 64, 2>(dy, h, weight, rms, dx, workspace, rows, C, grid, s); }
        else if (C <= 128 * 2 * 8) { grid = bwd_grid(rows, 4); e = launch_bwd_fast<T, 128, 2>(dy, h, weight, rms, dx, workspace, rows, C, grid, s); }
        else if (C <= 256 * 2 * 8) { grid = bwd_grid(rows, 2); e = launch_bwd_fast<T, 256, 2>(dy, h, weight, rms, dx, workspace, rows, C, grid, s); }
        else { grid = bwd_grid(rows, 1); e = launch_bwd_fast<T, 512, 2>(dy, h, weight, rms, dx, workspace, rows, C, grid, s); }
    } else {
        grid = bwd_grid(rows, 1);
        rmsnorm_bwd_generic_kernel<T><<<grid, 256, 0, s>>>(dy, h, weight, rms, dx, workspace, rows, C);
        count_launch();
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) return e;
    if (dw != nullptr) {
        rmsnorm_dw_reduce_kernel<T><<<(C + 127) / 128, 128, 0, s>>>(workspace, dw, grid, C);
        count_launch();
        e = cudaGetLastError();
    }
    return e;
}

cudaError_t add_rmsnorm_fwd(const void* x, const void* residual, const void* weight, void* y, void* h_out, float* rms,
                            int64_t rows, int C, float eps, int dtype, cudaStream_t s) {
    if (dtype == L32_BF16)
        return add_rmsnorm_fwd_t<__nv_bfloat16>((const __nv_bfloat16*)x, (const __nv_bfloat16*)residual, (const __nv_bfloat16*)weight,
                                                (__nv_bfloat16*)y, (__nv_bfloat16*)h_out, rms, rows, C, eps, s);
    return add_rmsnorm_fwd_t<__half>((const __half*)x, (const __half*)residual, (const __half*)weight, (__half*)y,
                                     (__half*)h_out, rms, rows, C, eps, s);
}

cudaError_t rmsnorm_bwd(const void* dy, const void* h, const void* weight, const float* rms, void* dx, void* dw,
                        float* workspace, int64_t rows, int C, int dtype, cudaStream_t s) {
    if (dtype == L32_BF16)
        return rmsnorm_bwd_t<__nv_bfloat16>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)h, (const __nv_bfloat16*)weight, rms,
                                            (__nv_bfloat16*)dx, (__nv_bfloat16*)dw, workspace, rows, C, s);
    return rmsnorm_bwd_t<__half>((const __half*)dy, (const __half*)h, (const __half*)weight, rms, (__half*)dx, (__half*)dw,
                                 workspace, rows, C, s);
}

}  // namespace l32
