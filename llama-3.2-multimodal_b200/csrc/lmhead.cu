// lm_head + shifted cross entropy (reference Model/model.py:429-438: logits = lm_head(hidden),
// CrossEntropyLoss(ignore_index)(shift_logits, shift_labels)).  The GEMM (gemm_sm100.cu, EPI_CE) stores the logits and leaves
// per-row, per-column-tile (max, sum exp) pairs plus the target logit; the kernels here finish the job:
//   ce_rows_kernel      one warp per row: combine the tile statistics -> lse[row], loss[row] = lse - target (0 when ignored)
//   ce_mean_kernel      one CTA: loss = sum(loss[row]) / #valid rows   (fixed tree order: deterministic)
//   ce_dlogits_kernel   dlogits = (softmax(logits) - onehot(label)) * grad_loss / #valid, from the STORED logits and lse
//                       (HBM-bound: one read + one write of [rows, vocab] 16-bit values; may run in place)
#include "l32_internal.cuh"

namespace l32 {
namespace {

__global__ void __launch_bounds__(256) ce_rows_kernel(const float2* __restrict__ partials, const float* __restrict__ target,
                                                      const long long* __restrict__ labels, long long ignore_index,
                                                      int64_t rows, int tiles_n, int vocab, float* __restrict__ lse,
                                                      float* __restrict__ loss_rows) {
    pdl_wait_prior_grid();
    const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float2* pr = partials + static_cast<size_t>(row) * tiles_n;
    float m = -INFINITY, ssum = 0.f;
    for (int t = lane; t < tiles_n; t += 32) {
        const float2 v = pr[t];
        const float nm = fmaxf(m, v.x);
        ssum = ssum * __expf(m - nm) + v.y * __expf(v.x - nm);
        m = nm;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, ssum, o);
        const float nm = fmaxf(m, om);
        // lanes without any tile carry (-inf, 0): exp(-inf - nm) = 0 unless nm is -inf too (then both sums are 0)
        ssum = (nm == -INFINITY) ? 0.f : ssum * __expf(m - nm) + os * __expf(om - nm);
        m = nm;
    }
    if (lane == 0) {
        const float l = m + __logf(ssum);
        lse[row] = l;
        const long long lab = labels[row];
        const bool valid = lab != ignore_index && lab >= 0 && lab < vocab;
        loss_rows[row] = valid ? (l - target[row]) : 0.f;
    }
}

__global__ void __launch_bounds__(1024) ce_mean_kernel(const float* __restrict__ loss_rows, const long long* __restrict__ labels,
                                                       long long ignore_index, int64_t rows, int vocab,
                                                       float* __restrict__ loss_and_count) {
    __shared__ float ssum[32];
    __shared__ float scnt[32];
    pdl_wait_prior_grid();
    float a = 0.f, c = 0.f;
    for (int64_t r = threadIdx.x; r < rows; r += 1024) {
        const long long lab = labels[r];
        if (lab != ignore_index && lab >= 0 && lab < vocab) {
            a += loss_rows[r];
            c += 1.f;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = a; scnt[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x < 32) {
        a = ssum[threadIdx.x];
        c = scnt[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (threadIdx.x == 0) {
            loss_and_count[0] = c > 0.f ? a / c : __int_as_float(0x7fc00000);   // mean over valid rows (torch: nan when none)
            loss_and_count[1] = c;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) ce_dlogits_kernel(const T* logits, const float* __restrict__ lse,
                                                         const long long* __restrict__ labels, long long ignore_index,
                                                         const float* __restrict__ loss_and_count, const float* __restrict__ grad_loss, T* dlogits,
                                                         int64_t rows, int vocab) {
    pdl_wait_prior_grid();
    const int nvec = vocab >> 3;
    const float cnt = loss_and_count[1];
    const float gl = grad_loss != nullptr ? grad_loss[0] : 1.f;   // upstream gradient of the scalar loss (device scalar)
    const float gscale = cnt > 0.f ? gl / cnt : 0.f;
    for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
        const long long lab = labels[row];
        const bool valid = lab != ignore_index && lab >= 0 && lab < vocab;
        const float g = valid ? gscale : 0.f;
        const float l = lse[row];
        const T* src = logits + static_cast<size_t>(row) * vocab;
        T* dst = dlogits + static_cast<size_t>(row) * vocab;
        for (int v = blockIdx.x * 256 + threadIdx.x; v < nvec; v += gridDim.x * 256) {
            uint4 q = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(v) * 8);
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = Pack2<T>::unpack(w[j]);
                const int col = v * 8 + 2 * j;
                f.x = (__expf(f.x - l) - (col == lab ? 1.f : 0.f)) * g;
                f.y = (__expf(f.y - l) - (col + 1 == lab ? 1.f : 0.f)) * g;
                w[j] = Pack2<T>::pack(f.x, f.y);
            }
            *reinterpret_cast<uint4*>(dst + static_cast<size_t>(v) * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

cudaLaunchConfig_t pdl_cfg(dim3 grid, dim3 block, cudaStream_t s, cudaLaunchAttribute* attr) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.stream = s;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cfg;
}

}  // namespace

cudaError_t ce_reduce(const void* partials, const float* target, const long long* labels, long long ignore_index, int64_t rows,
                      int tiles_n, int vocab, float* lse, float* loss_rows, float* loss_and_count, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = pdl_cfg(dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), s, attr);
    cudaError_t e = cudaLaunchKernelEx(&cfg, ce_rows_kernel, static_cast<const float2*>(partials), target, labels, ignore_index, rows,
                                       tiles_n, vocab, lse, loss_rows);
    if (e != cudaSuccess) return e;
    count_launch();
    cfg = pdl_cfg(dim3(1), dim3(1024), s, attr);
    const float* lr = loss_rows;
    e = cudaLaunchKernelEx(&cfg, ce_mean_kernel, lr, labels, ignore_index, rows, vocab, loss_and_count);
    if (e == cudaSuccess) count_launch();
    return e;
}

cudaError_t ce_backward_logits(const void* logits, const float* lse, const long long* labels, long long ignore_index,
                               const float* loss_and_count, const float* grad_loss, void* dlogits, int64_t rows, int vocab,
                               int dtype,
                               cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    int gx = (vocab / 8 + 255) / 256;
    if (gx > 64) gx = 64;
    const int64_t cap_rows = static_cast<int64_t>(num_sms()) * 16 / gx + 1;
    const unsigned gy = static_cast<unsigned>(rows < cap_rows ? rows : cap_rows);
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = pdl_cfg(dim3(static_cast<unsigned>(gx), gy), dim3(256), s, attr);
    cudaError_t e;
    if (dtype == L32_BF16)
        e = cudaLaunchKernelEx(&cfg, ce_dlogits_kernel<__nv_bfloat16>, static_cast<const __nv_bfloat16*>(logits), lse, labels,
                               ignore_index, loss_and_count, grad_loss, static_cast<__nv_bfloat16*>(dlogits), rows, vocab);
    else
        e = cudaLaunchKernelEx(&cfg, ce_dlogits_kernel<__half>, static_cast<const __half*>(logits), lse, labels, ignore_index,
                               loss_and_count, grad_loss, static_cast<__half*>(dlogits), rows, vocab);
    if (e == cudaSuccess) count_launch();
    return e;
}

}  // namespace l32
