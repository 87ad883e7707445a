// Tensor-parallel glue kernels (one process per GPU; all buffers named "peer" are symmetric allocations mapped into
// this process, reached over NVLink 5 / NVSwitch with plain loads and stores).
//
// The reference has no distributed code; these kernels serve the sharding BASELINE.json's
// north_star asks for -- gate/up column-parallel, down row-parallel -- together with the two fused GEMM variants in
// gemm_sm100.cu (all-gather pulled into the A operand, reduce-scatter pushed from the epilogue):
//   tp_signal_kernel          flag[index] := value on every rank (release at system scope)
//   tp_reduce_partials_kernel y = sum over ranks of the partial slots this rank received (+ optional addend),
//                             after waiting until every peer has raised its "partials written" flag
#include "l32_internal.cuh"

namespace l32 {
namespace {

struct PeerFlags {
    uint32_t* ptr[kMaxTpWorld];
};

__global__ void tp_signal_kernel(PeerFlags flags, int world, int index, uint32_t value, uint32_t* zero8) {
    const int d = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait_prior_grid();   // the kernel whose completion is being announced
    if (zero8 != nullptr && d < kMaxTpWorld) zero8[d] = 0u;   // arrival counters of the next fused all-gather
    if (d < world) {
        __threadfence_system();   // everything this GPU wrote before (earlier kernels of the stream) is visible first
        st_release_sys_u32(flags.ptr[d] + index, value);
    }
}

L32_DEVICE uint4 ldg_stream_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

// slots: [world][rows_per_rank][hidden] (slot s = the partial rank s pushed for this rank's rows); y, addend: [rows][hidden].
// HBM-bound: world reads + 1 write of rows*hidden 16-bit elements; fp32 accumulation in rank order (deterministic).
template <typename T>
__global__ void __launch_bounds__(256) tp_reduce_partials_kernel(const T* slots, const uint32_t* flags, uint32_t epoch,
                                                                 int world, int rank, const T* addend, T* __restrict__ y,
                                                                 int64_t rows, int64_t slot_rows, int hidden, unsigned long long timeout_ns) {
    pdl_wait_prior_grid();   // this rank's own slot was written by the down GEMM launched before
    if (threadIdx.x < world && static_cast<int>(threadIdx.x) != rank) wait_flag_ge<true>(&flags[threadIdx.x], epoch, timeout_ns);
    __syncthreads();
    const int64_t nvec = rows * hidden / 8;
    const int64_t slot_vec = slot_rows * hidden / 8;
    for (int64_t v = blockIdx.x * 256ll + threadIdx.x; v < nvec; v += gridDim.x * 256ll) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        if (addend != nullptr) {
            const uint4 a = ldg_stream_v4(addend + v * 8);
            const uint32_t av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = Pack2<T>::unpack(av[j]);
                acc[2 * j] = f.x;
                acc[2 * j + 1] = f.y;
            }
        }
        for (int s = 0; s < world; ++s) {
            const uint4 a = ldg_stream_v4(slots + (s * slot_vec + v) * 8);
            const uint32_t av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = Pack2<T>::unpack(av[j]);
                acc[2 * j] += f.x;
                acc[2 * j + 1] += f.y;
            }
        }
        uint4 o;
        o.x = Pack2<T>::pack(acc[0], acc[1]);
        o.y = Pack2<T>::pack(acc[2], acc[3]);
        o.z = Pack2<T>::pack(acc[4], acc[5]);
        o.w = Pack2<T>::pack(acc[6], acc[7]);
        *reinterpret_cast<uint4*>(y + v * 8) = o;
    }
}

// Plain peer copy (pull when src is peer memory, push when dst is): the NVLink bandwidth reference the fused kernels are
// compared with.  `warps` warps per CTA, 16-byte vectors, kUnroll loads in flight per lane.
// seg_bytes < 8192 emulates a GEMM epilogue: the buffer is viewed as rows of 8192 bytes and consecutive work items cover
// one `seg_bytes` segment of a row, then the same column block of the NEXT row (stride 8 KiB), ... -- i.e. contiguous
// runs of only seg_bytes on the wire.  seg_bytes = 0: fully linear copy.
template <int kUnroll>
__global__ void __launch_bounds__(1024) tp_peer_copy_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src,
                                                            long long nvec, int seg_bytes) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long vps = seg_bytes > 0 ? seg_bytes / 16 : 1;
    const long long nrows = nvec * 16 / 8192;
    auto offset = [&](long long v) -> long long {
        if (seg_bytes <= 0) return v * 16;
        const long long seg = v / vps, in = v % vps;
        return (seg % nrows) * 8192 + (seg / nrows) * seg_bytes + in * 16;
    };
    for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < nvec; v += stride * kUnroll) {
        uint4 buf[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (v + u * stride < nvec) buf[u] = ld_relaxed_sys_v4(src + offset(v + u * stride));
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (v + u * stride < nvec) *reinterpret_cast<uint4*>(dst + offset(v + u * stride)) = buf[u];
    }
}

}  // namespace

cudaError_t tp_peer_copy(void* dst, const void* src, size_t bytes, int ctas, int warps, int unroll, int seg_bytes,
                         cudaStream_t s) {
    const long long nvec = static_cast<long long>(bytes / 16);
    if (nvec == 0) return cudaSuccess;
    auto* d = static_cast<uint8_t*>(dst);
    auto* sp = static_cast<const uint8_t*>(src);
    if (unroll >= 16) tp_peer_copy_kernel<16><<<ctas, warps * 32, 0, s>>>(d, sp, nvec, seg_bytes);
    else if (unroll >= 8) tp_peer_copy_kernel<8><<<ctas, warps * 32, 0, s>>>(d, sp, nvec, seg_bytes);
    else tp_peer_copy_kernel<4><<<ctas, warps * 32, 0, s>>>(d, sp, nvec, seg_bytes);
    count_launch();
    return cudaGetLastError();
}

static cudaLaunchConfig_t pdl_config(unsigned grid, unsigned block, cudaStream_t s, cudaLaunchAttribute* attr) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.stream = s;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cfg;
}

cudaError_t tp_signal(void* const* peer_flags, int world, int index, uint32_t value, uint32_t* zero8, cudaStream_t s) {
    PeerFlags f;
    for (int i = 0; i < kMaxTpWorld; ++i) f.ptr[i] = i < world ? static_cast<uint32_t*>(peer_flags[i]) : nullptr;
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = pdl_config(1, 32, s, attr);
    cudaError_t e = cudaLaunchKernelEx(&cfg, tp_signal_kernel, f, world, index, value, zero8);
    if (e == cudaSuccess) count_launch();
    return e;
}

cudaError_t tp_reduce_partials(const void* slots, const uint32_t* flags, uint32_t epoch, int world, int rank,
                               const void* addend, void* y, int64_t rows, int64_t slot_rows, int hidden, int dtype,
                               cudaStream_t s) {
    if (rows == 0) return cudaSuccess;
    const int64_t nvec = rows * hidden / 8;
    int64_t blocks = (nvec + 255) / 256;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = pdl_config(static_cast<unsigned>(blocks), 256, s, attr);
    cudaError_t e;
    if (dtype == L32_BF16)
        e = cudaLaunchKernelEx(&cfg, tp_reduce_partials_kernel<__nv_bfloat16>, static_cast<const __nv_bfloat16*>(slots), flags,
                               epoch, world, rank, static_cast<const __nv_bfloat16*>(addend), static_cast<__nv_bfloat16*>(y),
                               rows, slot_rows, hidden, static_cast<unsigned long long>(spin_timeout_ns()));
    else
        e = cudaLaunchKernelEx(&cfg, tp_reduce_partials_kernel<__half>, static_cast<const __half*>(slots), flags, epoch, world,
                               rank, static_cast<const __half*>(addend), static_cast<__half*>(y), rows, slot_rows, hidden,
                               static_cast<unsigned long long>(spin_timeout_ns()));
    if (e == cudaSuccess) count_launch();
    return e;
}

}  // namespace l32
