// K5: weight-streaming small-M (KV-cached decode) feed-forward kernels for sm_100a.
//
// For tokens <= 128 the feed-forward is HBM-bound: every step has to stream 3*H*I 16-bit weights (352 MB at the
// 11B shape) while the activations are a few hundred KB.  Replaces, for that regime, the reference's scalar
// kernels swiglu_forward_kernel / swiglu_down_forward_kernel (reference Tools/swiglu/swiglu.cu:58-100, :228-272)
// and the two F.linear + F.silu calls of the live path (reference Tools/swiglu/FusedSwiglu.py:18-20,
// Model/model.py:217).
//
// Design ("swap-AB" on tcgen05): the WEIGHT rows are the UMMA M operand (128 rows per CTA), the tokens are the
// UMMA N operand (padded to a multiple of 16), the accumulator D[weight row, token] lives in TMEM.
//   * one CTA = one block of 128 weight rows x one K split; a cluster of `splits` CTAs shares a row block and
//     reduces its fp32 partials through distributed shared memory (no atomics, no workspace, deterministic);
//   * the host sizes the grid so that ALL CTAs are co-resident (2-4 per SM): equal work per CTA and a fair
//     share of HBM bandwidth means no wave-quantisation tail although 224 or 256 units never divide by 148 SMs;
//   * warp 0 streams weight tiles (TMA, evict-first) through a deep shared-memory ring, warp 2 the activation tiles (TMA,
//     evict-last) through a SEPARATE shallow ring: the weights come from HBM and need many bytes in flight, the token
//     tiles come from L2 and need little lookahead, so they do not take ring depth away from the weights (at 64 tokens a
//     joint ring had 4 stages of 24 KB, the split rings hold 6 x 16 KB of weights + 2 x 8 KB of tokens in the same space);
//     warp 1 issues tcgen05.mma, all four warps run the epilogue;
//   * gate and up rows of the same 64 act columns sit in one 128-row A tile (two 64-row TMA boxes), so
//     SiLU(g)*u is applied on chip and gate / up never exist in HBM;
//   * programmatic dependent launch: the kernel prefetches its first weight stages BEFORE waiting for the
//     previous kernel of the stream (weights never depend on it), so the down projection starts pulling HBM
//     while the gate/up kernel drains, and the gate/up kernel while the Add-RMSNorm drains.
#include "l32_internal.cuh"

#include <cstdlib>
#include <cstring>

namespace l32 {
namespace {

constexpr int kBlockK = 64;                 // one 128-byte swizzle atom of 16-bit elements
constexpr int kUmmaK = 16;
constexpr int kRowsA = 128;                 // UMMA M: weight rows per CTA
constexpr int kMaxThreads = 512;             // 4 x kParts warps; launched with 128 * parts threads (a one-warp-per-scheduler
                                             // epilogue ran at low IPC and is a pure tail: no HBM traffic while it runs)
constexpr int kABytes = kRowsA * kBlockK * 2;   // 16 KiB weight tile per stage
constexpr int kMaxStages = 12;               // weight ring
constexpr int kMaxXStages = 8;               // token-tile ring
constexpr int kBarrierBytes = (2 * kMaxStages + 2 * kMaxXStages + 1) * 8 + 16;

enum : int { DEC_LINEAR = 0, DEC_SWIGLU = 1 };

constexpr int kMaxLinearGroup = 3;
struct DecodeParams {
    CUtensorMap map_w[kMaxLinearGroup];   // weight matrices [rows, K]: [0] = gate (or the only matrix), [1] = up;
                                          // grouped linear (ngroup > 1): one per problem
    // Grouped linear (DEC_LINEAR, ngroup > 1): several y_i = x w_i^T with the SAME x -- e.g. the q / k / v projections of a
    // decode step -- as one launch: row blocks [rb_start[i], rb_start[i + 1]) belong to problem i.  No bias / addend.
    int ngroup;
    int rb_start[kMaxLinearGroup + 1];
    int rows_g[kMaxLinearGroup];
    void* out_g[kMaxLinearGroup];
    CUtensorMap map_x;      // activations [tokens, K]
    void* out;              // [tokens, rows_out]
    const void* bias[2];    // optional per-output-row bias ([0] gate / linear, [1] up)
    const void* addend;     // linear only, optional [tokens, rows_out]: out = a w^T + bias + addend
    void* cache[2];         // SwiGLU only, optional: pre-activation gate / up projections [tokens, rows_out]
    int tokens, n_pad;      // n_pad = UMMA N = tokens rounded up to a multiple of 16 (<= 128)
    int rows_out;           // inter (SwiGLU) or out_features (linear)
    int k;                  // reduction length
    int splits;             // K splits = cluster size
    int stages;             // depth of the weight ring (16 KB per stage)
    int xstages;            // depth of the token-tile ring (n_pad * 128 bytes per stage)
    int dbg_nox;            // experiments only (L32_DECODE_DEBUG_NOX=1, wrong results): token tiles are loaded for the first
                            // `xstages` k-blocks only -- what the step costs without the token tiles' L2 traffic
    int rotate;             // 1: every row block starts its K loop at a different k-block (spreads the L2 reads of
                            // the shared activation tiles over time instead of all CTAs hitting the same lines)
    long long ldo;          // row pitch of out, elements
    uint32_t idesc;
    uint32_t tmem_cols;
};

L32_DEVICE float ld_dsmem_f32(uint32_t local_addr, uint32_t cta_rank) {
    float v;
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %1, %2;\n\t"
        "ld.shared::cluster.f32 %0, [ra];\n\t"
        "}\n"
        : "=f"(v)
        : "r"(local_addr), "r"(cta_rank)
        : "memory");
    return v;
}

template <int kEpi, typename T>
__global__ void __launch_bounds__(kMaxThreads) ffn_decode_kernel(const __grid_constant__ DecodeParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];   // no static shared memory in this kernel: the window starts 1024-aligned
    const int stages = p.stages, xstages = p.xstages;
    const uint32_t b_bytes = static_cast<uint32_t>(p.n_pad) * 128u;
    uint8_t* smem_x = smem + stages * kABytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_x + xstages * b_bytes);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* xfull_bar = empty_bar + kMaxStages;
    uint64_t* xempty_bar = xfull_bar + kMaxXStages;
    uint64_t* tfull_bar = xempty_bar + kMaxXStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
    if ((smem_u32(smem) & 1023u) != 0u) __trap();       // the 128-byte swizzle needs 1024-aligned tiles

    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const uint32_t lane = lane_id();
    const int split = static_cast<int>(cluster_ctarank());
    const int row_block = blockIdx.x / p.splits;

    // grouped linear: which problem this row block belongs to, and its row block inside that problem
    int prob = 0, rb_local = row_block;
    if (kEpi == DEC_LINEAR && p.ngroup > 1) {
        while (prob + 1 < p.ngroup && row_block >= p.rb_start[prob + 1]) ++prob;
        rb_local = row_block - p.rb_start[prob];
    }
    const int rows_out = (kEpi == DEC_LINEAR && p.ngroup > 1) ? p.rows_g[prob] : p.rows_out;
    const long long ldo = (kEpi == DEC_LINEAR && p.ngroup > 1) ? static_cast<long long>(p.rows_g[prob]) : p.ldo;
    // this CTA's slice of the reduction
    const int nkb = (p.k + kBlockK - 1) / kBlockK;
    const int kb_base = nkb / p.splits, kb_rem = nkb % p.splits;
    const int kb0 = split * kb_base + min(split, kb_rem);
    const int cnt = kb_base + (split < kb_rem ? 1 : 0);
    const int rot = (p.rotate && cnt > 0) ? static_cast<int>((static_cast<uint32_t>(row_block) * 40503u) % static_cast<uint32_t>(cnt)) : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.map_w[prob]);
        if constexpr (kEpi == DEC_SWIGLU) tma_prefetch_desc(&p.map_w[1]);
        tma_prefetch_desc(&p.map_x);
        for (int i = 0; i < stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < xstages; ++i) {
            mbar_init(&xfull_bar[i], 1);
            mbar_init(&xempty_bar[i], 1);
        }
        mbar_init(tfull_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc<1>(tmem_slot, p.tmem_cols);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    // Let the next kernel of the stream get scheduled as soon as SM resources free up; it orders itself
    // behind this grid with griddepcontrol.wait.
    pdl_launch_dependents();

    if (warp == 0) {
        if (elect_one()) {   // elect.sync: single-lane branch the compiler can see (no ELECT / BRA.U.ANY loop per TMA / MMA)
            // ---------------------------------------------------------------- TMA producer: weights
            // (running stage / k-block counters: no integer divisions in the per-stage instruction stream)
            const int row0 = rb_local * (kEpi == DEC_SWIGLU ? 64 : kRowsA);
            int s = 0, kr = rot;
            uint32_t phase = 0;
            for (int kb = 0; kb < cnt; ++kb) {                  // weights do not depend on the previous kernel: no PDL wait here
                if (kb >= stages) mbar_wait(&empty_bar[s], phase ^ 1u);
                uint8_t* sa = smem + s * kABytes;
                mbar_arrive_expect_tx(&full_bar[s], kABytes);
                const int kcol = (kb0 + kr) * kBlockK;
                if constexpr (kEpi == DEC_SWIGLU) {
                    tma_load_2d(sa, &p.map_w[0], &full_bar[s], kcol, row0, kEvictFirst);
                    tma_load_2d(sa + kABytes / 2, &p.map_w[1], &full_bar[s], kcol, row0, kEvictFirst);
                } else {
                    tma_load_2d(sa, &p.map_w[prob], &full_bar[s], kcol, row0, kEvictFirst);
                }
                if (++kr == cnt) kr = 0;
                if (++s == stages) { s = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 2) {
        if (elect_one()) {   // elect.sync: single-lane branch the compiler can see (no ELECT / BRA.U.ANY loop per TMA / MMA)
            // ---------------------------------------------------------------- TMA producer: token tiles (L2-resident)
            pdl_wait_prior_grid();                              // the activations are the previous kernel's output
            int s = 0, kr = rot;
            uint32_t phase = 0;
            for (int kb = 0; kb < cnt; ++kb) {
                if (p.dbg_nox && kb >= xstages) break;
                if (kb >= xstages) mbar_wait(&xempty_bar[s], phase ^ 1u);
                mbar_arrive_expect_tx(&xfull_bar[s], b_bytes);
                tma_load_2d(smem_x + s * b_bytes, &p.map_x, &xfull_bar[s], (kb0 + kr) * kBlockK, 0, kEvictLast);
                if (++kr == cnt) kr = 0;
                if (++s == xstages) { s = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {   // elect.sync: single-lane branch the compiler can see (no ELECT / BRA.U.ANY loop per TMA / MMA)
            // ---------------------------------------------------------------- MMA issuer
            // This one thread's instruction stream (waits, 4 x tcgen05.mma, commits per 16 KB of weights) paces the stage: the
            // shared-memory descriptors are built ONCE and advanced by adding to their 14-bit address field (>> 4 units;
            // shared-memory addresses are < 256 KB, so the field never carries into the next one).
            const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(smem), 0, 1024);
            const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(smem_x), 0, 1024);
            const uint32_t a_step = kABytes >> 4, b_step = b_bytes >> 4;
            uint32_t accumulate = 0;
            int s = 0, sx = 0;
            uint32_t phase = 0, xphase = 0;
            uint64_t a_desc = a_desc0, b_desc = b_desc0;
            for (int kb = 0; kb < cnt; ++kb) {
                if (!(p.dbg_nox && kb >= xstages)) mbar_wait(&xfull_bar[sx], xphase);
                mbar_wait(&full_bar[s], phase);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                    umma_f16<1>(tmem_base, a_desc + static_cast<uint64_t>(k * (kUmmaK * 2 / 16)),
                                b_desc + static_cast<uint64_t>(k * (kUmmaK * 2 / 16)), p.idesc, accumulate);
                    accumulate = 1;
                }
                umma_commit<1>(&empty_bar[s]);
                umma_commit<1>(&xempty_bar[sx]);
                if (++s == stages) {
                    s = 0;
                    phase ^= 1u;
                    a_desc = a_desc0;
                } else {
                    a_desc += a_step;
                }
                if (++sx == xstages) {
                    sx = 0;
                    xphase ^= 1u;
                    b_desc = b_desc0;
                } else {
                    b_desc += b_step;
                }
            }
            umma_commit<1>(tfull_bar);
        }
    }
    __syncwarp();

    // -------------------------------------------------------------------- epilogue (all four warps)
    // TMEM lane = weight row of this block, TMEM column = token.  Shared memory is reused as exchange space (the
    // ring is idle: every TMA load has landed and every MMA that read it has retired once tfull fires).
    pdl_wait_prior_grid();   // `out` may still be read by the previous kernel of the stream
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    // Warp w reads TMEM lanes 32 * (w % 4) .. + 31 (hardware rule); the `nparts` warps of a lane quarter split the tokens in
    // units of 8 (one tcgen05.ld.x8 each).
    float* part = reinterpret_cast<float*>(smem);
    const uint32_t quarter = warp & 3u;
    const int pt = static_cast<int>(warp >> 2);
    const int nparts = static_cast<int>(blockDim.x >> 7);
    const int row = static_cast<int>(quarter * 32 + lane);
    const uint32_t taddr = tmem_base + ((quarter * 32u) << 16);
    const int units = p.n_pad >> 3;            // even: n_pad is a multiple of 16
    T* out = static_cast<T*>((kEpi == DEC_LINEAR && p.ngroup > 1) ? p.out_g[prob] : p.out);

    if (p.splits == 1) {
        // ---- fast path: the whole reduction ran in this CTA
        if constexpr (kEpi == DEC_SWIGLU) {
            // lanes 0-63 hold gate rows, lanes 64-127 the up rows of the same 64 act columns.  Token chunks of 16
            // alternate between the two halves: the gate warps finish the even chunks (up values through shared
            // memory), the up warps the odd ones (gate values through shared memory); chunk pairs alternate between the
            // two warps of a lane quarter.
            const bool is_gate = quarter < 2;
            const int r = row & 63;
            const int col = row_block * 64 + r;
            const bool col_ok = col < p.rows_out;
            float bg = 0.f, bu = 0.f;
            if (col_ok) {
                if (p.bias[0] != nullptr) bg = static_cast<float>(static_cast<const T*>(p.bias[0])[col]);
                if (p.bias[1] != nullptr) bu = static_cast<float>(static_cast<const T*>(p.bias[1])[col]);
            }
            for (int uu = pt; uu < (units >> 1); uu += nparts) {          // units the OTHER kind finishes: export
                const int u = 2 * uu + (is_gate ? 1 : 0);
                uint32_t v[8];
                tmem_ld_32x32b_x8(taddr + u * 8, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) part[(u * 8 + j) * 64 + r] = __uint_as_float(v[j]);
            }
            __syncthreads();
            for (int uu = pt; uu < (units >> 1); uu += nparts) {
                const int u = 2 * uu + (is_gate ? 0 : 1);
                uint32_t v[8];
                tmem_ld_32x32b_x8(taddr + u * 8, v);
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = part[(u * 8 + j) * 64 + r];
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float mine = __uint_as_float(v[j]);
                    const float g = (is_gate ? mine : o[j]) + bg;
                    const float uu_ = (is_gate ? o[j] : mine) + bu;
                    const int n = u * 8 + j;
                    if (col_ok && n < p.tokens) {
                        const size_t o_idx = static_cast<size_t>(n) * p.ldo + col;
                        out[o_idx] = static_cast<T>(silu_f32(g) * uu_);
                        if (p.cache[0] != nullptr) {
                            static_cast<T*>(p.cache[0])[o_idx] = static_cast<T>(g);
                            static_cast<T*>(p.cache[1])[o_idx] = static_cast<T>(uu_);
                        }
                    }
                }
            }
        } else {
            const int col = rb_local * kRowsA + row;
            const bool col_ok = col < rows_out;
            float b = 0.f;
            if (col_ok && p.bias[0] != nullptr) b = static_cast<float>(static_cast<const T*>(p.bias[0])[col]);
            for (int u = pt; u < units; u += nparts) {
                uint32_t v[8];
                tmem_ld_32x32b_x8(taddr + u * 8, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = u * 8 + j;
                    if (col_ok && n < p.tokens) {
                        const size_t o_idx = static_cast<size_t>(n) * ldo + col;
                        float r = __uint_as_float(v[j]) + b;
                        if (p.addend != nullptr) r = static_cast<float>(static_cast<T>(r)) + static_cast<float>(static_cast<const T*>(p.addend)[o_idx]);
                        out[o_idx] = static_cast<T>(r);
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        if (warp == 1) {
            tc_fence_after();
            tmem_dealloc<1>(tmem_base, p.tmem_cols);
        }
        return;
    }

    // ---- split-K path: park the fp32 partial as part[token][row], reduce across the cluster through DSMEM
    if (cnt > 0) {
        for (int u = pt; u < units; u += nparts) {
            uint32_t v[8];
            tmem_ld_32x32b_x8(taddr + u * 8, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) part[(u * 8 + j) * kRowsA + row] = __uint_as_float(v[j]);
        }
    } else {
        for (int n = pt; n < p.n_pad; n += nparts) part[n * kRowsA + row] = 0.f;
    }
    tc_fence_before();
    __syncwarp();
    cluster_sync_all();   // partials of every split visible cluster-wide

    // each CTA of the cluster finishes a slice of the tokens: sum the splits (fixed order), fused epilogue, store
    const int per = (p.tokens + p.splits - 1) / p.splits;
    const int n_lo = split * per;
    const int n_hi = min(p.tokens, n_lo + per);
    const uint32_t part_addr = smem_u32(part);
    auto sum_splits = [&](uint32_t addr) {
        float v[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) v[s] = (s < p.splits) ? ld_dsmem_f32(addr, s) : 0.f;
        return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    };
    if constexpr (kEpi == DEC_SWIGLU) {
        const int r = row & 63;
        const int col = row_block * 64 + r;
        const bool col_ok = col < p.rows_out;
        float bg = 0.f, bu = 0.f;
        if (col_ok) {
            if (p.bias[0] != nullptr) bg = static_cast<float>(static_cast<const T*>(p.bias[0])[col]);
            if (p.bias[1] != nullptr) bu = static_cast<float>(static_cast<const T*>(p.bias[1])[col]);
        }
        for (int n = n_lo + (row >> 6) + 2 * pt; n < n_hi; n += 2 * nparts) {
            const float g = sum_splits(part_addr + (n * kRowsA + r) * 4) + bg;
            const float u = sum_splits(part_addr + (n * kRowsA + 64 + r) * 4) + bu;
            if (col_ok) {
                const size_t o_idx = static_cast<size_t>(n) * p.ldo + col;
                out[o_idx] = static_cast<T>(silu_f32(g) * u);
                if (p.cache[0] != nullptr) {
                    static_cast<T*>(p.cache[0])[o_idx] = static_cast<T>(g);
                    static_cast<T*>(p.cache[1])[o_idx] = static_cast<T>(u);
                }
            }
        }
    } else {
        const int col = rb_local * kRowsA + row;
        const bool col_ok = col < rows_out;
        float b = 0.f;
        if (col_ok && p.bias[0] != nullptr) b = static_cast<float>(static_cast<const T*>(p.bias[0])[col]);
        auto store_out = [&](int nn, float r) {
            const size_t o_idx = static_cast<size_t>(nn) * ldo + col;
            if (p.addend != nullptr) r = static_cast<float>(static_cast<T>(r)) + static_cast<float>(static_cast<const T*>(p.addend)[o_idx]);
            out[o_idx] = static_cast<T>(r);
        };
        int n = n_lo + pt;
        for (; n + nparts < n_hi; n += 2 * nparts) {          // two tokens of this warp's strided sequence at a time
            const float a0 = sum_splits(part_addr + (n * kRowsA + row) * 4);
            const float a1 = sum_splits(part_addr + ((n + nparts) * kRowsA + row) * 4);
            if (col_ok) {
                store_out(n, a0 + b);
                store_out(n + nparts, a1 + b);
            }
        }
        if (n < n_hi) {
            const float a0 = sum_splits(part_addr + (n * kRowsA + row) * 4);
            if (col_ok) store_out(n, a0 + b);
        }
    }
    __syncwarp();
    cluster_sync_all();   // nobody leaves while a peer may still read its partial
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, p.tmem_cols);
    }
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    if (v == nullptr || *v == '\0') return dflt;
    return atoi(v);
}

template <int kEpi, typename T>
int launch_decode(const DecodeParams& p, int grid, int threads, size_t smem_bytes, cudaStream_t s) {
    auto* kernel = ffn_decode_kernel<kEpi, T>;
    static size_t configured_dev[kMaxDevices] = {};   // per instantiation and device: dynamic smem opted into so far
    size_t& configured = configured_dev[current_device_slot()];
    if (smem_bytes > configured) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return static_cast<int>(e);
        configured = 227 * 1024;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(static_cast<unsigned>(threads));
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(p.splits);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = env_int("L32_DECODE_PDL", 1) ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p);
    if (e == cudaSuccess) count_launch();
    return static_cast<int>(e);
}

// Common host path.  rows_per_block = output rows one CTA produces (64 for SwiGLU: 64 gate + 64 up weight rows).
struct LinearGroup {          // grouped linear: `count` problems y_i[tokens, rows[i]] = x w_i^T sharing x and K
    int count;
    const void* w[kMaxLinearGroup];
    void* y[kMaxLinearGroup];
    int rows[kMaxLinearGroup];
};

template <int kEpi>
int decode_gemm(const void* x, const void* w0, const void* w1, const void* bias0, const void* bias1, const void* addend, void* out,
                void* cache0, void* cache1, int tokens, int k, int rows_out, int dtype, cudaStream_t s,
                const LinearGroup* group = nullptr) {
    if (tokens <= 0 || tokens > 128 || k <= 0 || rows_out <= 0) return L32_ERR_BAD_SHAPE;
    if ((k % 8) != 0 || (rows_out % 8) != 0) return L32_ERR_BAD_SHAPE;
    if (!is_aligned16(x) || !is_aligned16(w0) || (w1 != nullptr && !is_aligned16(w1))) return L32_ERR_BAD_ALIGN;
    constexpr int rows_per_block = (kEpi == DEC_SWIGLU) ? 64 : kRowsA;

    DecodeParams p;
    memset(&p, 0, sizeof(p));
    p.tokens = tokens;
    p.n_pad = ((tokens + 15) / 16) * 16;   // UMMA N: any multiple of 16 (M = 128); a smaller token tile leaves more ring stages
    p.rows_out = rows_out;
    p.k = k;
    p.ldo = rows_out;
    p.out = out;
    p.bias[0] = bias0;
    p.bias[1] = bias1;
    p.addend = addend;
    p.cache[0] = cache0;
    p.cache[1] = cache1;
    p.idesc = make_idesc_f16(dtype == L32_BF16, kRowsA, static_cast<uint32_t>(p.n_pad), false, false);
    p.tmem_cols = 32u;                     // TMEM allocations are powers of two >= 32 columns
    while (p.tmem_cols < static_cast<uint32_t>(p.n_pad)) p.tmem_cols *= 2u;

    int row_blocks = (rows_out + rows_per_block - 1) / rows_per_block;
    if (group != nullptr && group->count > 1) {
        if (kEpi != DEC_LINEAR || group->count > kMaxLinearGroup || bias0 != nullptr || addend != nullptr) return L32_ERR_BAD_SHAPE;
        p.ngroup = group->count;
        row_blocks = 0;
        for (int i = 0; i < group->count; ++i) {
            if (group->rows[i] <= 0 || (group->rows[i] % 8) != 0 || group->w[i] == nullptr || group->y[i] == nullptr) return L32_ERR_BAD_SHAPE;
            if (!is_aligned16(group->w[i])) return L32_ERR_BAD_ALIGN;
            p.rb_start[i] = row_blocks;
            p.rows_g[i] = group->rows[i];
            p.out_g[i] = group->y[i];
            row_blocks += (group->rows[i] + kRowsA - 1) / kRowsA;
        }
        p.rb_start[group->count] = row_blocks;
    }
    const int nkb = (k + kBlockK - 1) / kBlockK;
    const int sms = num_sms();
    // K splits (cluster size): enough CTAs to keep every SM streaming, each with at least 4 k-blocks.
    int splits = 1;
    while (splits < 8 && row_blocks * splits < sms && nkb / (splits * 2) >= 4) splits *= 2;
    if (const int v = env_int("L32_DECODE_SPLITS", 0)) splits = v;
    if (const int v = env_int(kEpi == DEC_SWIGLU ? "L32_DECODE_SPLITS_SWIGLU" : "L32_DECODE_SPLITS_LINEAR", 0)) splits = v;
    if (splits < 1 || splits > 8 || (splits & (splits - 1)) != 0 || splits > nkb) return L32_ERR_BAD_SHAPE;
    p.splits = splits;
    p.rotate = env_int("L32_DECODE_ROTATE", 1);
    p.dbg_nox = env_int("L32_DECODE_DEBUG_NOX", 0);
    const int grid = row_blocks * splits;

    // Ring depths: all CTAs co-resident (up to 4 per SM), shared memory split evenly between them.  The token ring is
    // shallow (its tiles come from L2; 3 stages = two k-blocks of lookahead, 2 when the tile is 8 KB or more), the weight
    // ring takes everything else.
    int per_sm = (grid + sms - 1) / sms;
    if (per_sm > 4) per_sm = 4;
    if (const int v = env_int("L32_DECODE_CTAS_PER_SM", 0)) per_sm = v;
    const int b_bytes = p.n_pad * 128;
    const int budget = (228 * 1024) / per_sm - 1024 /* driver-reserved */ - kBarrierBytes;
    const int kb_per_cta = (nkb + splits - 1) / splits;
    int xstages = b_bytes >= 8192 ? 2 : 3;
    if (const int v = env_int("L32_DECODE_XSTAGES", 0)) xstages = v;
    if (xstages > kb_per_cta) xstages = kb_per_cta;
    if (xstages > kMaxXStages) xstages = kMaxXStages;
    if (xstages < 1) xstages = 1;
    int stages = (budget - xstages * b_bytes) / kABytes;
    if (stages > kb_per_cta) stages = kb_per_cta;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) stages = 2;
    if (const int v = env_int("L32_DECODE_STAGES", 0)) stages = v < kMaxStages ? v : kMaxStages;
    if (env_int("L32_DECODE_XSTAGES", 0) == 0) {   // what the weight ring leaves over deepens the token ring
        const int spare = (budget - stages * kABytes) / b_bytes;
        if (spare > xstages) xstages = spare < kMaxXStages ? spare : kMaxXStages;
        if (xstages > kb_per_cta) xstages = kb_per_cta;
    }
    // the fp32 partial [n_pad][128] of the epilogue reuses the rings (contiguous: weights first, then tokens)
    while (stages * kABytes + xstages * b_bytes < p.n_pad * kRowsA * 4) ++stages;
    if (stages < 1 || stages > kMaxStages) return L32_ERR_BAD_SHAPE;
    p.stages = stages;
    p.xstages = xstages;
    const size_t smem_bytes = static_cast<size_t>(stages) * kABytes + static_cast<size_t>(xstages) * b_bytes + kBarrierBytes;
    if (smem_bytes > 227 * 1024) return L32_ERR_BAD_SHAPE;

    int rc = make_tensor_map_2d(&p.map_x, x, static_cast<uint64_t>(tokens), static_cast<uint64_t>(k), static_cast<uint64_t>(k),
                                static_cast<uint32_t>(p.n_pad), kBlockK, dtype);
    if (rc != L32_OK) return rc;
    if (p.ngroup > 1) {
        for (int i = 0; i < p.ngroup; ++i) {
            rc = make_tensor_map_2d(&p.map_w[i], group->w[i], static_cast<uint64_t>(group->rows[i]), static_cast<uint64_t>(k),
                                    static_cast<uint64_t>(k), rows_per_block, kBlockK, dtype);
            if (rc != L32_OK) return rc;
        }
    } else {
        rc = make_tensor_map_2d(&p.map_w[0], w0, static_cast<uint64_t>(rows_out), static_cast<uint64_t>(k), static_cast<uint64_t>(k),
                                rows_per_block, kBlockK, dtype);
        if (rc != L32_OK) return rc;
    }
    if constexpr (kEpi == DEC_SWIGLU) {
        rc = make_tensor_map_2d(&p.map_w[1], w1, static_cast<uint64_t>(rows_out), static_cast<uint64_t>(k),
                                static_cast<uint64_t>(k), rows_per_block, kBlockK, dtype);
        if (rc != L32_OK) return rc;
    }
    // epilogue parallelism: 8 warps above 16 tokens (measured: 4 -> 8 warps takes 2 us off a 64-token step, 16 add nothing)
    int threads = p.n_pad > 16 ? 256 : 128;
    if (const int v = env_int("L32_DECODE_THREADS", 0)) threads = v;
    if (threads < 128 || threads > kMaxThreads || (threads % 128) != 0) return L32_ERR_BAD_SHAPE;
    if (dtype == L32_BF16) return launch_decode<kEpi, __nv_bfloat16>(p, grid, threads, smem_bytes, s);
    return launch_decode<kEpi, __half>(p, grid, threads, smem_bytes, s);
}

}  // namespace

int ffn_decode_swiglu(const void* x, const void* w_gate, const void* w_up, const void* b_gate, const void* b_up, void* act,
                      void* gate_cache, void* up_cache, int tokens, int hidden, int inter, int dtype, cudaStream_t s) {
    return decode_gemm<DEC_SWIGLU>(x, w_gate, w_up, b_gate, b_up, nullptr, act, gate_cache, up_cache, tokens, hidden, inter, dtype, s);
}

int ffn_decode_linear_group(const void* a, const void* const* w, void* const* y, const int* out_features, int count, int tokens,
                            int in_features, int dtype, cudaStream_t s) {
    if (count < 1 || count > kMaxLinearGroup) return L32_ERR_BAD_SHAPE;
    LinearGroup g;
    g.count = count;
    for (int i = 0; i < count; ++i) {
        g.w[i] = w[i];
        g.y[i] = y[i];
        g.rows[i] = out_features[i];
    }
    return decode_gemm<DEC_LINEAR>(a, w[0], nullptr, nullptr, nullptr, nullptr, y[0], nullptr, nullptr, tokens, in_features,
                                   out_features[0], dtype, s, count > 1 ? &g : nullptr);
}

int ffn_decode_linear(const void* a, const void* w, const void* bias, const void* addend, void* y, int tokens, int in_features,
                      int out_features, int dtype, cudaStream_t s) {
    return decode_gemm<DEC_LINEAR>(a, w, nullptr, bias, nullptr, addend, y, nullptr, nullptr, tokens, in_features, out_features, dtype, s);
}

}  // namespace l32
