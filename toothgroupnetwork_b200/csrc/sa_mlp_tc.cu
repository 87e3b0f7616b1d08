// sa_mlp_tc.cu -- tcgen05 (5th-gen tensor core) engine of the fused set-abstraction body.
//
// Same contract as the fp32 engine in sa_mlp.cu, but the per-layer contraction
//      D[128 rows, N] = A[128 rows, K] * W[N, K]^T
// runs on the tensor cores with the accumulator in TENSOR MEMORY:
//   * a CTA is 4 independent TILE GROUPS of 128 threads; a group owns one 128-row tile (128/K
//     neighbourhoods) at a time, thread r of the group owns row r end to end: it gathers its
//     neighbour's [xyz_rel | feats] row, and after each layer reads ITS accumulator lane back with
//     tcgen05.ld (32x32b: warp w owns TMEM lanes 32(w%4)..+31), applies bias + ReLU and writes
//     the next layer's A operand -- the grouped tensor and the inter-layer activations never
//     leave the SM.  The four groups share the weights but nothing else (own TMEM columns, own
//     operand buffers, own mbarrier, named barriers instead of __syncthreads), so while one group
//     waits for its gather or its MMAs the others fill the issue slots and the tensor pipe;
//   * operands live in shared memory in the canonical K-major, no-swizzle UMMA layout
//     (8x16-byte core matrices; element (r,k) at (k/4)*LBO + r*16 + (k%4)*4 with SBO = 128 B), which a
//     row-per-thread writer fills with conflict-free 16-byte stores;
//   * fp32 fidelity: inputs are split a = a_hi + a_lo with a_hi the tf32 truncation, and each
//     K-step issues three kind::tf32 MMAs (hi*hi + lo*hi + hi*lo) into the same accumulator
//     ("3xTF32"), so the result carries ~2^-21 relative error instead of tf32's 2^-10 -- the
//     reference's own default for these 1x1 convolutions is plain TF32;
//   * one elected thread per group issues the MMAs and tcgen05.commit's the group's mbarrier;
//   * the max over the K rows of a neighbourhood is a CREDUX per output channel when K is 16 or 32
//     (the rows of a neighbourhood are the lanes of one warp), a shared-memory pass otherwise.
// Weights are split and laid out once per CTA; CTAs are persistent over tiles.
#include <algorithm>

#include "common.cuh"
#include "sa_mlp.cuh"
#include "tgn_b200.h"

namespace tgn {
namespace {

constexpr int kRows = 128;                            // rows of a tile = threads of a tile group
constexpr int kGroups = 4;
constexpr int kThreads = kRows * kGroups;
constexpr uint32_t kChunkStrideA = kRows * 16;        // LBO of the A operand: one 16-byte K-chunk of all rows
constexpr unsigned FULL = 0xffffffffu;

struct TcLayout {
    int kpad[kSaMaxLayers];          // K of layer l, multiple of 8
    int npad[kSaMaxLayers];          // N of layer l, multiple of 16
    uint32_t w_hi[kSaMaxLayers], w_lo[kSaMaxLayers], bias[kSaMaxLayers];   // byte offsets in dynamic smem
    uint32_t a_hi[kGroups], a_lo[kGroups];   // per group A operands; a_hi doubles as fp32 staging for generic K
    uint32_t part[kGroups];          // per group: up to 8 per-warp partial maxima rows of npad floats
    uint32_t ones;                   // constant [128 x 8] A tile that adds the bias through the MMA
    uint32_t misc;                   // tmem base (u32) @0, group mbarriers (u64) @8+8g
    uint32_t total;
    uint32_t tmem_cols, group_cols;
    int tiles_per_cloud, gpt, stage_stride;
};

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;                 // descriptor version for sm_100
    return d;                        // layout type 0 (no swizzle), base offset 0
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(kRows >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(kRows) : "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// a = hi + lo with hi the tf32 truncation of a (exact split)
__device__ __forceinline__ void split_tf32(float a, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(a) & 0xFFFFE000u;
    lo = __float_as_uint(__fsub_rn(a, __uint_as_float(hi)));
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
sa_mlp_tc_kernel(const SaParams p, const TcLayout lay)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = tid / kRows;                 // tile group
    const int r = tid - g * kRows;             // row inside the tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + lay.misc);
    const uint32_t bar = sbase + lay.misc + 8 + 8 * g;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + lay.misc), "r"(lay.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (r == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    // ---- weights: split into tf32 hi/lo, canonical K-major layout (LBO = npad*16, SBO = 128) -------
    for (int l = 0; l < p.L; ++l) {
        const int cin = p.ch[l], cout = p.ch[l + 1], kp = lay.kpad[l], np = lay.npad[l];
        for (int e = tid; e < np * kp; e += kThreads) {
            const int n = e / kp, k = e - n * kp;
            const float w = (n < cout && k < cin) ? __ldg(p.W[l] + static_cast<size_t>(n) * cin + k) : 0.f;
            uint32_t hi, lo;
            split_tf32(w, hi, lo);
            const uint32_t off = static_cast<uint32_t>(k >> 2) * (np * 16) + n * 16 + (k & 3) * 4;
            *reinterpret_cast<uint32_t*>(smem + lay.w_hi[l] + off) = hi;
            *reinterpret_cast<uint32_t*>(smem + lay.w_lo[l] + off) = lo;
        }
        // bias as a B operand of one extra K=8 step: column k=0 holds bias_hi, k=1 bias_lo (rest 0);
        // multiplied by the constant "ones" A tile below it adds bias_hi + bias_lo to every row.
        for (int e = tid; e < np * 8; e += kThreads) {
            const int n = e >> 3, k = e & 7;
            uint32_t hi = 0u, lo = 0u;
            if (n < cout) split_tf32(__ldg(p.bias[l] + n), hi, lo);
            const uint32_t off = static_cast<uint32_t>(k >> 2) * (np * 16) + n * 16 + (k & 3) * 4;
            *reinterpret_cast<uint32_t*>(smem + lay.bias[l] + off) = k == 0 ? hi : (k == 1 ? lo : 0u);
        }
    }
    // constant A tile [128 rows x 8]: columns 0 and 1 are 1.0 (exact in tf32), the rest 0
    for (int e = tid; e < kRows * 8; e += kThreads) {
        const int rr = e >> 3, k = e & 7;
        *reinterpret_cast<float*>(smem + lay.ones + static_cast<uint32_t>(k >> 2) * kChunkStrideA + rr * 16 + (k & 3) * 4) =
            k < 2 ? 1.0f : 0.0f;
    }
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot + g * lay.group_cols;                       // this group's accumulator columns
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t a_hi = sbase + lay.a_hi[g], a_lo = sbase + lay.a_lo[g];
    uint32_t phase = 0;

    const int cin0 = p.ch[0];
    const int cout_last = p.ch[p.L];
    const int total_tiles = lay.tiles_per_cloud * p.B;
    const int tile_step = gridDim.x * kGroups;
    const bool lane_max = (p.K == 16 || p.K == 32 || p.K == 64 || p.K == 128);
    float* part = reinterpret_cast<float*>(smem + lay.part[g]);
    const uint64_t desc_a_hi = make_smem_desc(a_hi, kChunkStrideA, 128);
    const uint64_t desc_a_lo = make_smem_desc(a_lo, kChunkStrideA, 128);
    const int xoff = p.xyz_first ? 0 : p.D;       // first channel of the xyz_rel block
    const int foff = p.xyz_first ? 3 : 0;         // first channel of the feature block
    // neighbour index of this thread's row in tile `t` (-1 past the end / past the rows of the tile)
    auto load_index = [&](int t) -> int {
        if (t >= total_tiles) return -1;
        const int tb = t / lay.tiles_per_cloud;
        const int ts0 = (t - tb * lay.tiles_per_cloud) * lay.gpt;
        if (r >= min(lay.gpt, p.S - ts0) * p.K) return -1;
        return __ldg(p.gidx + (static_cast<size_t>(tb) * p.S + ts0) * p.K + r);
    };
    // the 16-wide zero-padded [xyz_rel | feats] (or [feats | xyz_rel]) row of this thread in tile `t`
    auto load_row16 = [&](int t, int j, float (&out)[16]) {
        const bool ok = t < total_tiles && j >= 0 && j < p.N;
        const int tb = ok ? t / lay.tiles_per_cloud : 0;
        const int ts = ok ? (t - tb * lay.tiles_per_cloud) * lay.gpt + r / p.K : 0;
        const int jj = ok ? j : 0;
        const float* px = p.xyz + 3 * (static_cast<size_t>(tb) * p.N + jj);
        const float* pc = p.new_xyz + 3 * (static_cast<size_t>(tb) * p.S + ts);
        const float* pf = p.feats ? p.feats + (static_cast<size_t>(tb) * p.N + jj) * p.D : px;
        const float rel0 = __fsub_rn(__ldg(px), __ldg(pc)), rel1 = __fsub_rn(__ldg(px + 1), __ldg(pc + 1)),
                    rel2 = __fsub_rn(__ldg(px + 2), __ldg(pc + 2));
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const int xc = c - xoff;
            float v = xc == 0 ? rel0 : (xc == 1 ? rel1 : rel2);
            if (static_cast<unsigned>(xc) > 2u) {
                const unsigned fc = static_cast<unsigned>(c - foff);
                v = fc < static_cast<unsigned>(p.D) ? __ldg(pf + fc) : 0.f;
            }
            out[c] = ok ? v : 0.f;
        }
    };
    const bool fast_rows = lay.kpad[0] == 16;     // the whole padded input row fits 16 registers
    float row[16];
    int j_next;
    {
        const int t0 = blockIdx.x * kGroups + g;
        if (fast_rows) {
            load_row16(t0, load_index(t0), row);
            j_next = load_index(t0 + tile_step);
        } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) row[c] = 0.f;
            j_next = load_index(t0);
        }
    }

    for (int tile = blockIdx.x * kGroups + g; tile < total_tiles; tile += tile_step) {
        const int b = tile / lay.tiles_per_cloud;
        const int s0 = (tile - b * lay.tiles_per_cloud) * lay.gpt;
        const int groups = min(lay.gpt, p.S - s0);
        const int rows = groups * p.K;

        // ---- layer-0 operand: this thread's grouped row --------------------------------------------
        if (fast_rows) {
            // the 16-wide padded row was loaded into registers while the previous tile computed
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) split_tf32(row[4 * kc + i], hi[i], lo[i]);
                const uint32_t off = kc * kChunkStrideA + r * 16;
                st_shared_v4(a_hi + off, hi[0], hi[1], hi[2], hi[3]);
                st_shared_v4(a_lo + off, lo[0], lo[1], lo[2], lo[3]);
            }
            load_row16(tile + tile_step, j_next, row);          // consumed one tile later: latency hidden
            j_next = load_index(tile + 2 * tile_step);
        } else {
            int j = j_next;                               // index prefetched while the previous tile computed
            const int s = s0 + (r < rows ? r / p.K : 0);
            if (r >= rows || j < 0 || j >= p.N) j = -1;
            const int jj = j < 0 ? 0 : j;
            const float* px = p.xyz + 3 * (static_cast<size_t>(b) * p.N + jj);
            const float* pc = p.new_xyz + 3 * (static_cast<size_t>(b) * p.S + s);
            const float* pf = p.feats ? p.feats + (static_cast<size_t>(b) * p.N + jj) * p.D : px;
            const float rel0 = __fsub_rn(__ldg(px), __ldg(pc)), rel1 = __fsub_rn(__ldg(px + 1), __ldg(pc + 1)),
                        rel2 = __fsub_rn(__ldg(px + 2), __ldg(pc + 2));
            for (int kc = 0; kc < lay.kpad[0] / 4; ++kc) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = 4 * kc + i;
                    const int xc = c - xoff;
                    float v = xc == 0 ? rel0 : (xc == 1 ? rel1 : rel2);
                    if (static_cast<unsigned>(xc) > 2u) {
                        const unsigned fc = static_cast<unsigned>(c - foff);
                        v = fc < static_cast<unsigned>(p.D) ? __ldg(pf + fc) : 0.f;
                    }
                    if (j < 0) v = 0.f;
                    split_tf32(v, hi[i], lo[i]);
                }
                const uint32_t off = kc * kChunkStrideA + r * 16;
                st_shared_v4(a_hi + off, hi[0], hi[1], hi[2], hi[3]);
                st_shared_v4(a_lo + off, lo[0], lo[1], lo[2], lo[3]);
            }
            j_next = load_index(tile + tile_step);
        }

        for (int l = 0; l < p.L; ++l) {
            const int kp = lay.kpad[l], np = lay.npad[l];
            proxy_fence_async();           // this thread's operand stores -> visible to the tensor core
            tc_fence_before();
            group_sync(g);
            if ((warp & 3) == 0) {
                // The whole first warp of the group runs the issue code (warp-uniform descriptor maths
                // stays on the uniform datapath); one lane issues.  The start-address field is the low
                // 14 bits (units of 16 B): a K-step advances it by a constant.
                tc_fence_after();
                const uint32_t idesc = make_idesc_tf32(np);
                const uint32_t lbo_b = static_cast<uint32_t>(np) * 16;
                uint64_t d_ahi = desc_a_hi, d_alo = desc_a_lo;
                uint64_t d_bhi = make_smem_desc(sbase + lay.w_hi[l], lbo_b, 128);
                uint64_t d_blo = make_smem_desc(sbase + lay.w_lo[l], lbo_b, 128);
                const uint64_t step_a = (2 * kChunkStrideA) >> 4, step_b = (2 * lbo_b) >> 4;
                const int nks = kp / 8;
                for (int ks = 0; ks < nks; ++ks) {
                    if (lane == 0) {
                        mma_tf32_ss(tmem_base, d_ahi, d_bhi, idesc, ks > 0);
                        mma_tf32_ss(tmem_base, d_alo, d_bhi, idesc, true);
                        mma_tf32_ss(tmem_base, d_ahi, d_blo, idesc, true);
                    }
                    d_ahi += step_a; d_alo += step_a; d_bhi += step_b; d_blo += step_b;
                }
                if (lane == 0) {
                    mma_tf32_ss(tmem_base, make_smem_desc(sbase + lay.ones, kChunkStrideA, 128),
                                make_smem_desc(sbase + lay.bias[l], lbo_b, 128), idesc, true);     // + bias
                    mma_commit(bar);       // arrives on the group's mbarrier when the MMAs above have completed
                }
                __syncwarp();
            }
            mbar_wait_relaxed(bar, phase);
            phase ^= 1;
            tc_fence_after();

            const bool last = (l == p.L - 1);
            for (int c0 = 0; c0 < np; c0 += 32) {
                uint32_t v[32];
                const bool full = np - c0 >= 32;             // np is a multiple of 16: a chunk is 32 or 16 wide
                if (full) tmem_ld32(tmem_row + c0, v);
                else tmem_ld16(tmem_row + c0, v);
                if (!last) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        if (i < 16 || full) {
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                split_tf32(fmaxf(__uint_as_float(v[i + u]), 0.f), hi[u], lo[u]);
                            const uint32_t off = static_cast<uint32_t>((c0 + i) >> 2) * kChunkStrideA + r * 16;
                            st_shared_v4(a_hi + off, hi[0], hi[1], hi[2], hi[3]);
                            st_shared_v4(a_lo + off, lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                } else if (lane_max) {
                    // 32 rows of a neighbourhood (or 2 x 16) are the lanes of this warp: one CREDUX per
                    // channel gives the warp's partial maximum.  Post-ReLU values are >= 0, so signed-int
                    // order == float order; rows beyond `rows` contribute 0, the identity.
                    const int wq = warp & 3;
                    const float live = r < rows ? 1.f : 0.f;  // x * live: exact for the rows that count, 0 else
                    if (p.K != 16) {
                        float* dst = part + wq * np + c0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (i < 16 || full) {
                                const float x = fmaxf(__uint_as_float(v[i]), 0.f) * live;
                                const int mx = __reduce_max_sync(FULL, __float_as_int(x));
                                if (lane == 0) dst[i] = __int_as_float(mx);
                            }
                        }
                    } else {
                        float* da = part + (2 * wq) * np + c0;
                        float* db = part + (2 * wq + 1) * np + c0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (i < 16 || full) {
                                const float x = fmaxf(__uint_as_float(v[i]), 0.f) * live;
                                const int ma = __reduce_max_sync(FULL, lane < 16 ? __float_as_int(x) : 0);
                                const int mb = __reduce_max_sync(FULL, lane >= 16 ? __float_as_int(x) : 0);
                                if (lane == 0) { da[i] = __int_as_float(ma); db[i] = __int_as_float(mb); }
                            }
                        }
                    }
                } else {
                    float* stage = reinterpret_cast<float*>(smem + lay.a_hi[g]);
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i < 16 || full)
                            stage[r * lay.stage_stride + c0 + i] = fmaxf(__uint_as_float(v[i]), 0.f);
                }
            }
            tc_fence_before();             // TMEM reads done before the next layer's MMAs overwrite D
        }

        if (lane_max) {
            // ---- combine the per-warp partial maxima of each neighbourhood, channel-first store -----------
            group_sync(g);
            const int np = lay.npad[p.L - 1];
            const int ppn = p.K == 16 ? 1 : p.K / 32;          // partial rows per neighbourhood
            float* ob = p.out + (static_cast<size_t>(b) * p.out_c_total + p.out_c_offset) * p.S;
            for (int e = r; e < groups * cout_last; e += kRows) {
                const int co = e / groups, sg = e - co * groups;
                float m = part[(sg * ppn) * np + co];
                for (int q = 1; q < ppn; ++q) m = fmaxf(m, part[(sg * ppn + q) * np + co]);
                ob[static_cast<size_t>(co) * p.S + s0 + sg] = m;
            }
        } else {
            // ---- generic K: max over the K rows of each neighbourhood through shared memory -------------
            group_sync(g);
            const float* stage = reinterpret_cast<const float*>(smem + lay.a_hi[g]);
            float* ob = p.out + (static_cast<size_t>(b) * p.out_c_total + p.out_c_offset) * p.S;
            for (int e = r; e < groups * cout_last; e += kRows) {
                const int sg = e / cout_last, co = e - sg * cout_last;
                float m = stage[(sg * p.K) * lay.stage_stride + co];
                for (int k = 1; k < p.K; ++k) m = fmaxf(m, stage[(sg * p.K + k) * lay.stage_stride + co]);
                ob[static_cast<size_t>(co) * p.S + s0 + sg] = m;
            }
            group_sync(g);                 // staging area becomes the next tile's A operand
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(lay.tmem_cols) : "memory");
    }
}

bool make_layout(const SaParams& p, TcLayout& lay)
{
    if (p.K > kRows || p.K < 1) return false;
    int kmax = 0, nmax = 0, nlast = 0;
    uint32_t off = 0;
    auto take = [&off](uint32_t bytes, uint32_t align) {
        off = (off + align - 1) / align * align;
        const uint32_t o = off;
        off += bytes;
        return o;
    };
    for (int l = 0; l < p.L; ++l) {
        lay.kpad[l] = l == 0 ? (p.ch[0] + 7) / 8 * 8 : lay.npad[l - 1];
        lay.npad[l] = (p.ch[l + 1] + 15) / 16 * 16;
        if (lay.npad[l] > 128 || lay.kpad[l] > 128) return false;
        kmax = std::max(kmax, lay.kpad[l]);
        nmax = std::max(nmax, lay.npad[l]);
        nlast = lay.npad[l];
    }
    lay.stage_stride = nlast + 1;
    const bool lane_max = (p.K == 16 || p.K == 32 || p.K == 64 || p.K == 128);
    const uint32_t a_bytes = static_cast<uint32_t>(kmax / 4) * kChunkStrideA;
    const uint32_t stage_bytes = lane_max ? 0u : static_cast<uint32_t>(kRows) * lay.stage_stride * 4;
    for (int g = 0; g < kGroups; ++g) {
        lay.a_hi[g] = take(std::max(a_bytes, stage_bytes), 128);
        lay.a_lo[g] = take(a_bytes, 128);
        lay.part[g] = take(lane_max ? 8u * nlast * 4u : 0u, 16);
    }
    for (int l = 0; l < p.L; ++l) {
        const uint32_t wb = static_cast<uint32_t>(lay.npad[l]) * lay.kpad[l] * 4;
        lay.w_hi[l] = take(wb, 128);
        lay.w_lo[l] = take(wb, 128);
        lay.bias[l] = take(lay.npad[l] * 32, 128);       // [npad x 8] B tile: bias_hi | bias_lo | 0...
    }
    lay.ones = take(kRows * 32, 128);
    lay.misc = take(8 + 8 * kGroups, 16);
    lay.total = off;
    lay.group_cols = nmax <= 32 ? 32 : (nmax <= 64 ? 64 : 128);
    lay.tmem_cols = lay.group_cols * kGroups;          // 128, 256 or 512: a power of two >= 32
    lay.gpt = kRows / p.K;
    lay.tiles_per_cloud = (p.S + lay.gpt - 1) / lay.gpt;
    return lay.total <= 220 * 1024;
}

}  // namespace

bool sa_mlp_tc_supported(const SaParams& p)
{
    TcLayout lay{};
    return make_layout(p, lay);
}

int sa_mlp_tc_launch(SaParams p, cudaStream_t st)
{
    TcLayout lay{};
    if (!make_layout(p, lay)) { set_error("sa_group_mlp_max: shape not supported by the tcgen05 engine"); return TGN_ERR_INVALID; }
    static uint32_t configured = 0;
    if (lay.total > configured) {
        const cudaError_t e = cudaFuncSetAttribute(sa_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(lay.total));
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return TGN_ERR_CUDA; }
        configured = lay.total;
    }
    // one persistent CTA (4 tile groups, the whole TMEM budget of its accumulators) per SM
    const long long tiles = static_cast<long long>(lay.tiles_per_cloud) * p.B;
    const int grid = static_cast<int>(std::min<long long>((tiles + kGroups - 1) / kGroups, sm_count()));
    sa_mlp_tc_kernel<<<grid, kThreads, lay.total, st>>>(p, lay);
    return check_launch("sa_mlp_tc_kernel");
}

}  // namespace tgn
