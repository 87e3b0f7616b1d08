// sa_mlp_tc8.cu -- tcgen05 engine of the fused set-abstraction body with EIGHT tile groups per SM.
//
// The 3xTF32 engine (sa_mlp_tc.cu) is latency-bound: four 128-row tiles in flight per SM leave the issue slots 29 % and
// the tensor pipe 36 % busy (profiles/r1f), and its MMA stream is 5.5x the useful work (last layer padded from 64 to 128
// channels for the transposed form, a K=8 bias MMA per layer).  Each tile group reserves 128 TMEM columns there because
// the inner layers' operands live in tensor memory next to a 128-column transposed accumulator.  This engine trades
// that for concurrency:
//   * every layer runs NON-transposed, D[row, channel] = X[row, k] * W[channel, k], so a group's accumulator is at most
//     64 columns and EIGHT groups (1024 threads, <= 64 registers each) share the 512 TMEM columns of an SM: twice the
//     tiles in flight;
//   * operands are split into THREE bf16 parts (x = x1 + x2 + x3 exactly) and every K=16 step issues six kind::f16 MMAs
//     (x1w1, x1w2, x2w1, x2w2, x1w3, x3w1): ~2^-23 per product, tighter than 3xTF32's 2^-21, same MMA count per K;
//     the activation parts of a tile live in shared memory (24 KB per group at 32 channels);
//   * biases are added in the epilogue (exact fp32), not through an extra MMA;
//   * the max over the K rows of a neighbourhood: K = 32 rows are the 32 lanes of one warp (K = 16: a half warp), and the
//     post-ReLU values are non-negative floats, whose bit patterns order like integers -- one REDUX.MAX per channel;
//     the group's results are staged through its (now idle) operand buffer and leave as 16-byte rows of the channel-first output.
// Shapes: C_in <= 16, every width <= 64, K in {16, 32}; everything else stays on the other engines.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "sa_mlp.cuh"
#include "tc_common.cuh"
#include "tgn_b200.h"

namespace tgn {
namespace {

using namespace tc;

constexpr int kRows = 128;
constexpr uint32_t kChunk = kRows * 16;          // bytes of one 16-byte K-chunk (8 bf16) of all 128 rows: LBO of the A operand
constexpr unsigned FULL = 0xffffffffu;

struct Lay8 {
    int kpad[kSaMaxLayers];          // multiple of 16
    int npad[kSaMaxLayers];          // multiple of 16, <= 64
    uint32_t w[kSaMaxLayers];        // byte offset of the layer's weight operand: three parts of npad x kpad bf16, part stride w_part
    uint32_t w_part[kSaMaxLayers];
    uint32_t bias[kSaMaxLayers];     // float[npad]
    uint32_t act;                    // first group's operand buffer; group stride act_stride; part stride act_part
    uint32_t act_stride, act_part;
    uint32_t stage, stage_stride;    // per group staging of the pooled outputs: float[npad_last][gpt]
    uint32_t misc, total;
    int groups;
    int tiles_per_cloud, gpt;
    uint32_t tpc_magic;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ void split3(float a, float b, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
    p1 = pack_bf16(a, b);
    const float ra = __fsub_rn(a, __uint_as_float(p1 << 16)), rb = __fsub_rn(b, __uint_as_float(p1 & 0xFFFF0000u));
    p2 = pack_bf16(ra, rb);
    p3 = pack_bf16(__fsub_rn(ra, __uint_as_float(p2 << 16)), __fsub_rn(rb, __uint_as_float(p2 & 0xFFFF0000u)));
}
__device__ __forceinline__ void group_sync_n(int g, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(threads) : "memory"); }

// 16 activation values (columns c0..c0+15 of row r) -> three bf16 parts, two 16-byte K-chunks each
__device__ __forceinline__ void store_parts16(const float (&x)[16], uint32_t buf, uint32_t part_stride, int c0, int r)
{
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t p1[4], p2[4], p3[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) split3(x[8 * h + 2 * u], x[8 * h + 2 * u + 1], p1[u], p2[u], p3[u]);
        const uint32_t off = static_cast<uint32_t>((c0 >> 3) + h) * kChunk + r * 16;
        st_shared_v4(buf + off, p1[0], p1[1], p1[2], p1[3]);
        st_shared_v4(buf + part_stride + off, p2[0], p2[1], p2[2], p2[3]);
        st_shared_v4(buf + 2 * part_stride + off, p3[0], p3[1], p3[2], p3[3]);
    }
}

template <int kGroups>
__global__ void __launch_bounds__(kRows * kGroups, 1)
sa_mlp_tc8_kernel(const SaParams p, const Lay8 lay)
{
    constexpr int kThreads = kRows * kGroups;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = tid / kRows;
    const int r = tid - g * kRows;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + lay.misc);
    const uint32_t bar = sbase + lay.misc + 8 + 8 * g;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + lay.misc), "r"(64 * kGroups) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (r == 0) { mbar_init(bar, 1); mbar_fence_init(); }

    // ---- weights (three bf16 parts, B operand [np x kp], LBO = np*16) and biases, once per CTA ---------------------------
    for (int l = 0; l < p.L; ++l) {
        const int cin = p.ch[l], cout = p.ch[l + 1], kp = lay.kpad[l], np = lay.npad[l];
        for (int e = tid; e < np * kp; e += kThreads) {
            const int n = e / kp, k = e - n * kp;
            const float w = (n < cout && k < cin) ? __ldg(p.W[l] + static_cast<size_t>(n) * cin + k) : 0.f;
            const __nv_bfloat16 w1 = __float2bfloat16_rn(w);
            const float r1 = __fsub_rn(w, __bfloat162float(w1));
            const __nv_bfloat16 w2 = __float2bfloat16_rn(r1);
            const __nv_bfloat16 w3 = __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(w2)));
            const uint32_t off = static_cast<uint32_t>(k >> 3) * (np * 16) + n * 16 + (k & 7) * 2;
            *reinterpret_cast<__nv_bfloat16*>(smem + lay.w[l] + off) = w1;
            *reinterpret_cast<__nv_bfloat16*>(smem + lay.w[l] + lay.w_part[l] + off) = w2;
            *reinterpret_cast<__nv_bfloat16*>(smem + lay.w[l] + 2 * lay.w_part[l] + off) = w3;
        }
        for (int n = tid; n < np; n += kThreads)
            *reinterpret_cast<float*>(smem + lay.bias[l] + 4 * n) = n < cout ? __ldg(p.bias[l] + n) : 0.f;
    }
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot + g * 64;
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t xbuf = sbase + lay.act + g * lay.act_stride;
    uint32_t phase = 0;

    const int cout_last = p.ch[p.L];
    const int total_tiles = lay.tiles_per_cloud * p.B;
    const int tile_step = gridDim.x * kGroups;
    const int r_div_k = r / p.K;
    auto cloud_of = [&](int t) -> int {
        int q = static_cast<int>(__umulhi(static_cast<unsigned>(t), lay.tpc_magic));
        if (t - q * lay.tiles_per_cloud >= lay.tiles_per_cloud) ++q;
        return q;
    };
    auto load_index = [&](int t) -> int {
        if (t >= total_tiles) return -1;
        const int tb = cloud_of(t);
        const int ts0 = (t - tb * lay.tiles_per_cloud) * lay.gpt;
        if (r >= min(lay.gpt, p.S - ts0) * p.K) return -1;
        return __ldg(p.gidx + (static_cast<size_t>(tb) * p.S + ts0) * p.K + r);
    };
    auto load_row16 = [&](int t, int j, float (&out)[16]) {
#pragma unroll
        for (int c = 0; c < 16; ++c) out[c] = 0.f;
        if (t < total_tiles && j >= 0 && j < p.N) {
            const int tb = cloud_of(t);
            const int ts = (t - tb * lay.tiles_per_cloud) * lay.gpt + r_div_k;
            const float* px = p.xyz + 3 * (static_cast<size_t>(tb) * p.N + j);
            const float* pc = p.new_xyz + 3 * (static_cast<size_t>(tb) * p.S + ts);
            const float* pf = p.feats ? p.feats + (static_cast<size_t>(tb) * p.N + j) * p.D : px;
            const float rel0 = __fsub_rn(__ldg(px), __ldg(pc)), rel1 = __fsub_rn(__ldg(px + 1), __ldg(pc + 1)),
                        rel2 = __fsub_rn(__ldg(px + 2), __ldg(pc + 2));
            if (p.xyz_first) {
                out[0] = rel0; out[1] = rel1; out[2] = rel2;
#pragma unroll
                for (int i = 0; i < 13; ++i)
                    if (i < p.D) out[3 + i] = __ldg(pf + i);
            } else {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    if (c < p.D) out[c] = __ldg(pf + c);
                    else if (c == p.D) out[c] = rel0;
                    else if (c == p.D + 1) out[c] = rel1;
                    else if (c == p.D + 2) out[c] = rel2;
                }
            }
        }
    };
    float row[16];
    int j_next;
    {
        const int t0 = blockIdx.x * kGroups + g;
        load_row16(t0, load_index(t0), row);
        j_next = load_index(t0 + tile_step);
    }
    // lanes of this thread's neighbourhood inside its warp (K = 32: the whole warp; K = 16: its half)
    const unsigned nb_mask = p.K == 32 ? FULL : (lane < 16 ? 0x0000FFFFu : 0xFFFF0000u);
    const int nb_in_tile = p.K == 32 ? (warp & 3) : 2 * (warp & 3) + (lane >> 4);        // neighbourhood index inside the tile
    const int lane_in_nb = lane & (p.K - 1);
    float* stage = reinterpret_cast<float*>(smem + lay.stage + g * lay.stage_stride);     // [npad_last channels][gpt]
    const bool vec_out = (p.S & 3) == 0;

    for (int tile = blockIdx.x * kGroups + g; tile < total_tiles; tile += tile_step) {
        const int b = cloud_of(tile);
        const int s0 = (tile - b * lay.tiles_per_cloud) * lay.gpt;
        const int groups = min(lay.gpt, p.S - s0);

        store_parts16(row, xbuf, lay.act_part, 0, r);
        load_row16(tile + tile_step, j_next, row);
        j_next = load_index(tile + 2 * tile_step);

        for (int l = 0; l < p.L; ++l) {
            const int kp = lay.kpad[l], np = lay.npad[l];
            const bool last = (l == p.L - 1);
            proxy_fence_async();
            tc_fence_before();
            group_sync_n(g, kRows);
            if ((warp & 3) == 0) {
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t idesc = make_idesc_bf16(np);
                    const uint32_t lbo_w = static_cast<uint32_t>(np) * 16;
                    const uint32_t wb = sbase + lay.w[l];
                    for (int ks = 0; ks < kp / 16; ++ks) {
                        uint64_t dx[3], dw[3];
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            dx[i] = make_smem_desc(xbuf + i * lay.act_part + ks * 2 * kChunk, kChunk);
                            dw[i] = make_smem_desc(wb + i * lay.w_part[l] + ks * 2 * lbo_w, lbo_w);
                        }
                        mma_bf16_ss(tmem_base, dx[0], dw[0], idesc, ks > 0);
                        mma_bf16_ss(tmem_base, dx[0], dw[1], idesc, true);
                        mma_bf16_ss(tmem_base, dx[1], dw[0], idesc, true);
                        mma_bf16_ss(tmem_base, dx[1], dw[1], idesc, true);
                        mma_bf16_ss(tmem_base, dx[0], dw[2], idesc, true);
                        mma_bf16_ss(tmem_base, dx[2], dw[0], idesc, true);
                    }
                    mma_commit(bar);
                }
                __syncwarp();
            }
            mbar_wait_suspend(bar, phase);
            phase ^= 1;
            tc_fence_after();
            const float* bias = reinterpret_cast<const float*>(smem + lay.bias[l]);
            if (!last) {
                for (int c0 = 0; c0 < np; c0 += 16) {
                    uint32_t v[32];
                    tmem_ld16(tmem_row + c0, v);
                    float x[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) x[i] = fmaxf(__uint_as_float(v[i]) + bias[c0 + i], 0.f);
                    store_parts16(x, xbuf, lay.act_part, c0, r);
                }
            } else {
                // max over the neighbourhood's rows (= lanes): post-ReLU values are >= 0, so their bit patterns order like
                // integers and one REDUX.MAX per channel does it; lane (channel mod K) of the neighbourhood stages the result
                for (int c0 = 0; c0 < np; c0 += 16) {
                    uint32_t v[32];
                    tmem_ld16(tmem_row + c0, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int bits = __float_as_int(fmaxf(__uint_as_float(v[i]) + bias[c0 + i], 0.f));
                        const int m = __reduce_max_sync(nb_mask, bits);
                        if (lane_in_nb == ((c0 + i) & (p.K - 1))) stage[(c0 + i) * lay.gpt + nb_in_tile] = __int_as_float(m);
                    }
                }
                tc_fence_before();
                group_sync_n(g, kRows);
                // channel-first output rows: thread (c, q) writes neighbourhoods 4q..4q+3 of channel c (16 bytes)
                const int quads = lay.gpt >> 2;
                if (r < cout_last * quads) {
                    const int c = r / quads, q4 = (r - c * quads) * 4;
                    float* ob = p.out + (static_cast<size_t>(b) * p.out_c_total + p.out_c_offset + c) * p.S + s0 + q4;
                    const float4 o = *reinterpret_cast<const float4*>(stage + c * lay.gpt + q4);
                    if (vec_out && q4 + 4 <= groups) {
                        *reinterpret_cast<float4*>(ob) = o;
                    } else {
                        if (q4 < groups) ob[0] = o.x;
                        if (q4 + 1 < groups) ob[1] = o.y;
                        if (q4 + 2 < groups) ob[2] = o.z;
                        if (q4 + 3 < groups) ob[3] = o.w;
                    }
                }
            }
            tc_fence_before();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(64 * kGroups) : "memory");
    }
}

bool make_lay8(const SaParams& p, Lay8& lay, int groups)
{
    lay.groups = groups;
    if (p.L < 1 || p.L > kSaMaxLayers) return false;
    if (!(p.K == 16 || p.K == 32)) return false;
    if (p.ch[0] > 16) return false;
    uint32_t off = 0;
    auto take = [&off](uint32_t bytes, uint32_t align) {
        off = (off + align - 1) / align * align;
        const uint32_t o = off;
        off += bytes;
        return o;
    };
    int kmax = 16;
    for (int l = 0; l < p.L; ++l) {
        lay.npad[l] = (p.ch[l + 1] + 15) / 16 * 16;
        lay.kpad[l] = l == 0 ? 16 : lay.npad[l - 1];
        if (lay.npad[l] > 64) return false;
        kmax = std::max(kmax, lay.kpad[l]);
    }
    lay.gpt = kRows / p.K;
    lay.act_part = static_cast<uint32_t>(kmax / 8) * kChunk;
    lay.act_stride = 3 * lay.act_part;
    lay.act = take(lay.act_stride * groups, 1024);
    lay.stage_stride = static_cast<uint32_t>(lay.npad[p.L - 1]) * lay.gpt * 4;
    lay.stage = take(lay.stage_stride * groups, 16);
    for (int l = 0; l < p.L; ++l) {
        lay.w_part[l] = static_cast<uint32_t>(lay.npad[l]) * lay.kpad[l] * 2;
        lay.w[l] = take(3 * lay.w_part[l], 128);
        lay.bias[l] = take(static_cast<uint32_t>(lay.npad[l]) * 4, 16);
    }
    lay.misc = take(8 + 8 * groups, 16);
    lay.total = off;
    lay.tiles_per_cloud = (p.S + lay.gpt - 1) / lay.gpt;
    lay.tpc_magic = static_cast<uint32_t>(std::min<unsigned long long>((1ull << 32) / static_cast<unsigned long long>(lay.tiles_per_cloud), 0xFFFFFFFFull));
    return lay.total <= 226 * 1024;
}

template <int kGroups>
int launch8(const SaParams& p, const Lay8& lay, cudaStream_t st)
{
    const int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(sa_mlp_tc8_kernel<kGroups>), lay.total);
    if (rc_attr != TGN_OK) return rc_attr;
    const long long tiles = static_cast<long long>(lay.tiles_per_cloud) * p.B;
    const int grid = static_cast<int>(std::min<long long>((tiles + kGroups - 1) / kGroups, sm_count()));
    sa_mlp_tc8_kernel<kGroups><<<grid, kRows * kGroups, lay.total, st>>>(p, lay);
    return check_launch("sa_mlp_tc8_kernel");
}

}  // namespace

bool sa_mlp_tc8_supported(const SaParams& p)
{
    Lay8 lay{};
    return make_lay8(p, lay, 8) || make_lay8(p, lay, 4);
}

int sa_mlp_tc8_launch(SaParams p, cudaStream_t st)
{
    Lay8 lay{};
    if (make_lay8(p, lay, 8)) return launch8<8>(p, lay, st);
    if (make_lay8(p, lay, 4)) return launch8<4>(p, lay, st);
    set_error("sa_group_mlp_max: shape not supported by the eight-group tcgen05 engine");
    return TGN_ERR_INVALID;
}

}  // namespace tgn
