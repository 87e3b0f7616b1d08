// sa_mlp_tcw.cu -- tcgen05 engine of the fused set-abstraction body for WIDE hidden layers (up to 128).
//
// The 3xTF32 engine in sa_mlp_tc.cu keeps every inter-layer operand as two tf32 tiles; at 128 channels
// that is 128 KB per 128-row tile and nothing fits.  The reference's real pointnet_pp SA1
// (PointNetSetAbstractionMsg(1024,[.025,.05],[32,64],6,[[128,128],[128,128]]), pointnet_pp.py:13) is
// exactly that shape, 15x the FLOPs of the BASELINE form.  This engine keeps the same structure (tile
// groups of 128 threads, thread = row, accumulators in TMEM, last layer transposed so that the max over
// the K neighbours stays inside one thread) with two changes:
//   * layer 0 (the raw [xyz_rel | feats] row, <= 16 channels) stays 3xTF32;
//   * every later layer splits its operands into TWO BF16 parts, x = x1 + x2 with x1 = bf16(x) and
//     x2 = bf16(x - x1) (16 mantissa bits kept), and issues three kind::f16 MMAs per K-step of 16
//     (x1*w1 + x2*w1 + x1*w2) into the fp32 accumulator: ~2^-17 relative error per product -- 60x
//     tighter than the TF32 the reference's own convolutions run in by default, inside
//     the 1e-4 bar (measured ~1e-5 end to end), at half the operand bytes and half the MMA count of 3xTF32;
//   * biases are added in the epilogue in fp32 (exact), not through an extra MMA.
// Two tile groups per CTA at 128 channels (operand tile 64 KB each, 80 KB of weights), four when the layers
// are narrow enough; persistent over tiles.  On narrow shapes the 3xTF32 engine is faster (bench shape
// 9->[32,32,64]: 3.8 ms against 5.3 ms here) and more accurate, so the dispatcher prefers it when it fits.
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
constexpr int kMaxGroups = 4;          // tile groups per CTA: 4 when the operand tiles are small enough, else 2
constexpr uint32_t kChunk = kRows * 16;               // bytes of one 16-byte K-chunk of all 128 rows (LBO of a 128-row operand)

struct WLayout {
    int kpad[kSaMaxLayers];          // K of layer l (layer 0: multiple of 8, 16 at most; later: multiple of 16)
    int npad[kSaMaxLayers];          // N of layer l, multiple of 16
    uint32_t w_hi[kSaMaxLayers], w_lo[kSaMaxLayers];   // byte offsets of the weight operands
    uint32_t act[kMaxGroups];        // per group operand buffer: [hi | lo], lo at act_lo_off
    int groups;
    uint32_t act_lo_off;
    uint32_t misc, total;
    int tiles_per_cloud, gpt;
    uint32_t tpc_magic;
};

// (a, b) -> packed bf16 pair, a in the LOW half (the lower K index), round to nearest even
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// x = x1 + x2 (two bf16 parts each) for a pair of values: hi word, lo word
__device__ __forceinline__ void split_bf16_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16(a, b);
    const float a1 = __uint_as_float(hi << 16), b1 = __uint_as_float(hi & 0xFFFF0000u);
    lo = pack_bf16(__fsub_rn(a, a1), __fsub_rn(b, b1));
}
// Inner-layer epilogue of one thread (row r): NP accumulator columns -> + bias, ReLU, bf16 hi/lo split,
// 16-byte stores (8 K-elements each) into the next layer's K-major operand.
template <int NP>
__device__ __forceinline__ void inner_epilogue(uint32_t tmem_row, uint32_t x_hi, uint32_t x_lo, int r, const float* __restrict__ bias,
                                               int cout)
{
#pragma unroll
    for (int c0 = 0; c0 < NP; c0 += 32) {
        uint32_t v[32];
        const bool full = NP - c0 >= 32;
        if (full) tmem_ld32(tmem_row + c0, v);
        else tmem_ld16(tmem_row + c0, v);
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            if (i < 16 || full) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = c0 + i + 2 * u;
                    const float b0 = c < cout ? __ldg(bias + c) : 0.f, b1 = c + 1 < cout ? __ldg(bias + c + 1) : 0.f;
                    const float a = fmaxf(__uint_as_float(v[i + 2 * u]) + b0, 0.f);
                    const float b = fmaxf(__uint_as_float(v[i + 2 * u + 1]) + b1, 0.f);
                    split_bf16_pair(a, b, hi[u], lo[u]);
                }
                const uint32_t off = static_cast<uint32_t>((c0 + i) >> 3) * kChunk + r * 16;
                st_shared_v4(x_hi + off, hi[0], hi[1], hi[2], hi[3]);
                st_shared_v4(x_lo + off, lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }
}

// max(v[B], ..., v[B+15]) with 3-input FMNMX (no ReLU here: the bias is added after the max)
template <int B>
__device__ __forceinline__ float max16(const uint32_t (&v)[32]) {
    float m = fmaxf(__uint_as_float(v[B]), __uint_as_float(v[B + 1]));
#pragma unroll
    for (int i = 2; i < 16; i += 2) m = max3(m, __uint_as_float(v[B + i]), __uint_as_float(v[B + i + 1]));
    return m;
}

template <int kGroups>
__global__ void __launch_bounds__(kRows * kGroups, 1)
sa_mlp_tcw_kernel(const SaParams p, const WLayout lay)
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
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + lay.misc), "r"(kRows * kGroups) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (r == 0) { mbar_init(bar, 1); mbar_fence_init(); }

    // ---- weights, once per CTA -----------------------------------------------------------------------
    // layer 0: tf32 hi/lo B operand [np x 16] (LBO = np*16); inner layers: bf16 hi/lo B operand [np x kp]
    // (LBO = np*16); last layer: bf16 hi/lo A operand [128 x kp] (LBO = 128*16), zero rows beyond cout.
    for (int l = 0; l < p.L; ++l) {
        const int cin = p.ch[l], cout = p.ch[l + 1], kp = lay.kpad[l];
        const int rows_w = l == p.L - 1 ? kRows : lay.npad[l];
        for (int e = tid; e < rows_w * kp; e += kThreads) {
            const int n = e / kp, k = e - n * kp;
            const float w = (n < cout && k < cin) ? __ldg(p.W[l] + static_cast<size_t>(n) * cin + k) : 0.f;
            if (l == 0) {
                uint32_t hi, lo;
                split_tf32(w, hi, lo);
                const uint32_t off = static_cast<uint32_t>(k >> 2) * (rows_w * 16) + n * 16 + (k & 3) * 4;
                *reinterpret_cast<uint32_t*>(smem + lay.w_hi[l] + off) = hi;
                *reinterpret_cast<uint32_t*>(smem + lay.w_lo[l] + off) = lo;
            } else {
                const __nv_bfloat16 w1 = __float2bfloat16_rn(w);
                const __nv_bfloat16 w2 = __float2bfloat16_rn(w - __bfloat162float(w1));
                const uint32_t off = static_cast<uint32_t>(k >> 3) * (rows_w * 16) + n * 16 + (k & 7) * 2;
                *reinterpret_cast<__nv_bfloat16*>(smem + lay.w_hi[l] + off) = w1;
                *reinterpret_cast<__nv_bfloat16*>(smem + lay.w_lo[l] + off) = w2;
            }
        }
    }
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot + g * kRows;
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t x_hi = sbase + lay.act[g], x_lo = x_hi + lay.act_lo_off;
    const uint32_t a0_lo = x_hi + 4 * kChunk;       // layer 0: four tf32 chunks of hi, then four of lo
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

    for (int tile = blockIdx.x * kGroups + g; tile < total_tiles; tile += tile_step) {
        const int b = cloud_of(tile);
        const int s0 = (tile - b * lay.tiles_per_cloud) * lay.gpt;
        const int groups = min(lay.gpt, p.S - s0);
        const int rows = groups * p.K;

        // ---- layer-0 operand (tf32 hi/lo) from the row prefetched one tile ago ---------------------------------
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) split_tf32(row[4 * kc + i], hi[i], lo[i]);
            const uint32_t off = kc * kChunk + r * 16;
            st_shared_v4(x_hi + off, hi[0], hi[1], hi[2], hi[3]);
            st_shared_v4(a0_lo + off, lo[0], lo[1], lo[2], lo[3]);
        }
        load_row16(tile + tile_step, j_next, row);
        j_next = load_index(tile + 2 * tile_step);

        for (int l = 0; l < p.L; ++l) {
            const int kp = lay.kpad[l], np = lay.npad[l];
            const bool last = (l == p.L - 1);
            proxy_fence_async();
            tc_fence_before();
            group_sync(g);
            if ((warp & 3) == 0) {
                tc_fence_after();
                const uint64_t step = (2 * kChunk) >> 4;         // two 16-byte K-chunks of a 128-row operand per MMA
                if (l == 0) {
                    // D[row, n] += X0[row, k] * W0[n, k], 3xTF32, K = 8 per MMA
                    const uint32_t idesc = make_idesc_tf32(np);
                    const uint32_t lbo_w = static_cast<uint32_t>(np) * 16;
                    uint64_t d_xh = make_smem_desc(x_hi, kChunk), d_xl = make_smem_desc(a0_lo, kChunk);
                    uint64_t d_wh = make_smem_desc(sbase + lay.w_hi[0], lbo_w), d_wl = make_smem_desc(sbase + lay.w_lo[0], lbo_w);
                    const uint64_t step_w = (2 * lbo_w) >> 4;
                    for (int ks = 0; ks < kp / 8; ++ks) {
                        if (lane == 0) {
                            mma_tf32_ss(tmem_base, d_xh, d_wh, idesc, ks > 0);
                            mma_tf32_ss(tmem_base, d_xl, d_wh, idesc, true);
                            mma_tf32_ss(tmem_base, d_xh, d_wl, idesc, true);
                        }
                        d_xh += step; d_xl += step; d_wh += step_w; d_wl += step_w;
                    }
                } else if (!last) {
                    // D[row, n] += X[row, k] * W[n, k], three bf16 terms, K = 16 per MMA
                    const uint32_t idesc = make_idesc_bf16(np);
                    const uint32_t lbo_w = static_cast<uint32_t>(np) * 16;
                    uint64_t d_xh = make_smem_desc(x_hi, kChunk), d_xl = make_smem_desc(x_lo, kChunk);
                    uint64_t d_wh = make_smem_desc(sbase + lay.w_hi[l], lbo_w), d_wl = make_smem_desc(sbase + lay.w_lo[l], lbo_w);
                    const uint64_t step_w = (2 * lbo_w) >> 4;
                    for (int ks = 0; ks < kp / 16; ++ks) {
                        if (lane == 0) {
                            mma_bf16_ss(tmem_base, d_xh, d_wh, idesc, ks > 0);
                            mma_bf16_ss(tmem_base, d_xl, d_wh, idesc, true);
                            mma_bf16_ss(tmem_base, d_xh, d_wl, idesc, true);
                        }
                        d_xh += step; d_xl += step; d_wh += step_w; d_wl += step_w;
                    }
                } else {
                    // D^T[channel, row] += W[channel, k] * X[row, k], three bf16 terms
                    const uint32_t idesc = make_idesc_bf16(kRows);
                    uint64_t d_xh = make_smem_desc(x_hi, kChunk), d_xl = make_smem_desc(x_lo, kChunk);
                    uint64_t d_wh = make_smem_desc(sbase + lay.w_hi[l], kChunk), d_wl = make_smem_desc(sbase + lay.w_lo[l], kChunk);
                    for (int ks = 0; ks < kp / 16; ++ks) {
                        if (lane == 0) {
                            mma_bf16_ss(tmem_base, d_wh, d_xh, idesc, ks > 0);
                            mma_bf16_ss(tmem_base, d_wh, d_xl, idesc, true);
                            mma_bf16_ss(tmem_base, d_wl, d_xh, idesc, true);
                        }
                        d_xh += step; d_xl += step; d_wh += step; d_wl += step;
                    }
                }
                if (lane == 0) mma_commit(bar);
                __syncwarp();
            }
            mbar_wait_suspend(bar, phase);
            phase ^= 1;
            tc_fence_after();

            if (!last) {
                switch (np) {
                    case 16: inner_epilogue<16>(tmem_row, x_hi, x_lo, r, p.bias[l], p.ch[l + 1]); break;
                    case 32: inner_epilogue<32>(tmem_row, x_hi, x_lo, r, p.bias[l], p.ch[l + 1]); break;
                    case 48: inner_epilogue<48>(tmem_row, x_hi, x_lo, r, p.bias[l], p.ch[l + 1]); break;
                    case 64: inner_epilogue<64>(tmem_row, x_hi, x_lo, r, p.bias[l], p.ch[l + 1]); break;
                    case 80: inner_epilogue<80>(tmem_row, x_hi, x_lo, r, p.bias[l], p.ch[l + 1]); break;
                    case 96: inner_epilogue<96>(tmem_row, x_hi, x_lo, r, p.bias[l], p.ch[l + 1]); break;
                    case 112: inner_epilogue<112>(tmem_row, x_hi, x_lo, r, p.bias[l], p.ch[l + 1]); break;
                    default: inner_epilogue<128>(tmem_row, x_hi, x_lo, r, p.bias[l], p.ch[l + 1]); break;
                }
            } else if ((warp & 3) * 32 < cout_last) {
                // thread r = output channel; columns = rows of the tile; max over the K columns of a neighbourhood,
                // then + bias and ReLU (both monotone, so they commute with the max)
                const bool store = r < cout_last;
                const float bv = store ? __ldg(p.bias[l] + r) : 0.f;
                float* ob = p.out + (static_cast<size_t>(b) * p.out_c_total + p.out_c_offset + (store ? r : 0)) * p.S + s0;
                float acc = -INFINITY;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (32 * q < rows) {                              // tile-uniform
                        uint32_t v[32];
                        tmem_ld32(tmem_row + 32 * q, v);
                        const float m_lo = max16<0>(v), m_hi = max16<16>(v);
                        if (p.K == 16) {
                            if (store) {
                                ob[2 * q] = fmaxf(m_lo + bv, 0.f);
                                if (2 * q + 1 < groups) ob[2 * q + 1] = fmaxf(m_hi + bv, 0.f);
                            }
                        } else {
                            acc = max3(acc, m_lo, m_hi);
                            const int cpn = p.K >> 5;
                            if (((q + 1) & (cpn - 1)) == 0) {
                                if (store) ob[q / cpn] = fmaxf(acc + bv, 0.f);
                                acc = -INFINITY;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(kRows * kGroups) : "memory");
    }
}

bool make_wlayout_groups(const SaParams& p, WLayout& lay, int kGroups)
{
    lay.groups = kGroups;
    if (p.L < 2 || p.L > kSaMaxLayers) return false;                       // one layer: the other engines
    if (!(p.K == 16 || p.K == 32 || p.K == 64 || p.K == 128)) return false;
    if (p.ch[0] > 16) return false;                                        // the gathered row is built in 16 registers
    uint32_t off = 0;
    auto take = [&off](uint32_t bytes, uint32_t align) {
        off = (off + align - 1) / align * align;
        const uint32_t o = off;
        off += bytes;
        return o;
    };
    int kmax = 0;
    for (int l = 0; l < p.L; ++l) {
        lay.npad[l] = (p.ch[l + 1] + 15) / 16 * 16;
        lay.kpad[l] = l == 0 ? 16 : lay.npad[l - 1];
        if (lay.npad[l] > 128) return false;
        if (l > 0) kmax = std::max(kmax, lay.kpad[l]);
    }
    // operand buffer of a group: bf16 hi then lo, kmax/8 chunks each; layer 0 needs 8 tf32 chunks in the same space
    const uint32_t half = std::max<uint32_t>(static_cast<uint32_t>(kmax / 8) * kChunk, 4 * kChunk);
    lay.act_lo_off = half;
    for (int g = 0; g < kGroups; ++g) lay.act[g] = take(2 * half, 1024);
    for (int l = 0; l < p.L; ++l) {
        const int rows_w = l == p.L - 1 ? kRows : lay.npad[l];
        const uint32_t wb = static_cast<uint32_t>(rows_w) * lay.kpad[l] * (l == 0 ? 4 : 2);
        lay.w_hi[l] = take(wb, 128);
        lay.w_lo[l] = take(wb, 128);
    }
    lay.misc = take(8 + 8 * kGroups, 16);
    lay.total = off;
    lay.gpt = kRows / p.K;
    lay.tiles_per_cloud = (p.S + lay.gpt - 1) / lay.gpt;
    lay.tpc_magic = static_cast<uint32_t>(std::min<unsigned long long>((1ull << 32) / static_cast<unsigned long long>(lay.tiles_per_cloud), 0xFFFFFFFFull));
    return lay.total <= 224 * 1024;
}

// four tile groups when they fit (narrow layers), else two
bool make_wlayout(const SaParams& p, WLayout& lay)
{
    return make_wlayout_groups(p, lay, 4) || make_wlayout_groups(p, lay, 2);
}

template <int kGroups>
int launch_tcw(const SaParams& p, const WLayout& lay, cudaStream_t st)
{
    const int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(sa_mlp_tcw_kernel<kGroups>), lay.total);
    if (rc_attr != TGN_OK) return rc_attr;
    const long long tiles = static_cast<long long>(lay.tiles_per_cloud) * p.B;
    const int grid = static_cast<int>(std::min<long long>((tiles + kGroups - 1) / kGroups, sm_count()));
    sa_mlp_tcw_kernel<kGroups><<<grid, kRows * kGroups, lay.total, st>>>(p, lay);
    return check_launch("sa_mlp_tcw_kernel");
}

}  // namespace

bool sa_mlp_tcw_supported(const SaParams& p)
{
    WLayout lay{};
    return make_wlayout(p, lay);
}

int sa_mlp_tcw_launch(SaParams p, cudaStream_t st)
{
    WLayout lay{};
    if (!make_wlayout(p, lay)) { set_error("sa_group_mlp_max: shape not supported by the wide tcgen05 engine"); return TGN_ERR_INVALID; }
    return lay.groups == 4 ? launch_tcw<4>(p, lay, st) : launch_tcw<2>(p, lay, st);
}

}  // namespace tgn
