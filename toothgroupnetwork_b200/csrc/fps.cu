// fps.cu -- farthest point sampling for sm_100a.
//
// Replaces pointops/src/sampling/sampling_cuda_kernel.cu:14-171 of the reference (one CTA per
// cloud re-streaming xyz+tmp from L2 every iteration, 11 block barriers per iteration).
//
// Design (DESIGN.md "FPS"):
//  * A cloud is owned by a thread-block CLUSTER of CS CTAs.  Every point (x,y,z and its running
//    minimum t) lives in REGISTERS for the whole kernel: thread u holds SLOTS points, packed two
//    per 64-bit register so the distance update runs on FADD2/FMUL2/FFMA2.  HBM is touched once
//    to load the cloud and once per sample to write idx.
//  * Per iteration each WARP reduces its maximum with two CREDUX ops, the winning lane finds its
//    slot, and publishes one candidate (t, j, x, y, z) straight into the mailbox of every CTA of
//    the cluster with st.async, which also completes bytes on that CTA's mbarrier.  There is no
//    __syncthreads in the loop: a warp continues as soon as the mailbox phase completes.
//  * Bit-exact tie-break.  The reference picks, among equal maxima, the point with the smallest
//    (bitrev(j mod BS), j div BS) where BS is ITS block size (cuda_utils.h:11-14) and j the
//    cloud-local index: thread tid scans j = tid, tid+BS, .. with a strict '>' (first maximum
//    wins) and the shared-memory tree lets the lower entry win ties, which orders threads by the
//    bit-reversed tid (sampling_cuda_kernel.cu:5-10,49-59,64-123).  Here thread u owns residue
//    r = u mod BS and the contiguous slot range [q*SLOTS, (q+1)*SLOTS) with q = u div BS, so the
//    order is (bitrev(r), q, slot) and a thread-level priority known without scanning suffices
//    to elect the one lane that has to look at its slots.
//  * Distance arithmetic is the reference's SASS sequence: t = dy*dy; t = fma(dx,dx,t);
//    d = fma(dz,dz,t); min; all IEEE-rn, so indices are bit-identical.
#include <algorithm>
#include <climits>
#include <cmath>

#include "common.cuh"
#include "tgn_b200.h"

namespace tgn {
namespace {

constexpr int kCandWords = 8;      // one mailbox entry = 32 bytes (20 used)
constexpr int kCandBytes = 20;     // bytes completed on the mbarrier per candidate

__device__ __forceinline__ int bitrev_low(int v, int bits) {
    return bits ? static_cast<int>(__brev(static_cast<unsigned>(v)) >> (32 - bits)) : 0;
}
// Total order of points under the reference's tie-break (smaller wins).
__device__ __forceinline__ int point_key(int j, int bs_log2) {
    return (bitrev_low(j & ((1 << bs_log2) - 1), bs_log2) << 21) | (j >> bs_log2);
}

template <int T, int SLOTS, int CS>
__global__ void __launch_bounds__(T, 1)
fps_resident_kernel(const float* __restrict__ xyz, const int* __restrict__ offset,
                    const int* __restrict__ new_offset, float* tmp, int* __restrict__ idx, int bs_log2)
{
    static_assert(SLOTS % 2 == 0, "slots are processed in packed pairs");
    constexpr int NW = T / 32;
    constexpr int NCAND = CS * NW;
    constexpr int PAIRS = SLOTS / 2;
    constexpr int NG = (PAIRS % 4 == 0) ? 4 : ((PAIRS % 2 == 0) ? 2 : 1);   // max accumulators
    constexpr int PPG = PAIRS / NG;
    constexpr unsigned FULL = 0xffffffffu;

    __shared__ __align__(16) uint32_t mailbox[2][NCAND * kCandWords];
    __shared__ __align__(8) uint64_t bars[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (CS > 1) ? static_cast<int>(cluster_ctarank()) : 0;
    const int cloud = blockIdx.x / CS;
    const int start_n = cloud ? offset[cloud - 1] : 0;
    const int n = offset[cloud] - start_n;
    const int start_m = cloud ? new_offset[cloud - 1] : 0;
    const int m = new_offset[cloud] - start_m;
    if (m <= 0 || n <= 0) return;                       // uniform across the cluster
    if (rank == 0 && tid == 0) idx[start_m] = start_n;  // sampling_cuda_kernel.cu:39
    if (m == 1) return;

    const uint32_t bar_base = smem_u32(&bars[0]);   // bars[p] lives at bar_base + 8*p
    if (tid == 0) {
        mbar_init(bar_base, CS > 1 ? 1 : NW);
        mbar_init(bar_base + 8, CS > 1 ? 1 : NW);
        mbar_fence_init();
        if (CS > 1) {
            mbar_arrive_expect_tx(bar_base, NCAND * kCandBytes);
            mbar_arrive_expect_tx(bar_base + 8, NCAND * kCandBytes);
        }
    }
    if (CS > 1) cluster_sync_all(); else __syncthreads();

    // ---- load this thread's points into registers ------------------------------------------
    const int bs = 1 << bs_log2;
    const int u = rank * T + tid;
    const int r = u & (bs - 1);
    const int q = u >> bs_log2;
    const int tprio = (bitrev_low(r, bs_log2) << 14) | q;
    const float* cxyz = xyz + 3 * static_cast<size_t>(start_n);
    float* ctmp = tmp ? tmp + start_n : nullptr;

    uint64_t X[PAIRS], Y[PAIRS], Z[PAIRS];
    float t[SLOTS];
#pragma unroll
    for (int p = 0; p < PAIRS; ++p) {
        float c[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = (q * SLOTS + 2 * p + h) * bs + r;
            const bool ok = j < n;
            c[h][0] = ok ? __ldg(cxyz + 3 * static_cast<size_t>(j) + 0) : 0.f;
            c[h][1] = ok ? __ldg(cxyz + 3 * static_cast<size_t>(j) + 1) : 0.f;
            c[h][2] = ok ? __ldg(cxyz + 3 * static_cast<size_t>(j) + 2) : 0.f;
            t[2 * p + h] = ok ? (ctmp ? ctmp[j] : 1e10f) : -1.0f;   // pads can never be a maximum
        }
        X[p] = pack2(c[0][0], c[1][0]);
        Y[p] = pack2(c[0][1], c[1][1]);
        Z[p] = pack2(c[0][2], c[1][2]);
    }
    float ox = __ldg(cxyz + 0), oy = __ldg(cxyz + 1), oz = __ldg(cxyz + 2);

    for (int it = 1; it < m; ++it) {
        const int par = it & 1;
        const uint32_t bar = bar_base + 8 * par;
        const uint64_t OX = pack2(ox, ox), OY = pack2(oy, oy), OZ = pack2(oz, oz);
        float acc[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) acc[g] = -1.0f;
#pragma unroll
        for (int p = 0; p < PAIRS; ++p) {
            const uint64_t dx = sub2(X[p], OX), dy = sub2(Y[p], OY), dz = sub2(Z[p], OZ);
            uint64_t d = mul2(dy, dy);
            d = fma2(dx, dx, d);
            d = fma2(dz, dz, d);
            float dl, dh;
            unpack2(d, dl, dh);
            t[2 * p] = fminf(dl, t[2 * p]);
            t[2 * p + 1] = fminf(dh, t[2 * p + 1]);
            acc[p / PPG] = max3(acc[p / PPG], t[2 * p], t[2 * p + 1]);
        }
        float best = acc[0];
#pragma unroll
        for (int g = 1; g < NG; ++g) best = fmaxf(best, acc[g]);

        // ---- warp candidate: max value, then smallest thread priority among the tied lanes ----
        const int bi = __float_as_int(best);            // t >= 0 or -1: signed-int order == float order
        const int wmax = __reduce_max_sync(FULL, bi);
        const int wpri = __reduce_min_sync(FULL, bi == wmax ? tprio : INT_MAX);
        if (bi == wmax && tprio == wpri) {              // exactly one lane
            // Which slot?  The accumulator groups narrow the search to PPG pairs; within the
            // group the lowest matching slot wins (the reference's strict '>' keeps the first).
            int gsel = NG - 1;
#pragma unroll
            for (int g = NG - 2; g >= 0; --g)
                if (acc[g] == best) gsel = g;
            int psel = 0;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (g == gsel) {
                    psel = (g + 1) * PPG - 1;
#pragma unroll
                    for (int p = (g + 1) * PPG - 2; p >= g * PPG; --p)
                        if (t[2 * p] == best || t[2 * p + 1] == best) psel = p;
                }
            }
            uint64_t px = 0, py = 0, pz = 0;
            float tl = 0.f;
            switch (psel) {
#define TGN_FPS_PAIR(P_) case P_: if (P_ < PAIRS) { px = X[P_ < PAIRS ? P_ : 0]; py = Y[P_ < PAIRS ? P_ : 0]; pz = Z[P_ < PAIRS ? P_ : 0]; tl = t[P_ < PAIRS ? 2 * P_ : 0]; } break;
                TGN_FPS_PAIR(0) TGN_FPS_PAIR(1) TGN_FPS_PAIR(2) TGN_FPS_PAIR(3) TGN_FPS_PAIR(4) TGN_FPS_PAIR(5)
                TGN_FPS_PAIR(6) TGN_FPS_PAIR(7) TGN_FPS_PAIR(8) TGN_FPS_PAIR(9) TGN_FPS_PAIR(10) TGN_FPS_PAIR(11)
#undef TGN_FPS_PAIR
                default: break;
            }
            static_assert(PAIRS <= 12, "extend the pair switch");
            const bool low = (tl == best);
            const int sel = 2 * psel + (low ? 0 : 1);
            float xl, xh, yl, yh, zl, zh;
            unpack2(px, xl, xh); unpack2(py, yl, yh); unpack2(pz, zl, zh);
            const float sx = low ? xl : xh, sy = low ? yl : yh, sz = low ? zl : zh;
            const int j = (q * SLOTS + sel) * bs + r;
            const int e = (rank * NW + warp) * kCandWords;
            if (CS > 1) {
                const uint32_t slot = smem_u32(&mailbox[par][e]);
#pragma unroll
                for (int dst = 0; dst < CS; ++dst) {
                    const uint32_t ra = map_to_cta(slot, dst), rb = map_to_cta(bar, dst);
                    st_async_v4(ra, static_cast<uint32_t>(bi), static_cast<uint32_t>(j), __float_as_uint(sx),
                                __float_as_uint(sy), rb);
                    st_async_b32(ra + 16, __float_as_uint(sz), rb);
                }
            } else {
                uint32_t* slot = &mailbox[par][e];
                *reinterpret_cast<uint4*>(slot) = make_uint4(static_cast<uint32_t>(bi), static_cast<uint32_t>(j),
                                                             __float_as_uint(sx), __float_as_uint(sy));
                slot[4] = __float_as_uint(sz);
                mbar_arrive(bar);             // release.cta: the stores above are visible to waiters
            }
        }

        // ---- wait for every warp of the cluster, then pick the global winner -------------------
        mbar_wait(bar, ((it - 1) >> 1) & 1);
        int cval = INT_MIN, ckey = INT_MAX, cent = 0;
#pragma unroll
        for (int e0 = 0; e0 < NCAND; e0 += 32) {
            const int e = e0 + lane;
            if (NCAND % 32 == 0 || e < NCAND) {
                const uint2 vj = *reinterpret_cast<const uint2*>(&mailbox[par][e * kCandWords]);
                const int v = static_cast<int>(vj.x), k = point_key(static_cast<int>(vj.y), bs_log2);
                if (v > cval || (v == cval && k < ckey)) { cval = v; ckey = k; cent = e; }
            }
        }
        const int gmax = __reduce_max_sync(FULL, cval);
        const int gkey = __reduce_min_sync(FULL, cval == gmax ? ckey : INT_MAX);
        const int src = __ffs(__ballot_sync(FULL, cval == gmax && ckey == gkey)) - 1;
        const uint4 w0 = *reinterpret_cast<const uint4*>(&mailbox[par][cent * kCandWords]);
        const uint32_t w1 = mailbox[par][cent * kCandWords + 4];
        const int jstar = __shfl_sync(FULL, static_cast<int>(w0.y), src);
        ox = __uint_as_float(__shfl_sync(FULL, w0.z, src));
        oy = __uint_as_float(__shfl_sync(FULL, w0.w, src));
        oz = __uint_as_float(__shfl_sync(FULL, w1, src));
        if (tid == 0) {
            if (rank == 0) idx[start_m + it] = start_n + jstar;
            // Re-arm this parity's barrier for iteration it+2.  Nobody can complete bytes on that
            // phase before receiving this CTA's candidates of iteration it+1, which are sent later.
            if (CS > 1 && it + 2 < m) mbar_arrive_expect_tx(bar, NCAND * kCandBytes);
        }
    }

    if (ctmp) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int j = (q * SLOTS + s) * bs + r;
            if (j < n) ctmp[j] = t[s];
        }
    }
    if (CS > 1) cluster_sync_all();   // no CTA leaves while a peer may still address its shared memory
}

// --------------------------------------------------------------------------------------------
// Streaming kernel for clouds that do not fit the register-resident kernels (n > 8*12288) or
// when a cluster cannot be scheduled: one CTA of 1024 threads per cloud, xyz and tmp re-read
// from L2 each iteration as in the reference, but with one barrier per iteration instead of
// eleven and the same exact tie-break.  tmp is required.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
fps_stream_kernel(const float* __restrict__ xyz, const int* __restrict__ offset, const int* __restrict__ new_offset,
                  float* __restrict__ tmp, int* __restrict__ idx, int bs_log2)
{
    constexpr int T = 1024, NW = T / 32;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ int2 wbuf[2][NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x;
    const int start_n = cloud ? offset[cloud - 1] : 0;
    const int n = offset[cloud] - start_n;
    const int start_m = cloud ? new_offset[cloud - 1] : 0;
    const int m = new_offset[cloud] - start_m;
    if (m <= 0 || n <= 0) return;
    if (tid == 0) idx[start_m] = start_n;
    const int bs = 1 << bs_log2;
    const int r = tid & (bs - 1), q = tid >> bs_log2, chunks = T >> bs_log2;
    const int ns = (n + bs - 1) >> bs_log2;             // slots per residue
    const int spc = (ns + chunks - 1) / chunks;         // slots per chunk
    const int tprio = (bitrev_low(r, bs_log2) << 14) | q;
    const float* cxyz = xyz + 3 * static_cast<size_t>(start_n);
    float* ctmp = tmp + start_n;
    int old = 0;
    for (int it = 1; it < m; ++it) {
        const float ox = cxyz[3 * static_cast<size_t>(old)], oy = cxyz[3 * static_cast<size_t>(old) + 1],
                    oz = cxyz[3 * static_cast<size_t>(old) + 2];
        float best = -1.0f;
        int bj = 0;
        const int s_end = min((q + 1) * spc, ns);
        for (int sg = q * spc; sg < s_end; ++sg) {
            const int j = sg * bs + r;
            if (j < n) {
                const float dx = cxyz[3 * static_cast<size_t>(j)] - ox, dy = cxyz[3 * static_cast<size_t>(j) + 1] - oy,
                            dz = cxyz[3 * static_cast<size_t>(j) + 2] - oz;
                float d = __fmul_rn(dy, dy);
                d = __fmaf_rn(dx, dx, d);
                d = __fmaf_rn(dz, dz, d);
                const float v = fminf(d, ctmp[j]);
                ctmp[j] = v;
                if (v > best) { best = v; bj = j; }
            }
        }
        const int bi = __float_as_int(best);
        const int wmax = __reduce_max_sync(FULL, bi);
        const int wpri = __reduce_min_sync(FULL, bi == wmax ? tprio : INT_MAX);
        if (bi == wmax && tprio == wpri) wbuf[it & 1][warp] = make_int2(bi, bj);
        __syncthreads();
        const int2 c = wbuf[it & 1][lane];              // NW == 32: one candidate per lane
        const int key = point_key(c.y, bs_log2);
        const int gmax = __reduce_max_sync(FULL, c.x);
        const int gkey = __reduce_min_sync(FULL, c.x == gmax ? key : INT_MAX);
        const int src = __ffs(__ballot_sync(FULL, c.x == gmax && key == gkey)) - 1;
        old = __shfl_sync(FULL, c.y, src);
        if (tid == 0) idx[start_m + it] = start_n + old;
    }
}

// cuda_utils.h:11-14 of the reference, literally (double log, truncation).
int ref_block_log2(int n)
{
    const int p = static_cast<int>(std::log(static_cast<double>(n)) / std::log(2.0));
    int bs = std::max(std::min(1 << p, 1024), 1);
    int l = 0;
    while ((1 << l) < bs) ++l;
    return l;
}

template <int T, int SLOTS, int CS>
int launch_resident(int b, const float* xyz, const int* offset, const int* new_offset, float* tmp, int* idx,
                    int bs_log2, cudaStream_t stream)
{
    auto kern = fps_resident_kernel<T, SLOTS, CS>;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(b) * CS);
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, xyz, offset, new_offset, tmp, idx, bs_log2);
    if (e != cudaSuccess) {
        set_error("fps_resident_kernel<%d,%d,%d> launch failed: %s", T, SLOTS, CS, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return TGN_ERR_CUDA;
    }
    return check_launch("fps_resident_kernel");
}

struct FpsConfig { int T, SLOTS, CS; };

// Smallest resident configuration that holds n_max points at cluster size cs (or {0,0,0}).
FpsConfig pick_config(int n_max, int bs_log2, int cs)
{
    static const FpsConfig table[] = {
        {128, 4, 1}, {256, 4, 1}, {512, 4, 1}, {1024, 4, 1}, {1024, 8, 1},
        {512, 4, 2}, {512, 8, 2}, {512, 12, 2}, {512, 16, 2}, {512, 24, 2},
        {512, 4, 4}, {512, 8, 4}, {512, 12, 4}, {512, 16, 4}, {512, 24, 4},
        {512, 4, 8}, {512, 8, 8}, {512, 12, 8}, {512, 16, 8}, {512, 24, 8},
    };
    const int bs = 1 << bs_log2;
    for (const FpsConfig& c : table) {
        if (c.CS != cs) continue;
        const int v = c.T * c.CS;
        if (v < bs) continue;
        const long long cap_slots = static_cast<long long>(v / bs) * c.SLOTS;     // slots per residue
        if (cap_slots * bs >= n_max && cap_slots >= (n_max + bs - 1) / bs) return c;
    }
    return {0, 0, 0};
}

#define TGN_FPS_CASE(T_, S_, C_)                                                                      \
    if (c.T == T_ && c.SLOTS == S_ && c.CS == C_)                                                      \
        return launch_resident<T_, S_, C_>(b, xyz, offset, new_offset, tmp, idx, bs_log2, stream);

int dispatch_resident(const FpsConfig& c, int b, const float* xyz, const int* offset, const int* new_offset,
                      float* tmp, int* idx, int bs_log2, cudaStream_t stream)
{
    TGN_FPS_CASE(128, 4, 1) TGN_FPS_CASE(256, 4, 1) TGN_FPS_CASE(512, 4, 1) TGN_FPS_CASE(1024, 4, 1)
    TGN_FPS_CASE(1024, 8, 1)
    TGN_FPS_CASE(512, 4, 2) TGN_FPS_CASE(512, 8, 2) TGN_FPS_CASE(512, 12, 2) TGN_FPS_CASE(512, 16, 2)
    TGN_FPS_CASE(512, 24, 2)
    TGN_FPS_CASE(512, 4, 4) TGN_FPS_CASE(512, 8, 4) TGN_FPS_CASE(512, 12, 4) TGN_FPS_CASE(512, 16, 4)
    TGN_FPS_CASE(512, 24, 4)
    TGN_FPS_CASE(512, 4, 8) TGN_FPS_CASE(512, 8, 8) TGN_FPS_CASE(512, 12, 8) TGN_FPS_CASE(512, 16, 8)
    TGN_FPS_CASE(512, 24, 8)
    set_error("no resident FPS kernel for T=%d SLOTS=%d CS=%d", c.T, c.SLOTS, c.CS);
    return TGN_ERR_INVALID;
}

}  // namespace

int fps_dispatch(int b, int n_max, const float* xyz, const int* offset, const int* new_offset, float* tmp, int* idx,
                 int mode, cudaStream_t stream)
{
    if (b <= 0) return TGN_OK;
    if (n_max <= 0) { set_error("furthestsampling: n_max must be positive"); return TGN_ERR_INVALID; }
    const int bs_log2 = ref_block_log2(n_max);
    if (mode == -1) {
        if (!tmp) { set_error("furthestsampling: the streaming kernel needs a tmp buffer"); return TGN_ERR_INVALID; }
        fps_stream_kernel<<<b, 1024, 0, stream>>>(xyz, offset, new_offset, tmp, idx, bs_log2);
        return check_launch("fps_stream_kernel");
    }
    FpsConfig cfg{0, 0, 0};
    if (mode > 0) {
        cfg = pick_config(n_max, bs_log2, mode);
    } else {
        // Throughput when the batch can fill the machine with the smallest feasible cluster,
        // latency (wider clusters) when only a few clouds are in flight.
        const int sms = sm_count();
        int first = 0;
        for (int cs = 1; cs <= 8; cs *= 2)
            if (pick_config(n_max, bs_log2, cs).T) { first = cs; break; }
        if (first) {
            int cs = first;
            while (cs < 8 && static_cast<long long>(b) * cs * 2 <= sms && n_max / (cs * 2) >= 1024 &&
                   pick_config(n_max, bs_log2, cs * 2).T)
                cs *= 2;
            cfg = pick_config(n_max, bs_log2, cs);
        }
    }
    if (!cfg.T) {
        if (!tmp) {
            set_error("furthestsampling: n_max=%d exceeds the register-resident kernels and no tmp buffer was given", n_max);
            return TGN_ERR_INVALID;
        }
        fps_stream_kernel<<<b, 1024, 0, stream>>>(xyz, offset, new_offset, tmp, idx, bs_log2);
        return check_launch("fps_stream_kernel");
    }
    return dispatch_resident(cfg, b, xyz, offset, new_offset, tmp, idx, bs_log2, stream);
}

}  // namespace tgn

extern "C" {

int tgn_furthestsampling(int b, int n_max, const float* xyz, const int* offset, const int* new_offset, float* tmp,
                         int* idx, int mode, void* stream)
{
    return tgn::fps_dispatch(b, n_max, xyz, offset, new_offset, tmp, idx, mode, static_cast<cudaStream_t>(stream));
}

void furthestsampling_cuda_launcher(int b, int n, const float* xyz, const int* offset, const int* new_offset, float* tmp,
                                    int* idx)
{
    (void)tgn::fps_dispatch(b, n, xyz, offset, new_offset, tmp, idx, 0, static_cast<cudaStream_t>(0));
}

}  // extern "C"
