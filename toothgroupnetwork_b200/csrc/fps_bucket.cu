// fps_bucket.cu -- bucket-pruned farthest point sampling for sm_100a (bit-exact).
//
// The brute-force FPS update touches every point every iteration although, after a few
// samples, a new sample only lowers the running minimum of the points NEAR it.  This kernel
// keeps the reference's result bit for bit and skips the rest:
//
//  * prologue kernel: the cloud is sorted by a 21-bit Morton code (stable LSD radix sort, three
//    7-bit passes, one CTA per cloud) so that every 32 consecutive points -- a BUCKET, exactly one
//    coalesced 512-byte float4 row -- are spatial neighbours; each bucket gets a bounding sphere.
//  * main kernel (one CTA per cloud, several CTAs per SM): per iteration
//      (a) every bucket is tested against the new sample o:  with D = |c_b - o| and the inflated
//          radius r_b, every point p of the bucket has |p - o| >= D - r_b, so if
//          (D - r_b - slack)^2 * (1 - 1e-5) > max_t(bucket) the update min(t, |p-o|^2) cannot change
//          any t of the bucket (the slack terms dominate every fp32 rounding error involved; see
//          DESIGN.md "pruning is conservative") and the bucket is skipped;
//      (b) one warp per surviving bucket re-evaluates its 32 points with the reference's exact
//          arithmetic (FMUL dy*dy, FFMA dx*dx+., FFMA dz*dz+., FMNMX), writes back changed minima and
//          refreshes the bucket's cached candidate (max t, tie key, coordinates);
//      (c) the block arg-max runs over the cached candidates of ALL buckets (two CREDUX levels).
//    Ties are resolved by the reference's order (bitrev(j mod BS), j div BS) on ORIGINAL indices,
//    carried in the w component of each sorted point, so the result does not depend on the
//    internal order.
//  * the clouds live in L2 (20 B/point), the bucket table in shared memory; an iteration is three
//    block barriers and one L2 round trip, and its latency is hidden by the other CTAs of the SM --
//    a 126 MB L2 holds ~300 clouds of 24k points at once, which is what makes "one CTA per cloud,
//    many CTAs per SM" possible on B200.
//
// Work drops from N point-updates per sample to roughly N*(0.001 + 0.07/sqrt(k) + 1/k) at sample k
// for surface-like clouds (a ~16x reduction over 1024 samples of a 24k cloud).
#include <algorithm>
#include <climits>

#include "common.cuh"
#include "fps.cuh"
#include "tgn_b200.h"

namespace tgn {
namespace {

constexpr int kT = 256;            // sort kernel
constexpr int kNW = kT / 32;
constexpr int kMT = 512;           // main kernel
constexpr int kMNW = kMT / 32;
constexpr int kBatch = 4;          // buckets a warp keeps in flight (memory-level parallelism)
constexpr unsigned FULL = 0xffffffffu;

struct BucketWs {
    float4* pts4;     // [b][stride]   sorted (x, y, z, bits(original local index)); pads have index -1
    float* tval;      // [b][stride]   running minima in sorted order; pads -1
    uint2* key_a;     // [b][stride]   radix ping
    uint2* key_b;     // [b][stride]   radix pong
    float4* bsphere;  // [b][nbmax]    bucket centre + inflated radius (absolute slack included)
    int2* bvk;        // [b][nbmax]    initial cached candidate (value bits, tie key)
    float4* bxyzj;    // [b][nbmax]    initial cached candidate coordinates + original index
    float* scale;     // [b]           max |coordinate| of the cloud
    int stride, nbmax;
};

__device__ __forceinline__ int bitrev_low(int v, int bits) {
    return bits ? static_cast<int>(__brev(static_cast<unsigned>(v)) >> (32 - bits)) : 0;
}
__device__ __forceinline__ int point_key(int j, int bs_log2) {
    return (bitrev_low(j & ((1 << bs_log2) - 1), bs_log2) << 21) | (j >> bs_log2);
}
// Squared skip threshold of a bucket: with r' = inflated radius + absolute slack and M = max t of the
// bucket, the bucket cannot change when D = |c - o| satisfies D > r' + sqrt(M * (1 + 1e-5)); both
// sides are positive, so the test is D^2 > theta2 with theta2 rounded UP by the extra factors.
__device__ __forceinline__ float skip_threshold2(float r_slack, float max_t) {
    if (max_t < 0.f) return 0.f;                       // bucket of pads only (cannot happen)
    const float th = (r_slack + sqrtf(max_t * 1.00002f)) * 1.000002f;
    return th * th * 1.000002f;
}
__device__ __forceinline__ unsigned spread3(unsigned x) {   // 10 bits -> every third bit
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// One stable LSD radix pass on 7 bits.  Warp w owns a contiguous range of the input; ranks inside
// a 32-element step come from __match_any_sync, so the pass is deterministic.
__device__ void radix_pass7(const uint2* __restrict__ src, uint2* __restrict__ dst, int n, int shift,
                            int (*hist)[128], int* dig_base)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per_warp = ((n + kNW - 1) / kNW + 31) & ~31;
    const int lo = warp * per_warp, hi = min(n, lo + per_warp);
    for (int i = tid; i < kNW * 128; i += kT) (&hist[0][0])[i] = 0;
    __syncthreads();
    for (int base = lo; base < hi; base += 32) {
        const int i = base + lane;
        const bool ok = i < hi;
        const int d = ok ? static_cast<int>((src[i].x >> shift) & 127u) : 128 + lane;   // unique dummy digits
        const unsigned m = __match_any_sync(FULL, d);
        if (ok && lane == __ffs(m) - 1) hist[warp][d] += __popc(m);
    }
    __syncthreads();
    // digit-major, warp-minor exclusive scan
    if (tid < 128) {
        int s = 0;
        for (int w = 0; w < kNW; ++w) s += hist[w][tid];
        dig_base[tid] = s;
    }
    __syncthreads();
    if (warp == 0) {                       // exclusive scan of 128 digit totals, 4 per lane
        int v[4], s = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k] = dig_base[lane * 4 + k]; s += v[k]; }
        int inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        int run = inc - s;
#pragma unroll
        for (int k = 0; k < 4; ++k) { dig_base[lane * 4 + k] = run; run += v[k]; }
    }
    __syncthreads();
    if (tid < 128) {
        int run = dig_base[tid];
        for (int w = 0; w < kNW; ++w) { const int c = hist[w][tid]; hist[w][tid] = run; run += c; }
    }
    __syncthreads();
    for (int base = lo; base < hi; base += 32) {
        const int i = base + lane;
        const bool ok = i < hi;
        uint2 e = make_uint2(0u, 0u);
        if (ok) e = src[i];
        const int d = ok ? static_cast<int>((e.x >> shift) & 127u) : 128 + lane;
        const unsigned m = __match_any_sync(FULL, d);
        const int rank = __popc(m & ((1u << lane) - 1u));
        int cur = 0;
        if (ok) cur = hist[warp][d];
        __syncwarp();
        if (ok) {
            dst[cur + rank] = e;
            if (lane == __ffs(m) - 1) hist[warp][d] = cur + __popc(m);
        }
        __syncwarp();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kT)
fps_bucket_sort_kernel(const float* __restrict__ xyz, const int* __restrict__ offset, const float* __restrict__ tmp,
                       BucketWs ws, int bs_log2)
{
    __shared__ float red[6][kNW];
    __shared__ float box[6];
    __shared__ int hist[kNW][128];
    __shared__ int dig_base[128];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x;
    const int start_n = cloud ? offset[cloud - 1] : 0;
    const int n = offset[cloud] - start_n;
    if (n <= 0) return;
    const float* cx = xyz + 3 * static_cast<size_t>(start_n);
    float4* pts4 = ws.pts4 + static_cast<size_t>(cloud) * ws.stride;
    float* tval = ws.tval + static_cast<size_t>(cloud) * ws.stride;
    uint2* ka = ws.key_a + static_cast<size_t>(cloud) * ws.stride;
    uint2* kb = ws.key_b + static_cast<size_t>(cloud) * ws.stride;

    // ---- bounding box --------------------------------------------------------------------------------
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int j = tid; j < n; j += kT) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(cx + 3 * static_cast<size_t>(j) + a);
            mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float lo = warp_min(mn[a]), hi = warp_max(mx[a]);
        if (lane == 0) { red[a][warp] = lo; red[3 + a][warp] = hi; }
    }
    __syncthreads();
    if (tid < 6) {
        float v = red[tid][0];
        for (int w = 1; w < kNW; ++w) v = tid < 3 ? fminf(v, red[tid][w]) : fmaxf(v, red[tid][w]);
        box[tid] = v;
    }
    __syncthreads();
    float inv[3];
    float scale = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float ext = box[3 + a] - box[a];
        inv[a] = ext > 0.f ? 127.999f / ext : 0.f;
        scale = fmaxf(scale, fmaxf(fabsf(box[a]), fabsf(box[3 + a])));
    }
    if (tid == 0) ws.scale[cloud] = scale;

    // ---- Morton keys + stable radix sort ----------------------------------------------------------------
    for (int j = tid; j < n; j += kT) {
        unsigned code = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(cx + 3 * static_cast<size_t>(j) + a);
            const unsigned q = min(127u, static_cast<unsigned>(fmaxf((v - box[a]) * inv[a], 0.f)));
            code |= spread3(q) << a;
        }
        ka[j] = make_uint2(code, static_cast<unsigned>(j));
    }
    __syncthreads();
    radix_pass7(ka, kb, n, 0, hist, dig_base);
    radix_pass7(kb, ka, n, 7, hist, dig_base);
    radix_pass7(ka, kb, n, 14, hist, dig_base);
    const uint2* sorted = kb;

    // ---- sorted points and their running minima -----------------------------------------------------------
    const int npad = (n + 31) & ~31;
    for (int p = tid; p < npad; p += kT) {
        if (p < n) {
            const int j = static_cast<int>(sorted[p].y);
            pts4[p] = make_float4(__ldg(cx + 3 * static_cast<size_t>(j)), __ldg(cx + 3 * static_cast<size_t>(j) + 1),
                                  __ldg(cx + 3 * static_cast<size_t>(j) + 2), __int_as_float(j));
            tval[p] = tmp ? tmp[start_n + j] : 1e10f;
        } else {
            pts4[p] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
            tval[p] = -1.0f;
        }
    }
    __syncthreads();

    // ---- bucket spheres and initial candidates ------------------------------------------------------------
    const int nb = npad >> 5;
    float4* bs = ws.bsphere + static_cast<size_t>(cloud) * ws.nbmax;
    int2* bvk = ws.bvk + static_cast<size_t>(cloud) * ws.nbmax;
    float4* bxj = ws.bxyzj + static_cast<size_t>(cloud) * ws.nbmax;
    for (int bk = warp; bk < nb; bk += kNW) {
        const float4 P = pts4[bk * 32 + lane];
        const float tv = tval[bk * 32 + lane];
        const int j = __float_as_int(P.w);
        const bool ok = j >= 0;
        const float lx = warp_min(ok ? P.x : INFINITY), hx = warp_max(ok ? P.x : -INFINITY);
        const float ly = warp_min(ok ? P.y : INFINITY), hy = warp_max(ok ? P.y : -INFINITY);
        const float lz = warp_min(ok ? P.z : INFINITY), hz = warp_max(ok ? P.z : -INFINITY);
        const float ccx = 0.5f * (lx + hx), ccy = 0.5f * (ly + hy), ccz = 0.5f * (lz + hz);
        const float ddx = P.x - ccx, ddy = P.y - ccy, ddz = P.z - ccz;
        const float r = warp_max(ok ? sqrtf(ddx * ddx + ddy * ddy + ddz * ddz) : 0.f);
        const int bi = __float_as_int(tv);
        const int wmax = __reduce_max_sync(FULL, bi);
        const int key = (bi == wmax && ok) ? point_key(j, bs_log2) : INT_MAX;
        const int wkey = __reduce_min_sync(FULL, key);
        if (lane == 0) bs[bk] = make_float4(ccx, ccy, ccz, r * 1.00001f + 4e-6f * scale);
        if (bi == wmax && key == wkey) { bvk[bk] = make_int2(wmax, wkey); bxj[bk] = P; }
    }
}

__global__ void __launch_bounds__(kMT)
fps_bucket_kernel(const float* __restrict__ xyz, const int* __restrict__ offset, const int* __restrict__ new_offset,
                  float* tmp, int* __restrict__ idx, BucketWs ws, int bs_log2)
{
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ int nact;
    __shared__ int4 wres[kMNW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x;
    const int start_n = cloud ? offset[cloud - 1] : 0;
    const int n = offset[cloud] - start_n;
    const int start_m = cloud ? new_offset[cloud - 1] : 0;
    const int m = new_offset[cloud] - start_m;
    if (m <= 0 || n <= 0) return;
    if (tid == 0) idx[start_m] = start_n;                      // sampling_cuda_kernel.cu:39
    if (m == 1) return;

    const int nb = ((n + 31) & ~31) >> 5;
    float4* sph = reinterpret_cast<float4*>(dyn);                                   // [nbmax] centre, theta^2
    float4* cxyzj = sph + ws.nbmax;                                                 // [nbmax] candidate x,y,z,j
    int2* cvk = reinterpret_cast<int2*>(cxyzj + ws.nbmax);                          // [nbmax] candidate value bits, key
    float* rad = reinterpret_cast<float*>(cvk + ws.nbmax);                          // [nbmax] inflated radius + slack
    int* alist = reinterpret_cast<int*>(rad + ws.nbmax);                            // [nbmax] active buckets

    const float4* pts4 = ws.pts4 + static_cast<size_t>(cloud) * ws.stride;
    float* tval = ws.tval + static_cast<size_t>(cloud) * ws.stride;
    {
        const float4* gs = ws.bsphere + static_cast<size_t>(cloud) * ws.nbmax;
        const int2* gv = ws.bvk + static_cast<size_t>(cloud) * ws.nbmax;
        const float4* gx = ws.bxyzj + static_cast<size_t>(cloud) * ws.nbmax;
        for (int bk = tid; bk < nb; bk += kMT) {
            const float4 c = gs[bk];
            const int2 v = gv[bk];
            rad[bk] = c.w;
            sph[bk] = make_float4(c.x, c.y, c.z, skip_threshold2(c.w, __int_as_float(v.x)));
            cvk[bk] = v;
            cxyzj[bk] = gx[bk];
        }
        if (tid == 0) nact = 0;
    }
    float ox = __ldg(xyz + 3 * static_cast<size_t>(start_n)), oy = __ldg(xyz + 3 * static_cast<size_t>(start_n) + 1),
          oz = __ldg(xyz + 3 * static_cast<size_t>(start_n) + 2);
    __syncthreads();

    const int nb_round = (nb + kMT - 1) / kMT * kMT;
    for (int it = 1; it < m; ++it) {
        // ---- (a) which buckets can change? --------------------------------------------------------------
        for (int bk = tid; bk < nb_round; bk += kMT) {
            bool act = false;
            if (bk < nb) {
                const float4 c = sph[bk];
                const float dx = c.x - ox, dy = c.y - oy, dz = c.z - oz;
                act = !(dx * dx + dy * dy + dz * dz > c.w);        // D^2 > theta^2  ->  nothing can change
            }
            const unsigned mask = __ballot_sync(FULL, act);
            if (mask) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&nact, __popc(mask));
                base = __shfl_sync(FULL, base, 0);
                if (act) alist[base + __popc(mask & ((1u << lane) - 1u))] = bk;
            }
        }
        __syncthreads();
        const int na = nact;

        // ---- (b) exact update of the surviving buckets: one warp per bucket, kBatch loads in flight ---------
        for (int a0 = warp; a0 < na; a0 += kMNW * kBatch) {
            int bk[kBatch];
            float4 P[kBatch];
            float tv[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int a = a0 + u * kMNW;
                bk[u] = a < na ? alist[a] : -1;
                if (bk[u] >= 0) {
                    P[u] = __ldg(pts4 + bk[u] * 32 + lane);
                    tv[u] = __ldcg(tval + bk[u] * 32 + lane);
                }
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                if (bk[u] < 0) continue;                      // warp-uniform
                const float dx = P[u].x - ox, dy = P[u].y - oy, dz = P[u].z - oz;
                float d = __fmul_rn(dy, dy);
                d = __fmaf_rn(dx, dx, d);
                d = __fmaf_rn(dz, dz, d);
                const float nt = fminf(d, tv[u]);
                if (nt < tv[u]) __stcg(tval + bk[u] * 32 + lane, nt);
                const int bi = __float_as_int(nt);          // pads stay at -1
                const int wmax = __reduce_max_sync(FULL, bi);
                const int key = (bi == wmax) ? point_key(__float_as_int(P[u].w), bs_log2) : INT_MAX;
                const int wkey = __reduce_min_sync(FULL, key);
                if (bi == wmax && key == wkey) {
                    cvk[bk[u]] = make_int2(wmax, wkey);
                    cxyzj[bk[u]] = P[u];
                    const float4 c = sph[bk[u]];
                    sph[bk[u]] = make_float4(c.x, c.y, c.z, skip_threshold2(rad[bk[u]], nt));
                }
            }
        }
        __syncthreads();
        if (tid == 0) nact = 0;

        // ---- (c) arg-max over the cached candidates of all buckets -----------------------------------------
        int bv = INT_MIN, bkey = INT_MAX, bbk = 0;
        for (int b2 = tid; b2 < nb; b2 += kMT) {
            const int2 c = cvk[b2];
            if (c.x > bv || (c.x == bv && c.y < bkey)) { bv = c.x; bkey = c.y; bbk = b2; }
        }
        {
            const int wv = __reduce_max_sync(FULL, bv);
            const int wk = __reduce_min_sync(FULL, bv == wv ? bkey : INT_MAX);
            if (bv == wv && bkey == wk) wres[warp] = make_int4(wv, wk, bbk, 0);
        }
        __syncthreads();
        {
            int4 c = make_int4(INT_MIN, INT_MAX, 0, 0);
            if (lane < kMNW) c = wres[lane];
            const int gv = __reduce_max_sync(FULL, c.x);
            const int gk = __reduce_min_sync(FULL, c.x == gv ? c.y : INT_MAX);
            const int src = __ffs(__ballot_sync(FULL, c.x == gv && c.y == gk)) - 1;
            const int bstar = __shfl_sync(FULL, c.z, src);
            const float4 w = cxyzj[bstar];
            ox = w.x; oy = w.y; oz = w.z;
            if (tid == 0) idx[start_m + it] = start_n + __float_as_int(w.w);
        }
    }

    if (tmp) {
        __syncthreads();
        for (int p = tid; p < n; p += kMT) {
            const int j = __float_as_int(__ldg(&pts4[p].w));
            tmp[start_n + j] = __ldcg(tval + p);
        }
    }
}

bool g_pool_configured = false;

}  // namespace

// Largest cloud the bucket kernel takes (bucket table in shared memory: 48 bytes per 32 points).
int fps_bucket_max_points() { return 4096 * 32; }

int fps_bucket_launch(int b, int n_max, const float* xyz, const int* offset, const int* new_offset, float* tmp, int* idx,
                      int bs_log2, cudaStream_t stream)
{
    BucketWs ws{};
    ws.stride = (n_max + 31) & ~31;
    ws.nbmax = ws.stride / 32;
    const size_t pts = static_cast<size_t>(b) * ws.stride;
    const size_t nbt = static_cast<size_t>(b) * ws.nbmax;
    // one stream-ordered allocation, carved
    size_t off = 0;
    auto take = [&off](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~static_cast<size_t>(255); return o; };
    const size_t o_pts = take(pts * sizeof(float4)), o_t = take(pts * sizeof(float)), o_ka = take(pts * sizeof(uint2)),
                 o_kb = take(pts * sizeof(uint2)), o_bs = take(nbt * sizeof(float4)), o_bv = take(nbt * sizeof(int2)),
                 o_bx = take(nbt * sizeof(float4)), o_sc = take(static_cast<size_t>(b) * sizeof(float));
    if (!g_pool_configured) {            // keep freed blocks in the pool instead of returning them to the OS
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        (void)cudaGetLastError();
        g_pool_configured = true;
    }
    unsigned char* base = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&base), off, stream);
    if (e != cudaSuccess) { set_error("furthestsampling: workspace of %zu bytes: %s", off, cudaGetErrorString(e)); (void)cudaGetLastError(); return TGN_ERR_CUDA; }
    ws.pts4 = reinterpret_cast<float4*>(base + o_pts);
    ws.tval = reinterpret_cast<float*>(base + o_t);
    ws.key_a = reinterpret_cast<uint2*>(base + o_ka);
    ws.key_b = reinterpret_cast<uint2*>(base + o_kb);
    ws.bsphere = reinterpret_cast<float4*>(base + o_bs);
    ws.bvk = reinterpret_cast<int2*>(base + o_bv);
    ws.bxyzj = reinterpret_cast<float4*>(base + o_bx);
    ws.scale = reinterpret_cast<float*>(base + o_sc);

    fps_bucket_sort_kernel<<<b, kT, 0, stream>>>(xyz, offset, tmp, ws, bs_log2);
    int rc = check_launch("fps_bucket_sort_kernel");
    if (rc == TGN_OK) {
        const size_t smem = static_cast<size_t>(ws.nbmax) * (sizeof(float4) * 2 + sizeof(int2) + sizeof(float) + sizeof(int));
        static size_t configured = 0;
        if (smem > 48 * 1024 && smem > configured) {
            e = cudaFuncSetAttribute(fps_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); rc = TGN_ERR_CUDA; }
            else configured = smem;
        }
        if (rc == TGN_OK) {
            fps_bucket_kernel<<<b, kMT, smem, stream>>>(xyz, offset, new_offset, tmp, idx, ws, bs_log2);
            rc = check_launch("fps_bucket_kernel");
        }
    }
    (void)cudaFreeAsync(base, stream);
    return rc;
}

}  // namespace tgn
