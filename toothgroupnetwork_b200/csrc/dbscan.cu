// dbscan.cu -- DBSCAN labels on the device, equal to scikit-learn's.
//
// The reference clusters the offset-moved foreground points of every two-stage forward on the host:
// ops_utils.get_clustering_labels (ops_utils.py:86-144) calls sklearn.cluster.DBSCAN(eps=0.03, min_samples=30).fit(...),
// 0.14 s (sparse predictions) to 0.9 s (a trained network: every tooth collapsed onto its centroid, ~1500 neighbours per
// point) for 23 000 points -- several times the two networks around it.
//
// sklearn's result is a deterministic function of the eps-graph (sklearn/cluster/_dbscan.py + _dbscan_inner.pyx):
//   * neighbourhood: KDTree over the float64 copy of the points, point j is a neighbour of i when
//     ((dx*dx + dy*dy) + dz*dz) <= eps*eps in float64 (reduced distance, no square root); i is its own neighbour;
//   * core: at least min_samples neighbours;
//   * clusters: connected components of the core points under that relation, numbered in the order of their smallest
//     member index (the outer loop of dbscan_inner visits points in index order and floods one component at a time);
//   * border points (non-core with a core neighbour) take the first cluster that reaches them = the smallest label among
//     their core neighbours; everything else is noise (-1).
// Here: a uniform grid with cells >= eps, float64 distance predicate in the same operation order without FMA contraction,
// a lock-free union-find that always hooks the larger root under the smaller (so a component's root IS its smallest core
// index), a scan over the roots for the numbering, and a last pass for the border points.  One warp per point, lanes over
// the candidates of the 27 neighbouring cells.
#include <algorithm>
#include <cfloat>
#include <climits>

#include "common.cuh"
#include "tgn_b200.h"

namespace tgn {
namespace {

constexpr int kMaxDim = 64;                       // cells per axis (cells grow beyond eps when the cloud is wider than 64 eps)
constexpr int kT = 256;
constexpr unsigned FULL = 0xffffffffu;

struct DbGrid {
    double lo[3], h[3];
    int dim[3], ncells;
};

struct DbWs {
    DbGrid* grid;
    int* cell_start;      // [kMaxDim^3 + 1]
    int* cell_fill;       // [kMaxDim^3]
    int* cell_of;         // [n]
    float4* sorted;       // [n] x, y, z, original index (bits)
    int* parent;          // [n]
    int* cid;             // [n] cluster id of a root
    int* n_clusters;      // [1]
};

__host__ __device__ inline size_t align_up(size_t v) { return (v + 255) / 256 * 256; }

inline size_t ws_layout(int n, unsigned char* base, DbWs* w)
{
    const size_t cells = static_cast<size_t>(kMaxDim) * kMaxDim * kMaxDim;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
    const size_t o_grid = take(sizeof(DbGrid)), o_start = take((cells + 1) * 4), o_fill = take(cells * 4), o_cell = take(static_cast<size_t>(n) * 4),
                 o_sorted = take(static_cast<size_t>(n) * 16), o_parent = take(static_cast<size_t>(n) * 4), o_cid = take(static_cast<size_t>(n) * 4),
                 o_nc = take(4);
    if (w) {
        w->grid = reinterpret_cast<DbGrid*>(base + o_grid);
        w->cell_start = reinterpret_cast<int*>(base + o_start);
        w->cell_fill = reinterpret_cast<int*>(base + o_fill);
        w->cell_of = reinterpret_cast<int*>(base + o_cell);
        w->sorted = reinterpret_cast<float4*>(base + o_sorted);
        w->parent = reinterpret_cast<int*>(base + o_parent);
        w->cid = reinterpret_cast<int*>(base + o_cid);
        w->n_clusters = reinterpret_cast<int*>(base + o_nc);
    }
    return off;
}

// bounding box -> grid geometry; one block
__global__ void __launch_bounds__(1024) db_setup_kernel(int n, const float* __restrict__ xyz, double eps, DbWs w)
{
    __shared__ float s_lo[3][32], s_hi[3][32];
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = threadIdx.x; i < n; i += blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = xyz[3 * static_cast<size_t>(i) + a];
            if (isfinite(v)) { lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
        }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o; o >>= 1) { lo[a] = fminf(lo[a], __shfl_xor_sync(FULL, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(FULL, hi[a], o)); }
        if ((threadIdx.x & 31) == 0) { s_lo[a][threadIdx.x >> 5] = lo[a]; s_hi[a][threadIdx.x >> 5] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        DbGrid g;
        g.ncells = 1;
        for (int a = 0; a < 3; ++a) {
            float l = FLT_MAX, h = -FLT_MAX;
            for (int q = 0; q < (blockDim.x >> 5); ++q) { l = fminf(l, s_lo[a][q]); h = fmaxf(h, s_hi[a][q]); }
            if (!(l <= h)) { l = 0.f; h = 0.f; }
            const double extent = static_cast<double>(h) - static_cast<double>(l);
            const double cell = fmax(eps * (1.0 + 1e-9), extent / (kMaxDim - 1));           // always >= eps
            g.lo[a] = l;
            g.h[a] = cell > 0.0 ? cell : 1.0;
            g.dim[a] = min(kMaxDim, static_cast<int>(extent / g.h[a]) + 1);
            g.ncells *= g.dim[a];
        }
        *w.grid = g;
        *w.n_clusters = 0;
    }
}

__device__ __forceinline__ int cell_coord(float v, double lo, double h, int dim)
{
    if (!isfinite(v)) return 0;
    const int c = static_cast<int>((static_cast<double>(v) - lo) / h);
    return min(max(c, 0), dim - 1);
}

__global__ void db_count_kernel(int n, const float* __restrict__ xyz, DbWs w)
{
    const DbGrid g = *w.grid;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int cx = cell_coord(xyz[3 * static_cast<size_t>(i)], g.lo[0], g.h[0], g.dim[0]),
                  cy = cell_coord(xyz[3 * static_cast<size_t>(i) + 1], g.lo[1], g.h[1], g.dim[1]),
                  cz = cell_coord(xyz[3 * static_cast<size_t>(i) + 2], g.lo[2], g.h[2], g.dim[2]);
        const int cell = (cz * g.dim[1] + cy) * g.dim[0] + cx;
        w.cell_of[i] = cell;
        atomicAdd(w.cell_start + cell + 1, 1);                     // counts land one slot up: the scan below makes them starts
    }
}

// inclusive scan of the per-cell counts (one block); cell_start[0] = 0
__global__ void __launch_bounds__(1024) db_scan_kernel(DbWs w)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int ncells = w.grid->ncells;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ncells; base += 1024) {
        const int i = base + threadIdx.x;
        int v = i < ncells ? w.cell_start[i + 1] : 0;
        const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, v, o); if (lane >= o) v += t; }
        if (lane == 31) warp_sums[wp] = v;
        __syncthreads();
        if (wp == 0) {
            int s = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, s, o); if (lane >= o) s += t; }
            warp_sums[lane] = s;
        }
        __syncthreads();
        const int total = v + (wp ? warp_sums[wp - 1] : 0) + carry;
        if (i < ncells) { w.cell_start[i + 1] = total; w.cell_fill[i] = 0; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = total;
        __syncthreads();
    }
}

__global__ void db_scatter_kernel(int n, const float* __restrict__ xyz, DbWs w)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int cell = w.cell_of[i];
        const int pos = w.cell_start[cell] + atomicAdd(w.cell_fill + cell, 1);
        w.sorted[pos] = make_float4(xyz[3 * static_cast<size_t>(i)], xyz[3 * static_cast<size_t>(i) + 1], xyz[3 * static_cast<size_t>(i) + 2], __int_as_float(i));
        w.parent[i] = i;
        w.cid[i] = -1;
    }
}

// sklearn's predicate: float64, (dx*dx + dy*dy) + dz*dz <= eps*eps, no contraction
__device__ __forceinline__ bool within(const float4& a, const float4& b, double r2)
{
    const double dx = static_cast<double>(a.x) - static_cast<double>(b.x), dy = static_cast<double>(a.y) - static_cast<double>(b.y),
                 dz = static_cast<double>(a.z) - static_cast<double>(b.z);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) <= r2;
}

__device__ __forceinline__ int uf_find(int* parent, int x)
{
    volatile int* p = parent;
    int px = p[x];
    while (px != x) {
        const int gp = p[px];
        if (gp != px) p[x] = gp;                                   // path halving: gp is an ancestor in any interleaving
        x = px;
        px = gp;
    }
    return x;
}

// The same walk through L1-cached loads and without writes, for the hot comparison "are i and j already together?".  A stale
// line can only show an OLDER parent pointer, which still names an ancestor of x (hooks are never undone, parent[x] <= x always),
// so the walk ends at some ancestor r of x: "r(i) == r(j)" then proves i and j share a component (no false positive); a false
// negative just falls through to uf_union, whose compare-and-swap sees the real value.
__device__ __forceinline__ int uf_find_cached(const int* parent, int x)
{
    int px = __ldca(parent + x);
    while (px != x) { x = px; px = __ldca(parent + x); }
    return x;
}

// hook the larger root under the smaller: a component's root is its smallest member
__device__ __forceinline__ void uf_union(int* parent, int a, int b)
{
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicCAS(parent + a, a, b);
        if (old == a) return;
        a = old;
    }
}

// PASS 0: core flags   PASS 1: union of core neighbours   PASS 2: border labels
template <int PASS>
__global__ void __launch_bounds__(kT) db_neighbour_kernel(int n, double r2, int min_samples, DbWs w, unsigned char* core, int* labels)
{
    const DbGrid g = *w.grid;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int pos = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pos < n; pos += warps) {
        const float4 me = w.sorted[pos];
        const int i = __float_as_int(me.w);
        if (PASS == 1 && !core[i]) continue;
        if (PASS == 2 && core[i]) continue;
        const int cell = w.cell_of[i];
        const int cx = cell % g.dim[0], cy = (cell / g.dim[0]) % g.dim[1], cz = cell / (g.dim[0] * g.dim[1]);
        int count = 0, best = INT_MAX;
        bool done = false;
        int ri = PASS == 1 ? uf_find(w.parent, i) : 0;               // an ancestor of i, refreshed after every union this lane makes
        for (int dz = -1; dz <= 1 && !done; ++dz) {
            const int z = cz + dz;
            if (z < 0 || z >= g.dim[2]) continue;
            for (int dy = -1; dy <= 1 && !done; ++dy) {
                const int y = cy + dy;
                if (y < 0 || y >= g.dim[1]) continue;
                // the three x-neighbours are one contiguous run of the sorted array
                const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dim[0] - 1);
                const int row = (z * g.dim[1] + y) * g.dim[0];
                const int s = w.cell_start[row + x0], e = w.cell_start[row + x1 + 1];
                for (int q = s + lane; q < e + ((32 - ((e - s) & 31)) & 31); q += 32) {      // warp-uniform trip count
                    bool hit = false;
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (q < e) { o = w.sorted[q]; hit = within(me, o, r2); }
                    if (PASS == 0) {
                        count += __popc(__ballot_sync(FULL, hit));
                        if (count >= min_samples) { done = true; break; }
                    } else if (PASS == 1) {
                        const int j = __float_as_int(o.w);
                        // in a converged cloud every point has ~1500 core neighbours that were joined long ago: the common case is
                        // two short cached walks and a compare (ncu before: 7 % issue utilisation, the kernel waited on L2 loads)
                        if (hit && j < i && core[j] && uf_find_cached(w.parent, j) != ri) {
                            uf_union(w.parent, i, j);
                            ri = uf_find(w.parent, i);
                        }
                    } else {
                        const int j = __float_as_int(o.w);
                        if (hit && core[j]) best = min(best, labels[j]);          // core labels were written by db_number_kernel
                    }
                }
            }
        }
        if (PASS == 0 && lane == 0) core[i] = count >= min_samples ? 1 : 0;
        if (PASS == 2) {
            for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(FULL, best, o));
            if (lane == 0) labels[i] = best == INT_MAX ? -1 : best;
        }
    }
}

// clusters numbered by their smallest core index: exclusive scan of "is a root" in index order (one block), then every core
// point takes its root's number
__global__ void __launch_bounds__(1024) db_number_kernel(int n, DbWs w, const unsigned char* __restrict__ core, int* __restrict__ labels)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int flag = (i < n && core[i] && w.parent[i] == i) ? 1 : 0;
        int v = flag;
        const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, v, o); if (lane >= o) v += t; }
        if (lane == 31) warp_sums[wp] = v;
        __syncthreads();
        if (wp == 0) {
            int s = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, s, o); if (lane >= o) s += t; }
            warp_sums[lane] = s;
        }
        __syncthreads();
        const int incl = v + (wp ? warp_sums[wp - 1] : 0) + carry;
        if (flag) w.cid[i] = incl - 1;
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *w.n_clusters = carry;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 1024)
        if (core[i]) labels[i] = w.cid[uf_find(w.parent, i)];
}

}  // namespace
}  // namespace tgn

extern "C" {

size_t tgn_dbscan_bytes(int n) { return tgn::ws_layout(std::max(n, 1), nullptr, nullptr); }

int tgn_dbscan(int n, const float* xyz, double eps, int min_samples, void* workspace, int* labels, unsigned char* core, int* n_clusters, void* stream)
{
    using namespace tgn;
    if (n < 0 || !(eps > 0.0) || min_samples < 1) { set_error("dbscan: bad arguments n=%d eps=%g min_samples=%d", n, eps, min_samples); return TGN_ERR_INVALID; }
    if (n == 0) return TGN_OK;
    if (!xyz || !workspace || !labels || !core) { set_error("dbscan: null argument"); return TGN_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DbWs w;
    ws_layout(n, static_cast<unsigned char*>(workspace), &w);
    const size_t cells = static_cast<size_t>(kMaxDim) * kMaxDim * kMaxDim;
    if (cudaMemsetAsync(w.cell_start, 0, (cells + 1) * 4, st) != cudaSuccess) { set_error("dbscan: memset failed"); return TGN_ERR_CUDA; }
    db_setup_kernel<<<1, 1024, 0, st>>>(n, xyz, eps, w);
    const int blocks = std::max(1, std::min((n + kT - 1) / kT, 8 * sm_count()));
    db_count_kernel<<<blocks, kT, 0, st>>>(n, xyz, w);
    db_scan_kernel<<<1, 1024, 0, st>>>(w);
    db_scatter_kernel<<<blocks, kT, 0, st>>>(n, xyz, w);
    const int wblocks = std::max(1, std::min((n + kT / 32 - 1) / (kT / 32), 16 * sm_count()));
    const double r2 = eps * eps;
    db_neighbour_kernel<0><<<wblocks, kT, 0, st>>>(n, r2, min_samples, w, core, labels);
    db_neighbour_kernel<1><<<wblocks, kT, 0, st>>>(n, r2, min_samples, w, core, labels);
    db_number_kernel<<<1, 1024, 0, st>>>(n, w, core, labels);
    db_neighbour_kernel<2><<<wblocks, kT, 0, st>>>(n, r2, min_samples, w, core, labels);
    if (n_clusters && cudaMemcpyAsync(n_clusters, w.n_clusters, 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        set_error("dbscan: copy of the cluster count failed");
        return TGN_ERR_CUDA;
    }
    return check_launch("dbscan");
}

}  // extern "C"
