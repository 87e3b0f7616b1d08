// ballquery.cuh -- internal interface between the ball-query kernels.
#pragma once
#include <cuda_runtime.h>

namespace tgn {

// Per-call scratch of the grid ball query (ballquery_grid.cu), all per cloud b.
struct BqGridWs {
    float4* ga;        // [B][(N+1)/2]  cell-sorted points in pairs: (x0, x1, y0, y1)
    float4* gb;        // [B][(N+1)/2]  (z0, z1, |p0|^2, |p1|^2)
    int2* gj;          // [B][(N+1)/2]  original indices (j0, j1)
    int* cell_start;   // [B][kMaxCells + 1]
    float4* org;       // [B]      grid origin x, y, z, 1 / cell size
    int4* dim;         // [B]      nx, ny, nz, cells
    float4* bnd;       // [B]      max |p|^2, max |coordinate|
    int* flag;         // [B]      1: answered by the grid kernel, 0: left to the tile kernel
};

int bq_grid_max_points();
// Byte size of the scratch; fills `offsets` with the byte offsets of the arrays (as pointers from 0).
size_t bq_grid_workspace_bytes(int B, int N, BqGridWs* offsets);
// force: 0 = per-cloud estimate decides, 1 = every cloud takes the grid kernel.
// order: bit 0 / bit 1 = |new_xyz|^2 / |xyz|^2 rounded as torch's contiguous reduce does (see ballquery.cu).
int bq_grid_launch(int B, int N, int S, float r2, int nsample, const float* xyz, const float* new_xyz, void* group_idx,
                   bool idx64, int force, int order, const BqGridWs& ws, cudaStream_t st);

}  // namespace tgn
