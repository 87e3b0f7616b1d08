// fps.cuh -- internal interface between the FPS kernels.
#pragma once
#include <cuda_runtime.h>

namespace tgn {

// Bucket-pruned FPS (fps_bucket.cu): any batch, clouds up to fps_bucket_max_points() points.
int fps_bucket_max_points();
// shape: warps per cloud (16, 8, 4, 2, 1) or 0 = by batch size.
int fps_bucket_launch(int b, int n_max, const float* xyz, const int* offset, const int* new_offset, float* tmp, int* idx,
                      int bs_log2, int shape, cudaStream_t stream);

}  // namespace tgn
