// Host helpers shared by the tensor-core launchers: TMA tensor-map encoding (driver entry point resolved at
// run time, so the extension links without libcuda on the GPU-less build box), SM count, last-error string.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ts {
// 2-D bf16 row-major tensor [rows, cols] with row pitch `pitch_elems`; box = box_rows x box_cols, 128 B swizzle
// (box_cols must be 64) or no swizzle (swizzle=0).
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                      uint32_t box_cols, uint32_t box_rows, int swizzle128 = 1);
// 3-D bf16 tensor [d2, d1, d0] (d0 contiguous), strides in elements.
int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                      uint64_t stride2, uint32_t box0, uint32_t box1, uint32_t box2, int swizzle128 = 1);
int sm_count(int dev);
void set_last_error(const char* msg);
}  // namespace ts
