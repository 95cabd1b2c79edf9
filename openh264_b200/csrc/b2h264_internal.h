// b2h264_internal.h — shared by the .cu translation units of libopenh264_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2h264.h"

// records one kernel launch (for b2h264_launch_count) and returns the launch status
int b2h264_launched();

// TMA descriptor (CUtensorMap, 128 bytes) over a stack of n byte planes; 0 on success (k_pixel.cu)
int b2h264_make_tmap_planes(void* out, const void* base, uint64_t w, uint64_t h, uint64_t n, uint64_t stride_y, uint64_t stride_plane,
                            uint32_t box_w, uint32_t box_h);
