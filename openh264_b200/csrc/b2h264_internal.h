// b2h264_internal.h — shared by the .cu translation units of libopenh264_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2h264.h"

// records one kernel launch (for b2h264_launch_count) and returns the launch status
int b2h264_launched();
