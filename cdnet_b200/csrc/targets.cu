// targets.cu -- placeholder until the target-transform kernels land (next commit).
#include "internal.h"

extern "C" size_t cdnet_center_points_workspace_bytes(int, int, int, int) { return 0; }
extern "C" int cdnet_center_points(const int32_t*, int32_t*, int, int, int, int, void*, size_t, void*) { return 3; }
extern "C" size_t cdnet_encode_targets_workspace_bytes(int, int, int) { return 0; }
extern "C" int cdnet_encode_targets(const uint8_t*, const uint8_t*, uint8_t*, uint16_t*, int64_t*, int32_t*, int32_t*,
                                    int, int, int, int, void*, size_t, void*) { return 3; }
