// cube_tables.cuh — the start cube of Polyhedron::build_cube (polyhedron.rs:268-392) as constant tables,
// shared by the warp-per-cell kernel (clip.cu) and the thread-per-cell kernel (clip_thread.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tess {
namespace {

// Start cube: {flip, target, next} of the 24 half-edges (ids of polyhedron.rs:97-199)
__constant__ uint32_t kCubeEdges[24] = {0x100301u, 0x0f0002u, 0x140103u, 0x050200u, 0x110205u, 0x030106u, 0x170507u, 0x090604u,
                                        0x120609u, 0x07050au, 0x16040bu, 0x0d0708u, 0x13070du, 0x0b040eu, 0x15000fu, 0x01030cu,
                                        0x000211u, 0x040612u, 0x080713u, 0x0c0310u, 0x020015u, 0x0e0416u, 0x0a0517u, 0x060114u};

// ... and, per cube vertex, the three half-edges that START there (source(e) = target(flip(e)))
__constant__ unsigned char kCubeVout[24] = {2, 15, 21, 3, 6, 20, 0, 5, 17, 1, 12, 16, 11, 14, 22, 7, 10, 23, 4, 9, 18, 8, 13, 19};

}  // namespace
}  // namespace tess
