/* kernels.h — launch wrappers exported by kernels.cu to engine.cu. */
#ifndef RVPT_KERNELS_H
#define RVPT_KERNELS_H

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "device_scene.h"

namespace rvpt
{
cudaError_t configure_kernels(size_t max_dynamic_smem);
size_t frame_smem_bytes(size_t scene_bytes, uint32_t n_nodes, uint32_t n_tris, bool oct);
cudaError_t occupancy(int* frame_ctas_per_sm, int* primary_ctas_per_sm, int* bounce_ctas_per_sm,
                      bool smem, bool oct, size_t scene_bytes, uint32_t n_nodes, uint32_t n_tris);
cudaError_t launch_frame(const FrameParams& p, bool smem, bool oct, int grid, cudaStream_t st);
cudaError_t launch_primary(const FrameParams& p, bool smem, int grid, cudaStream_t st);
cudaError_t launch_bounce(const FrameParams& p, int b, bool smem, int grid, cudaStream_t st);
cudaError_t launch_modes(const FrameParams& p, bool smem, int grid, cudaStream_t st);
cudaError_t launch_untile(const void* src, void* dst, uint32_t words, uint32_t W, uint32_t H,
                          uint32_t tiles_x, uint32_t n_tiles, uint32_t nranks, uint32_t first_rank,
                          uint32_t n_src_ranks, uint32_t n_local_padded, cudaStream_t st);
cudaError_t launch_tile(const void* raster, void* tiles, uint32_t words, uint32_t W, uint32_t H,
                        uint32_t tiles_x, uint32_t n_tiles, uint32_t nranks, uint32_t rank,
                        uint32_t n_local_padded, cudaStream_t st);
cudaError_t launch_u8_to_f32(const void* src, void* dst, uint64_t n, cudaStream_t st);
cudaError_t launch_f32_to_u8(const void* src, void* dst, uint64_t n, cudaStream_t st);
cudaError_t launch_selftest(int op, const float* in, size_t n, float* out, cudaStream_t st);
int threads_per_cta();
} /* namespace rvpt */

#endif
