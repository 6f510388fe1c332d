// dv_kernels.h -- host-visible launch wrappers of the divANS kernels.
#pragma once
#include "dv_common.cuh"

namespace dv {
constexpr int DECODE_BLOCK_THREADS = 64;

void launch_frame(const FrameParams &p, uint8_t *payload, uint64_t payload_cap_bytes, cudaStream_t st);   // frame + payload scan + demux (3 launches)
void launch_decode32(const DecodeParams &p, uint32_t n_blocks, cudaStream_t st);
void launch_decode16(const DecodeParams &p, uint32_t n_blocks, cudaStream_t st);
int decode_max_blocks_per_sm32();
void launch_decode8(const DecodeParams &p, uint32_t n_blocks, cudaStream_t st);   // 8-lane engine: 4 streams per warp (dv8_kernels.cu)
int decode_max_blocks_per_sm8();
int decode_groups_per_block8();
int decode_max_blocks_per_sm16();
void launch_encode_model(const EncodeParams &p, uint32_t n_blocks, cudaStream_t st);   // groups of 16 lanes
void launch_encode_flush_mux(const EncodeParams &p, cudaStream_t st);                  // reverse rANS + mux/CRC (2 launches)
int encode_max_blocks_per_sm();
void launch_rcp15_init(uint64_t *tab, cudaStream_t st);
}  // namespace dv
