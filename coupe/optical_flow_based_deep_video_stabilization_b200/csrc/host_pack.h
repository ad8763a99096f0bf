// Host-side float32 -> bf16 packing of the network input for ofs_net_stabilize_host (see host_pack.cpp).
#pragma once

#include <cstddef>
#include <cstdint>

namespace ofs {

struct HostPacker;                                   // a few worker threads converting one array at a time
HostPacker* host_packer_create(int threads);
void host_packer_destroy(HostPacker* p);
void host_packer_start(HostPacker* p, const float* src, uint16_t* dst, size_t n);   // returns at once
void host_packer_wait(HostPacker* p);                                               // until the started job is done
void host_cvt_f32_to_bf16(const float* src, uint16_t* dst, size_t n);               // cvt.rn.bf16.f32, single thread

}  // namespace ofs
