// crt_kernels.h — host-visible launch interface of crt_kernels.cu
#pragma once
#include <cuda_runtime.h>
#include "crt_common.h"

namespace crtb {

// Device pointers of one uploaded batch, passed to every kernel by value.
struct DevBatch {
	const uint8_t *blobs;        // blob arena (each blob 16-byte aligned)
	uint8_t *symbols;            // decoded entropy symbols (clers, logs), each block 16-byte aligned
	const MeshDesc *mesh;
	const TunDesc *tun;
	uint8_t *tunrec;             // ntun x TUN_REC_BYTES dictionaries
	uint32_t *tun_used;          // ntun: used dictionary text bytes (rounded to 16)
	const uint32_t *group_ends;
	int32_t *status;             // per mesh: 0 or CRT_E_*
	uint32_t *vertex_count;      // per mesh: vertices the CLERS automaton created
	uint32_t *regular;           // per mesh: 1 when most of its CLERS symbols sat in long VERTEX / LEFT runs (k_clers_cta), else 0;
	                             // bit 31 while k_clers_cta has left the mesh to k_clers_lf (irregular stream)
	unsigned long long *tun_bits; // per entropy block: sum of its decoded symbols = bits its values occupy (zeroed per decode)
};

// Per-slot state of the CLERS automaton (v1: all in global memory).
struct ClersScratch {
	EdgeA *ea; EdgeB *eb; uint32_t *order; uint32_t *delayed;
	uint32_t cap;                // entries per slot
	uint32_t slots;              // concurrent meshes
};

int launch_tun_tables(const DevBatch &B, int ntun, cudaStream_t s);
int launch_tun_decode(const DevBatch &B, const Tile *tiles, uint32_t ntiles, uint64_t *states, uint32_t *ticket, int sms, cudaStream_t s);
int launch_mesh_unpack(const DevBatch &B, const Tile *tiles, uint32_t ntiles, const uint32_t *heads, uint32_t nchains, uint64_t *states, uint32_t *ticket, int sms, cudaStream_t s);
int launch_clers(const DevBatch &B, const uint32_t *order, uint32_t nwork, const ClersScratch &scratch, uint32_t *ticket, int sms, cudaStream_t s);
int launch_clers_cta(const DevBatch &B, const uint32_t *order, uint32_t nwork, const ClersScratch &scratch, uint32_t *ticket, int sms, bool defer, cudaStream_t s);
int launch_delta_mesh(const DevBatch &B, const uint2 *work, uint32_t nwork, bool split, cudaStream_t s);
int launch_adj_build(const DevBatch &B, const Tile *tiles, uint32_t ntiles, int part, cudaStream_t s);
int launch_scan_u32(const DevBatch &B, const Tile *tiles, uint32_t ntiles, uint64_t *states, uint32_t *ticket, int sms, cudaStream_t s);
int launch_normal_estimate(const DevBatch &B, const Tile *tiles, uint32_t ntiles, cudaStream_t s);
int launch_dequant(const DevBatch &B, const Tile *tiles, uint32_t ntiles, cudaStream_t s);
int launch_cloud_fused(const DevBatch &B, const Tile *tiles, uint32_t ntiles, const uint32_t *heads, uint32_t nchains, uint64_t *states, uint32_t *ticket, int sms, cudaStream_t s);

}  // namespace crtb
