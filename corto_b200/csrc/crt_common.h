// crt_common.h — descriptors shared by the host directory walk (crt_walk.cpp), the C ABI (crt_api.cu) and the
// CUDA kernels (crt_kernels.cu).  Everything here is POD and is copied to the device verbatim.
//
// Vocabulary follows the reference's .crt wire format: a blob holds, per mesh, a CLERS symbol
// stream + split bitstream (index attribute) and per vertex attribute one BITS block of raw residual bits plus
// one (CORRELATED) or N (per-component) entropy-coded blocks of bit-lengths ("logs").
#pragma once
#include <stdint.h>

namespace crtb {

constexpr int MAX_ATTR = 8;        // attributes per mesh handled by this build (reference: unbounded std::map)
constexpr int MAX_COMP = 4;        // components per attribute handled by the kernels
constexpr int TUN_TABLE_BYTES = 8192;   // dictionary text capacity, src/tunstall.cpp:138
constexpr int TUN_REC_BYTES = 1024 + TUN_TABLE_BYTES;   // per block: 256 packed (offset | len<<16) entries + text

// tile sizes of the scan-based kernels (elements per CTA iteration)
constexpr int TUN_TILE = 2048;     // compressed bytes per tile (256 threads x 8)
constexpr int BIT_TILE = 1024;     // vertices per tile of the fused unpack kernels (256 threads x 4)
constexpr int SCAN_TILE = 1024;    // elements per tile of the generic u32 scans

enum Format { F_UINT32 = 0, F_INT32, F_UINT16, F_INT16, F_UINT8, F_INT8, F_FLOAT, F_DOUBLE };
enum Strategy { S_PARALLEL = 1, S_CORRELATED = 2 };
enum Codec { CODEC_GENERIC = 1, CODEC_NORMAL = 2, CODEC_COLOR = 3 };
enum NormalPrediction { N_DIFF = 0, N_ESTIMATED = 1, N_BORDER = 2 };
enum Cler { C_VERTEX = 0, C_LEFT, C_RIGHT, C_END, C_BOUNDARY, C_DELAY, C_SPLIT };

// One entropy-coded block (src/cstream.cpp:111-128) or, for Stream::NONE, one raw run (cstream.cpp:68-73).
struct TunDesc {
	uint64_t probs_off;   // byte offset (blob arena) of the nsym (symbol, prob) pairs
	uint64_t data_off;    // byte offset (blob arena) of the compressed bytes / raw bytes
	uint64_t out_off;     // byte offset (symbol arena) of the decoded symbols
	uint32_t nsym;        // 0 for raw
	uint32_t size;        // decoded symbol count
	uint32_t csize;       // compressed byte count (== size for raw)
	uint32_t raw;         // 1: Stream::NONE, copy through
	uint32_t tile0;       // first tile of this block in the Tunstall tile list
	uint32_t pad;
};

struct AttrDesc {
	int32_t codec;            // Codec
	int32_t N;                // header component count (normals: 3)
	int32_t ncomp;            // components carried by the stream (normals: 2)
	int32_t strategy;         // Strategy bits from the header
	float q;                  // quantisation step (normals: 2^(bits-1))
	int32_t out_format;       // Format, or -1 when unbound (parse + skip)
	int32_t out_components;   // colour only (3 or 4)
	int32_t prediction;       // normals only: NormalPrediction byte from the stream
	int32_t qc[4];            // colour only: per-channel steps from the stream
	uint64_t bits_off;        // byte offset (blob arena) of the BITS words
	uint32_t bits_nwords;
	int32_t ntun;             // 1 (CORRELATED / normals) or ncomp
	int32_t tun[MAX_COMP];    // TunDesc indices
	uint32_t count;           // symbols per log stream (nvert, or boundary count for BORDER normals)
	uint32_t pad0;
	uint64_t out_ptr;         // device address of this mesh's slice of the output arena (0 = unbound)
	uint64_t work_ptr;        // device address of scratch: normals int32 diffs[2*nvert] / colour u8 values[N*nvert]
};

struct MeshDesc {
	uint64_t blob_off;        // byte offset of the blob in the blob arena (16-byte aligned)
	uint32_t blob_len;
	uint32_t nvert, nface;
	uint32_t nattr;
	uint32_t ngroups, group0; // group end-face indices live in group_ends[group0 .. group0+ngroups)
	int32_t clers_tun;        // TunDesc index of the CLERS stream, -1 for point clouds
	uint32_t nclers;
	uint64_t split_off;       // byte offset (blob arena) of the split-index BITS words
	uint32_t split_nwords;
	uint32_t max_front;       // hint written by the encoder (index_attribute.h:84); not trusted
	uint32_t max_group_faces; // largest group -> bound on the front size (3 edges per face)
	int32_t index16;          // 1: uint16 indices
	uint64_t index_ptr;       // device address of this mesh's slice of the index arena (0 = unbound)
	uint64_t pred_ptr;        // device scratch: uint4 (a,b,c,-) per vertex  (decoder.cpp:171)
	uint64_t face_ptr;        // device scratch: u32 faces when index is unbound or u16 but normals need them (else == index_ptr)
	int32_t position_attr;    // index of the "position" attribute, -1 if none
	int32_t normal_attr;      // index of a bound ESTIMATED/BORDER normal attribute needing postDelta, else -1
	uint64_t csr_ptr;         // device scratch (zeroed per decode) for normal estimation: u32 cnt[nvert] | bnd[nvert] | cidx[nvert+1] | novf
	uint64_t adj_ptr;         // device scratch: float4 fnormal[nface] (unnormalised face normals) | u32 adj8[8*nvert] (incident faces, 8 slots per vertex) | uint2 ovf[3*nface] (overflow pairs)
	AttrDesc attr[MAX_ATTR];
};

// Work tiles.  Each scan-style kernel walks a flat tile list in ticket order; `first` marks the first tile of a
// look-back chain (a chain = one Tunstall block / all component streams of one attribute / one scan segment).
struct Tile {
	uint32_t a;       // kernel-specific: TunDesc index | mesh index
	uint32_t b;       // kernel-specific: attribute index
	uint32_t tile;    // tile index inside the stream
	uint32_t first;   // 1: no predecessor in the chain
};

// Front edge of the CLERS automaton, split in a write-once half and a mutable half (src/decoder.cpp:32-39).
struct EdgeA { uint32_t v0, v1, v2, deleted; };
struct EdgeB { uint32_t prev, next; };

}  // namespace crtb
