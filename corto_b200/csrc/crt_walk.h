// crt_walk.h — host side "K0": header parse (crt::Decoder ctor, src/decoder.cpp:41-89) and the stream-directory
// walk (the block headers Decoder::decode touches before any payload: index_attribute.h:83-99,
// cstream.cpp:111-128, cstream.h:283-291, normal_attribute.cpp:178-180, color_attribute.h:55-58).
// O(#blocks) pointer chasing, no payload byte is read.  Unlike the reference every length is checked against the
// blob size, so a truncated / corrupt .crt is rejected here instead of reading out of bounds on the GPU.
#pragma once
#include <stdint.h>
#include <string>
#include <utility>
#include <vector>

namespace crtb {

struct ParsedAttr {
	std::string name;
	int codec = 0;
	float q = 0.f;
	int N = 0, format = 0, strategy = 0;
};

struct Block {                 // one entropy-coded block
	uint32_t probs_off = 0, nsym = 0, size = 0, csize = 0, data_off = 0;
	bool raw = false;
};

struct AttrStreams {
	uint32_t bits_off = 0, bits_nwords = 0;
	std::vector<Block> blocks; // 1 (CORRELATED / normal) or N
	int prediction = 0;        // normals
	int qc[4] = {4, 4, 4, 8};  // colour (ctor defaults color_attribute.h:31-34)
};

typedef std::vector<std::pair<std::string, std::string>> Props;

struct ParsedMesh {
	const uint8_t *blob = nullptr;
	uint32_t len = 0;
	uint32_t version = 0;
	int entropy = 1;
	Props exif;                        // std::map order == wire order (sorted by key)
	std::vector<ParsedAttr> attrs;     // wire order == std::map<std::string,...> order (sorted by name)
	uint32_t nvert = 0, nface = 0;
	uint32_t body = 0;                 // offset of the group table
	// filled by walk_directory
	std::vector<uint32_t> group_ends;
	std::vector<Props> group_props;
	uint32_t max_front = 0;
	Block clers;
	uint32_t split_off = 0, split_nwords = 0;
	std::vector<AttrStreams> streams;  // one per attr
	bool walked = false;
	// walk tapes (crt_walk.cpp: Cur): `record` collects the bytes the two functions below read; with `tape` set and blob ==
	// nullptr they read from the tape instead (the blob itself lives in device memory only)
	std::vector<uint8_t> *record = nullptr;
	const uint8_t *tape = nullptr;
	uint32_t tape_len = 0, tape_body = 0;
	int find(const char *name) const;
};

// Both return CRT_OK or a negative CRT_E_* code with a message in err.
int parse_header(const uint8_t *blob, int len, ParsedMesh &m, std::string &err);
int walk_directory(ParsedMesh &m, std::string &err);

}  // namespace crtb
