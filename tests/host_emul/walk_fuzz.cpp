// tests/host_emul/walk_fuzz.cpp — the host directory walk (corto_b200/csrc/crt_walk.cpp) under AddressSanitizer / UBSan:
// every truncation and a few thousand random byte edits of a fixture, each in a heap buffer of EXACTLY the blob's size, so a
// single byte read past the end aborts the run.  The reference has no bounds checks at all; this is the evidence that
// the replacement's walk has them.  Driven by tests/test_host.py::test_walk_under_sanitizers.  Not part of the product.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../corto_b200/csrc/crt_walk.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 32); }

static int try_blob(const uint8_t *src, size_t n) {
	uint8_t *buf = (uint8_t *)malloc(n ? n : 1);      // exact size: ASan traps the first out-of-bounds byte
	memcpy(buf, src, n);
	crtb::ParsedMesh m;
	std::string err;
	int rc = crtb::parse_header(buf, (int)n, m, err);
	if(rc == 0) rc = crtb::walk_directory(m, err);
	free(buf);
	return rc;
}

// The same walk REPLAYED from a tape (crt_walk.cpp: Cur, replay mode — how a rank that holds the blob in device memory only gets
// its directory): the tape is in a heap buffer of exactly its size, the blob pointer is null.
static int try_tape(const uint8_t *tape, size_t tn, size_t blob_len, crtb::ParsedMesh *out = nullptr) {
	uint8_t *buf = (uint8_t *)malloc(tn ? tn : 1);
	memcpy(buf, tape, tn);
	crtb::ParsedMesh m;
	m.tape = buf; m.tape_len = (uint32_t)tn;
	std::string err;
	int rc = crtb::parse_header(nullptr, (int)blob_len, m, err);
	if(rc == 0) rc = crtb::walk_directory(m, err);
	if(out && rc == 0) { *out = m; out->tape = nullptr; }
	free(buf);
	return rc;
}

static bool same_directory(const crtb::ParsedMesh &a, const crtb::ParsedMesh &b) {
	if(a.nvert != b.nvert || a.nface != b.nface || a.max_front != b.max_front || a.split_off != b.split_off || a.split_nwords != b.split_nwords) return false;
	if(a.group_ends != b.group_ends || a.group_props != b.group_props || a.exif != b.exif || a.streams.size() != b.streams.size()) return false;
	if(a.clers.data_off != b.clers.data_off || a.clers.csize != b.clers.csize || a.clers.size != b.clers.size) return false;
	for(size_t i = 0; i < a.streams.size(); i++) {
		const crtb::AttrStreams &x = a.streams[i], &y = b.streams[i];
		if(x.bits_off != y.bits_off || x.bits_nwords != y.bits_nwords || x.prediction != y.prediction || x.blocks.size() != y.blocks.size()) return false;
		for(size_t k = 0; k < x.blocks.size(); k++)
			if(x.blocks[k].data_off != y.blocks[k].data_off || x.blocks[k].csize != y.blocks[k].csize || x.blocks[k].size != y.blocks[k].size ||
			   x.blocks[k].probs_off != y.blocks[k].probs_off || x.blocks[k].nsym != y.blocks[k].nsym) return false;
	}
	return true;
}

int main(int argc, char **argv) {
	if(argc < 3) return 2;
	FILE *f = fopen(argv[1], "rb");
	if(!f) return 2;
	std::vector<uint8_t> blob;
	uint8_t tmp[65536];
	size_t k;
	while((k = fread(tmp, 1, sizeof tmp, f)) > 0) blob.insert(blob.end(), tmp, tmp + k);
	fclose(f);
	const int trials = atoi(argv[2]);
	const size_t n = blob.size();
	if(try_blob(blob.data(), n) != 0) { fprintf(stderr, "the intact blob was rejected\n"); return 1; }
	long accepted = 0, rejected = 0;
	for(size_t cut = 0; cut < n; cut += (cut < 256 ? 1 : 1 + n/509)) {             // truncations: none may be accepted
		if(try_blob(blob.data(), cut) == 0) { fprintf(stderr, "truncation at %zu of %zu accepted\n", cut, n); return 1; }
		rejected++;
	}
	std::vector<uint8_t> bad;
	for(int t = 0; t < trials; t++) {
		bad = blob;
		const int edits = 1 + (int)(rnd() % 3);
		for(int e = 0; e < edits; e++) {
			const size_t pos = (rnd() % 10 < 7) ? rnd() % (n < 256 ? n : 256) : rnd() % n;
			const uint32_t how = rnd() % 4;
			bad[pos] = how == 0 ? 0x00 : how == 1 ? 0xFF : (uint8_t)rnd();
		}
		if(try_blob(bad.data(), n) == 0) accepted++; else rejected++;
	}
	// ---- walk tapes: record on the intact blob, replay without it; truncated and edited tapes must fail or stay in bounds
	{
		crtb::ParsedMesh direct, replayed;
		std::vector<uint8_t> tape;
		std::string err;
		direct.record = &tape;
		if(crtb::parse_header(blob.data(), (int)n, direct, err) || crtb::walk_directory(direct, err)) { fprintf(stderr, "recording walk failed\n"); return 1; }
		if(try_tape(tape.data(), tape.size(), n, &replayed) != 0 || !same_directory(direct, replayed)) { fprintf(stderr, "tape replay differs from the direct walk\n"); return 1; }
		for(size_t cut = 0; cut < tape.size(); cut++)
			if(try_tape(tape.data(), cut, n) == 0) { fprintf(stderr, "tape truncated at %zu of %zu accepted\n", cut, tape.size()); return 1; }
		long tacc = 0, trej = 0;
		std::vector<uint8_t> bt;
		for(int t = 0; t < trials; t++) {
			bt = tape;
			const int edits = 1 + (int)(rnd() % 3);
			for(int e = 0; e < edits; e++) { const uint32_t how = rnd() % 4; bt[rnd() % bt.size()] = how == 0 ? 0x00 : how == 1 ? 0xFF : (uint8_t)rnd(); }
			if(try_tape(bt.data(), bt.size(), (rnd() & 1) ? n : (size_t)(rnd() % (n + 1))) == 0) tacc++; else trej++;
		}
		printf("tape %zu bytes: accepted %ld rejected %ld\n", tape.size(), tacc, trej);
	}
	printf("accepted %ld rejected %ld\n", accepted, rejected);
	return 0;
}
