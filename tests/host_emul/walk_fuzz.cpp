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
	printf("accepted %ld rejected %ld\n", accepted, rejected);
	return 0;
}
