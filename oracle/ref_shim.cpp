// oracle/ref_shim.cpp — TEST INFRASTRUCTURE, not product code.
//
// A thin extern "C" shim over the UNMODIFIED reference sources, compiled where they lie under
// /root/reference by oracle/Makefile into oracle/_ref/libcorto_ref.so (git-ignored, travels to the
// GPU box).  It is the ground truth the CUDA path and the C restatement (oracle/crt_oracle.c) are
// pinned against, the fixture generator (reference Encoder), and the CPU baseline that bench.py
// times (`--impl reference`, `cpu_baseline.kind = "reference"`).
//
// Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may load it.
//
// Reference entry points used (all under /root/reference):
//   crt::Encoder  include/corto/encoder.h:36-97   (add*, addGroup, encode)
//   crt::Decoder  include/corto/decoder.h:38-73   (ctor, set*, decode)
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>
#include <atomic>
#include <map>
#include <string>

#include "encoder.h"
#include "decoder.h"

using namespace crt;

extern "C" {

void ref_free(void *p) { free(p); }

// Encode one mesh / point cloud with the reference Encoder.  Any optional array may be NULL.
// pos_bits>0 -> addPositionsBits, else addPositions(q=pos_q).  Returns a malloc'ed 16-byte aligned
// blob (caller ref_free()s it) or NULL on error (message in err, if given).
unsigned char *ref_encode(uint32_t nvert, uint32_t nface,
                          const float *pos, const uint32_t *faces,
                          int pos_bits, float pos_q,
                          const float *uv, float uv_q,
                          const float *normals, int normal_bits, int normal_pred,
                          const unsigned char *colors, int color_comps, const int *color_bits,
                          const float *radius, float radius_q, int radius_strategy,
                          const int *group_ends, int ngroups,
                          int entropy,
                          uint32_t *out_len, uint32_t *out_nvert, uint32_t *out_nface,
                          char *err, int errlen) {
	try {
		Encoder enc(nvert, nface, (Stream::Entropy)entropy);
		for(int g = 0; g < ngroups; g++) {
			std::map<std::string, std::string> props;
			if(g & 1) props["material"] = "m" + std::to_string(g);
			enc.addGroup(group_ends[g], props);
		}
		if(nface) {
			if(pos_bits > 0) enc.addPositionsBits(pos, (uint32_t *)faces, pos_bits);
			else enc.addPositions(pos, faces, pos_q);
		} else {
			if(pos_bits > 0) enc.addPositionsBits(pos, pos_bits);
			else enc.addPositions(pos, pos_q);
		}
		if(normals) enc.addNormals(normals, normal_bits, (NormalAttr::Prediction)normal_pred);
		if(colors) {
			if(color_comps == 3) enc.addColors3(colors, color_bits[0], color_bits[1], color_bits[2]);
			else enc.addColors(colors, color_bits[0], color_bits[1], color_bits[2], color_bits[3]);
		}
		if(uv) enc.addUvs(uv, uv_q);
		if(radius) enc.addAttribute("radius", (const char *)radius, VertexAttribute::FLOAT, 1, radius_q, radius_strategy);
		enc.encode();
		uint32_t len = enc.stream.size();
		unsigned char *blob = (unsigned char *)aligned_alloc(16, (len + 15) & ~15u);
		memcpy(blob, enc.stream.data(), len);
		*out_len = len;
		*out_nvert = enc.nvert;
		*out_nface = enc.nface;
		return blob;
	} catch(const char *e) {
		if(err) { strncpy(err, e, errlen - 1); err[errlen - 1] = 0; }
		return NULL;
	}
}

// Header info via the reference Decoder ctor.  attr_mask bit0 position, bit1 normal, bit2 color, bit3 uv, bit4 radius.
int ref_info(const unsigned char *blob, int len, uint32_t *nvert, uint32_t *nface, int *attr_mask, int *color_comps) {
	try {
		Decoder dec(len, blob);
		*nvert = dec.nvert;
		*nface = dec.nface;
		int m = 0;
		if(dec.hasAttr("position")) m |= 1;
		if(dec.hasAttr("normal")) m |= 2;
		if(dec.hasAttr("color")) { m |= 4; *color_comps = dec.data["color"]->N; }
		if(dec.hasAttr("uv")) m |= 8;
		if(dec.hasAttr("radius")) m |= 16;
		*attr_mask = m;
		return 0;
	} catch(const char *) { return -1; }
}

static int decode_one(const unsigned char *blob, int len,
                      float *pos, void *index, int index16,
                      float *normals32, int16_t *normals16,
                      unsigned char *colors, int color_comps,
                      float *uv, float *radius,
                      unsigned char *clers_out, uint32_t *nclers, uint32_t *prediction_out, int *groups_out) {
	Decoder dec(len, blob);
	if(pos) dec.setPositions(pos);
	if(normals32) dec.setNormals(normals32);
	else if(normals16) dec.setNormals(normals16);
	if(colors) dec.setColors(colors, color_comps);
	if(uv) dec.setUvs(uv);
	if(radius) dec.setAttribute("radius", (char *)radius, VertexAttribute::FLOAT);
	if(index && dec.nface) {
		if(index16) dec.setIndex((uint16_t *)index);
		else dec.setIndex((uint32_t *)index);
	}
	dec.decode();
	if(nclers) *nclers = (uint32_t)dec.index.clers.size();
	if(clers_out) memcpy(clers_out, dec.index.clers.data(), dec.index.clers.size());
	if(prediction_out && dec.nface) memcpy(prediction_out, dec.index.prediction.data(), dec.index.prediction.size()*12);
	if(groups_out) for(size_t g = 0; g < dec.index.groups.size(); g++) groups_out[g] = dec.index.groups[g].end;
	return (int)dec.index.groups.size();
}

// Decode with the reference Decoder into caller-allocated buffers (NULL = leave that attribute unbound).
// Returns number of groups, or -1 on a reference exception.
int ref_decode(const unsigned char *blob, int len,
               float *pos, void *index, int index16,
               float *normals32, int16_t *normals16,
               unsigned char *colors, int color_comps,
               float *uv, float *radius) {
	try {
		return decode_one(blob, len, pos, index, index16, normals32, normals16, colors, color_comps, uv, radius, 0, 0, 0, 0);
	} catch(const char *) { return -1; }
}

// Same, plus intermediate pins (public members of crt::IndexAttribute): clers bytes, prediction triples, group ends.
int ref_decode_debug(const unsigned char *blob, int len,
                     float *pos, void *index, int index16,
                     float *normals32, int16_t *normals16,
                     unsigned char *colors, int color_comps,
                     float *uv, float *radius,
                     unsigned char *clers_out, uint32_t *nclers, uint32_t *prediction_out, int *groups_out) {
	try {
		return decode_one(blob, len, pos, index, index16, normals32, normals16, colors, color_comps, uv, radius,
		                  clers_out, nclers, prediction_out, groups_out);
	} catch(const char *) { return -1; }
}

// Generic attributes decoded into caller buffers with an explicit output Format (VertexAttribute::INT32 / UINT32 / FLOAT,
// vertex_attribute.h:184-230) through Decoder::setAttribute(name, buffer, format) (decoder.cpp:96-102).  NULL = unbound.
int ref_decode_fmt(const unsigned char *blob, int len, void *pos, int pos_fmt, void *uv, int uv_fmt, void *radius, int radius_fmt) {
	try {
		Decoder dec(len, blob);
		if(pos) dec.setAttribute("position", (char *)pos, (VertexAttribute::Format)pos_fmt);
		if(uv) dec.setAttribute("uv", (char *)uv, (VertexAttribute::Format)uv_fmt);
		if(radius) dec.setAttribute("radius", (char *)radius, (VertexAttribute::Format)radius_fmt);
		std::vector<uint32_t> index((size_t)dec.nface*3 + 3);        // decodeFaces writes the index unconditionally (decoder.cpp:246-249)
		if(dec.nface) dec.setIndex(index.data());
		dec.decode();
		return 0;
	} catch(const char *) { return -1; }
}

// CPU baseline: decode `n` blobs `repeats` times with `nthreads` host threads (one crt::Decoder per
// blob, as a user of the single-threaded reference would), all attributes bound (float normals,
// u32 index, colours with their own component count), outputs pre-allocated per thread and reused.
// Returns the best wall time of one full pass in seconds.
double ref_decode_bench(int n, const unsigned char *const *blobs, const int *lens, int nthreads, int repeats) {
	uint32_t maxv = 0, maxf = 0;
	for(int i = 0; i < n; i++) {
		Decoder d(lens[i], blobs[i]);
		if(d.nvert > maxv) maxv = d.nvert;
		if(d.nface > maxf) maxf = d.nface;
	}
	struct Bufs { std::vector<float> pos, nrm, uv, rad; std::vector<uint32_t> idx; std::vector<unsigned char> col; };
	std::vector<Bufs> bufs(nthreads);
	for(auto &b: bufs) {
		b.pos.resize(maxv*3); b.nrm.resize(maxv*3); b.uv.resize(maxv*2); b.rad.resize(maxv);
		b.idx.resize((size_t)maxf*3 + 3); b.col.resize(maxv*4);
	}
	double best = 1e30;
	for(int r = 0; r < repeats; r++) {
		std::atomic<int> next(0);
		auto t0 = std::chrono::steady_clock::now();
		std::vector<std::thread> th;
		for(int t = 0; t < nthreads; t++) {
			th.emplace_back([&, t]() {
				Bufs &b = bufs[t];
				for(;;) {
					int i = next.fetch_add(1);
					if(i >= n) break;
					Decoder dec(lens[i], blobs[i]);
					dec.setPositions(b.pos.data());
					if(dec.hasAttr("normal")) dec.setNormals(b.nrm.data());
					if(dec.hasAttr("color")) dec.setColors(b.col.data(), dec.data["color"]->N);
					if(dec.hasAttr("uv")) dec.setUvs(b.uv.data());
					if(dec.hasAttr("radius")) dec.setAttribute("radius", (char *)b.rad.data(), VertexAttribute::FLOAT);
					if(dec.nface) dec.setIndex(b.idx.data());
					dec.decode();
				}
			});
		}
		for(auto &t: th) t.join();
		double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		if(dt < best) best = dt;
	}
	return best;
}

} // extern "C"
