// include/corto_b200/decoder.h — source-compatible C++ facade: the reference's `crt::Decoder` / `crt::VertexAttribute` /
// `crt::NormalAttr` / `crt::ColorAttr` surface (include/corto/decoder.h:38-73, vertex_attribute.h:28-67,
// normal_attribute.h:38-122, color_attribute.h:26-60) implemented over the C ABI of libcorto_b200.so.
//
//   #include <corto_b200/decoder.h>          // instead of <corto/decoder.h>
//   crt::Decoder decoder(size, data);         // header parse on the host
//   decoder.setPositions(coords); decoder.setNormals(normals); decoder.setIndex(index);
//   decoder.decode();                         // blob H2D -> CUDA kernels -> outputs D2H, synchronous
//
// Same names, same argument meaning, same error behaviour: failures throw `const char *` like the reference
// (src/decoder.cpp:44,51,274).  What is NOT here: the encoder side (quantize/encode members) and the CPU decode virtuals
// (decode / deltaDecode / postDelta / dequantize, vertex_attribute.h:48-66): attribute objects are plain descriptors, the
// decode runs on the device with the codec the stream names.  setAttribute(name, buf, attr*) keeps the reference's ownership
// rule; an object whose codec() differs from the stream's (a user subclass with its own decode) cannot be honoured and is
// refused with a thrown message instead of being silently ignored.
// The same header is reachable under the reference's own include names (include/corto/decoder.h, corto.h, ...), so user
// code compiles unchanged with -I<this repo>/include.
#ifndef CORTO_B200_DECODER_H
#define CORTO_B200_DECODER_H

#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#include <map>
#include <string>
#include <vector>

#include "../corto_b200.h"

typedef unsigned char uchar;

namespace crt {

struct Face { uint32_t a, b, c; Face() {} Face(uint32_t v0, uint32_t v1, uint32_t v2): a(v0), b(v1), c(v2) {} };

// The small fixed-size vectors the attribute statics take and return (include/corto/point.h:29-185: Point2s / Point2i /
// Point2f / Point3s / Point3i / Point3f); only what a decoder-side caller touches: construction, operator[], norm().
template <typename S, int D> class PointN {
	S v[D];
public:
	PointN() {}
	PointN(S x, S y) { static_assert(D == 2, "two components"); v[0] = x; v[1] = y; }
	PointN(S x, S y, S z) { static_assert(D == 3, "three components"); v[0] = x; v[1] = y; v[2] = z; }
	explicit PointN(const S *x) { for(int k = 0; k < D; k++) v[k] = x[k]; }
	S &operator[](int k) { return v[k]; }
	const S &operator[](int k) const { return v[k]; }
	bool operator==(const PointN &o) const { for(int k = 0; k < D; k++) if(v[k] != o.v[k]) return false; return true; }
	bool operator!=(const PointN &o) const { return !(*this == o); }
	S norm() const { S s = v[0]*v[0]; for(int k = 1; k < D; k++) s = s + v[k]*v[k]; return (S)sqrt((double)s); }   // point.h:59,111
};
typedef PointN<int16_t, 2> Point2s;
typedef PointN<int32_t, 2> Point2i;
typedef PointN<float, 2> Point2f;
typedef PointN<int16_t, 3> Point3s;
typedef PointN<int32_t, 3> Point3i;
typedef PointN<float, 3> Point3f;

struct Group {                                   // include/corto/index_attribute.h:40-46
	uint32_t end;
	std::map<std::string, std::string> properties;
	Group(): end(0) {}
	Group(uint32_t e): end(e) {}
};

class IndexAttribute {                           // index_attribute.h:48-60 (decode-side members)
public:
	uint32_t *faces32;
	uint16_t *faces16;
	std::vector<Group> groups;
	uint32_t max_front;
	IndexAttribute(): faces32(nullptr), faces16(nullptr), max_front(0) {}
};

class VertexAttribute {                          // vertex_attribute.h:28-46
public:
	enum Format { UINT32 = 0, INT32, UINT16, INT16, UINT8, INT8, FLOAT, DOUBLE };
	enum Strategy { PARALLEL = 0x1, CORRELATED = 0x2 };
	enum CODEC { GENERIC_CODEC = 1, NORMAL_CODEC = 2, COLOR_CODEC = 3, CUSTOM_CODEC = 100 };
	char *buffer;
	int N;
	float q;
	int strategy;
	Format format;
	uint32_t size;
	int bits;
	VertexAttribute(): buffer(nullptr), N(0), q(0.0f), strategy(0), format(INT32), size(0), bits(0) {}
	virtual ~VertexAttribute() {}
	virtual int codec() = 0;
};

template <class T> class GenericAttr: public VertexAttribute {   // vertex_attribute.h:71-77
public:
	GenericAttr(int dim) { N = dim; }
	virtual int codec() { return GENERIC_CODEC; }
};

class NormalAttr: public VertexAttribute {       // normal_attribute.h:38-58
public:
	enum Prediction { DIFF = 0x0, ESTIMATED = 0x1, BORDER = 0x2 };
	uint32_t prediction;
	NormalAttr(int bits = 10) { N = 3; q = powf(2.0f, (float)(bits - 1)); prediction = DIFF; strategy |= VertexAttribute::CORRELATED; }
	virtual int codec() { return NORMAL_CODEC; }

	// Octahedral mapping, host side (normal_attribute.h:75-122) — the same arithmetic the device kernels emulate (crt_device.cuh:
	// to_octa / to_sphere): fp32 throughout, one divisor for both components, truncating float -> int conversions.
	static Point2i toOcta(Point3f n, int unit) {                       // normal_attribute.h:75-85
		const float l1 = fabsf(n[0]) + fabsf(n[1]) + fabsf(n[2]);
		float x = n[0]/l1, y = n[1]/l1;
		if(n[2] < 0) {
			const float fx = 1.0f - fabsf(y), fy = 1.0f - fabsf(x);
			x = n[0] < 0 ? -fx : fx;
			y = n[1] < 0 ? -fy : fy;
		}
		return Point2i((int)(x*unit), (int)(y*unit));
	}
	static Point2i toOcta(Point3i n, int unit) {                       // normal_attribute.h:87-102 (integer variant)
		const int l1 = abs(n[0]) + abs(n[1]) + abs(n[2]);
		if(l1 == 0) return Point2i(0, 0);
		int x = n[0]*unit/l1, y = n[1]*unit/l1;
		if(n[2] < 0) {
			const int fx = (int)(unit - fabs(y)), fy = (int)(unit - fabs(x));
			x = n[0] < 0 ? -fx : fx;
			y = n[1] < 0 ? -fy : fy;
		}
		return Point2i(x, y);
	}
	static Point3f toSphere(Point2i v, int unit) {                     // normal_attribute.h:104-112
		float n[3];
		octa_to_vector(v[0], v[1], unit, n);
		return Point3f(n[0], n[1], n[2]);
	}
	static Point3s toSphere(Point2s v, int unit) {                     // normal_attribute.h:114-122
		float n[3];
		octa_to_vector(v[0], v[1], unit, n);
		return Point3s((int16_t)(n[0]*32767), (int16_t)(n[1]*32767), (int16_t)(n[2]*32767));
	}
private:
	static void octa_to_vector(int vx, int vy, int unit, float n[3]) {
		n[0] = (float)vx; n[1] = (float)vy; n[2] = (float)(unit - abs(vx) - abs(vy));
		if(n[2] < 0) {                                                 // lower hemisphere: fold back; sgn(0) counts as -1
			n[0] = (float)((vx > 0 ? 1 : -1)*(unit - abs(vy)));
			n[1] = (float)((vy > 0 ? 1 : -1)*(unit - abs(vx)));
		}
		const float len = (float)sqrt((double)(n[0]*n[0] + n[1]*n[1] + n[2]*n[2]));
		n[0] /= len; n[1] /= len; n[2] /= len;
	}
};

class ColorAttr: public GenericAttr<uchar> {     // color_attribute.h:26-42
public:
	int qc[4];
	int out_components;
	ColorAttr(int components = 4): GenericAttr<uchar>(components), out_components(4) { qc[0] = qc[1] = qc[2] = 4; qc[3] = 8; }
	virtual int codec() { return COLOR_CODEC; }
	void setQ(int r_bits, int g_bits, int b_bits, int a_bits) {
		qc[0] = 1 << (8 - r_bits); qc[1] = 1 << (8 - g_bits); qc[2] = 1 << (8 - b_bits); qc[3] = 1 << (8 - a_bits);
	}
};

class Decoder {                                  // include/corto/decoder.h:38-73
public:
	uint32_t nvert, nface;
	std::map<std::string, std::string> exif;
	std::map<std::string, VertexAttribute *> data;
	IndexAttribute index;

	Decoder(int len, const uchar *input): nvert(0), nface(0), h(nullptr) {
		h = crt_new_decoder(len, input);
		if(!h) throw_last();
		nvert = crt_nvert(h); nface = crt_nface(h);
		for(int i = 0; i < crt_nexif(h); i++) { const char *v = nullptr; const char *k = crt_exif(h, i, &v); exif[k] = v; }
		for(int i = 0; i < crt_nattr(h); i++) {
			int codec, comps, fmt, strat; float q;
			const char *name = crt_attr_info(h, i, &codec, &q, &comps, &fmt, &strat);
			VertexAttribute *attr;
			switch(codec) {                      // src/decoder.cpp:72-81
			case VertexAttribute::NORMAL_CODEC: attr = new NormalAttr(); break;
			case VertexAttribute::COLOR_CODEC: attr = new ColorAttr(comps); break;
			default: attr = new GenericAttr<int>(comps);
			}
			attr->q = q; attr->format = (VertexAttribute::Format)fmt; attr->strategy = strat;
			data[name] = attr;
		}
	}
	~Decoder() {
		for(auto it: data) delete it.second;
		if(h) crt_delete_decoder(h);
	}

	bool hasAttr(const char *name) { return data.count(name) != 0; }
	bool setPositions(float *buffer) { return setAttribute("position", (char *)buffer, VertexAttribute::FLOAT); }
	bool setNormals(float *buffer)   { return setAttribute("normal", (char *)buffer, VertexAttribute::FLOAT); }
	bool setNormals(int16_t *buffer) { return setAttribute("normal", (char *)buffer, VertexAttribute::INT16); }
	bool setUvs(float *buffer)       { return setAttribute("uv", (char *)buffer, VertexAttribute::FLOAT); }
	bool setColors(uchar *buffer, int components = 4) {                       // src/decoder.cpp:116-123
		if(data.find("color") == data.end()) return false;
		ColorAttr *attr = dynamic_cast<ColorAttr *>(data["color"]);
		attr->format = VertexAttribute::UINT8; attr->buffer = (char *)buffer; attr->out_components = components;
		return true;
	}
	bool setAttribute(const char *name, char *buffer, VertexAttribute::Format format) {   // src/decoder.cpp:96-102
		if(data.find(name) == data.end()) return false;
		VertexAttribute *attr = data[name];
		attr->format = format; attr->buffer = buffer;
		return true;
	}
	bool setAttribute(const char *name, char *buffer, VertexAttribute *attr) {            // src/decoder.cpp:104-114 (takes ownership)
		if(data.find(name) == data.end()) return false;
		VertexAttribute *found = data[name];
		if(attr->codec() != found->codec()) {                                             // a user codec: the device cannot run its virtuals
			delete attr;                                                                  // (ownership was transferred)
#ifndef NO_EXCEPTIONS
			throw "corto_b200: custom attribute codecs are not supported (the decode runs on the device with the stream's codec)";
#else
			return false;
#endif
		}
		attr->q = found->q; attr->strategy = found->strategy; attr->N = found->N; attr->buffer = buffer;
		delete data[name];
		data[name] = attr;
		return true;
	}
	void setIndex(uint32_t *buffer) { index.faces32 = buffer; }
	void setIndex(uint16_t *buffer) { index.faces16 = buffer; }

	void decode() {                                                           // src/decoder.cpp:126-131
		for(auto it: data) {
			VertexAttribute *a = it.second;
			if(!a->buffer) continue;
			ColorAttr *c = dynamic_cast<ColorAttr *>(a);
			if(c) crt_set_colors(h, (uchar *)a->buffer, c->out_components);
			else crt_set_attribute(h, it.first.c_str(), a->buffer, (int)a->format);
		}
		if(index.faces16) crt_set_index16(h, index.faces16);
		else if(index.faces32) crt_set_index32(h, index.faces32);
		if(crt_decode(h) != CRT_OK) throw_last();
		// what the stream carried (reference: filled in by the attribute decode() methods)
		for(auto it: data) {
			NormalAttr *n = dynamic_cast<NormalAttr *>(it.second);
			if(n) n->prediction = (uint32_t)crt_normal_prediction(h);
			ColorAttr *c = dynamic_cast<ColorAttr *>(it.second);
			if(c) crt_color_q(h, c->qc);
		}
		int ng = crt_ngroups(h);
		std::vector<int> ends(ng > 0 ? ng : 1);
		crt_groups(h, ends.data());
		index.groups.clear();
		for(int g = 0; g < ng; g++) {
			Group grp((uint32_t)ends[g]);
			for(int i = 0; i < crt_group_nprops(h, g); i++) { const char *v = nullptr; const char *k = crt_group_prop(h, g, i, &v); grp.properties[k] = v; }
			index.groups.push_back(grp);
		}
	}

private:
	crt_decoder *h;
	Decoder(const Decoder &);
	Decoder &operator=(const Decoder &);
	static void throw_last() {
		static thread_local std::string msg;      // the reference throws string literals; keep the pointer alive
		msg = crt_last_error();
#ifndef NO_EXCEPTIONS
		throw msg.c_str();
#endif
	}
};

}  // namespace crt
#endif  // CORTO_B200_DECODER_H
