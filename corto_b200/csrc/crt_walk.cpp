// crt_walk.cpp — see crt_walk.h.  Wire format: SURVEY §8.0.
#include "crt_walk.h"
#include "crt_common.h"
#include "../../include/corto_b200.h"
#include <string.h>
#include <algorithm>
#include <vector>

namespace crtb {

namespace {
// Byte cursor over a blob.  Three modes: direct (host blob), recording (host blob; every byte the walk READS — never the
// payload it skips — is appended to a tape), replay (no blob: the reads are served from such a tape while `p` tracks the
// position in the blob the tape came from).  The walk is a deterministic function of the bytes it reads, so a replay visits
// the same positions and yields the same directory: that is how a rank that holds a blob only in DEVICE memory gets its
// directory without copying the payload back (crt_walk_tape / crt_batch_create_device).
struct Cur {
	const uint8_t *b; uint32_t len; uint32_t p; bool bad;
	std::vector<uint8_t> *rec = nullptr;          // recording
	const uint8_t *tape = nullptr; uint32_t tape_len = 0, tape_p = 0;   // replay (b == nullptr)
	bool need(uint64_t n) { if(bad || (uint64_t)p + n > len) { bad = true; return false; } return true; }
	// the next n bytes at p (n <= 65535): pointer valid until the next call
	const uint8_t *rd(uint32_t n) {
		if(!need(n)) return nullptr;
		const uint8_t *src;
		if(b) { src = b + p; if(rec) rec->insert(rec->end(), src, src + n); }
		else { if((uint64_t)tape_p + n > tape_len) { bad = true; return nullptr; } src = tape + tape_p; tape_p += n; }
		p += n;
		return src;
	}
	uint32_t u8() { const uint8_t *s = rd(1); return s ? s[0] : 0; }
	uint32_t u16() { const uint8_t *s = rd(2); return s ? (uint32_t)(s[0] | (s[1] << 8)) : 0; }                 // cstream.h:250-257
	uint32_t u32() { const uint8_t *s = rd(4); return s ? (s[0] | (s[1] << 8) | (s[2] << 16) | ((uint32_t)s[3] << 24)) : 0; }  // :259-270
	std::string str() {                                                                                       // :277-280: u16 length incl. NUL
		uint32_t n = u16();
		const uint8_t *s = rd(n);
		if(!s) return std::string();
		return std::string((const char *)s, strnlen((const char *)s, n));
	}
	void skip(uint64_t n) { if(need(n)) p += (uint32_t)n; }
	bool bits(uint32_t &off, uint32_t &nwords) {                                                              // :283-291
		nwords = u32();
		if(p & 3) skip(4 - (p & 3));
		off = p;
		skip((uint64_t)nwords*4);
		return !bad;
	}
};

int entropy_block(Cur &c, int entropy, Block &blk, std::string &err) {
	if(entropy == 0) {                    // Stream::NONE, cstream.cpp:68-73
		blk.raw = true; blk.nsym = 0;
		blk.size = blk.csize = c.u32();
		blk.data_off = c.p;
		c.skip(blk.size);
	} else if(entropy == 1) {             // Stream::TUNSTALL, cstream.cpp:111-128
		blk.raw = false;
		blk.nsym = c.u8();
		blk.probs_off = c.p;
		c.skip(2*blk.nsym);
		blk.size = c.u32();
		blk.csize = c.u32();
		blk.data_off = c.p;
		c.skip(blk.csize);
		if(!c.bad && blk.size && ((blk.nsym == 0) || (blk.nsym > 1 && blk.csize == 0))) { err = "corrupt entropy block"; return CRT_E_TRUNCATED; }
	} else { err = "Unknown entropy"; return CRT_E_ENTROPY; }
	if(c.bad) { err = "blob truncated inside an entropy block"; return CRT_E_TRUNCATED; }
	return CRT_OK;
}
}  // namespace

int ParsedMesh::find(const char *name) const {
	for(size_t i = 0; i < attrs.size(); i++) if(attrs[i].name == name) return (int)i;
	return -1;
}

int parse_header(const uint8_t *blob, int len, ParsedMesh &m, std::string &err) {
	if((!blob && !m.tape) || len < 0) { err = "null blob"; return CRT_E_ARG; }
	if((uintptr_t)blob & 3) { err = "Memory must be alignegned on 4 bytes."; return CRT_E_ALIGN; }      // decoder.cpp:43-44 (sic)
	Cur c{blob, (uint32_t)len, 0, false};
	c.rec = m.record; c.tape = m.tape; c.tape_len = m.tape_len;
	uint32_t magic = c.u32();
	if(c.bad || magic != 0x787A6300u) { err = "Not a crt file."; return CRT_E_MAGIC; }                    // decoder.cpp:48-52
	m.blob = blob; m.len = (uint32_t)len;
	m.version = c.u32();
	m.entropy = (int)c.u8();
	uint32_t nexif = c.u32();
	for(uint32_t i = 0; i < nexif && !c.bad; i++) {
		std::string k = c.str(), v = c.str();
		// std::map semantics (decoder.cpp:58-59): last value for a repeated key wins
		bool dup = false;
		for(auto &kv: m.exif) if(kv.first == k) { kv.second = v; dup = true; }
		if(!dup) m.exif.push_back({k, v});
	}
	uint32_t nattr = c.u32();
	if(!c.bad && nattr > (uint32_t)MAX_ATTR) { err = "too many attributes for this build"; return CRT_E_LIMIT; }
	for(uint32_t i = 0; i < nattr && !c.bad; i++) {
		ParsedAttr a;
		a.name = c.str();
		a.codec = (int)c.u32();
		uint32_t qb = c.u32(); memcpy(&a.q, &qb, 4);
		a.N = (int)c.u8(); a.format = (int)c.u8(); a.strategy = (int)c.u8();
		if(a.codec != CODEC_NORMAL && a.codec != CODEC_COLOR) a.codec = CODEC_GENERIC;                     // decoder.cpp:76-79 default branch
		m.attrs.push_back(a);
	}
	// The reference keeps its attributes in a std::map<std::string, ...> (decoder.cpp:72-86): every pass, the stream walk
	// included, visits them in byte-wise name order (decoder.cpp:168), and a repeated name keeps only its LAST header entry.
	// Encoder-written files are sorted already; a hand-built header gets the same treatment here.
	{
		std::vector<ParsedAttr> uniq;
		for(const ParsedAttr &a: m.attrs) {
			bool dup = false;
			for(ParsedAttr &u: uniq) if(u.name == a.name) { u = a; dup = true; }
			if(!dup) uniq.push_back(a);
		}
		std::stable_sort(uniq.begin(), uniq.end(), [](const ParsedAttr &x, const ParsedAttr &y) { return x.name < y.name; });
		m.attrs.swap(uniq);
	}
	m.nvert = c.u32();
	m.nface = c.u32();
	m.body = c.p; m.tape_body = c.tape_p;
	if(c.bad) { err = "blob truncated inside the header"; return CRT_E_TRUNCATED; }
	for(auto &a: m.attrs) {
		int nc = a.codec == CODEC_NORMAL ? 2 : a.N;
		if(nc < 1 || nc > MAX_COMP) { err = "attribute '" + a.name + "' has an unsupported component count"; return CRT_E_LIMIT; }
	}
	return CRT_OK;
}

int walk_directory(ParsedMesh &m, std::string &err) {
	Cur c{m.blob, m.len, m.body, false};
	c.rec = m.record; c.tape = m.tape; c.tape_len = m.tape_len; c.tape_p = m.tape_body;
	m.group_ends.clear(); m.group_props.clear(); m.streams.clear();
	uint32_t ngroups = c.u32();                                 // index_attribute.h:89-99
	if(!c.bad && (uint64_t)ngroups*5 > m.len) { err = "corrupt group table"; return CRT_E_TRUNCATED; }
	for(uint32_t g = 0; g < ngroups && !c.bad; g++) {
		m.group_ends.push_back(c.u32());
		uint32_t np = c.u8();
		Props props;
		for(uint32_t k = 0; k < np && !c.bad; k++) { std::string key = c.str(), val = c.str(); props.push_back({key, val}); }
		m.group_props.push_back(props);
	}
	if(c.bad) { err = "blob truncated inside the group table"; return CRT_E_TRUNCATED; }
	if(m.nface > 0) {                                           // index_attribute.h:83-87
		m.max_front = c.u32();
		int rc = entropy_block(c, m.entropy, m.clers, err);
		if(rc) return rc;
		if(!c.bits(m.split_off, m.split_nwords)) { err = "blob truncated inside the split bitstream"; return CRT_E_TRUNCATED; }
	}
	for(const ParsedAttr &a: m.attrs) {                         // decoder.cpp:168-169 (map order)
		AttrStreams s;
		int nblocks;
		if(a.codec == CODEC_NORMAL) { s.prediction = (int)c.u8(); nblocks = 1; }                           // normal_attribute.cpp:179-181
		else if(a.codec == CODEC_COLOR) { for(int k = 0; k < a.N; k++) s.qc[k] = (int)c.u8(); nblocks = a.N; }   // color_attribute.h:55-58
		else nblocks = (a.strategy & S_CORRELATED) ? 1 : a.N;                                           // vertex_attribute.h:153-158
		if(!c.bits(s.bits_off, s.bits_nwords)) { err = "blob truncated inside a bitstream"; return CRT_E_TRUNCATED; }
		for(int k = 0; k < nblocks; k++) {
			Block b;
			int rc = entropy_block(c, m.entropy, b, err);
			if(rc) return rc;
			if(b.size > m.nvert) { err = "attribute '" + a.name + "' carries more values than vertices"; return CRT_E_TRUNCATED; }
			s.blocks.push_back(b);
		}
		m.streams.push_back(s);
	}
	m.walked = true;
	return CRT_OK;
}

}  // namespace crtb
