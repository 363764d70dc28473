// crt_device.cuh — the sequential cores of the decode path as __host__ __device__ functions, so the exact code
// the kernels run can also be exercised on the CPU by tests/host_emul (logic check without a GPU).  They are
// NOT exported from the product library as a CPU path; only kernels call them there.
//
// Semantics pinned here: no FMA contraction (explicit _rn intrinsics on the device),
// IEEE div / sqrt, x86 cvttss2si float->int (NaN / out of range -> INT_MIN), wrapping int32 arithmetic,
// x86's "indefinite" quiet NaN (0xFFC00000) as the result of 0/0.
#pragma once
#include <stdint.h>
#include <math.h>
#include "crt_common.h"

#ifdef __CUDACC__
#define CRT_HD __host__ __device__ __forceinline__
#else
#define CRT_HD inline
#endif

namespace crtb {

// ---- exact float helpers -------------------------------------------------------------------------------
CRT_HD float f_mul(float a, float b) {
#ifdef __CUDA_ARCH__
	return __fmul_rn(a, b);
#else
	return a*b;
#endif
}
CRT_HD float f_add(float a, float b) {
#ifdef __CUDA_ARCH__
	return __fadd_rn(a, b);
#else
	return a + b;
#endif
}
CRT_HD float f_sub(float a, float b) {
#ifdef __CUDA_ARCH__
	return __fsub_rn(a, b);
#else
	return a - b;
#endif
}
CRT_HD float f_bits(uint32_t u) {
#ifdef __CUDA_ARCH__
	return __uint_as_float(u);
#else
	float f; memcpy(&f, &u, 4); return f;
#endif
}
// divss: 0/0, inf/inf give the x86 default NaN (sign bit set); CUDA would give 0x7FFFFFFF.
CRT_HD float f_div(float a, float b) {
#ifdef __CUDA_ARCH__
	float r = __fdiv_rn(a, b);
#else
	float r = a/b;
#endif
	if(r != r) r = f_bits(0xFFC00000u);
	return r;
}
// Point3f::norm(), include/corto/point.h:111: fp32 sum left to right, sqrt through double == correctly rounded sqrtf.
CRT_HD float f_norm3(float x, float y, float z) {
	float s = f_add(f_add(f_mul(x, x), f_mul(y, y)), f_mul(z, z));
#ifdef __CUDA_ARCH__
	return __fsqrt_rn(s);
#else
	return (float)sqrt((double)s);
#endif
}
CRT_HD float i2f(int32_t v) {
#ifdef __CUDA_ARCH__
	return __int2float_rn(v);
#else
	return (float)v;
#endif
}
// cvttss2si
CRT_HD int32_t f2i_x86(float f) {
	if(!(f >= -2147483648.0f && f < 2147483648.0f)) return (int32_t)0x80000000;
	return (int32_t)f;
}
CRT_HD int16_t f2s_x86(float f) { return (int16_t)(uint16_t)(uint32_t)f2i_x86(f); }
// float -> uint32 as gcc x86-64 does it: cvttss2si r64, keep the low half
CRT_HD uint32_t f2u_x86(float f) {
	if(!(f >= -9223372036854775808.0f && f < 9223372036854775808.0f)) return 0u;
	return (uint32_t)(uint64_t)(int64_t)f;
}
CRT_HD int32_t iabs_wrap(int32_t v) { return v < 0 ? (int32_t)(0u - (uint32_t)v) : v; }
CRT_HD float f_abs(float v) { return fabsf(v); }

// NormalAttr::toOcta(Point3f,int) — include/corto/normal_attribute.h:75-85
CRT_HD void to_octa(float x, float y, float z, int unit, int32_t &ox, int32_t &oy) {
	float s = f_add(f_add(f_abs(x), f_abs(y)), f_abs(z));
	float px = f_div(x, s), py = f_div(y, s);
	if(z < 0) {
		float ax = f_sub(1.0f, f_abs(py)), ay = f_sub(1.0f, f_abs(px));
		px = ax; py = ay;
		if(x < 0) px = -px;
		if(y < 0) py = -py;
	}
	float u = i2f(unit);
	ox = f2i_x86(f_mul(px, u));
	oy = f2i_x86(f_mul(py, u));
}

// NormalAttr::toSphere(Point2i,int) — normal_attribute.h:104-112 (all int32 arithmetic wraps)
CRT_HD void to_sphere(int32_t vx, int32_t vy, int unit, float &nx, float &ny, float &nz) {
	int32_t z = (int32_t)((uint32_t)unit - (uint32_t)iabs_wrap(vx) - (uint32_t)iabs_wrap(vy));
	nx = i2f(vx); ny = i2f(vy); nz = i2f(z);
	if(nz < 0) {
		int32_t ax = (int32_t)((uint32_t)unit - (uint32_t)iabs_wrap(vy));
		int32_t ay = (int32_t)((uint32_t)unit - (uint32_t)iabs_wrap(vx));
		nx = i2f((vx > 0) ? ax : (int32_t)(0u - (uint32_t)ax));
		ny = i2f((vy > 0) ? ay : (int32_t)(0u - (uint32_t)ay));
	}
	float len = f_norm3(nx, ny, nz);
	nx = f_div(nx, len); ny = f_div(ny, len); nz = f_div(nz, len);
}

// ---- bit reader ---------------------------------------------------------------------------------------
// MSB-first n-bit field (0..32) at absolute bit `pos` of a little-endian u32 word array: the random-access
// form of BitStream::read (src/bitstream.cpp:103-121).  Words past `nwords` read as 0 (a corrupt stream must
// not fault); like the reference's lazy refill, the word after the last needed one is never touched.
CRT_HD uint32_t getbits(const uint32_t *w, uint32_t nwords, uint64_t pos, int n) {
	if(n <= 0) return 0;
	uint64_t i = pos >> 5;
	int o = (int)(pos & 31);
	uint32_t hi = i < nwords ? w[i] : 0u;
	if(o + n <= 32) return (hi << o) >> (32 - n);
	uint32_t lo = (i + 1) < nwords ? w[i + 1] : 0u;
	uint64_t win = ((uint64_t)hi << 32) | lo;
	return (uint32_t)((win << o) >> (64 - n));
}

// `(1<<diff)>>1` with an int shift as x86 evaluates it (include/corto/cstream.h:344):
// diff=31 -> 0xC0000000, diff=32 -> 0.
CRT_HD uint32_t array_bias(int d) {
	int32_t one = (int32_t)(1u << (d & 31));
	return (uint32_t)(one >> 1);
}
// decodeValues sign fold, cstream.h:309-315
CRT_HD int32_t fold_value(uint32_t raw, int d) {
	int32_t val = (int32_t)raw;
	int32_t middle = (int32_t)(1u << ((d - 1) & 31));
	if(val < middle) val = (int32_t)(0u - (uint32_t)val - (uint32_t)middle);
	return val;
}

// ---- Tunstall dictionary (src/tunstall.cpp:125-256) -----------------------------------------------------
// Scratch the caller provides (shared memory in the kernel): qprob[512], widx[512], wlen[512], head[256],
// text[TUN_TABLE_BYTES].  Result: entry[256] = offset | len<<16, text bytes; returns used text bytes.
// Rows: slot s belongs to symbol s % n; head[r] = oldest live slot of row r.  Each round extends the most
// probable head (first strict maximum) by every symbol; the round that reaches 256 words is cut short and
// then keeps its parent.
struct TunScratch {
	uint32_t qprob[512];
	uint16_t widx[512];
	uint16_t wlen[512];
	uint16_t head[256];
};

CRT_HD uint32_t tun_build_seq(const uint8_t *probs /* (sym,prob) pairs */, uint32_t n, TunScratch &S, uint8_t *text, uint32_t *entry) {
	for(int i = 0; i < 512; i++) S.qprob[i] = 0;
	uint32_t pos = 0, slots = 0, nwords;
	uint32_t p0 = (uint32_t)probs[1] << 8, p1 = (uint32_t)probs[3] << 8;
	uint32_t run = 2, pr = (p0*p0) >> 16, max_run = 255u/(n - 1);
	while(pr > p1 && run < max_run) { pr = (pr*p0) >> 16; run++; }
	if(run >= 16) {
		text[pos++] = probs[0];
		for(uint32_t k = 1; k < n; k++) {
			for(uint32_t i = 0; i + 1 < run; i++) text[pos++] = probs[0];
			text[pos++] = probs[2*k];
		}
		S.head[0] = (uint16_t)((run - 1)*n);
		for(uint32_t k = 1; k < n; k++) S.head[k] = (uint16_t)k;
		for(uint32_t c = 0; c < run; c++) {
			for(uint32_t k = 1; k < n; k++) {
				uint32_t s = k + c*n, pk = (uint32_t)probs[2*k + 1] << 8;
				S.qprob[s] = (c == 0) ? pk : ((pr*pk) >> 16);
				S.widx[s] = (uint16_t)(k*run - c);
				S.wlen[s] = (uint16_t)(c + 1);
			}
			pr = (c == 0) ? p0 : ((pr*p0) >> 16);
		}
		uint32_t s0 = (run - 1)*n;
		S.qprob[s0] = pr; S.widx[s0] = 0; S.wlen[s0] = (uint16_t)run;
		nwords = 1 + run*(n - 1);
		slots = run*n;
	} else {
		for(uint32_t k = 0; k < n; k++) {
			S.head[k] = (uint16_t)k;
			S.qprob[slots] = (uint32_t)probs[2*k + 1] << 8;
			S.widx[slots] = (uint16_t)pos; S.wlen[slots] = 1; slots++;
			text[pos++] = probs[2*k];
		}
		nwords = n;
	}
	while(nwords < 256) {
		uint32_t best = 0, bestp = 0;
		for(uint32_t k = 0; k < n; k++) {
			uint32_t p = S.qprob[S.head[k]];
			if(p > bestp) { bestp = p; best = k; }
		}
		uint32_t parent = S.head[best], pp = S.qprob[parent], poff = S.widx[parent], plen = S.wlen[parent];
		const uint32_t room = 256 - nwords;        // the child that makes word 256 ends the round (tunstall.cpp:234-235)
		const uint32_t m = room < n ? room : n;    // children this round
		for(uint32_t k = 0; k < m; k++) {
			if(slots < 512) {
				S.qprob[slots] = (pp*((uint32_t)probs[2*k + 1] << 8)) >> 16;
				S.widx[slots] = (uint16_t)pos; S.wlen[slots] = (uint16_t)(plen + 1);
			}
			slots++;
			if(pos + plen + 1 <= (uint32_t)TUN_TABLE_BYTES) {
				for(uint32_t j = 0; j < plen; j++) text[pos + j] = text[poff + j];
				text[pos + plen] = probs[2*k];
			}
			pos += plen + 1;
		}
		if(room > n) S.head[best] = (uint16_t)(S.head[best] + n);   // parent retires only if the loop ran to completion (:237-238)
		nwords += n - 1;
	}
	if(slots > 512) slots = 512;
	uint32_t word = 0;
	for(uint32_t s = 0; s < slots && word < 256; s++) {
		if(S.head[s % n] > s) continue;
		entry[word++] = (uint32_t)S.widx[s] | ((uint32_t)S.wlen[s] << 16);
	}
	for(; word < 256; word++) entry[word] = 0;
	return pos < (uint32_t)TUN_TABLE_BYTES ? pos : (uint32_t)TUN_TABLE_BYTES;
}

// ---- CLERS automaton (src/decoder.cpp:204-358) -----------------------------------------------------------
// One call decodes ALL groups of one mesh sequentially (cler cursor, vertex counter and split-bit cursor carry
// across groups, decoder.cpp:173-178; the front / FIFO / delayed stack restart per group, :207-221).
// Memory (global in the v1 kernel, plain arrays in host emulation):
//   ea/eb [cap]  front edges;  order [cap] FIFO of edge ids (faceorder);  delayed [cap] LIFO
//   faces: u32 or u16 triples;  pred: uint4 (a,b,c,0) per vertex (prediction[0] = 0xFFFFFFFF, never read)
// Returns 0 or CRT_E_TOPOLOGY (-5) when the stream runs dry / is inconsistent.
struct ClersIO {
	const uint8_t *clers; uint32_t nclers;
	const uint32_t *split; uint32_t split_nwords;
	const uint32_t *group_ends; uint32_t ngroups;
	uint32_t nvert, nface;
	EdgeA *ea; EdgeB *eb; uint32_t *order; uint32_t *delayed; uint32_t cap;
	uint32_t *faces32; uint16_t *faces16;
	uint32_t *pred;     // 4 x u32 per vertex
};

CRT_HD int ilog2_u32(uint32_t p) { int k = 0; while(p >>= 1) k++; return k; }   // src/cstream.cpp:31-35

CRT_HD void clers_put_face(const ClersIO &io, uint32_t at, uint32_t a, uint32_t b, uint32_t c) {
	if(io.faces16) { io.faces16[at] = (uint16_t)a; io.faces16[at + 1] = (uint16_t)b; io.faces16[at + 2] = (uint16_t)c; }
	else if(io.faces32) { io.faces32[at] = a; io.faces32[at + 1] = b; io.faces32[at + 2] = c; }
}
CRT_HD void clers_put_pred(const ClersIO &io, uint32_t v, uint32_t a, uint32_t b, uint32_t c) {
	uint32_t *p = io.pred + (size_t)v*4;
	p[0] = a; p[1] = b; p[2] = c; p[3] = 0;
}

CRT_HD int clers_decode_seq(const ClersIO &io, uint32_t *vertex_count_out) {
	uint32_t cler = 0, vertex_count = 0;
	uint64_t splitpos = 0;
	const int splitbits = ilog2_u32(io.nvert) + 1;
	uint32_t start;
	for(uint32_t g = 0; g < io.ngroups; g++) {
		uint32_t end = io.group_ends[g]*3;
		if(end > io.nface*3) end = io.nface*3;
		start = g ? io.group_ends[g - 1]*3 : 0;      // decoder.cpp:174-177: each group restarts at the previous group's end
		if(start > io.nface*3) start = io.nface*3;
		uint32_t nfront = 0, norder = 0, cursor = 0, ndelayed = 0;
		uint32_t new_edge = 0xFFFFFFFFu;
		while(start < end) {
			if(new_edge == 0xFFFFFFFFu && cursor >= norder && ndelayed == 0) {
				if(cler >= io.nclers || nfront + 3 > io.cap) return -5;
				uint32_t last = vertex_count - 1;
				uint32_t vi[3];
				uint32_t mask = 0;
				uint32_t c = io.clers[cler++];
				if(c == C_SPLIT) { mask = getbits(io.split, io.split_nwords, splitpos, 3); splitpos += 3; }
				for(int k = 0; k < 3; k++) {
					uint32_t v;
					if(mask & (1u << k)) {
						v = getbits(io.split, io.split_nwords, splitpos, splitbits); splitpos += (uint64_t)splitbits;
						if(v >= io.nvert) return -5;
					} else {
						if(vertex_count >= io.nvert) return -5;
						clers_put_pred(io, vertex_count, last, last, last);
						last = v = vertex_count++;
					}
					vi[k] = v;
				}
				clers_put_face(io, start, vi[0], vi[1], vi[2]);
				start += 3;
				uint32_t b = nfront;
				io.ea[b + 0] = EdgeA{vi[1], vi[2], vi[0], 0}; io.eb[b + 0] = EdgeB{b + 2, b + 1};
				io.ea[b + 1] = EdgeA{vi[2], vi[0], vi[1], 0}; io.eb[b + 1] = EdgeB{b + 0, b + 2};
				io.ea[b + 2] = EdgeA{vi[0], vi[1], vi[2], 0}; io.eb[b + 2] = EdgeB{b + 1, b + 0};
				io.order[norder++] = b; io.order[norder++] = b + 1; io.order[norder++] = b + 2;
				nfront += 3;
				continue;
			}
			uint32_t f;
			if(new_edge != 0xFFFFFFFFu) { f = new_edge; new_edge = 0xFFFFFFFFu; }
			else if(cursor < norder) f = io.order[cursor++];
			else f = io.delayed[--ndelayed];

			const EdgeA e = io.ea[f];
			if(e.deleted) continue;
			if(cler >= io.nclers) return -5;
			const uint32_t c = io.clers[cler++];
			if(c == C_BOUNDARY) continue;
			const EdgeB el = io.eb[f];
			if(nfront + 2 > io.cap) return -5;
			new_edge = nfront;
			uint32_t opposite;
			if(c == C_VERTEX || c == C_SPLIT) {
				if(c == C_SPLIT) {
					opposite = getbits(io.split, io.split_nwords, splitpos, splitbits); splitpos += (uint64_t)splitbits;
					if(opposite >= io.nvert) return -5;
				} else {
					if(vertex_count >= io.nvert) return -5;
					clers_put_pred(io, vertex_count, e.v1, e.v0, e.v2);
					opposite = vertex_count++;
				}
				io.eb[el.prev].next = new_edge;
				io.eb[el.next].prev = new_edge + 1;
				io.ea[nfront] = EdgeA{e.v0, opposite, e.v1, 0}; io.eb[nfront] = EdgeB{el.prev, new_edge + 1};
				io.ea[nfront + 1] = EdgeA{opposite, e.v1, e.v0, 0}; io.eb[nfront + 1] = EdgeB{new_edge, el.next};
				io.order[norder++] = nfront + 1;
				nfront += 2;
			} else if(c == C_LEFT) {
				const EdgeB pl = io.eb[el.prev];            // previous_edge copy (decoder.cpp:288)
				opposite = io.ea[el.prev].v0;
				io.ea[el.prev].deleted = 1;
				io.eb[pl.prev].next = new_edge;
				io.eb[el.next].prev = new_edge;
				io.ea[nfront] = EdgeA{opposite, e.v1, e.v0, 0}; io.eb[nfront] = EdgeB{pl.prev, el.next};
				nfront += 1;
			} else if(c == C_RIGHT) {
				const EdgeB nl = io.eb[el.next];            // next_edge copy (decoder.cpp:289)
				opposite = io.ea[el.next].v1;
				io.ea[el.next].deleted = 1;
				io.eb[nl.next].prev = new_edge;
				io.eb[el.prev].next = new_edge;
				io.ea[nfront] = EdgeA{e.v0, opposite, e.v1, 0}; io.eb[nfront] = EdgeB{el.prev, nl.next};
				nfront += 1;
			} else if(c == C_DELAY) {
				io.delayed[ndelayed++] = f;
				new_edge = 0xFFFFFFFFu;
				continue;
			} else if(c == C_END) {
				const EdgeB pl = io.eb[el.prev];
				const EdgeB nl = io.eb[el.next];
				opposite = io.ea[el.prev].v0;
				io.ea[el.prev].deleted = 1;
				io.ea[el.next].deleted = 1;
				io.eb[pl.prev].next = nl.next;
				io.eb[nl.next].prev = pl.prev;
				new_edge = 0xFFFFFFFFu;
			} else return -5;
			clers_put_face(io, start, e.v1, e.v0, opposite);
			start += 3;
		}
	}
	*vertex_count_out = vertex_count;
	return 0;
}


// ---- CLERS automaton, ring-cached (the kernel's hot version) -----------------------------------------------------
// Same machine as clers_decode_seq.  The automaton is a pointer chase whose cost is memory latency, so the hot state
// lives in fast memory (shared memory in k_clers): the most recent R front edges and Q FIFO entries are mirrored in
// rings indexed by id & (R-1); global memory stays authoritative (write-through) and serves the rare reach-back
// (99.8 % of front accesses fall within the last 4096 edges).  The edge created last, which is the next
// one processed 62 % of the time (decoder.cpp:262-264), never leaves registers.  CLERS symbols arrive through a
// double-buffered 8-byte register window (the stream must be 8-byte aligned and padded by 16 readable bytes).
struct ClersRing { EdgeA *ra; EdgeB *rb; uint32_t *rq; uint32_t R, Q; };   // R, Q powers of two

CRT_HD uint64_t load_u64(const uint8_t *p) { return *(const uint64_t *)p; }

CRT_HD int clers_decode_ring(const ClersIO &io, const ClersRing &rg, uint32_t *vertex_count_out) {
	const uint32_t R = rg.R, RM = rg.R - 1, Q = rg.Q, QM = rg.Q - 1;
	EdgeA *const ra = rg.ra; EdgeB *const rb = rg.rb; uint32_t *const rq = rg.rq;
	uint32_t cler = 0, vertex_count = 0;
	uint64_t cw = io.nclers ? load_u64(io.clers) : 0, cw_next = io.nclers > 8 ? load_u64(io.clers + 8) : 0;
	uint64_t splitpos = 0;
	const int splitbits = ilog2_u32(io.nvert) + 1;
	const uint32_t NONE = 0xFFFFFFFFu;
#define CRT_NEXT_CLER(c)                                                                              \
	do {                                                                                              \
		if(cler >= io.nclers) return -5;                                                              \
		c = (uint32_t)(cw & 0xffu); cw >>= 8; cler++;                                                 \
		if((cler & 7u) == 0) { cw = cw_next; cw_next = (cler + 8 < io.nclers) ? load_u64(io.clers + cler + 8) : 0; } \
	} while(0)
	for(uint32_t g = 0; g < io.ngroups; g++) {
		uint32_t end = io.group_ends[g]*3;
		if(end > io.nface*3) end = io.nface*3;
		uint32_t start = g ? io.group_ends[g - 1]*3 : 0;
		if(start > io.nface*3) start = io.nface*3;
		uint32_t nfront = 0, norder = 0, cursor = 0, ndelayed = 0;
		bool have = false;                 // current edge (the one created last) is held in registers
		uint32_t cf = 0; EdgeA ce = {0, 0, 0, 0}; EdgeB cl = {0, 0};
		while(start < end) {
			if(!have && cursor >= norder && ndelayed == 0) {
				if(nfront + 3 > io.cap) return -5;
				uint32_t last = vertex_count - 1;
				uint32_t vi[3];
				uint32_t mask = 0, c;
				CRT_NEXT_CLER(c);
				if(c == C_SPLIT) { mask = getbits(io.split, io.split_nwords, splitpos, 3); splitpos += 3; }
				for(int k = 0; k < 3; k++) {
					uint32_t v;
					if(mask & (1u << k)) {
						v = getbits(io.split, io.split_nwords, splitpos, splitbits); splitpos += (uint64_t)splitbits;
						if(v >= io.nvert) return -5;
					} else {
						if(vertex_count >= io.nvert) return -5;
						clers_put_pred(io, vertex_count, last, last, last);
						last = v = vertex_count++;
					}
					vi[k] = v;
				}
				clers_put_face(io, start, vi[0], vi[1], vi[2]);
				start += 3;
				const uint32_t b = nfront;
				const EdgeA a0 = {vi[1], vi[2], vi[0], 0}, a1 = {vi[2], vi[0], vi[1], 0}, a2 = {vi[0], vi[1], vi[2], 0};
				const EdgeB b0 = {b + 2, b + 1}, b1 = {b + 0, b + 2}, b2 = {b + 1, b + 0};
				io.ea[b] = a0; io.eb[b] = b0; ra[b & RM] = a0; rb[b & RM] = b0;
				io.ea[b + 1] = a1; io.eb[b + 1] = b1; ra[(b + 1) & RM] = a1; rb[(b + 1) & RM] = b1;
				io.ea[b + 2] = a2; io.eb[b + 2] = b2; ra[(b + 2) & RM] = a2; rb[(b + 2) & RM] = b2;
				for(uint32_t k = 0; k < 3; k++) { io.order[norder] = b + k; rq[norder & QM] = b + k; norder++; }
				nfront += 3;
				continue;
			}
			uint32_t f; EdgeA e; EdgeB el;
			if(have) { f = cf; e = ce; el = cl; have = false; }
			else {
				if(cursor < norder) { f = (cursor + Q >= norder) ? rq[cursor & QM] : io.order[cursor]; cursor++; }
				else f = io.delayed[--ndelayed];
				const bool inw = f + R >= nfront;
				e = inw ? ra[f & RM] : io.ea[f];
				if(e.deleted) continue;
				el = inw ? rb[f & RM] : io.eb[f];
			}
			uint32_t c;
			CRT_NEXT_CLER(c);
			if(c == C_BOUNDARY) continue;
			if(nfront + 2 > io.cap) return -5;
			const uint32_t ne = nfront;
			uint32_t opposite;
			if(c == C_VERTEX || c == C_SPLIT) {
				if(c == C_SPLIT) {
					opposite = getbits(io.split, io.split_nwords, splitpos, splitbits); splitpos += (uint64_t)splitbits;
					if(opposite >= io.nvert) return -5;
				} else {
					if(vertex_count >= io.nvert) return -5;
					clers_put_pred(io, vertex_count, e.v1, e.v0, e.v2);
					opposite = vertex_count++;
				}
				io.eb[el.prev].next = ne;     if(el.prev + R >= ne) rb[el.prev & RM].next = ne;
				io.eb[el.next].prev = ne + 1; if(el.next + R >= ne) rb[el.next & RM].prev = ne + 1;
				const EdgeA a0 = {e.v0, opposite, e.v1, 0}, a1 = {opposite, e.v1, e.v0, 0};
				const EdgeB b0 = {el.prev, ne + 1}, b1 = {ne, el.next};
				io.ea[ne] = a0; io.eb[ne] = b0; ra[ne & RM] = a0; rb[ne & RM] = b0;
				io.ea[ne + 1] = a1; io.eb[ne + 1] = b1; ra[(ne + 1) & RM] = a1; rb[(ne + 1) & RM] = b1;
				io.order[norder] = ne + 1; rq[norder & QM] = ne + 1; norder++;
				nfront += 2;
				cf = ne; ce = a0; cl = b0; have = true;
			} else if(c == C_LEFT) {
				const uint32_t p = el.prev;
				const bool pw = p + R >= ne;
				const EdgeB pl = pw ? rb[p & RM] : io.eb[p];
				opposite = pw ? ra[p & RM].v0 : io.ea[p].v0;
				io.ea[p].deleted = 1;         if(pw) ra[p & RM].deleted = 1;
				io.eb[pl.prev].next = ne;     if(pl.prev + R >= ne) rb[pl.prev & RM].next = ne;
				io.eb[el.next].prev = ne;     if(el.next + R >= ne) rb[el.next & RM].prev = ne;
				const EdgeA a0 = {opposite, e.v1, e.v0, 0};
				const EdgeB b0 = {pl.prev, el.next};
				io.ea[ne] = a0; io.eb[ne] = b0; ra[ne & RM] = a0; rb[ne & RM] = b0;
				nfront += 1;
				cf = ne; ce = a0; cl = b0; have = true;
			} else if(c == C_RIGHT) {
				const uint32_t n = el.next;
				const bool nw = n + R >= ne;
				const EdgeB nl = nw ? rb[n & RM] : io.eb[n];
				opposite = nw ? ra[n & RM].v1 : io.ea[n].v1;
				io.ea[n].deleted = 1;         if(nw) ra[n & RM].deleted = 1;
				io.eb[nl.next].prev = ne;     if(nl.next + R >= ne) rb[nl.next & RM].prev = ne;
				io.eb[el.prev].next = ne;     if(el.prev + R >= ne) rb[el.prev & RM].next = ne;
				const EdgeA a0 = {e.v0, opposite, e.v1, 0};
				const EdgeB b0 = {el.prev, nl.next};
				io.ea[ne] = a0; io.eb[ne] = b0; ra[ne & RM] = a0; rb[ne & RM] = b0;
				nfront += 1;
				cf = ne; ce = a0; cl = b0; have = true;
			} else if(c == C_DELAY) {
				io.delayed[ndelayed++] = f;
				continue;
			} else if(c == C_END) {
				const uint32_t p = el.prev, n = el.next;
				const bool pw = p + R >= ne, nw = n + R >= ne;
				const EdgeB pl = pw ? rb[p & RM] : io.eb[p];
				const EdgeB nl = nw ? rb[n & RM] : io.eb[n];
				opposite = pw ? ra[p & RM].v0 : io.ea[p].v0;
				io.ea[p].deleted = 1;         if(pw) ra[p & RM].deleted = 1;
				io.ea[n].deleted = 1;         if(nw) ra[n & RM].deleted = 1;
				io.eb[pl.prev].next = nl.next; if(pl.prev + R >= ne) rb[pl.prev & RM].next = nl.next;
				io.eb[nl.next].prev = pl.prev; if(nl.next + R >= ne) rb[nl.next & RM].prev = pl.prev;
			} else return -5;
			clers_put_face(io, start, e.v1, e.v0, opposite);
			start += 3;
		}
	}
#undef CRT_NEXT_CLER
	*vertex_count_out = vertex_count;
	return 0;
}

}  // namespace crtb
