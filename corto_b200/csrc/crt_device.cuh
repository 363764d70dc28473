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

struct uint4_t { uint32_t x, y, z, w; };
struct uint2_t { uint32_t x, y; };

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
	uint8_t *fl;        // v4: flag byte per edge (global backing of the flag ring); aliases `order`, which v4 does not use
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


CRT_HD uint64_t load_u64(const uint8_t *p) { return *(const uint64_t *)p; }

// ---- CLERS automaton, v3: lazy front + staged outputs (the kernel's hot version) ---------------------------------
// ncu on a straightforward ring-cached port showed the machine is INSTRUCTION-latency bound: one warp, ~150 dependent
// instructions per symbol at 4.4-5.6 cycles each, plus L2 round trips whenever the front reaches back past the ring.
// v3 removes work from the per-symbol chain instead of hiding latency:
//  * the edge created last is processed next 62 % of the time (decoder.cpp:262-264) and is then never referenced
//    again, so it lives only in registers: it gets NO id and NO record, and the two links pointing at it are not
//    written ("deferred": lp = prev's .next, ln = next's .prev) unless the next symbol is BOUNDARY / DELAY, which
//    materialise it.  Deferred fields are never read while the edge is in flight (LEFT reads prev's .prev/.v0, RIGHT
//    next's .next/.v1, END both) except in a 2-edge loop, where the edge is materialised first.  Edge ids are
//    internal (never output), so only materialised edges consume ids: a ring of R slots reaches ~3x further back.
//  * rings (front edges, FIFO) and outputs (faces, predictions) are written to fast memory only; every `budget`
//    symbols the WARP (all lanes) drains them to global memory with coalesced stores: faces/predictions as final
//    outputs, ring entries about to leave the window as the reach-back backing store.  An id below the flushed limit
//    (eflush / qflush) is served from global memory, anything newer from the ring.
// Mem policy RG supplies the ring accessors (shared memory in the kernel, plain arrays in tests/host_emul).
struct ClersState {
	uint32_t cler, vertex_count;
	uint64_t cw, cw_next, splitpos;
	uint32_t g, start, end;            // current group; face cursor / end of group, in FACES
	uint32_t nfront, norder, cursor, ndelayed;
	uint32_t have, lp, ln;             // current edge valid; deferred incoming links
	uint32_t cf, cv0, cv1, cv2, cprev, cnext;   // current edge; cf == CLERS_NOID while it has no record
	uint32_t eflush, qflush;           // edge ids / FIFO positions below these live in global memory
	uint32_t fflush, pflush;           // faces / predictions already drained to global memory
};
constexpr uint32_t CLERS_NOID = 0xFFFFFFFFu;

CRT_HD void clers_state_init(ClersState &S, const ClersIO &io) {
	S.cler = 0; S.vertex_count = 0; S.splitpos = 0;
	S.cw = io.nclers ? load_u64(io.clers) : 0; S.cw_next = io.nclers > 8 ? load_u64(io.clers + 8) : 0;
	S.g = 0; S.start = 0; S.end = 0;
	S.nfront = S.norder = S.cursor = S.ndelayed = 0;
	S.have = S.lp = S.ln = 0; S.cf = CLERS_NOID; S.cv0 = S.cv1 = S.cv2 = S.cprev = S.cnext = 0;
	S.eflush = S.qflush = S.fflush = S.pflush = 0;
}

// Slow paths kept out of line so the hot loop stays small (reach-back into global memory is rare by construction).
#ifdef __CUDACC__
#define CRT_COLD __device__ __host__ __noinline__
#else
#define CRT_COLD __attribute__((noinline))
#endif
struct EdgeRec { uint32_t v0, v1, v2, del, p, n; };
CRT_COLD static void clers_g_set_next(EdgeB *eb, uint32_t x, uint32_t v) { eb[x].next = v; }
CRT_COLD static void clers_g_set_prev(EdgeB *eb, uint32_t x, uint32_t v) { eb[x].prev = v; }
CRT_COLD static void clers_g_set_del(EdgeA *ea, uint32_t x) { ea[x].deleted = 1; }
CRT_COLD static EdgeRec clers_g_load(const EdgeA *ea, const EdgeB *eb, uint32_t x) {   // by value: nothing of the hot loop gets its address taken
	const EdgeA a = ea[x]; const EdgeB l = eb[x];
	EdgeRec r; r.v0 = a.v0; r.v1 = a.v1; r.v2 = a.v2; r.del = a.deleted; r.p = l.prev; r.n = l.next;
	return r;
}

// Runs at most `budget` symbols.  Returns 1 when all groups are done, 0 when the caller must drain the staging rings
// and call again (budget exhausted, or a group table that moves the face cursor), < 0 on a topology error (the state is
// then meaningless).  `splitbits` = ilog2(nvert)+1 (decoder.cpp:219), passed in so it is computed once per mesh.
template <class RG> CRT_HD int clers_run(const ClersIO &io, RG &rg, ClersState &S, int budget, int splitbits) {
	uint32_t cler = S.cler, vcount = S.vertex_count, start = S.start, end = S.end;
	uint32_t nfront = S.nfront, norder = S.norder, cursor = S.cursor, ndel = S.ndelayed;
	uint64_t cw = S.cw, cwn = S.cw_next, splitpos = S.splitpos;
	uint32_t have = S.have, lp = S.lp, ln = S.ln, f = S.cf, v0 = S.cv0, v1 = S.cv1, v2 = S.cv2, prev = S.cprev, next = S.cnext;
	uint32_t g = S.g, fflush = S.fflush;
	const uint32_t eflush = S.eflush, qflush = S.qflush, nclers = io.nclers, nvert = io.nvert, cap = io.cap;
	// symbols this run may consume: the budget, or what is left of the stream (running dry with faces missing = error)
	uint32_t n = nclers - cler;
	if(n > (uint32_t)budget) n = (uint32_t)budget;
	int rc = 0;
#define CRT_FETCH(c)                                                                                     \
	do {                                                                                                 \
		c = (uint32_t)cw & 0xffu; cw >>= 8; cler++;                                                      \
		if((cler & 7u) == 0) { cw = cwn; cwn = (cler + 8 < nclers) ? load_u64(io.clers + cler + 8) : 0; } \
	} while(0)
#define CRT_SET_NEXT(x, v) do { if((x) >= eflush) rg.stB_next(x, v); else clers_g_set_next(io.eb, x, v); } while(0)
#define CRT_SET_PREV(x, v) do { if((x) >= eflush) rg.stB_prev(x, v); else clers_g_set_prev(io.eb, x, v); } while(0)
#define CRT_SET_DEL(x)     do { if((x) >= eflush) rg.stA_del(x); else clers_g_set_del(io.ea, x); } while(0)
#define CRT_MATERIALISE()                                                                                \
	do {                                                                                                 \
		if(nfront >= cap) return -5;                                                                     \
		f = nfront++;                                                                                    \
		rg.stA(f, v0, v1, v2, 0); rg.stB(f, prev, next);                                                 \
		if(lp) CRT_SET_NEXT(prev, f);                                                                    \
		if(ln) CRT_SET_PREV(next, f);                                                                    \
		lp = ln = 0;                                                                                     \
	} while(0)
	for(;;) {
		if(!have) {
			if(start >= end) {                         // open the next group: fresh front (decoder.cpp:173-178, 207-221)
				if(g >= io.ngroups) { rc = 1; break; }
				uint32_t e = io.group_ends[g];
				if(e > io.nface) e = io.nface;
				uint32_t st = g ? io.group_ends[g - 1] : 0;
				if(st > io.nface) st = io.nface;
				if(st != start) {                      // malformed group table moves the cursor: drain staged faces first
					if(fflush != start) { rc = 0; break; }
					fflush = st;
				}
				g++;
				start = st; end = e;
				nfront = norder = cursor = ndel = 0;
				// the rings restart with the group: the flushed limits of the old front must go before any id of the new
				// front is looked up, so hand back to the caller (its drain resets them) and resume
				if(eflush | qflush) { S.eflush = 0; S.qflush = 0; rc = 0; break; }
				continue;
			}
			// next edge: FIFO (skipping edges deleted since they were queued, decoder.cpp:278-279), then the delayed stack
			uint32_t del = 1;
			while(cursor < norder) {
				f = (cursor >= qflush) ? rg.ldQ(cursor) : io.order[cursor];
				cursor++;
				if(f >= eflush) { rg.ldA(f, v0, v1, v2, del); if(!del) rg.ldB(f, prev, next); }
				else { const EdgeRec r = clers_g_load(io.ea, io.eb, f); v0 = r.v0; v1 = r.v1; v2 = r.v2; del = r.del; prev = r.p; next = r.n; }
				if(!del) break;
			}
			if(del && ndel) {
				f = io.delayed[--ndel];
				if(f >= eflush) { rg.ldA(f, v0, v1, v2, del); if(!del) rg.ldB(f, prev, next); }
				else { const EdgeRec r = clers_g_load(io.ea, io.eb, f); v0 = r.v0; v1 = r.v1; v2 = r.v2; del = r.del; prev = r.p; next = r.n; }
				if(del) continue;
			}
			if(!del) { lp = ln = 0; have = 1; }
			else {                                     // nothing pending: start triangle (decoder.cpp:224-259)
				if(n == 0) { rc = (cler >= nclers) ? -5 : 0; break; }
				n--;
				if(nfront + 3 > cap) return -5;
				uint32_t last = vcount - 1;
				uint32_t vi[3];
				uint32_t mask = 0, c;
				CRT_FETCH(c);
				if(c == C_SPLIT) { mask = getbits(io.split, io.split_nwords, splitpos, 3); splitpos += 3; }
				for(int k = 0; k < 3; k++) {
					uint32_t v;
					if(mask & (1u << k)) {
						v = getbits(io.split, io.split_nwords, splitpos, splitbits); splitpos += (uint64_t)splitbits;
						if(v >= nvert) return -5;
					} else {
						if(vcount >= nvert) return -5;
						rg.stP(vcount, last, last, last);
						last = v = vcount++;
					}
					vi[k] = v;
				}
				rg.stF(start, vi[0], vi[1], vi[2]);
				start += 1;
				const uint32_t b = nfront;
				rg.stA(b, vi[1], vi[2], vi[0], 0);     rg.stB(b, b + 2, b + 1);
				rg.stA(b + 1, vi[2], vi[0], vi[1], 0); rg.stB(b + 1, b + 0, b + 2);
				rg.stA(b + 2, vi[0], vi[1], vi[2], 0); rg.stB(b + 2, b + 1, b + 0);
				rg.stQ(norder, b); rg.stQ(norder + 1, b + 1); rg.stQ(norder + 2, b + 2);
				norder += 3; nfront += 3;
				continue;
			}
		}
		// ---- strip: the current edge stays in registers from symbol to symbol ----
		do {
			if(n == 0) { rc = (cler >= nclers) ? -5 : 0; goto save; }
			n--;
			uint32_t c;
			CRT_FETCH(c);
			if(c == C_VERTEX) {
				if(vcount >= nvert || nfront >= cap) return -5;
				rg.stP(vcount, v1, v0, v2);
				const uint32_t opp = vcount++, b = nfront++;
				rg.stA(b, opp, v1, v0, 0); rg.stB(b, CLERS_NOID, next);      // second new edge: persistent, queued
				CRT_SET_PREV(next, b);
				rg.stQ(norder, b); norder++;
				rg.stF(start, v1, v0, opp); start++;
				v2 = v1; v1 = opp; next = b; lp = 1; ln = 1; f = CLERS_NOID;     // first new edge: registers only
			} else if(c == C_LEFT) {
				if((lp | ln) && prev == next) CRT_MATERIALISE();                 // 2-edge loop: deferred fields would be read
				uint32_t pp, pn, pv0, t1, t2, t3;
				if(prev >= eflush) { rg.ldB(prev, pp, pn); rg.ldA(prev, pv0, t1, t2, t3); }
				else { const EdgeRec r = clers_g_load(io.ea, io.eb, prev); pv0 = r.v0; pp = r.p; }
				(void)pn; (void)t1; (void)t2; (void)t3;
				CRT_SET_DEL(prev);
				rg.stF(start, v1, v0, pv0); start++;
				v2 = v0; v0 = pv0; prev = pp; lp = 1; ln = 1; f = CLERS_NOID;
			} else if(c == C_RIGHT) {
				if((lp | ln) && prev == next) CRT_MATERIALISE();
				uint32_t np, nn, nv1, t0, t2, t3;
				if(next >= eflush) { rg.ldB(next, np, nn); rg.ldA(next, t0, nv1, t2, t3); }
				else { const EdgeRec r = clers_g_load(io.ea, io.eb, next); nv1 = r.v1; nn = r.n; }
				(void)np; (void)t0; (void)t2; (void)t3;
				CRT_SET_DEL(next);
				rg.stF(start, v1, v0, nv1); start++;
				v2 = v1; v1 = nv1; next = nn; lp = 1; ln = 1; f = CLERS_NOID;
			} else if(c == C_BOUNDARY) {
				if(f == CLERS_NOID) CRT_MATERIALISE();
				have = 0; break;
			} else if(c == C_DELAY) {
				if(f == CLERS_NOID) CRT_MATERIALISE();
				io.delayed[ndel++] = f;
				have = 0; break;
			} else if(c == C_END) {
				if((lp | ln) && prev == next) CRT_MATERIALISE();
				uint32_t pp, pn, np, nn, pv0, t1, t2, t3;
				if(prev >= eflush) { rg.ldB(prev, pp, pn); rg.ldA(prev, pv0, t1, t2, t3); }
				else { const EdgeRec r = clers_g_load(io.ea, io.eb, prev); pv0 = r.v0; pp = r.p; }
				if(next >= eflush) rg.ldB(next, np, nn); else { const EdgeRec r = clers_g_load(io.ea, io.eb, next); nn = r.n; }
				(void)pn; (void)np; (void)t1; (void)t2; (void)t3;
				CRT_SET_DEL(prev);
				CRT_SET_DEL(next);
				CRT_SET_NEXT(pp, nn);
				CRT_SET_PREV(nn, pp);
				rg.stF(start, v1, v0, pv0); start++;
				have = 0; break;
			} else if(c == C_SPLIT) {
				if(nfront >= cap) return -5;
				const uint32_t opp = getbits(io.split, io.split_nwords, splitpos, splitbits); splitpos += (uint64_t)splitbits;
				if(opp >= nvert) return -5;
				const uint32_t b = nfront++;
				rg.stA(b, opp, v1, v0, 0); rg.stB(b, CLERS_NOID, next);
				CRT_SET_PREV(next, b);
				rg.stQ(norder, b); norder++;
				rg.stF(start, v1, v0, opp); start++;
				v2 = v1; v1 = opp; next = b; lp = 1; ln = 1; f = CLERS_NOID;
			} else return -5;
		} while(start < end);
		have = 0;                                          // strip over (terminator) or group complete (front discarded, decoder.cpp:223)
	}
save:
	S.cler = cler; S.vertex_count = vcount; S.start = start; S.end = end;
	S.nfront = nfront; S.norder = norder; S.cursor = cursor; S.ndelayed = ndel;
	S.cw = cw; S.cw_next = cwn; S.splitpos = splitpos;
	S.have = have; S.lp = lp; S.ln = ln; S.cf = f; S.cv0 = v0; S.cv1 = v1; S.cv2 = v2; S.cprev = prev; S.cnext = next;
	S.g = g; S.fflush = fflush;
	return rc;
#undef CRT_FETCH
#undef CRT_SET_NEXT
#undef CRT_SET_PREV
#undef CRT_SET_DEL
#undef CRT_MATERIALISE
}

// ---- CLERS automaton, v4: leader / follower ------------------------------------------------------------------------
// The machine has two kinds of state: LINKS (which edge comes next: prev/next/deleted, the FIFO, the symbol stream) and
// LABELS (which vertices an edge carries: v0,v1,v2 -> faces, predictions).  Labels never influence links
// (decoder.cpp:204-358: every branch depends on the symbol, `deleted`, and the queues only), so the serial chain that
// limits throughput is the link machine alone.  v4 runs it as the LEADER in one warp and streams a log of what it did
// (one 32-bit word per event) through a shared-memory ring to the FOLLOWER in a second warp (its own SM sub-partition),
// which replays the log on labels only: allocates vertex ids, reads split indices, emits faces and predictions.
// Both are lane-0 machines; the remaining lanes of each warp do the coalesced drains.
// Log word: type << 28 | id.
enum { LG_TV = 0, LG_TS = 1, LG_V = 2, LG_S = 3, LG_L = 4, LG_R = 5, LG_E = 6, LG_P = 7, LG_M = 8, LG_G = 9 };
constexpr uint32_t CLERS_DEL = 1u;                // flag byte per edge: deleted
constexpr uint32_t CLERS_NQ = 2u;                 // flag byte per edge: "not queued": the implicit FIFO skips this edge (materialised register edges)
constexpr uint32_t CLERS_IDMASK = 0x0FFFFFFFu;    // ids fit the 28 payload bits of a log word
constexpr uint32_t CLERS_NOLINK = 0x0FFFFFFFu;    // placeholder for a link that is still deferred

struct LeadState {
	uint32_t cler; uint64_t cw, cw_next;
	uint32_t g, start, end;
	uint32_t nfront, scan, ndelayed;   // scan: next edge id the implicit FIFO looks at
	uint32_t have, lp, ln, cf, cprev, cnext;
	uint32_t eflush;
	uint32_t nlog, bad;
	uint32_t cwv;                      // 0: cw / cw_next are stale (a window step moved cler), re-prime on entry
};
CRT_HD void lead_init(LeadState &S, const ClersIO &io) {
	S.cler = 0; S.cw = io.nclers ? load_u64(io.clers) : 0; S.cw_next = io.nclers > 8 ? load_u64(io.clers + 8) : 0;
	S.g = 0; S.start = S.end = 0; S.nfront = S.scan = S.ndelayed = 0;
	S.have = S.lp = S.ln = 0; S.cf = CLERS_NOID; S.cprev = S.cnext = 0; S.eflush = 0; S.nlog = 0; S.bad = 0; S.cwv = 1;
}

CRT_COLD static uint2_t lead_g_load(const EdgeB *eb, uint32_t x) { const EdgeB l = eb[x]; return uint2_t{l.prev, l.next}; }
CRT_COLD static uint32_t lead_g_flag(const uint8_t *fl, uint32_t x) { return fl[x]; }
CRT_COLD static void lead_g_set_flag(uint8_t *fl, uint32_t x, uint32_t v) { fl[x] = (uint8_t)v; }

// Link machine.  The reference's faceorder FIFO (decoder.cpp:213-215) holds exactly the queued edges in creation order,
// and ids here are allocated in creation order too, so the FIFO is IMPLICIT: popping = scanning ids upward for the next
// edge that is queued and alive (a flag byte per edge: CLERS_DEL / CLERS_NQ).  No queue is stored.
// Consumes at most `budget` symbols (each yields at most 2 log words, a pop 1).  Returns 1 when all groups are done, 0 to
// be called again after the caller drained / waited for log space, 3 (only with vec) when a VERTEX/LEFT run starts and the
// caller should run its warp-wide window step, 4 (only with vec) when the implicit FIFO must be scanned (lead_pop_vector),
// < 0 on a topology error.  There are no exits from
// inside the hot paths: errors set a sticky flag and indices are clamped, the flag is reported at the chunk end.
template <class RG> CRT_HD int clers_lead(const ClersIO &io, RG &rg, LeadState &S, int budget, bool vec = false) {
	uint32_t cler = S.cler, start = S.start, end = S.end;
	uint32_t nfront = S.nfront, scan = S.scan, ndel = S.ndelayed, nlog = S.nlog, bad = S.bad;
	if(!S.cwv) {                                   // re-prime the 8-byte symbol window after a warp-wide step
		const uint32_t g8 = cler & ~7u;
		S.cw = cler < io.nclers ? load_u64(io.clers + g8) >> (8u*(cler & 7u)) : 0;
		S.cw_next = g8 + 8 < io.nclers ? load_u64(io.clers + g8 + 8) : 0;
		S.cwv = 1;
	}
	uint64_t cw = S.cw, cwn = S.cw_next;
	uint32_t have = S.have, lp = S.lp, ln = S.ln, f = S.cf, prev = S.cprev, next = S.cnext, g = S.g;
	const uint32_t eflush = S.eflush, nclers = io.nclers, cap = io.cap;
	uint32_t n = nclers - cler;
	if(n > (uint32_t)budget) n = (uint32_t)budget;
	int rc = 0;
#define LD_FETCH(c)                                                                                      \
	do {                                                                                                 \
		c = (uint32_t)cw & 0xffu; cw >>= 8; cler++;                                                      \
		if((cler & 7u) == 0) { cw = cwn; cwn = (cler + 8 < nclers) ? load_u64(io.clers + cler + 8) : 0; } \
	} while(0)
#define LD_LOG(t, id) do { rg.stLog(nlog, ((uint32_t)(t) << 28) | (id)); nlog++; } while(0)
#define LD_SET_NEXT(x, v) do { if((x) >= eflush) rg.stB_next(x, v); else clers_g_set_next(io.eb, x, v); } while(0)
#define LD_SET_PREV(x, v) do { if((x) >= eflush) rg.stB_prev(x, v); else clers_g_set_prev(io.eb, x, v); } while(0)
#define LD_LOADB(ID_, P_, Q_) do { if((ID_) >= eflush) rg.ldB(ID_, P_, Q_); else { const uint2_t t_ = lead_g_load(io.eb, ID_); P_ = t_.x; Q_ = t_.y; } } while(0)
#define LD_FLAG(ID_) (((ID_) >= eflush) ? rg.ldFl(ID_) : lead_g_flag(io.fl, ID_))
#define LD_SET_FLAG(ID_, V_) do { if((ID_) >= eflush) rg.stFl(ID_, V_); else lead_g_set_flag(io.fl, ID_, V_); } while(0)
// new id, clamped so that a corrupt stream cannot run past the scratch arrays
#define LD_NEWID(ID_) do { ID_ = nfront; bad |= (nfront >= cap); nfront += (nfront < cap); } while(0)
#define LD_MATERIALISE()                                                                                 \
	do {                                                                                                 \
		LD_NEWID(f);                                                                                     \
		rg.stB(f, prev, next); rg.stFl(f, CLERS_NQ);                                                     \
		if(lp) LD_SET_NEXT(prev, f);                                                                     \
		if(ln) LD_SET_PREV(next, f);                                                                     \
		lp = ln = 0;                                                                                     \
		LD_LOG(LG_M, f);                                                                                 \
	} while(0)
	for(;;) {
		if(!have) {
			if(start >= end) {                         // next group: fresh front (decoder.cpp:173-178, 207-221)
				if(g >= io.ngroups) { rc = 1; break; }
				uint32_t e = io.group_ends[g];
				if(e > io.nface) e = io.nface;
				uint32_t st = g ? io.group_ends[g - 1] : 0;
				if(st > io.nface) st = io.nface;
				g++;
				start = st; end = e;
				nfront = scan = ndel = 0;
				LD_LOG(LG_G, g - 1);
				if(eflush) { S.eflush = 0; rc = 0; break; }   // ring restarts: reset the flushed limit first
				continue;
			}
			uint32_t skip = 1;
			if(vec && scan < nfront) { rc = 4; break; }   // caller scans the flag bytes 32 at a time (k_clers_lf: lead_pop_vector)
			while(scan < nfront) {                     // implicit FIFO: next queued, alive edge in id order
				f = scan++;
				skip = LD_FLAG(f);
				if(!skip) break;
			}
			if(skip && ndel) {
				f = io.delayed[--ndel];
				skip = LD_FLAG(f) & CLERS_DEL;
				if(skip) continue;
			}
			if(!skip) { LD_LOADB(f, prev, next); lp = ln = 0; have = 1; LD_LOG(LG_P, f); }
			else {                                     // nothing pending: start triangle
				if(n == 0) break;
				n--;
				uint32_t c, b;
				LD_FETCH(c);
				b = nfront; bad |= (nfront + 3 > cap); nfront += (nfront + 3 <= cap) ? 3u : 0u;
				rg.stB(b, b + 2, b + 1); rg.stB(b + 1, b + 0, b + 2); rg.stB(b + 2, b + 1, b + 0);
				rg.stFl(b, 0); rg.stFl(b + 1, 0); rg.stFl(b + 2, 0);
				LD_LOG(c == C_SPLIT ? LG_TS : LG_TV, b);
				start += 1;
				continue;
			}
		}
		while(n) {
			// a run of VERTEX / LEFT symbols ahead: yield to the caller's warp-wide window step (k_clers_lf: lead_vector)
			if(vec && n >= 2 && (cler & 7u) <= 6u && ((uint32_t)cw & 0xfefeu) == 0 && end - start >= 2) { rc = 3; break; }
			n--;
			uint32_t c;
			LD_FETCH(c);
			if(c == C_VERTEX || c == C_SPLIT) {
				uint32_t b;
				LD_NEWID(b);
				rg.stB(b, CLERS_NOLINK, next); rg.stFl(b, 0);   // second new edge: persistent, queued; its prev link is deferred
				LD_SET_PREV(next, b);
				LD_LOG(c == C_VERTEX ? LG_V : LG_S, b);
				start++;
				next = b; lp = 1; ln = 1; f = CLERS_NOID;
			} else if(c == C_LEFT) {
				if((lp | ln) && prev == next) LD_MATERIALISE();
				uint32_t pp, pn;
				LD_LOADB(prev, pp, pn);
				(void)pn;
				LD_SET_FLAG(prev, CLERS_DEL);
				LD_LOG(LG_L, prev);
				start++;
				prev = pp; lp = 1; ln = 1; f = CLERS_NOID;
			} else if(c == C_RIGHT) {
				if((lp | ln) && prev == next) LD_MATERIALISE();
				uint32_t np, nn;
				LD_LOADB(next, np, nn);
				(void)np;
				LD_SET_FLAG(next, CLERS_DEL);
				LD_LOG(LG_R, next);
				start++;
				next = nn; lp = 1; ln = 1; f = CLERS_NOID;
			} else if(c == C_END) {
				if((lp | ln) && prev == next) LD_MATERIALISE();
				uint32_t pp, pn, np, nn;
				LD_LOADB(prev, pp, pn);
				LD_LOADB(next, np, nn);
				(void)pn;
				(void)np;
				LD_SET_FLAG(prev, CLERS_DEL);
				LD_SET_FLAG(next, CLERS_DEL);
				LD_SET_NEXT(pp, nn);
				LD_SET_PREV(nn, pp);
				LD_LOG(LG_E, prev);
				start++;
				have = 0; break;
			} else {                                       // BOUNDARY, DELAY (anything else is a corrupt stream: treated as BOUNDARY + flag)
				bad |= (c != C_BOUNDARY && c != C_DELAY);
				if(f == CLERS_NOID) LD_MATERIALISE();
				if(c == C_DELAY) io.delayed[ndel++] = f;
				have = 0; break;
			}
			if(start >= end) { have = 0; break; }
		}
		if(have) break;                                    // budget used up in the middle of a strip
	}
	if(rc == 0 && (bad || (cler >= nclers && !(start >= end && g >= io.ngroups)))) rc = -5;
	if(rc == 4 && bad) rc = -5;   // flagged, or the stream ran dry with faces missing
	S.cler = cler; S.start = start; S.end = end; S.nfront = nfront; S.scan = scan; S.ndelayed = ndel;
	S.cw = cw; S.cw_next = cwn; S.have = have; S.lp = lp; S.ln = ln; S.cf = f; S.cprev = prev; S.cnext = next; S.g = g; S.nlog = nlog; S.bad = bad;
	return rc;
#undef LD_FETCH
#undef LD_LOG
#undef LD_SET_NEXT
#undef LD_SET_PREV
#undef LD_LOADB
#undef LD_FLAG
#undef LD_SET_FLAG
#undef LD_NEWID
#undef LD_MATERIALISE
}

struct FollowState {
	uint32_t v0, v1, v2;               // labels of the current edge
	uint32_t vcount, nfaces;           // vertices created / faces emitted so far
	uint64_t splitpos;
	uint32_t gstart;                   // face index where the current group starts (for malformed group tables)
	uint32_t aflush, amax;             // label ids below aflush live in global memory; amax = ids written so far
	uint32_t fflush, pflush;           // faces / predictions already drained
	uint32_t tail;                     // log words consumed
};
CRT_HD void follow_init(FollowState &S) {
	S.v0 = S.v1 = S.v2 = 0; S.vcount = 0; S.nfaces = 0; S.splitpos = 0; S.gstart = 0; S.aflush = S.amax = 0; S.fflush = S.pflush = 0; S.tail = 0;
}
CRT_COLD static uint4_t follow_g_load(const EdgeA *ea, uint32_t x) { const EdgeA a = ea[x]; return uint4_t{a.v0, a.v1, a.v2, 0}; }

// Label machine: replays log words [S.tail, upto).  Stops early (returns 0 with S.tail < upto) when a staging ring is
// full (`room` entries left at call time) so the caller can drain; returns 0 normally, < 0 on an inconsistent stream.
template <class RG> CRT_HD int clers_follow(const ClersIO &io, RG &rg, FollowState &S, uint32_t upto, uint32_t room, int splitbits, bool vec = false) {
	uint32_t v0 = S.v0, v1 = S.v1, v2 = S.v2, vcount = S.vcount, nf = S.nfaces, tail = S.tail, amax = S.amax;
	uint64_t splitpos = S.splitpos;
	const uint32_t aflush = S.aflush, nvert = io.nvert, nface = io.nface;
	int rc = 0;
#define FW_LOADA(ID_, A_, B_, C_) do { if((ID_) >= aflush) rg.ldA(ID_, A_, B_, C_); else { const uint4_t t_ = follow_g_load(io.ea, ID_); A_ = t_.x; B_ = t_.y; C_ = t_.z; } } while(0)
	while(tail < upto && room >= 3) {
		const uint32_t w = rg.ldLog(tail);
		const uint32_t t = w >> 28, id = w & 0x0FFFFFFFu;
		if(vec && (t == LG_V || t == LG_L) && tail + 1 < upto) {     // two VERTEX/LEFT words in a row: yield to the warp-wide window step
			const uint32_t t2 = rg.ldLog(tail + 1) >> 28;
			if(t2 == LG_V || t2 == LG_L) { rc = 3; break; }
		}
		tail++;
		if(t == LG_V || t == LG_S) {
			uint32_t opp;
			if(t == LG_V) {
				if(vcount >= nvert) { rc = -5; break; }
				rg.stP(vcount, v1, v0, v2);
				opp = vcount++;
			} else {
				opp = getbits(io.split, io.split_nwords, splitpos, splitbits); splitpos += (uint64_t)splitbits;
				if(opp >= nvert) { rc = -5; break; }
			}
			if(nf >= nface) { rc = -5; break; }
			rg.stA(id, opp, v1, v0); amax = id + 1;
			rg.stF(nf, v1, v0, opp); nf++;
			v2 = v1; v1 = opp;
			room--;
		} else if(t == LG_L) {
			uint32_t a, b, c;
			FW_LOADA(id, a, b, c); (void)b; (void)c;
			if(nf >= nface) { rc = -5; break; }
			rg.stF(nf, v1, v0, a); nf++;
			v2 = v0; v0 = a;
			room--;
		} else if(t == LG_R) {
			uint32_t a, b, c;
			FW_LOADA(id, a, b, c); (void)a; (void)c;
			if(nf >= nface) { rc = -5; break; }
			rg.stF(nf, v1, v0, b); nf++;
			v2 = v1; v1 = b;
			room--;
		} else if(t == LG_P) {
			FW_LOADA(id, v0, v1, v2);
		} else if(t == LG_M) {
			rg.stA(id, v0, v1, v2); amax = id + 1;
		} else if(t == LG_E) {
			uint32_t a, b, c;
			FW_LOADA(id, a, b, c); (void)b; (void)c;
			if(nf >= nface) { rc = -5; break; }
			rg.stF(nf, v1, v0, a); nf++;
			room--;
		} else if(t == LG_TV || t == LG_TS) {          // start triangle (decoder.cpp:224-259)
			uint32_t last = vcount - 1, vi[3], mask = 0;
			if(t == LG_TS) { mask = getbits(io.split, io.split_nwords, splitpos, 3); splitpos += 3; }
			for(int k = 0; k < 3; k++) {
				uint32_t v;
				if(mask & (1u << k)) {
					v = getbits(io.split, io.split_nwords, splitpos, splitbits); splitpos += (uint64_t)splitbits;
					if(v >= nvert) { rc = -5; break; }
				} else {
					if(vcount >= nvert) { rc = -5; break; }
					rg.stP(vcount, last, last, last);
					last = v = vcount++;
				}
				vi[k] = v;
			}
			if(rc || nf >= nface) { rc = -5; break; }
			rg.stF(nf, vi[0], vi[1], vi[2]); nf++;
			rg.stA(id, vi[1], vi[2], vi[0]); rg.stA(id + 1, vi[2], vi[0], vi[1]); rg.stA(id + 2, vi[0], vi[1], vi[2]); amax = id + 3;
			room -= 3;
		} else if(t == LG_G) {                         // group restart: label ids start over; a malformed table may move the face cursor
			uint32_t st = id ? io.group_ends[id - 1] : 0;
			if(st > nface) st = nface;
			if(st != nf || S.aflush) { S.tail = tail - 1; S.gstart = st; rc = 2; break; }   // caller drains, resets aflush / nf, replays the word
		} else { rc = -5; break; }
	}
#undef FW_LOADA
	if(rc != 2) S.tail = tail;
	S.v0 = v0; S.v1 = v1; S.v2 = v2; S.vcount = vcount; S.nfaces = nf; S.splitpos = splitpos; S.amax = amax;
	return rc;
}

// ---- CLERS automaton, v7: ONE merged machine per mesh, run by a whole CTA (k_clers_cta) -----------------------------------
// v4-v6 split links and labels over two warps to shorten the serial chain; what bounds them is still one warp stepping
// through 32 symbols at a time (~40 dependent collectives per window).  v7 turns the closed form of a VERTEX/LEFT run into a
// CTA-wide scan: 256 threads take 256 symbols per step, thread = symbol, two barriers per step, links AND labels in the same
// pass (so there is no log and no second machine):
//   V at rank r among the V's:  new vertex x = vcount + r, new queued edge b = nfront + r with links (b+1 | deferred, b-1 | next)
//   L at rank k among the L's:  consumes chain[k], the k-th edge of the prev chain (prev, prev.prev, ...) — usually the
//                               consecutive ids prev, prev+1, ... of ONE earlier strip, checked with one load per L
//   labels: v1 = last vertex created before me, v0 = label of the edge the last L before me consumed, v2 = v1 / v0 of the symbol
//           before me — all functions of the two ranks.
// Everything else (BOUNDARY, DELAY, RIGHT, END, SPLIT, start triangles, runs shorter than the threshold) goes through the scalar
// machine below on thread 0, which shares the rings (labels 16 B + links 8 B + flag byte per materialised edge, implicit
// FIFO = scan of the flag bytes, 256 per step by the whole CTA).  Faces and predictions go straight to global memory.
struct MergedState {
	uint32_t cler, vcount;
	uint64_t splitpos;
	uint32_t g, start, end;            // current group; face cursor / end of group, in FACES
	uint32_t nfront, scan, ndel;       // ids allocated; next id the implicit FIFO looks at; delayed-stack depth
	uint32_t have, lp, ln, cf;         // current edge valid; deferred incoming links; its id (CLERS_NOID while it has no record)
	uint32_t v0, v1, v2, prev, next;   // current edge
	uint32_t eflush;                   // ids below live in the global backing store
	uint32_t bad;
};
CRT_HD void merged_init(MergedState &S) {
	S.cler = S.vcount = 0; S.splitpos = 0; S.g = S.start = S.end = 0; S.nfront = S.scan = S.ndel = 0;
	S.have = S.lp = S.ln = 0; S.cf = CLERS_NOID; S.v0 = S.v1 = S.v2 = S.prev = S.next = 0; S.eflush = 0; S.bad = 0;
}
CRT_COLD static void merged_g_storeA(EdgeA *ea, uint32_t x, uint32_t a, uint32_t b, uint32_t c) { ea[x] = EdgeA{a, b, c, 0}; }
CRT_COLD static void merged_g_storeB(EdgeB *eb, uint32_t x, uint32_t p, uint32_t n) { eb[x] = EdgeB{p, n}; }

// Scalar machine.  RG: ldA/stA (labels), ldB/stB/stB_prev/stB_next (links), ldFl/stFl (flags), sym(i) (symbol i of the stream;
// valid for i < nclers).  Consumes at most `budget` symbols.  Returns 1 when all groups are done, 0 to be called again (budget),
// 3 (only with vec) when at least `runmin` VERTEX/LEFT symbols lie ahead of a valid current edge (caller: CTA-wide window),
// 4 (only with vec) when the implicit FIFO has to be scanned (caller: CTA-wide pop), < 0 on a topology error.
// Errors set a sticky flag and indices are clamped (nothing is written out of bounds); the flag is reported when the chunk ends.
template <class RG> CRT_HD int clers_merged(const ClersIO &io, RG &rg, MergedState &S, int budget, bool vec, uint32_t runmin, int splitbits) {
	uint32_t cler = S.cler, vcount = S.vcount, start = S.start, end = S.end, g = S.g;
	uint32_t nfront = S.nfront, scan = S.scan, ndel = S.ndel, bad = S.bad;
	uint64_t splitpos = S.splitpos;
	uint32_t have = S.have, lp = S.lp, ln = S.ln, f = S.cf, v0 = S.v0, v1 = S.v1, v2 = S.v2, prev = S.prev, next = S.next;
	uint32_t eflush = S.eflush;
	const uint32_t nclers = io.nclers, nvert = io.nvert, cap = io.cap;
	uint32_t n = nclers - cler;
	if(n > (uint32_t)budget) n = (uint32_t)budget;
	int rc = 0;
#define MG_SET_NEXT(x, v) do { if((x) >= eflush) rg.stB_next(x, v); else clers_g_set_next(io.eb, x, v); } while(0)
#define MG_SET_PREV(x, v) do { if((x) >= eflush) rg.stB_prev(x, v); else clers_g_set_prev(io.eb, x, v); } while(0)
#define MG_LOADB(ID_, P_, Q_) do { if((ID_) >= eflush) rg.ldB(ID_, P_, Q_); else { const uint2_t t_ = lead_g_load(io.eb, ID_); P_ = t_.x; Q_ = t_.y; } } while(0)
#define MG_LOADA(ID_, A_, B_, C_) do { if((ID_) >= eflush) rg.ldA(ID_, A_, B_, C_); else { const uint4_t t_ = follow_g_load(io.ea, ID_); A_ = t_.x; B_ = t_.y; C_ = t_.z; } } while(0)
#define MG_FLAG(ID_) (((ID_) >= eflush) ? rg.ldFl(ID_) : lead_g_flag(io.fl, ID_))
#define MG_SET_FLAG(ID_, V_) do { if((ID_) >= eflush) rg.stFl(ID_, V_); else lead_g_set_flag(io.fl, ID_, V_); } while(0)
#define MG_NEWID(ID_) do { ID_ = nfront; bad |= (nfront >= cap); nfront += (nfront < cap); } while(0)
#define MG_MATERIALISE()                                                                                 \
	do {                                                                                                 \
		MG_NEWID(f);                                                                                     \
		rg.stB(f, prev, next); rg.stA(f, v0, v1, v2); rg.stFl(f, CLERS_NQ);                              \
		if(lp) MG_SET_NEXT(prev, f);                                                                     \
		if(ln) MG_SET_PREV(next, f);                                                                     \
		lp = ln = 0;                                                                                     \
	} while(0)
#define MG_FACE(A_, B_, C_) do { if(start < io.nface) clers_put_face(io, (size_t)start*3u, A_, B_, C_); else bad = 1; start++; } while(0)
	for(;;) {
		if(!have) {
			if(start >= end) {                         // next group: fresh front (decoder.cpp:173-178, 207-221)
				if(g >= io.ngroups) { rc = 1; break; }
				uint32_t e = io.group_ends[g];
				if(e > io.nface) e = io.nface;
				uint32_t st = g ? io.group_ends[g - 1] : 0;
				if(st > io.nface) st = io.nface;
				g++;
				start = st; end = e;
				nfront = scan = ndel = 0; eflush = 0;
				continue;
			}
			uint32_t skip = 1;
			if(vec && scan < nfront) { rc = 4; break; }   // caller scans the flag bytes 256 at a time
			while(scan < nfront) {                     // implicit FIFO: next queued, alive edge in id order (decoder.cpp:265-266, 278-279)
				f = scan++;
				skip = MG_FLAG(f);
				if(!skip) break;
			}
			if(skip && ndel) {                         // delayed stack (decoder.cpp:267-269)
				f = io.delayed[--ndel];
				skip = MG_FLAG(f) & CLERS_DEL;
				if(skip) continue;
			}
			if(!skip) { MG_LOADB(f, prev, next); MG_LOADA(f, v0, v1, v2); lp = ln = 0; have = 1; }
			else {                                     // nothing pending: start triangle (decoder.cpp:224-259)
				if(n == 0) break;
				n--;
				const uint32_t c = rg.sym(cler); cler++;
				uint32_t last = vcount - 1, vi[3], mask = 0;
				if(c == C_SPLIT) { mask = getbits(io.split, io.split_nwords, splitpos, 3); splitpos += 3; }
				else bad |= (c != C_VERTEX);
				for(int k = 0; k < 3; k++) {
					uint32_t v;
					if(mask & (1u << k)) {
						v = getbits(io.split, io.split_nwords, splitpos, splitbits); splitpos += (uint64_t)splitbits;
						if(v >= nvert) { bad = 1; v = 0; }
					} else if(vcount < nvert) {
						clers_put_pred(io, vcount, last, last, last);
						last = v = vcount++;
					} else { bad = 1; v = 0; }
					vi[k] = v;
				}
				MG_FACE(vi[0], vi[1], vi[2]);
				const uint32_t b = nfront;
				bad |= (nfront + 3 > cap); nfront += (nfront + 3 <= cap) ? 3u : 0u;
				if(!bad) {
					rg.stA(b, vi[1], vi[2], vi[0]);     rg.stB(b, b + 2, b + 1);     rg.stFl(b, 0);
					rg.stA(b + 1, vi[2], vi[0], vi[1]); rg.stB(b + 1, b + 0, b + 2); rg.stFl(b + 1, 0);
					rg.stA(b + 2, vi[0], vi[1], vi[2]); rg.stB(b + 2, b + 1, b + 0); rg.stFl(b + 2, 0);
				}
				continue;
			}
		}
		while(n) {
			if(vec && n >= runmin && end - start >= runmin) {      // a run of VERTEX / LEFT symbols ahead: CTA-wide window
				uint32_t k = 0;
				while(k < runmin && rg.sym(cler + k) <= (uint32_t)C_LEFT) k++;
				if(k == runmin) { rc = 3; break; }
			}
			n--;
			const uint32_t c = rg.sym(cler); cler++;
			if(c == C_VERTEX || c == C_SPLIT) {
				uint32_t opp, b;
				if(c == C_VERTEX) {
					if(vcount < nvert) { clers_put_pred(io, vcount, v1, v0, v2); opp = vcount++; } else { bad = 1; opp = 0; }
				} else {
					opp = getbits(io.split, io.split_nwords, splitpos, splitbits); splitpos += (uint64_t)splitbits;
					if(opp >= nvert) { bad = 1; opp = 0; }
				}
				MG_NEWID(b);
				rg.stA(b, opp, v1, v0); rg.stB(b, CLERS_NOLINK, next); rg.stFl(b, 0);   // second new edge: persistent, queued; its prev link is deferred
				MG_SET_PREV(next, b);
				MG_FACE(v1, v0, opp);
				v2 = v1; v1 = opp; next = b; lp = 1; ln = 1; f = CLERS_NOID;               // first new edge: registers only
			} else if(c == C_LEFT) {
				if((lp | ln) && prev == next) MG_MATERIALISE();                           // 2-edge loop: deferred fields would be read
				uint32_t pp, pn, a, t1, t2;
				MG_LOADB(prev, pp, pn); MG_LOADA(prev, a, t1, t2);
				(void)pn; (void)t1; (void)t2;
				MG_SET_FLAG(prev, CLERS_DEL);
				MG_FACE(v1, v0, a);
				v2 = v0; v0 = a; prev = pp; lp = 1; ln = 1; f = CLERS_NOID;
			} else if(c == C_RIGHT) {
				if((lp | ln) && prev == next) MG_MATERIALISE();
				uint32_t np, nn, t0, b1, t2;
				MG_LOADB(next, np, nn); MG_LOADA(next, t0, b1, t2);
				(void)np; (void)t0; (void)t2;
				MG_SET_FLAG(next, CLERS_DEL);
				MG_FACE(v1, v0, b1);
				v2 = v1; v1 = b1; next = nn; lp = 1; ln = 1; f = CLERS_NOID;
			} else if(c == C_END) {
				if((lp | ln) && prev == next) MG_MATERIALISE();
				uint32_t pp, pn, np, nn, a, t1, t2;
				MG_LOADB(prev, pp, pn); MG_LOADB(next, np, nn); MG_LOADA(prev, a, t1, t2);
				(void)pn; (void)np; (void)t1; (void)t2;
				MG_SET_FLAG(prev, CLERS_DEL);
				MG_SET_FLAG(next, CLERS_DEL);
				MG_SET_NEXT(pp, nn);
				MG_SET_PREV(nn, pp);
				MG_FACE(v1, v0, a);
				have = 0; break;
			} else {                                       // BOUNDARY, DELAY (anything else is a corrupt stream: treated as BOUNDARY + flag)
				bad |= (c != C_BOUNDARY && c != C_DELAY);
				if(f == CLERS_NOID) MG_MATERIALISE();
				if(c == C_DELAY) { if(ndel < cap) io.delayed[ndel++] = f; else bad = 1; }
				have = 0; break;
			}
			if(start >= end) { have = 0; break; }          // group complete: the front is discarded (decoder.cpp:223)
		}
		if(have) break;                                    // budget used up in the middle of a strip / yield to a window
	}
	if(rc == 0 && (bad || (cler >= nclers && !(start >= end && g >= io.ngroups)))) rc = -5;   // flagged, or the stream ran dry with faces missing
	if((rc == 3 || rc == 4) && bad) rc = -5;
	S.cler = cler; S.vcount = vcount; S.start = start; S.end = end; S.g = g; S.nfront = nfront; S.scan = scan; S.ndel = ndel; S.bad = bad;
	S.splitpos = splitpos; S.have = have; S.lp = lp; S.ln = ln; S.cf = f; S.v0 = v0; S.v1 = v1; S.v2 = v2; S.prev = prev; S.next = next;
	S.eflush = eflush;
	return rc;
#undef MG_SET_NEXT
#undef MG_SET_PREV
#undef MG_LOADB
#undef MG_LOADA
#undef MG_FLAG
#undef MG_SET_FLAG
#undef MG_NEWID
#undef MG_MATERIALISE
#undef MG_FACE
}

// Plain-array ring policy (tests/host_emul; the kernel has its own shared-memory policy with the same interface).
struct ArrayRings {
	uint4_t *ra; uint2_t *rb; uint32_t *rq; uint4_t *sf; uint4_t *sp;
	uint32_t RM, QM, FM, PM;
	CRT_HD void ldA(uint32_t id, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) const { const uint4_t v = ra[id & RM]; a = v.x; b = v.y; c = v.z; d = v.w; }
	CRT_HD void stA(uint32_t id, uint32_t a, uint32_t b, uint32_t c, uint32_t d) { ra[id & RM] = uint4_t{a, b, c, d}; }
	CRT_HD void stA_del(uint32_t id) { ra[id & RM].w = 1; }
	CRT_HD void ldB(uint32_t id, uint32_t &p, uint32_t &n) const { const uint2_t v = rb[id & RM]; p = v.x; n = v.y; }
	CRT_HD void stB(uint32_t id, uint32_t p, uint32_t n) { rb[id & RM] = uint2_t{p, n}; }
	CRT_HD void stB_prev(uint32_t id, uint32_t p) { rb[id & RM].x = p; }
	CRT_HD void stB_next(uint32_t id, uint32_t n) { rb[id & RM].y = n; }
	CRT_HD uint32_t ldQ(uint32_t i) const { return rq[i & QM]; }
	CRT_HD void stQ(uint32_t i, uint32_t v) { rq[i & QM] = v; }
	CRT_HD void stF(uint32_t face, uint32_t a, uint32_t b, uint32_t c) { sf[face & FM] = uint4_t{a, b, c, 0}; }
	CRT_HD void stP(uint32_t v, uint32_t a, uint32_t b, uint32_t c) { sp[v & PM] = uint4_t{a, b, c, 0}; }
	// v4 (leader / follower) extras: 3-word labels in `ra` with their own mask, log ring
	uint32_t *lg; uint32_t LM, AM;
	CRT_HD void ldA(uint32_t id, uint32_t &a, uint32_t &b, uint32_t &c) const { const uint4_t v = ra[id & AM]; a = v.x; b = v.y; c = v.z; }
	CRT_HD void stA(uint32_t id, uint32_t a, uint32_t b, uint32_t c) { ra[id & AM] = uint4_t{a, b, c, 0}; }
	CRT_HD void stLog(uint32_t i, uint32_t w) { lg[i & LM] = w; }
	CRT_HD uint32_t ldLog(uint32_t i) const { return lg[i & LM]; }
	uint8_t *rf;
	CRT_HD uint32_t ldFl(uint32_t id) const { return rf[id & RM]; }
	CRT_HD void stFl(uint32_t id, uint32_t v) { rf[id & RM] = (uint8_t)v; }
	// v7 (merged machine): the symbol stream
	const uint8_t *syms;
	CRT_HD uint32_t sym(uint32_t i) const { return syms[i]; }
};

}  // namespace crtb
