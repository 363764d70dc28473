/* oracle/crt_oracle.c — TEST INFRASTRUCTURE, not product code.
 *
 * A plain-C, single-threaded CPU restatement of the reference's .crt DECODE path, written from the
 * algorithm (not translated line by line), kept deliberately sequential and obvious.  It is the checker
 * the CUDA path is compared against on machines where /root/reference does not exist (the GPU box),
 * and bench.py's `cpu_baseline.kind = "port"` fallback.  Only tests/, __graft_entry__.smoke() and
 * bench.py's baseline legs may load liboracle.so; the product never links or calls it.
 *
 * PARITY PINNED: tests/test_oracle.py checks this file bit-for-bit against (a) the unmodified reference
 * compiled in place (oracle/_ref/libcorto_ref.so) on every fixture category of SURVEY §8c, (b) the committed
 * golden vectors under tests/golden/ produced by that reference (tests/golden/make_golden.py), and
 * (c) the reference's only shipped .crt (html/models/tarta.crt) digests.
 *
 * Every function cites the reference file:line (paths relative to /root/reference) it restates.
 * All integer arithmetic is 2's-complement wrapping (-fwrapv); all float arithmetic is IEEE binary32
 * without contraction (-ffp-contract=off), float->int conversions follow x86 cvttss2si.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ------------------------------------------------------------------------------------------------ */
/* Formats / enums — include/corto/vertex_attribute.h:30-33, normal_attribute.h:42-44, index_attribute.h:26 */
enum { F_UINT32 = 0, F_INT32, F_UINT16, F_INT16, F_UINT8, F_INT8, F_FLOAT, F_DOUBLE };
enum { S_PARALLEL = 1, S_CORRELATED = 2 };
enum { CODEC_GENERIC = 1, CODEC_NORMAL = 2, CODEC_COLOR = 3 };
enum { N_DIFF = 0, N_ESTIMATED = 1, N_BORDER = 2 };
enum { C_VERTEX = 0, C_LEFT, C_RIGHT, C_END, C_BOUNDARY, C_DELAY, C_SPLIT };

#define ORACLE_MAX_ATTR 16

typedef struct {
	char name[64];
	int codec;
	float q;
	int N;            /* header component count (3 for normals although the stream carries 2) */
	int format;       /* header format (encoder input); overwritten by the binding */
	int strategy;
} OAttrInfo;

typedef struct {
	uint32_t nvert, nface;
	int entropy;
	int nattr;
	OAttrInfo attr[ORACLE_MAX_ATTR];
	uint32_t body;    /* byte offset of the group table (first byte after nvert/nface) */
} OInfo;

/* One output binding per attribute, same order as OInfo.attr.  buffer==NULL: parse + skip. */
typedef struct {
	void *buffer;
	int format;          /* F_FLOAT / F_INT16 (normals) / F_UINT8 (colour) / F_INT32|F_UINT32 (generic) */
	int out_components;  /* colour only */
} OBind;

/* ------------------------------------------------------------------------------------------------ */
/* Byte cursor — include/corto/cstream.h:246-291 (little-endian scalars, u16-length strings, 4-byte
 * aligned embedded bitstreams). */
typedef struct { const uint8_t *base, *p; } Cur;

static uint32_t rd8(Cur *c)  { return *c->p++; }
static uint32_t rd16(Cur *c) { uint32_t v = c->p[0] | (c->p[1] << 8); c->p += 2; return v; }
static uint32_t rd32(Cur *c) { uint32_t v = c->p[0] | (c->p[1] << 8) | (c->p[2] << 16) | ((uint32_t)c->p[3] << 24); c->p += 4; return v; }
static const char *rdstr(Cur *c) { uint32_t n = rd16(c); const char *s = (const char *)c->p; c->p += n; return s; }

typedef struct { const uint32_t *w; uint32_t nwords; } Bits;

static Bits rdbits(Cur *c) {                       /* cstream.h:283-291 */
	Bits b;
	b.nwords = rd32(c);
	size_t off = (size_t)(c->p - c->base);
	if(off & 3) c->p += 4 - (off & 3);
	b.w = (const uint32_t *)c->p;
	c->p += (size_t)b.nwords*4;
	return b;
}

/* MSB-first n-bit field (0..32) starting at absolute bit `pos` — random-access form of
 * BitStream::read, src/bitstream.cpp:103-121.  The sequential reader refills lazily, so the word after
 * the last needed one is never touched; neither is it here. */
static uint32_t getbits(const Bits *b, uint64_t pos, int n) {
	if(n == 0) return 0;
	uint64_t w = pos >> 5;
	int o = (int)(pos & 31);
	uint32_t hi = b->w[w];
	if(o + n <= 32)
		return (hi << o) >> (32 - n);          /* o<=31, 32-n<=31 */
	uint32_t lo = b->w[w + 1];
	uint64_t win = ((uint64_t)hi << 32) | lo;
	return (uint32_t)((win << o) >> (64 - n));
}

/* ------------------------------------------------------------------------------------------------ */
/* Tunstall — src/tunstall.cpp:125-256 (createDecodingTables2) and :430-452 (decompress).
 * The dictionary is grown as `nsym` FIFO rows (row r = words ending in symbol r, kept in creation
 * order in slots r, r+n, r+2n, ...).  Each round pops the most probable row head (first strict max)
 * and appends its nsym one-symbol extensions; the last round is cut when the 256th word appears and then
 * the popped parent is NOT removed.  Probabilities are 16-bit fixed point (p<<8), products >>16. */
#define TUN_TABLE_CAP 16384   /* reference sizes it 8192 (tunstall.cpp:138); slack keeps a corrupt stream in bounds */

typedef struct {
	int nsym;
	uint8_t sym[256], prob[256];
	int index[512];
	int length[512];
	uint8_t table[TUN_TABLE_CAP];
} Tun;

static void tun_build(Tun *t) {
	const uint32_t n = (uint32_t)t->nsym;
	if(n <= 1) return;                               /* tunstall.cpp:127 */
	uint32_t qprob[1024] = {0};
	int *widx = t->index, *wlen = t->length;         /* slot -> (offset,len); compacted in place at the end */
	uint32_t head[256];
	uint8_t *buf = t->table;
	uint32_t pos = 0, slots = 0, nwords;

	/* how many times the top symbol can be repeated before it is less likely than the runner-up (:140-147) */
	uint32_t p0 = (uint32_t)t->prob[0] << 8, p1 = (uint32_t)t->prob[1] << 8;
	uint32_t run = 2, pr = (p0*p0) >> 16, max_run = 255/(n - 1);
	while(pr > p1 && run < max_run) { pr = (pr*p0) >> 16; run++; }

	if(run >= 16) {
		/* low-entropy start (:149-194): words  s0^run  and  s0^c s_k  (c = 0..run-1, k = 1..n-1) */
		buf[pos++] = t->sym[0];
		for(uint32_t k = 1; k < n; k++) {
			for(uint32_t i = 0; i + 1 < run; i++) buf[pos++] = t->sym[0];
			buf[pos++] = t->sym[k];
		}
		head[0] = (run - 1)*n;
		for(uint32_t k = 1; k < n; k++) head[k] = k;
		for(uint32_t c = 0; c < run; c++) {
			for(uint32_t k = 1; k < n; k++) {
				uint32_t s = k + c*n, pk = (uint32_t)t->prob[k] << 8;
				qprob[s] = (c == 0) ? pk : ((pr*pk) >> 16);
				widx[s] = (int)(k*run - c);
				wlen[s] = (int)(c + 1);
			}
			pr = (c == 0) ? p0 : ((pr*p0) >> 16);
		}
		uint32_t s0 = (run - 1)*n;
		qprob[s0] = pr; widx[s0] = 0; wlen[s0] = (int)run;
		nwords = 1 + run*(n - 1);
		slots = run*n;
	} else {
		for(uint32_t k = 0; k < n; k++) {            /* one-symbol words (:198-206) */
			head[k] = k;
			qprob[slots] = (uint32_t)t->prob[k] << 8;
			widx[slots] = (int)pos; wlen[slots] = 1; slots++;
			buf[pos++] = t->sym[k];
		}
		nwords = n;
	}

	while(nwords < 256) {                            /* :208-240 */
		uint32_t best = 0, bestp = 0;
		for(uint32_t k = 0; k < n; k++) {
			uint32_t p = qprob[head[k]];
			if(p > bestp) { bestp = p; best = k; }
		}
		uint32_t parent = head[best], pp = qprob[parent];
		uint32_t poff = (uint32_t)widx[parent], plen = (uint32_t)wlen[parent];
		uint32_t k = 0;
		for(; k < n; k++) {
			qprob[slots] = (pp*((uint32_t)t->prob[k] << 8)) >> 16;
			widx[slots] = (int)pos; wlen[slots] = (int)plen + 1; slots++;
			if(pos + plen + 1 <= TUN_TABLE_CAP) memmove(buf + pos, buf + poff, plen);
			pos += plen;
			if(pos < TUN_TABLE_CAP) buf[pos] = t->sym[k];
			pos++;
			if(nwords + k == 255) break;
		}
		if(k == n) head[best] += n;                  /* parent leaves the dictionary only after a full split */
		nwords += n - 1;
	}

	uint32_t word = 0;                                /* compaction (:242-255): keep slots not popped */
	for(uint32_t s = 0; s < slots; s++) {
		if(head[s % n] > s) continue;
		widx[word] = widx[s]; wlen[word] = wlen[s]; word++;
	}
}

/* src/cstream.cpp:66-87,111-128 + tunstall.cpp:430-452.  Returns a malloc'ed symbol array. */
static uint8_t *entropy_decompress(Cur *c, int entropy, uint32_t *out_size) {
	if(entropy == 0) {                               /* Stream::NONE: u32 size + raw bytes */
		uint32_t size = rd32(c);
		uint8_t *o = (uint8_t *)malloc(size + 1);
		memcpy(o, c->p, size); c->p += size;
		*out_size = size;
		return o;
	}
	Tun *t = (Tun *)malloc(sizeof(Tun));
	t->nsym = (int)rd8(c);
	for(int i = 0; i < t->nsym; i++) { t->sym[i] = c->p[0]; t->prob[i] = c->p[1]; c->p += 2; }
	tun_build(t);
	uint32_t size = rd32(c), csize = rd32(c);
	const uint8_t *data = c->p; c->p += csize;
	uint8_t *o = (uint8_t *)malloc((size_t)size + 1);
	*out_size = size;
	if(size) {
		if(t->nsym == 1) memset(o, t->sym[0], size);           /* tunstall.cpp:433-436 */
		else if(csize > 0) {
			uint32_t w = 0;
			for(uint32_t i = 0; i + 1 < csize; i++) {            /* every byte but the last: whole word */
				uint32_t len = (uint32_t)t->length[data[i]], st = (uint32_t)t->index[data[i]];
				for(uint32_t k = 0; k < len && w < size; k++) o[w++] = t->table[st + k];
			}
			uint32_t st = (uint32_t)t->index[data[csize - 1]];     /* last byte: exactly the remainder (:446-451) */
			for(uint32_t k = 0; w < size; k++) o[w++] = t->table[(st + k) % TUN_TABLE_CAP];
		}
	}
	free(t);
	return o;
}

/* ------------------------------------------------------------------------------------------------ */
/* Residual decoders.  out is laid out [i*N + c]; esize 4 -> int32 values, 1 -> uint8 values. */

/* `max` term of decodeArray, include/corto/cstream.h:344: `const uint64_t max = (1<<diff)>>1` with an int
 * shift: diff=31 gives INT_MIN>>1 = 0xC0000000 (sign-extended, low 32 bits matter), diff=32 shifts by
 * 32&31=0 on x86 so the term is 0. */
static uint32_t array_bias(int d) {
	int32_t one = (int32_t)((uint32_t)1 << (d & 31));
	return (uint32_t)(one >> 1);
}

/* cstream.h:324-360 — one log per vertex shared by the N components. */
static uint32_t decode_array(Cur *c, int entropy, void *values, int N, int esize) {
	Bits b = rdbits(c);
	uint32_t n;
	uint8_t *logs = entropy_decompress(c, entropy, &n);
	if(values) {
		uint64_t bit = 0;
		for(uint32_t i = 0; i < n; i++) {
			int d = logs[i];
			for(int k = 0; k < N; k++) {
				uint32_t v = 0;
				if(d) { v = getbits(&b, bit, d > 32 ? 32 : d) - array_bias(d); bit += (uint64_t)d; }
				if(esize == 4) ((int32_t *)values)[(size_t)i*N + k] = (int32_t)v;
				else ((uint8_t *)values)[(size_t)i*N + k] = (uint8_t)v;
			}
		}
	}
	free(logs);
	return n;
}

/* cstream.h:294-319 — one Tunstall block of logs per component, all components share ONE bitstream,
 * component c's fields follow all of component c-1's. */
static uint32_t decode_values(Cur *c, int entropy, void *values, int N, int esize) {
	Bits b = rdbits(c);
	uint64_t bit = 0;
	uint32_t n = 0;
	for(int k = 0; k < N; k++) {
		uint8_t *logs = entropy_decompress(c, entropy, &n);
		if(values) {
			for(uint32_t i = 0; i < n; i++) {
				int d = logs[i];
				int32_t val = 0;
				if(d) {
					val = (int32_t)getbits(&b, bit, d > 32 ? 32 : d);
					bit += (uint64_t)d;
					int32_t middle = (int32_t)((uint32_t)1 << ((d - 1) & 31));
					if(val < middle) val = -val - middle;
				}
				if(esize == 4) ((int32_t *)values)[(size_t)i*N + k] = val;
				else ((uint8_t *)values)[(size_t)i*N + k] = (uint8_t)val;
			}
		}
		free(logs);
	}
	return n;
}

/* ------------------------------------------------------------------------------------------------ */
/* x86 conversions: cvttss2si yields INT_MIN for NaN / out of range; (int16_t)f is the low
 * half of that; float->uint32 goes through the 64-bit form. */
static int32_t f2i(float f) {
	if(!(f >= -2147483648.0f && f < 2147483648.0f)) return INT32_MIN;
	return (int32_t)f;
}
static int16_t f2s(float f) { return (int16_t)(uint16_t)(uint32_t)f2i(f); }
static uint32_t f2u_via64(float f) {
	if(!(f >= -9223372036854775808.0f && f < 9223372036854775808.0f)) return 0;
	return (uint32_t)(uint64_t)(int64_t)f;
}
static int32_t iabs(int32_t v) { return v < 0 ? (int32_t)(0u - (uint32_t)v) : v; }   /* abs(INT_MIN)=INT_MIN */

/* include/corto/point.h:111 */
static float norm3(const float v[3]) { float s = v[0]*v[0] + v[1]*v[1]; s = s + v[2]*v[2]; return (float)sqrt((double)s); }

/* include/corto/normal_attribute.h:75-85 */
static void to_octa(const float v[3], int unit, int32_t out[2]) {
	float s = fabsf(v[0]) + fabsf(v[1]);
	s = s + fabsf(v[2]);
	float px = v[0]/s, py = v[1]/s;
	if(v[2] < 0) {
		float ax = 1.0f - fabsf(py), ay = 1.0f - fabsf(px);
		px = ax; py = ay;
		if(v[0] < 0) px = -px;
		if(v[1] < 0) py = -py;
	}
	out[0] = f2i(px*(float)unit);
	out[1] = f2i(py*(float)unit);
}

/* normal_attribute.h:104-112 (int inputs) */
static void to_sphere_i(int32_t vx, int32_t vy, int unit, float n[3]) {
	int32_t z = unit - iabs(vx) - iabs(vy);
	n[0] = (float)vx; n[1] = (float)vy; n[2] = (float)z;
	if(n[2] < 0) {
		n[0] = (float)(((vx > 0) ? 1 : -1)*(unit - iabs(vy)));
		n[1] = (float)(((vy > 0) ? 1 : -1)*(unit - iabs(vx)));
	}
	float len = norm3(n);
	n[0] /= len; n[1] /= len; n[2] /= len;
}
/* normal_attribute.h:114-122 (inputs truncated to int16 first) */
static void to_sphere_s(int16_t vx, int16_t vy, int unit, int16_t o[3]) {
	float n[3];
	to_sphere_i(vx, vy, unit, n);       /* |v| <= 32768: no wrap, identical arithmetic */
	o[0] = f2s(n[0]*32767.0f); o[1] = f2s(n[1]*32767.0f); o[2] = f2s(n[2]*32767.0f);
}

/* ------------------------------------------------------------------------------------------------ */
/* CLERS face reconstruction — src/decoder.cpp:204-358.  Front edges live in growable arrays. */
typedef struct { uint32_t v0, v1, v2, prev, next; uint8_t deleted; } FEdge;
typedef struct { FEdge *e; size_t n, cap; } Front;
typedef struct { int *v; size_t n, cap; } IVec;

static uint32_t front_push(Front *f, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t prev, uint32_t next) {
	if(f->n == f->cap) { f->cap = f->cap ? f->cap*2 : 1024; f->e = (FEdge *)realloc(f->e, f->cap*sizeof(FEdge)); }
	FEdge *e = &f->e[f->n];
	e->v0 = v0; e->v1 = v1; e->v2 = v2; e->prev = prev; e->next = next; e->deleted = 0;
	return (uint32_t)f->n++;
}
static void ivec_push(IVec *v, int x) {
	if(v->n == v->cap) { v->cap = v->cap ? v->cap*2 : 1024; v->v = (int *)realloc(v->v, v->cap*sizeof(int)); }
	v->v[v->n++] = x;
}
static int ilog2_u(uint64_t p) { int k = 0; while(p >>= 1) k++; return k; }   /* src/cstream.cpp:31-35 */

typedef struct {
	const uint8_t *clers; uint32_t cler;           /* cursor carried across groups (decoder.cpp:173-178) */
	Bits split; uint64_t splitpos;
	uint32_t vertex_count;
	uint32_t nvert;
	uint32_t *faces32; uint16_t *faces16;
	uint32_t *prediction;                          /* nvert x (a,b,c) */
} Topo;

static void put_face(Topo *t, uint32_t *start, uint32_t a, uint32_t b, uint32_t c) {
	if(t->faces16) { t->faces16[*start] = (uint16_t)a; t->faces16[*start + 1] = (uint16_t)b; t->faces16[*start + 2] = (uint16_t)c; }
	else if(t->faces32) { t->faces32[*start] = a; t->faces32[*start + 1] = b; t->faces32[*start + 2] = c; }
	*start += 3;
}

static int decode_faces(Topo *t, uint32_t start, uint32_t end) {
	Front fr = {0, 0, 0};
	IVec order = {0, 0, 0}, delayed = {0, 0, 0};
	size_t cursor = 0;
	int splitbits = ilog2_u(t->nvert) + 1;
	int new_edge = -1;
	int rc = 0;
	while(start < end) {
		if(new_edge == -1 && cursor >= order.n && delayed.n == 0) {      /* nothing pending: start triangle (:224-259) */
			uint32_t last = t->vertex_count - 1;
			uint32_t vi[3];
			int mask = 0;
			int c = t->clers[t->cler++];
			if(c == C_SPLIT) { mask = (int)getbits(&t->split, t->splitpos, 3); t->splitpos += 3; }
			for(int k = 0; k < 3; k++) {
				uint32_t v;
				if(mask & (1 << k)) { v = getbits(&t->split, t->splitpos, splitbits); t->splitpos += (uint64_t)splitbits; }
				else {
					uint32_t *p = t->prediction + (size_t)t->vertex_count*3;
					p[0] = p[1] = p[2] = last;
					last = v = t->vertex_count++;
				}
				vi[k] = v;
			}
			uint32_t s = start;
			put_face(t, &s, vi[0], vi[1], vi[2]);
			start = s;
			uint32_t base = (uint32_t)fr.n;
			ivec_push(&order, (int)front_push(&fr, vi[1], vi[2], vi[0], base + 2, base + 1));
			ivec_push(&order, (int)front_push(&fr, vi[2], vi[0], vi[1], base + 0, base + 2));
			ivec_push(&order, (int)front_push(&fr, vi[0], vi[1], vi[2], base + 1, base + 0));
			continue;
		}
		int f;
		if(new_edge != -1) { f = new_edge; new_edge = -1; }
		else if(cursor < order.n) f = order.v[cursor++];
		else f = delayed.v[--delayed.n];

		const FEdge e = fr.e[f];
		if(e.deleted) continue;                                            /* before a cler is consumed (:278-279) */
		int c = t->clers[t->cler++];
		if(c == C_BOUNDARY) continue;

		const FEdge pe = fr.e[e.prev], ne = fr.e[e.next];                  /* copies taken before mutation (:288-289) */
		new_edge = (int)fr.n;
		uint32_t opposite;
		if(c == C_VERTEX || c == C_SPLIT) {
			if(c == C_SPLIT) { opposite = getbits(&t->split, t->splitpos, splitbits); t->splitpos += (uint64_t)splitbits; }
			else {
				uint32_t *p = t->prediction + (size_t)t->vertex_count*3;
				p[0] = e.v1; p[1] = e.v0; p[2] = e.v2;
				opposite = t->vertex_count++;
			}
			fr.e[e.prev].next = (uint32_t)new_edge;
			fr.e[e.next].prev = (uint32_t)new_edge + 1;
			front_push(&fr, e.v0, opposite, e.v1, e.prev, (uint32_t)new_edge + 1);
			ivec_push(&order, (int)fr.n);
			front_push(&fr, opposite, e.v1, e.v0, (uint32_t)new_edge, e.next);
		} else if(c == C_LEFT) {
			fr.e[e.prev].deleted = 1;
			fr.e[pe.prev].next = (uint32_t)new_edge;
			fr.e[e.next].prev = (uint32_t)new_edge;
			opposite = pe.v0;
			front_push(&fr, opposite, e.v1, e.v0, pe.prev, e.next);
		} else if(c == C_RIGHT) {
			fr.e[e.next].deleted = 1;
			fr.e[ne.next].prev = (uint32_t)new_edge;
			fr.e[e.prev].next = (uint32_t)new_edge;
			opposite = ne.v1;
			front_push(&fr, e.v0, opposite, e.v1, e.prev, ne.next);
		} else if(c == C_DELAY) {
			ivec_push(&delayed, f);
			new_edge = -1;
			continue;
		} else if(c == C_END) {
			fr.e[e.prev].deleted = 1;
			fr.e[e.next].deleted = 1;
			fr.e[pe.prev].next = ne.next;
			fr.e[ne.next].prev = pe.prev;
			opposite = pe.v0;
			new_edge = -1;
		} else { rc = -2; break; }
		uint32_t s = start;
		put_face(t, &s, e.v1, e.v0, opposite);
		start = s;
	}
	free(fr.e); free(order.v); free(delayed.v);
	return rc;
}

/* ------------------------------------------------------------------------------------------------ */
/* Header — src/decoder.cpp:41-89. */
int crt_oracle_info(const uint8_t *blob, int len, OInfo *info) {
	(void)len;
	if((uintptr_t)blob & 3) return -1;
	Cur c = { blob, blob };
	if(rd32(&c) != 0x787A6300u) return -2;
	rd32(&c);                                      /* version */
	info->entropy = (int)rd8(&c);
	uint32_t nexif = rd32(&c);
	for(uint32_t i = 0; i < nexif; i++) { rdstr(&c); rdstr(&c); }
	info->nattr = (int)rd32(&c);
	if(info->nattr > ORACLE_MAX_ATTR) return -3;
	for(int i = 0; i < info->nattr; i++) {
		OAttrInfo *a = &info->attr[i];
		const char *nm = rdstr(&c);
		strncpy(a->name, nm, 63); a->name[63] = 0;
		a->codec = (int)rd32(&c);
		uint32_t qb = rd32(&c); memcpy(&a->q, &qb, 4);
		a->N = (int)rd8(&c); a->format = (int)rd8(&c); a->strategy = (int)rd8(&c);
	}
	/* std::map<std::string, VertexAttribute *> (decoder.cpp:72-86): a repeated name keeps its LAST header entry and every pass
	 * visits the attributes in byte-wise name order (decoder.cpp:168) — which is the wire order of encoder-written files. */
	{
		int n = 0;
		for(int i = 0; i < info->nattr; i++) {
			int dup = -1;
			for(int j = 0; j < n; j++) if(strcmp(info->attr[j].name, info->attr[i].name) == 0) dup = j;
			if(dup >= 0) info->attr[dup] = info->attr[i]; else info->attr[n++] = info->attr[i];
		}
		info->nattr = n;
		for(int i = 1; i < n; i++) {                 /* insertion sort, stable */
			OAttrInfo t = info->attr[i];
			int j = i;
			while(j > 0 && strcmp(info->attr[j - 1].name, t.name) > 0) { info->attr[j] = info->attr[j - 1]; j--; }
			info->attr[j] = t;
		}
	}
	info->nvert = rd32(&c);
	info->nface = rd32(&c);
	info->body = (uint32_t)(c.p - c.base);
	return 0;
}

/* Full decode — src/decoder.cpp:126-196.  `binds[i]` matches info.attr[i].  index: u32 or u16 array or NULL.
 * Optional pins: clers_out/nclers_out, prediction_out (nvert x 3 u32).  Returns 0, or <0 on error. */
int crt_oracle_decode(const uint8_t *blob, int len, const OBind *binds, void *index, int index16,
                      uint8_t *clers_out, uint32_t *nclers_out, uint32_t *prediction_out) {
	OInfo info;
	int rc = crt_oracle_info(blob, len, &info);
	if(rc) return rc;
	const uint32_t nvert = info.nvert, nface = info.nface;
	Cur c = { blob, blob + info.body };

	/* groups — include/corto/index_attribute.h:89-99 */
	uint32_t ngroups = rd32(&c);
	uint32_t *gend = (uint32_t *)malloc(sizeof(uint32_t)*(ngroups + 1));
	for(uint32_t g = 0; g < ngroups; g++) {
		gend[g] = rd32(&c);
		uint32_t np = rd8(&c);
		for(uint32_t k = 0; k < np; k++) { rdstr(&c); rdstr(&c); }
	}

	uint8_t *clers = NULL; uint32_t nclers = 0;
	Bits split = {0, 0};
	if(nface > 0) {                                  /* index_attribute.h:83-87 */
		rd32(&c);                                    /* max_front (a reserve() hint only) */
		clers = entropy_decompress(&c, info.entropy, &nclers);
		split = rdbits(&c);
		if(clers_out) memcpy(clers_out, clers, nclers);
		if(nclers_out) *nclers_out = nclers;
	}

	/* per attribute: entropy decode + bit unpack (decoder.cpp:168-169) */
	int32_t *ndiffs[ORACLE_MAX_ATTR]; int nprediction[ORACLE_MAX_ATTR]; int qc[ORACLE_MAX_ATTR][4];
	for(int a = 0; a < info.nattr; a++) {
		OAttrInfo *ai = &info.attr[a];
		ndiffs[a] = NULL; nprediction[a] = 0;
		if(ai->codec == CODEC_NORMAL) {              /* src/normal_attribute.cpp:178-185 */
			nprediction[a] = (int)rd8(&c);
			ndiffs[a] = (int32_t *)calloc((size_t)nvert*2 + 2, 4);
			decode_array(&c, info.entropy, ndiffs[a], 2, 4);
		} else if(ai->codec == CODEC_COLOR) {        /* include/corto/color_attribute.h:55-59 */
			qc[a][0] = qc[a][1] = qc[a][2] = 4; qc[a][3] = 8;   /* ctor defaults :31-34 */
			for(int k = 0; k < ai->N; k++) qc[a][k] = (int)rd8(&c);
			decode_values(&c, info.entropy, binds[a].buffer, ai->N, 1);
		} else {                                     /* include/corto/vertex_attribute.h:153-158 (always GenericAttr<int>) */
			if(ai->strategy & S_CORRELATED) decode_array(&c, info.entropy, binds[a].buffer, ai->N, 4);
			else decode_values(&c, info.entropy, binds[a].buffer, ai->N, 4);
		}
	}

	uint32_t *prediction = NULL;
	if(nface > 0) {                                  /* decoder.cpp:171-178 */
		prediction = (uint32_t *)calloc((size_t)nvert*3 + 3, 4);
		Topo t;
		t.clers = clers; t.cler = 0; t.split = split; t.splitpos = 0; t.vertex_count = 0; t.nvert = nvert;
		t.faces32 = index16 ? NULL : (uint32_t *)index; t.faces16 = index16 ? (uint16_t *)index : NULL;
		t.prediction = prediction;
		uint32_t start = 0;
		for(uint32_t g = 0; g < ngroups && rc == 0; g++) { rc = decode_faces(&t, start*3, gend[g]*3); start = gend[g]; }
		if(prediction_out) memcpy(prediction_out, prediction, (size_t)nvert*12);
	}

	/* deltaDecode (decoder.cpp:188-189 / :141-142) */
	for(int a = 0; a < info.nattr && rc == 0; a++) {
		OAttrInfo *ai = &info.attr[a];
		if(ai->codec == CODEC_NORMAL) {              /* normal_attribute.cpp:187-208 */
			if(!binds[a].buffer || nprediction[a] != N_DIFF) continue;
			int32_t *d = ndiffs[a];
			if(nface > 0) { for(uint32_t i = 1; i < nvert; i++) for(int k = 0; k < 2; k++) d[i*2 + k] += d[prediction[i*3]*2 + k]; }
			else for(uint32_t i = 2; i < nvert*2; i++) d[i] += d[i - 2];
			continue;
		}
		if(!binds[a].buffer) continue;               /* vertex_attribute.h:160-182 */
		const int N = ai->N;
		if(ai->codec == CODEC_COLOR) {
			uint8_t *v = (uint8_t *)binds[a].buffer;
			if(nface > 0 && (ai->strategy & S_PARALLEL)) {
				for(uint32_t i = 1; i < nvert; i++) { const uint32_t *p = prediction + (size_t)i*3;
					for(int k = 0; k < N; k++) v[i*N + k] = (uint8_t)(v[i*N + k] + v[p[0]*N + k] + v[p[1]*N + k] - v[p[2]*N + k]); }
			} else if(nface > 0) {
				for(uint32_t i = 1; i < nvert; i++) for(int k = 0; k < N; k++) v[i*N + k] = (uint8_t)(v[i*N + k] + v[prediction[i*3]*N + k]);
			} else for(uint32_t i = (uint32_t)N; i < nvert*N; i++) v[i] = (uint8_t)(v[i] + v[i - N]);
		} else {
			int32_t *v = (int32_t *)binds[a].buffer;
			if(nface > 0 && (ai->strategy & S_PARALLEL)) {
				for(uint32_t i = 1; i < nvert; i++) { const uint32_t *p = prediction + (size_t)i*3;
					for(int k = 0; k < N; k++) v[(size_t)i*N + k] += v[(size_t)p[0]*N + k] + v[(size_t)p[1]*N + k] - v[(size_t)p[2]*N + k]; }
			} else if(nface > 0) {
				for(uint32_t i = 1; i < nvert; i++) for(int k = 0; k < N; k++) v[(size_t)i*N + k] += v[(size_t)prediction[i*3]*N + k];
			} else for(size_t i = (size_t)N; i < (size_t)nvert*N; i++) v[i] += v[i - N];
		}
	}

	/* postDelta — normals ESTIMATED / BORDER need integer positions (normal_attribute.cpp:210-255) */
	if(nface > 0) for(int a = 0; a < info.nattr && rc == 0; a++) {
		OAttrInfo *ai = &info.attr[a];
		if(ai->codec != CODEC_NORMAL || !binds[a].buffer || nprediction[a] == N_DIFF) continue;
		int pa = -1;
		for(int k = 0; k < info.nattr; k++) if(!strcmp(info.attr[k].name, "position")) pa = k;
		if(pa < 0 || !binds[pa].buffer || !index) { rc = -4; break; }
		const int32_t *P = (const int32_t *)binds[pa].buffer;
		float *est = (float *)calloc((size_t)nvert*3 + 3, 4);
		int32_t *bnd = (int32_t *)calloc((size_t)nvert + 1, 4);
		for(uint32_t f = 0; f < nface; f++) {        /* estimateNormals :40-59, markBoundary :24-37 */
			uint32_t i0, i1, i2;
			if(index16) { const uint16_t *x = (const uint16_t *)index + (size_t)f*3; i0 = x[0]; i1 = x[1]; i2 = x[2]; }
			else { const uint32_t *x = (const uint32_t *)index + (size_t)f*3; i0 = x[0]; i1 = x[1]; i2 = x[2]; }
			float v0[3], a1[3], a2[3], n[3];
			for(int k = 0; k < 3; k++) { v0[k] = (float)P[(size_t)i0*3 + k]; a1[k] = (float)P[(size_t)i1*3 + k] - v0[k]; a2[k] = (float)P[(size_t)i2*3 + k] - v0[k]; }
			n[0] = a1[1]*a2[2] - a1[2]*a2[1];
			n[1] = a1[2]*a2[0] - a1[0]*a2[2];
			n[2] = a1[0]*a2[1] - a1[1]*a2[0];
			for(int k = 0; k < 3; k++) { est[(size_t)i0*3 + k] += n[k]; est[(size_t)i1*3 + k] += n[k]; est[(size_t)i2*3 + k] += n[k]; }
			if(nprediction[a] == N_BORDER) {
				bnd[i0] ^= (int32_t)i1; bnd[i0] ^= (int32_t)i2;
				bnd[i1] ^= (int32_t)i2; bnd[i1] ^= (int32_t)i0;
				bnd[i2] ^= (int32_t)i0; bnd[i2] ^= (int32_t)i1;
			}
		}
		const int unit = f2i(ai->q);
		const int32_t *d = ndiffs[a];
		uint32_t count = 0;
		for(uint32_t i = 0; i < nvert; i++) {        /* computeNormals :281-325 */
			float *e = est + (size_t)i*3;
			if(nprediction[a] == N_ESTIMATED || bnd[i]) {
				int32_t qn[2]; to_octa(e, unit, qn);
				int32_t dx = d[count*2], dy = d[count*2 + 1]; count++;
				if(binds[a].format == F_FLOAT) to_sphere_i(qn[0] + dx, qn[1] + dy, unit, (float *)binds[a].buffer + (size_t)i*3);
				else to_sphere_s((int16_t)(qn[0] + dx), (int16_t)(qn[1] + dy), unit, (int16_t *)binds[a].buffer + (size_t)i*3);
			} else if(binds[a].format == F_FLOAT) {
				float *o = (float *)binds[a].buffer + (size_t)i*3, len = norm3(e);
				o[0] = e[0]/len; o[1] = e[1]/len; o[2] = e[2]/len;
			} else {
				int16_t *o = (int16_t *)binds[a].buffer + (size_t)i*3;
				float len = norm3(e);
				if(!(len < 0.00001f)) {                /* tiny normal: output left untouched */
					len = 32767.0f/len;
					for(int k = 0; k < 3; k++) o[k] = f2s(e[k]*len);
				}
			}
		}
		free(est); free(bnd);
	}

	/* dequantize (decoder.cpp:194-195 / :145-146) */
	for(int a = 0; a < info.nattr && rc == 0; a++) {
		OAttrInfo *ai = &info.attr[a];
		if(!binds[a].buffer) continue;
		if(ai->codec == CODEC_NORMAL) {              /* normal_attribute.cpp:257-279 */
			if(nprediction[a] != N_DIFF) continue;
			const int unit = f2i(ai->q);
			const int32_t *d = ndiffs[a];
			for(uint32_t i = 0; i < nvert; i++) {
				if(binds[a].format == F_FLOAT) to_sphere_i(d[i*2], d[i*2 + 1], unit, (float *)binds[a].buffer + (size_t)i*3);
				else to_sphere_s((int16_t)d[i*2], (int16_t)d[i*2 + 1], unit, (int16_t *)binds[a].buffer + (size_t)i*3);
			}
		} else if(ai->codec == CODEC_COLOR) {        /* src/color_attribute.cpp:76-95 + point.h:214; in place, back to front */
			const int N = ai->N, oc = binds[a].out_components;
			uint8_t *buf = (uint8_t *)binds[a].buffer;
			for(uint32_t i = nvert; i-- > 0;) {
				uint8_t y[4] = {0, 0, 0, 255};
				for(int k = 0; k < N; k++) y[k] = buf[(size_t)i*N + k];
				uint8_t rgb[4] = { (uint8_t)(y[2] + y[0]), y[0], (uint8_t)(y[1] + y[0]), y[3] };
				for(int k = 0; k < oc; k++) buf[(size_t)i*oc + k] = (uint8_t)(rgb[k]*qc[a][k]);
			}
		} else {                                     /* vertex_attribute.h:184-230 */
			size_t n = (size_t)nvert*ai->N;
			if(binds[a].format == F_FLOAT) { int32_t *v = (int32_t *)binds[a].buffer; float *o = (float *)binds[a].buffer;
				for(size_t i = 0; i < n; i++) o[i] = (float)v[i]*ai->q; }
			else if(binds[a].format == F_INT32 || binds[a].format == F_UINT32) { uint32_t *v = (uint32_t *)binds[a].buffer;
				for(size_t i = 0; i < n; i++) v[i] = f2u_via64((float)v[i]*ai->q); }
			else rc = -5;
		}
	}

	for(int a = 0; a < info.nattr; a++) free(ndiffs[a]);
	free(prediction); free(clers); free(gend);
	return rc;
}

/* FNV-1a 64 over raw bytes: the digest SURVEY §8c quotes for html/models/tarta.crt outputs. */
uint64_t crt_oracle_fnv1a64(const uint8_t *p, uint64_t n) {
	uint64_t h = 0xcbf29ce484222325ull;
	for(uint64_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
	return h;
}

/* Tunstall table pin for kernel-level tests: build from (symbol,prob) pairs; returns used table bytes. */
int crt_oracle_tunstall_tables(const uint8_t *probs, int nsym, int *index256, int *lengths256, uint8_t *table8192) {
	Tun *t = (Tun *)calloc(1, sizeof(Tun));
	t->nsym = nsym;
	for(int i = 0; i < nsym; i++) { t->sym[i] = probs[2*i]; t->prob[i] = probs[2*i + 1]; }
	tun_build(t);
	int used = 0;
	for(int i = 0; i < 256; i++) { index256[i] = t->index[i]; lengths256[i] = t->length[i]; if(nsym > 1 && t->index[i] + t->length[i] > used) used = t->index[i] + t->length[i]; }
	memcpy(table8192, t->table, 8192);
	free(t);
	return used;
}

/* CPU baseline loop (kind "port"): best wall seconds is measured by the caller; this decodes n blobs once,
 * everything bound (float normals, u32 index), scratch outputs allocated once for the largest mesh. */
int crt_oracle_decode_all(int n, const uint8_t *const *blobs, const int *lens) {
	uint32_t maxv = 0, maxf = 0;
	OInfo info;
	for(int i = 0; i < n; i++) { if(crt_oracle_info(blobs[i], lens[i], &info)) return -1; if(info.nvert > maxv) maxv = info.nvert; if(info.nface > maxf) maxf = info.nface; }
	void *bufs[ORACLE_MAX_ATTR];
	for(int a = 0; a < ORACLE_MAX_ATTR; a++) bufs[a] = malloc((size_t)maxv*16 + 16);
	uint32_t *idx = (uint32_t *)malloc((size_t)maxf*12 + 12);
	int rc = 0;
	for(int i = 0; i < n && rc == 0; i++) {
		crt_oracle_info(blobs[i], lens[i], &info);
		OBind b[ORACLE_MAX_ATTR];
		for(int a = 0; a < info.nattr; a++) {
			b[a].buffer = bufs[a];
			b[a].format = info.attr[a].codec == CODEC_COLOR ? F_UINT8 : F_FLOAT;
			b[a].out_components = info.attr[a].N;
		}
		rc = crt_oracle_decode(blobs[i], lens[i], b, idx, 0, 0, 0, 0);
	}
	for(int a = 0; a < ORACLE_MAX_ATTR; a++) free(bufs[a]);
	free(idx);
	return rc;
}
