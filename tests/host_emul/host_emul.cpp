// tests/host_emul/host_emul.cpp — CPU exercise of the __host__ __device__ cores the kernels run
// (corto_b200/csrc/crt_device.cuh) plus the host directory walk, so their logic is checked without a GPU.
// Built and driven by tests/test_host_emul.py; compares against the oracle from Python.  Not part of the product.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include "../../corto_b200/csrc/crt_device.cuh"
#include "../../corto_b200/csrc/crt_walk.h"
#include "../../include/corto_b200.h"

using namespace crtb;

// ---- host transcriptions of the warp-wide window steps of k_clers_lf (lead_vector / follow_vector, crt_kernels.cu): same
// formulas, lanes as a loop.  They let the window logic be checked on the CPU together with the scalar machines.
static uint32_t lead_vector_host(const ClersIO &io, ArrayRings &rg, LeadState &S, uint32_t left) {
	const uint32_t cler = S.cler, nfront0 = S.nfront, next0 = S.cnext, nlog0 = S.nlog, eflush = S.eflush;
	uint32_t lim = std::min(std::min(32u, left), std::min(io.nclers - cler, S.end - S.start));
	uint32_t m = 0;
	while(m < lim && (io.clers[cler + m] == C_VERTEX || io.clers[cler + m] == C_LEFT)) m++;
	if(m < 2) return 0;
	uint32_t nV = 0, nL = 0;
	for(uint32_t j = 0; j < m; j++) { if(io.clers[cler + j] == C_VERTEX) nV++; else nL++; }
	uint32_t chain[33];
	uint32_t p = S.cprev; bool ok = nfront0 + nV <= io.cap;
	for(uint32_t k = 0; k < nL; k++) {
		chain[k] = p;
		if(p == next0) ok = false;
		uint32_t pp, pn;
		if(p >= eflush) rg.ldB(p, pp, pn); else { pp = io.eb[p].prev; pn = io.eb[p].next; }
		(void)pn;
		p = pp;
	}
	chain[nL] = p;
	if(!ok) return 0;
	uint32_t rv = 0, rl = 0;
	for(uint32_t j = 0; j < m; j++) {
		if(io.clers[cler + j] == C_VERTEX) {
			const uint32_t r = rv++, b = nfront0 + r;
			rg.stB(b, r + 1 < nV ? b + 1 : CLERS_NOLINK, r ? b - 1 : next0);
			rg.stFl(b, 0);
			rg.stLog(nlog0 + j, ((uint32_t)LG_V << 28) | b);
		} else {
			const uint32_t q = chain[rl++];
			if(q >= eflush) rg.stFl(q, CLERS_DEL); else io.fl[q] = CLERS_DEL;
			rg.stLog(nlog0 + j, ((uint32_t)LG_L << 28) | q);
		}
	}
	if(nV) { if(next0 >= eflush) rg.stB_prev(next0, nfront0); else io.eb[next0].prev = nfront0; }
	S.nfront = nfront0 + nV; S.cprev = chain[nL]; if(nV) S.cnext = nfront0 + nV - 1;
	S.lp = S.ln = 1; S.cf = CLERS_NOID; S.nlog = nlog0 + m; S.start += m;
	S.cler = cler + m; S.cwv = 0;
	if(S.start >= S.end) S.have = 0;
	return m;
}

static void lead_pop_host(const ClersIO &io, ArrayRings &rg, LeadState &S) {
	uint32_t found = CLERS_NOID;
	while(S.scan < S.nfront) {
		const uint32_t id = S.scan++;
		const uint32_t fl = id >= S.eflush ? rg.ldFl(id) : io.fl[id];
		if(fl == 0) { found = id; break; }
	}
	if(found != CLERS_NOID) {
		uint32_t p, q;
		if(found >= S.eflush) rg.ldB(found, p, q); else { p = io.eb[found].prev; q = io.eb[found].next; }
		S.cprev = p; S.cnext = q; S.lp = S.ln = 0; S.have = 1; S.cf = found;
		rg.stLog(S.nlog, ((uint32_t)LG_P << 28) | found); S.nlog++;
	}
}

static uint32_t follow_vector_host(const ClersIO &io, ArrayRings &rg, FollowState &F, uint32_t upto) {
	const uint32_t tail = F.tail, lim = std::min(32u, upto - tail);
	uint32_t ty[32], id[32], m = 0;
	for(uint32_t j = 0; j < lim; j++) { const uint32_t w = rg.ldLog(tail + j); ty[j] = w >> 28; id[j] = w & 0x0FFFFFFFu; }
	{
		bool seenV = false; uint32_t firstid = 0;
		while(m < lim && (ty[m] == LG_V || ty[m] == LG_L)) {
			if(ty[m] == LG_V && !seenV) { seenV = true; firstid = id[m]; }
			else if(ty[m] == LG_L && seenV && id[m] >= firstid) break;     // consumes an edge created inside this window
			m++;
		}
	}
	if(m < 2) return 0;
	uint32_t nV = 0;
	for(uint32_t j = 0; j < m; j++) nV += ty[j] == LG_V;
	if(F.vcount + nV > io.nvert || F.nfaces + m > io.nface) return 0;
	uint32_t a[32], x[32];
	uint32_t rv = 0;
	for(uint32_t j = 0; j < m; j++) {
		a[j] = 0; x[j] = F.vcount + rv;
		if(ty[j] == LG_V) rv++;
		else { uint32_t t1, t2; if(id[j] >= F.aflush) rg.ldA(id[j], a[j], t1, t2); else a[j] = io.ea[id[j]].v0; }
	}
	uint32_t v0b[32], v1b[32], v0a[32], v1a[32], v2b[32];
	uint32_t v0 = F.v0, v1 = F.v1;
	for(uint32_t j = 0; j < m; j++) {
		v0b[j] = v0; v1b[j] = v1;
		if(ty[j] == LG_V) v1 = x[j]; else v0 = a[j];
		v0a[j] = v0; v1a[j] = v1;
	}
	for(uint32_t j = 0; j < m; j++) v2b[j] = j == 0 ? F.v2 : (ty[j - 1] == LG_V ? v1b[j - 1] : v0b[j - 1]);
	for(uint32_t j = 0; j < m; j++) {
		rg.stF(F.nfaces + j, v1b[j], v0b[j], ty[j] == LG_V ? x[j] : a[j]);
		if(ty[j] == LG_V) { rg.stP(x[j], v1b[j], v0b[j], v2b[j]); rg.stA(id[j], x[j], v1b[j], v0b[j]); F.amax = id[j] + 1; }
	}
	F.v0 = v0a[m - 1]; F.v1 = v1a[m - 1]; F.v2 = ty[m - 1] == LG_V ? v1b[m - 1] : v0b[m - 1];
	F.vcount += nV; F.nfaces += m; F.tail += m;
	return m;
}

// ---- host transcriptions of the CTA-wide steps of k_clers_cta (cta_window / cta_pop, crt_clers_cta.cu): the same closed
// form — ranks among the V's and the L's of the run — with threads as a loop; W = symbols per window.
static uint32_t cta_window_host(const ClersIO &io, ArrayRings &rg, MergedState &S, uint32_t W, uint32_t R) {
	uint32_t done = 0;
	for(;;) {
		const uint32_t cler = S.cler, start = S.start, end = S.end;
		const uint32_t lim = std::min(W, std::min(io.nclers - cler, end - start));
		uint32_t m = 0;
		while(m < lim && rg.sym(cler + m) <= (uint32_t)C_LEFT) m++;
		std::vector<uint32_t> nVb(m + 1, 0), nLb(m + 1, 0);
		for(uint32_t j = 0; j < m; j++) { nVb[j + 1] = nVb[j] + (rg.sym(cler + j) == C_VERTEX); nLb[j + 1] = nLb[j] + (rg.sym(cler + j) == C_LEFT); }
		const uint32_t nV = nVb[m], nL = nLb[m];
		const uint32_t prev = S.prev, next = S.next, nfront = S.nfront, vcount = S.vcount, eflush = S.eflush;
		if(m < 2 || nfront + nV > io.cap || vcount + nV > io.nvert || nfront + nV > eflush + R) return done;
		// chain: fast path = consecutive ids, else a walk
		std::vector<uint32_t> chain(nL + 1);
		bool fast = true;
		for(uint32_t k = 0; k < nL && fast; k++) {
			const uint32_t id = prev + k;
			if(!(id >= eflush && id < nfront && id != next)) { fast = false; break; }
			uint32_t pk, pn; rg.ldB(id, pk, pn);
			chain[k] = id;
			if(k + 1 == nL) chain[nL] = pk; else if(pk != id + 1) fast = false;
		}
		if(nL == 0) chain[0] = prev;
		if(getenv("EMUL_STATS")) { static long nf = 0, ns = 0; (fast ? nf : ns)++; if(((nf + ns) & 255) == 0) fprintf(stderr, "windows fast %ld slow %ld (m=%u nL=%u prev=%u next=%u nfront=%u eflush=%u)\n", nf, ns, m, nL, prev, next, nfront, eflush); }
		if(!fast) {
			uint32_t q = prev; bool ok = true;
			for(uint32_t k = 0; k < nL; k++) {
				chain[k] = q;
				if(q == next || q >= nfront) { ok = false; break; }
				uint32_t pp, pq;
				if(q >= eflush) rg.ldB(q, pp, pq); else { pp = io.eb[q].prev; pq = io.eb[q].next; }
				(void)pq; q = pp;
			}
			chain[nL] = q;
			if(!ok) return done;
		}
		std::vector<uint32_t> aL(nL + 1, 0);
		for(uint32_t k = 0; k < nL; k++) { const uint32_t id = chain[k]; uint32_t t1, t2; if(id >= eflush) rg.ldA(id, aL[k], t1, t2); else aL[k] = io.ea[id].v0; }
		uint32_t e_v0 = S.v0, e_v1 = S.v1, e_v2 = S.v2;
		for(uint32_t i = 0; i < m; i++) {
			const bool isV = rg.sym(cler + i) == C_VERTEX;
			const uint32_t rv = nVb[i], rl = nLb[i];
			const uint32_t v0i = rl ? aL[rl - 1] : S.v0, v1i = rv ? vcount + rv - 1 : S.v1;
			uint32_t v2i;
			if(i == 0) v2i = S.v2;
			else if(rg.sym(cler + i - 1) == C_VERTEX) v2i = rv > 1 ? vcount + rv - 2 : S.v1;
			else v2i = rl > 1 ? aL[rl - 2] : S.v0;
			const uint32_t x = vcount + rv, a = isV ? 0 : aL[rl];
			clers_put_face(io, (size_t)(start + i)*3u, v1i, v0i, isV ? x : a);
			if(isV) {
				clers_put_pred(io, x, v1i, v0i, v2i);
				const uint32_t b = nfront + rv;
				rg.stA(b, x, v1i, v0i); rg.stB(b, rv + 1 < nV ? b + 1 : CLERS_NOLINK, rv ? b - 1 : next); rg.stFl(b, 0);
			} else {
				const uint32_t id = chain[rl];
				if(id >= eflush) rg.stFl(id, CLERS_DEL); else io.fl[id] = CLERS_DEL;
			}
			if(i == m - 1) { e_v0 = isV ? v0i : a; e_v1 = isV ? x : v1i; e_v2 = isV ? v1i : v0i; }
		}
		if(nV) { if(next >= eflush) rg.stB_prev(next, nfront); else io.eb[next].prev = nfront; S.next = nfront + nV - 1; }
		S.v0 = e_v0; S.v1 = e_v1; S.v2 = e_v2;
		S.nfront = nfront + nV; S.vcount = vcount + nV; S.prev = chain[nL];
		S.start = start + m; S.cler = cler + m; S.lp = S.ln = 1; S.cf = CLERS_NOID; S.have = S.start < end ? 1u : 0u;
		done += m;
		if(m < W || S.start >= end || S.cler >= io.nclers) return done;
		if(S.nfront + W > eflush + R) return done;
	}
}

static void cta_pop_host(const ClersIO &io, ArrayRings &rg, MergedState &S) {
	uint32_t found = CLERS_NOID;
	while(S.scan < S.nfront) {
		const uint32_t id = S.scan++;
		const uint32_t fl = id >= S.eflush ? rg.ldFl(id) : io.fl[id];
		if(fl == 0) { found = id; break; }
	}
	if(found != CLERS_NOID) {
		uint32_t p, q, a, b, c;
		if(found >= S.eflush) { rg.ldB(found, p, q); rg.ldA(found, a, b, c); }
		else { p = io.eb[found].prev; q = io.eb[found].next; a = io.ea[found].v0; b = io.ea[found].v1; c = io.ea[found].v2; }
		S.prev = p; S.next = q; S.v0 = a; S.v1 = b; S.v2 = c; S.lp = S.ln = 0; S.have = 1; S.cf = found;
	}
}

static std::vector<uint32_t> g_log;
extern "C" {
int emul_log(uint32_t *out, int cap) { int n = (int)g_log.size(); for(int i = 0; i < n && i < cap; i++) out[i] = g_log[i]; return n; }

// Tunstall dictionary through tun_build_seq: index[256], lengths[256], text[8192]; returns used bytes.
int emul_tunstall_tables(const uint8_t *probs, int nsym, int *index256, int *lengths256, uint8_t *text8192) {
	static TunScratch S;
	static uint32_t entry[256];
	memset(text8192, 0, 8192);
	uint32_t used = tun_build_seq(probs, (uint32_t)nsym, S, text8192, entry);
	for(int i = 0; i < 256; i++) { index256[i] = (int)(entry[i] & 0xffff); lengths256[i] = (int)(entry[i] >> 16); }
	return (int)used;
}

// Walk a blob and return every entropy block's (probs_off, nsym, size, csize, data_off) — 5 u32 per block.
int emul_walk_blocks(const uint8_t *blob, int len, uint32_t *out, int cap) {
	ParsedMesh pm; std::string err;
	if(parse_header(blob, len, pm, err) || walk_directory(pm, err)) return -1;
	int n = 0;
	auto put = [&](const Block &b) { if(n < cap) { uint32_t *o = out + n*5; o[0] = b.probs_off; o[1] = b.nsym; o[2] = b.size; o[3] = b.csize; o[4] = b.data_off; } n++; };
	if(pm.nface) put(pm.clers);
	for(auto &s: pm.streams) for(auto &b: s.blocks) put(b);
	return n;
}

// CLERS automaton through clers_decode_seq given the decoded cler bytes (from the oracle); outputs faces (u32)
// and prediction (3 u32 per vertex).  Returns the automaton's return code.
// ---- host transcriptions of the CURRENT CTA-wide steps (k_clers_cta v7.3, crt_clers_cta.cu): windows of any width take runs of
// >= 1 symbols, verify consecutive prev chains in the ring OR in the reach-back store, RETIRE the gate when the symbol after the
// run is BOUNDARY / DELAY and hand over to the pop; the pop CONSUMES a BOUNDARY / DELAY waiting on the popped edge and scans on.
// Returns 1 when the gate was retired (pop next), else 0; `progress` = symbols consumed.
static int cta_window2_host(const ClersIO &io, ArrayRings &rg, MergedState &S, uint32_t W, uint32_t R, bool &bail) {
	bail = false;
	uint32_t done = 0;
	for(;;) {
		const uint32_t cler = S.cler, start = S.start, end = S.end;
		const uint32_t lim = std::min(W, std::min(io.nclers - cler, end - start));
		uint32_t m = 0;
		while(m < lim && rg.sym(cler + m) <= (uint32_t)C_LEFT) m++;
		std::vector<uint32_t> nVb(m + 1, 0), nLb(m + 1, 0);
		for(uint32_t j = 0; j < m; j++) { nVb[j + 1] = nVb[j] + (rg.sym(cler + j) == C_VERTEX); nLb[j + 1] = nLb[j] + (rg.sym(cler + j) == C_LEFT); }
		const uint32_t nV = nVb[m], nL = nLb[m];
		const uint32_t prev = S.prev, next = S.next, nfront = S.nfront, vcount = S.vcount, eflush = S.eflush, ndel = S.ndel;
		if(m < 1 || nfront + nV + 1 > io.cap || vcount + nV > io.nvert) { bail = done == 0; return 0; }
		if(nfront + nV + 1 > eflush + R) return 0;                       // the caller writes ring entries back first
		auto loadB = [&](uint32_t id, uint32_t &p, uint32_t &n) { if(id >= eflush) rg.ldB(id, p, n); else { p = io.eb[id].prev; n = io.eb[id].next; } };
		auto loadA0 = [&](uint32_t id) { uint32_t a, t1, t2; if(id >= eflush) { rg.ldA(id, a, t1, t2); return a; } return io.ea[id].v0; };
		std::vector<uint32_t> chain(nL + 1);
		bool fast = true;
		for(uint32_t k = 0; k < nL && fast; k++) {                        // consecutive ids, in the ring or in the reach-back store
			const uint32_t id = prev + k;
			if(!(id < nfront && id != next)) { fast = false; break; }
			uint32_t pk, pn; loadB(id, pk, pn);
			chain[k] = id;
			if(k + 1 == nL) chain[nL] = pk; else if(pk != id + 1) fast = false;
		}
		if(nL == 0) chain[0] = prev;
		if(!fast) {
			uint32_t q = prev; bool ok = true;
			for(uint32_t k = 0; k < nL; k++) {
				chain[k] = q;
				if(q == next || q >= nfront) { ok = false; break; }
				uint32_t pp, pq; loadB(q, pp, pq); q = pp;
			}
			chain[nL] = q;
			if(!ok) { bail = done == 0; return 0; }
		}
		std::vector<uint32_t> aL(nL + 1, 0);
		for(uint32_t k = 0; k < nL; k++) aL[k] = loadA0(chain[k]);
		// the symbol after the run
		const uint32_t newprev = chain[nL], nextf = nV ? nfront + nV - 1 : next, gid = nfront + nV;
		uint32_t cm = 0xff;
		if(cler + m < io.nclers && start + m < end) cm = rg.sym(cler + m);
		const bool gend = (cm == C_BOUNDARY || (cm == C_DELAY && ndel < io.cap)) && newprev != nextf && newprev < nfront;
		uint32_t e_v0 = S.v0, e_v1 = S.v1, e_v2 = S.v2;
		for(uint32_t i = 0; i < m; i++) {
			const bool isV = rg.sym(cler + i) == C_VERTEX;
			const uint32_t rv = nVb[i], rl = nLb[i];
			const uint32_t v0i = rl ? aL[rl - 1] : S.v0, v1i = rv ? vcount + rv - 1 : S.v1;
			uint32_t v2i;
			if(i == 0) v2i = S.v2;
			else if(rg.sym(cler + i - 1) == C_VERTEX) v2i = rv > 1 ? vcount + rv - 2 : S.v1;
			else v2i = rl > 1 ? aL[rl - 2] : S.v0;
			const uint32_t x = vcount + rv, a = isV ? 0 : aL[rl];
			clers_put_face(io, (size_t)(start + i)*3u, v1i, v0i, isV ? x : a);
			if(isV) {
				clers_put_pred(io, x, v1i, v0i, v2i);
				const uint32_t b = nfront + rv;
				rg.stA(b, x, v1i, v0i); rg.stB(b, rv + 1 < nV ? b + 1 : (gend ? gid : CLERS_NOLINK), rv ? b - 1 : next); rg.stFl(b, 0);
			} else {
				const uint32_t id = chain[rl];
				if(id >= eflush) rg.stFl(id, CLERS_DEL); else io.fl[id] = CLERS_DEL;
			}
			if(i == m - 1) { e_v0 = isV ? v0i : a; e_v1 = isV ? x : v1i; e_v2 = isV ? v1i : v0i; }
		}
		if(gend) {                                                        // the gate gets a record and leaves the machine
			rg.stA(gid, e_v0, e_v1, e_v2); rg.stB(gid, newprev, nextf); rg.stFl(gid, CLERS_NQ);
			if(newprev >= eflush) rg.stB_next(newprev, gid); else io.eb[newprev].next = gid;
			if(nV == 0) { if(next >= eflush) rg.stB_prev(next, gid); else io.eb[next].prev = gid; }
			if(cm == C_DELAY) io.delayed[ndel] = gid;
		}
		if(nV) { if(next >= eflush) rg.stB_prev(next, nfront); else io.eb[next].prev = nfront; S.next = nextf; }
		S.v0 = e_v0; S.v1 = e_v1; S.v2 = e_v2;
		S.nfront = nfront + nV + (gend ? 1u : 0u); S.vcount = vcount + nV; S.prev = newprev;
		S.start = start + m; S.cler = cler + m + (gend ? 1u : 0u); S.lp = S.ln = 1; S.cf = CLERS_NOID;
		S.have = (!gend && S.start < end) ? 1u : 0u;
		if(gend && cm == C_DELAY) S.ndel = ndel + 1;
		done += m;
		if(gend) return 1;
		if(m < W || S.start >= end || S.cler >= io.nclers) return 0;
		if(S.nfront + W + 1 > eflush + R) return 0;
	}
}

static int cta_pop2_host(const ClersIO &io, ArrayRings &rg, MergedState &S) {
	uint32_t found = CLERS_NOID, c = 0xff;
	while(S.scan < S.nfront) {
		const uint32_t id = S.scan++;
		const uint32_t fl = id >= S.eflush ? rg.ldFl(id) : io.fl[id];
		if(fl != 0) continue;
		found = id;
		c = S.cler < io.nclers ? rg.sym(S.cler) : 0xffu;
		if(c == C_BOUNDARY || (c == C_DELAY && S.ndel < io.cap)) {        // consumed right here: the edge keeps its record
			if(c == C_DELAY) io.delayed[S.ndel++] = found;
			S.cler++;
			found = CLERS_NOID;
			continue;
		}
		break;
	}
	if(found != CLERS_NOID) {
		uint32_t p, q, a, b, c2;
		if(found >= S.eflush) { rg.ldB(found, p, q); rg.ldA(found, a, b, c2); }
		else { p = io.eb[found].prev; q = io.eb[found].next; a = io.ea[found].v0; b = io.ea[found].v1; c2 = io.ea[found].v2; }
		S.prev = p; S.next = q; S.v0 = a; S.v1 = b; S.v2 = c2; S.lp = S.ln = 0; S.have = 1; S.cf = found;
	}
	return found != CLERS_NOID && c <= (uint32_t)C_LEFT;
}

// ring_r == 0: clers_decode_seq; else clers_decode_ring with an R = ring_r edge ring and Q = ring_q FIFO ring (tiny rings
// force the reach-back paths that are rare with the kernel's sizes).
int emul_clers(const uint8_t *blob, int len, const uint8_t *clers_in, uint32_t nclers, uint32_t *faces, uint32_t *prediction, int ring_r, int ring_q) {
	std::vector<uint8_t> padded((size_t)nclers + 64, 0);
	memcpy(padded.data() + ((8 - ((uintptr_t)padded.data() & 7)) & 7), clers_in, nclers);
	const uint8_t *clers = padded.data() + ((8 - ((uintptr_t)padded.data() & 7)) & 7);
	ParsedMesh pm; std::string err;
	if(parse_header(blob, len, pm, err) || walk_directory(pm, err)) return -100;
	uint32_t maxg = 0, prev = 0;
	for(uint32_t e: pm.group_ends) { if(e > prev && e - prev > maxg) maxg = e - prev; prev = e; }
	uint32_t cap = 3*maxg + 16;
	std::vector<EdgeA> ea(cap); std::vector<EdgeB> eb(cap); std::vector<uint32_t> order(cap), delayed(cap), pred((size_t)pm.nvert*4 + 4);
	ClersIO io;
	io.fl = nullptr;
	io.clers = clers; io.nclers = nclers;
	io.split = (const uint32_t *)(blob + pm.split_off); io.split_nwords = pm.split_nwords;
	io.group_ends = pm.group_ends.data(); io.ngroups = (uint32_t)pm.group_ends.size();
	io.nvert = pm.nvert; io.nface = pm.nface;
	io.ea = ea.data(); io.eb = eb.data(); io.order = order.data(); io.delayed = delayed.data(); io.cap = cap;
	io.faces32 = faces; io.faces16 = nullptr; io.pred = pred.data();
	uint32_t vc = 0;
	int rc;
	if(ring_r == 0) rc = clers_decode_seq(io, &vc);
	else if(ring_r < 0) {
		// v3 machine (clers_run) with the kernel's drain protocol emulated serially; R = -ring_r, Q = ring_q, budget = R/8
		const uint32_t R = (uint32_t)(-ring_r), Q = (uint32_t)ring_q;
		const int budget = (int)(std::min(R, Q)/8);
		uint32_t ST = 1; while(ST < 3u*budget) ST <<= 1;
		std::vector<uint4_t> ra(R), sf(ST), sp(ST); std::vector<uint2_t> rb(R); std::vector<uint32_t> rq(Q);
		ArrayRings rg{ra.data(), rb.data(), rq.data(), sf.data(), sp.data(), R - 1, Q - 1, ST - 1, ST - 1};
		const uint32_t W = R - 3u*budget, QW = Q - 3u*budget;
		ClersState S; clers_state_init(S, io);
		for(int guard = 0; guard < (1 << 30); guard++) {
			rc = clers_run(io, rg, S, budget, ilog2_u32(io.nvert) + 1);
			for(uint32_t f = S.fflush; f < S.start; f++) { const uint4_t v = sf[f & (ST - 1)]; faces[(size_t)f*3] = v.x; faces[(size_t)f*3 + 1] = v.y; faces[(size_t)f*3 + 2] = v.z; }
			for(uint32_t v = S.pflush; v < S.vertex_count; v++) { const uint4_t x = sp[v & (ST - 1)]; pred[(size_t)v*4] = x.x; pred[(size_t)v*4 + 1] = x.y; pred[(size_t)v*4 + 2] = x.z; pred[(size_t)v*4 + 3] = 0; }
			S.fflush = S.start; S.pflush = S.vertex_count;
			const uint32_t e1 = S.nfront > W ? S.nfront - W : 0;
			if(e1 > S.eflush) { for(uint32_t id = S.eflush; id < e1; id++) { const uint4_t a = ra[id & (R - 1)]; const uint2_t l = rb[id & (R - 1)]; ea[id] = EdgeA{a.x, a.y, a.z, a.w}; eb[id] = EdgeB{l.x, l.y}; } S.eflush = e1; }
			const uint32_t q1 = S.norder > QW ? S.norder - QW : 0;
			if(q1 > S.qflush) { for(uint32_t i = std::max(S.qflush, S.cursor); i < q1; i++) order[i] = rq[i & (Q - 1)]; S.qflush = q1; }
			if(rc != 0) break;
		}
		vc = S.vertex_count;
		rc = rc < 0 ? rc : 0;
	} else if(ring_q < 0) {
		// v4 leader / follower (clers_lead + clers_follow) interleaved chunk by chunk; RB = RA = ring_r, Q = -ring_q
		const uint32_t R = (uint32_t)ring_r, Q = (uint32_t)(-ring_q);
		const int budget = (int)(std::min(R, Q)/8);
		uint32_t ST = 1; while(ST < 3u*budget + 8) ST <<= 1;
		uint32_t LGN = 1; while(LGN < 4u*budget + 16) LGN <<= 1;
		std::vector<uint4_t> ra(R), sf(ST), sp(ST); std::vector<uint2_t> rb(R); std::vector<uint32_t> rq(Q), lg(LGN);
		std::vector<uint8_t> rf(R);
		ArrayRings rg{ra.data(), rb.data(), rq.data(), sf.data(), sp.data(), R - 1, Q - 1, ST - 1, ST - 1, lg.data(), LGN - 1, R - 1, rf.data()};
		io.fl = (uint8_t *)order.data();
		const uint32_t W = R - 3u*budget;
		const int splitbits = ilog2_u32(io.nvert) + 1;
		g_log.clear();
		LeadState L; lead_init(L, io);
		FollowState F; follow_init(F);
		int lrc = 0; rc = 0;
		for(int guard = 0; guard < (1 << 30) && rc >= 0; guard++) {
			{   // one chunk, like the kernel: scalar machine + window steps
				const bool vec = getenv("EMUL_VEC") != nullptr;
				uint32_t left = (uint32_t)budget;
				lrc = 0;
				bool tried = false;
				while(lrc == 0 && left > 0) {
					if(vec && left >= 2 && !tried && L.have) { const uint32_t mm = lead_vector_host(io, rg, L, left); if(mm) { left = left > mm ? left - mm : 0; continue; } }
					const uint32_t c0 = L.cler;
					lrc = clers_lead(io, rg, L, tried ? 1 : (int)left, vec && !tried);
					const uint32_t used = L.cler - c0;
					left = left > used ? left - used : 0;
					tried = false;
					if(lrc == 3) { lrc = 0; const uint32_t mm = left >= 2 ? lead_vector_host(io, rg, L, left) : 0; if(mm) left = left > mm ? left - mm : 0; else tried = true; }
					else if(lrc == 4) { lrc = 0; lead_pop_host(io, rg, L); if(!L.have) tried = true; }
				}
			}
			for(uint32_t i = (uint32_t)g_log.size(); i < L.nlog; i++) g_log.push_back(lg[i & (LGN - 1)]);
			if(lrc < 0) { rc = lrc; break; }
			const uint32_t e1 = L.nfront > W ? L.nfront - W : 0;
			if(e1 > L.eflush) { for(uint32_t id = L.eflush; id < e1; id++) { const uint2_t l = rb[id & (R - 1)]; eb[id] = EdgeB{l.x, l.y}; io.fl[id] = rf[id & (R - 1)]; } L.eflush = e1; }
			while(F.tail < L.nlog && rc >= 0) {
				const uint32_t upto = std::min(L.nlog, F.tail + (uint32_t)budget);
				int frc = 0;
				{
					const bool vec = getenv("EMUL_VEC") != nullptr;
					while(frc == 0 && F.tail < upto) {
						frc = clers_follow(io, rg, F, upto, ST, splitbits, vec);
						if(frc != 3) break;
						frc = 0;
						if(follow_vector_host(io, rg, F, upto) == 0) frc = clers_follow(io, rg, F, F.tail + 1, ST, splitbits, false);
					}
				}
				for(uint32_t f = F.fflush; f < F.nfaces; f++) { const uint4_t v = sf[f & (ST - 1)]; faces[(size_t)f*3] = v.x; faces[(size_t)f*3 + 1] = v.y; faces[(size_t)f*3 + 2] = v.z; }
				for(uint32_t v = F.pflush; v < F.vcount; v++) { const uint4_t x = sp[v & (ST - 1)]; pred[(size_t)v*4] = x.x; pred[(size_t)v*4 + 1] = x.y; pred[(size_t)v*4 + 2] = x.z; pred[(size_t)v*4 + 3] = 0; }
				F.fflush = F.nfaces; F.pflush = F.vcount;
				if(frc == 2) { F.nfaces = F.fflush = F.gstart; F.aflush = 0; F.amax = 0; continue; }
				if(frc < 0) { rc = frc; break; }
				const uint32_t a1 = F.amax > W ? F.amax - W : 0;
				if(a1 > F.aflush) { for(uint32_t id = F.aflush; id < a1; id++) { const uint4_t a = ra[id & (R - 1)]; ea[id] = EdgeA{a.x, a.y, a.z, 0}; } F.aflush = a1; }
			}
			if(lrc == 1) break;
		}
		vc = F.vcount;
		rc = rc < 0 ? rc : 0;
	} else if(ring_q == 0) {
		// v7 merged machine (clers_merged) with the kernel's step protocol emulated serially: scalar chunks, CTA-wide windows of W
		// symbols (EMUL_W, default 256), CTA-wide pops, write-back of ring entries leaving the window; R = ring_r
		const uint32_t R = (uint32_t)ring_r;
		const uint32_t W = getenv("EMUL_W") ? (uint32_t)atoi(getenv("EMUL_W")) : 256u;
		const uint32_t runmin = getenv("EMUL_RUNMIN") ? (uint32_t)atoi(getenv("EMUL_RUNMIN")) : 4u;
		const bool vec = getenv("EMUL_VEC") != nullptr;
		const uint32_t KEEP = R - 2u*W;
		std::vector<uint4_t> ra(R); std::vector<uint2_t> rb(R); std::vector<uint8_t> rf(R);
		ArrayRings rg{};
		rg.ra = ra.data(); rg.rb = rb.data(); rg.rf = rf.data(); rg.RM = R - 1; rg.AM = R - 1; rg.syms = clers;
		io.fl = (uint8_t *)order.data();
		const int splitbits = ilog2_u32(io.nvert) + 1;
		const int budget = (int)std::min<uint32_t>(48u, W/3u);
		MergedState S; merged_init(S);
		int mode = 0; bool tried = false;
		rc = 0;
		if(getenv("EMUL_PROTO") && atoi(getenv("EMUL_PROTO")) == 2) {
			// the CURRENT step protocol of k_clers_cta: dispatch -> (window <-> pop) cycles, everything else one scalar symbol at a time
			const uint32_t KEEP2 = R - 2u*W, ROOM = W + 4u;
			for(long guard = 0; guard < (1l << 40) && mode == 0; guard++) {
				if(S.nfront + ROOM > S.eflush + R) {
					const uint32_t e1 = (S.nfront - KEEP2) & ~3u;
					for(uint32_t id = S.eflush; id < e1; id++) { const uint4_t a = ra[id & (R - 1)]; const uint2_t l = rb[id & (R - 1)]; ea[id] = EdgeA{a.x, a.y, a.z, 0}; eb[id] = EdgeB{l.x, l.y}; io.fl[id] = rf[id & (R - 1)]; }
					S.eflush = e1;
					continue;
				}
				const bool canpop = S.start < S.end && S.scan < S.nfront;
				const uint32_t c0 = S.cler < io.nclers ? rg.sym(S.cler) : 0xffu;
				int step = 2;
				if(S.have && !tried && c0 <= (uint32_t)C_LEFT) step = 0; else if(!S.have && canpop) step = 1;
				if(step == 2) {
					const uint32_t c1 = S.cler, s1 = S.start, g1 = S.g, h1 = S.have, n1 = S.ndel, q1 = S.scan;
					int r2 = clers_merged(io, rg, S, 1, false, 2u, splitbits);
					if(r2 == 0 && c1 == S.cler && s1 == S.start && g1 == S.g && h1 == S.have && n1 == S.ndel && q1 == S.scan) r2 = -5;
					tried = false; mode = r2;
					continue;
				}
				for(;;) {
					if(step == 0) { bool bail; const int r2 = cta_window2_host(io, rg, S, W, R, bail); tried = bail; if(!r2) break; step = 1; }
					else { const int r2 = cta_pop2_host(io, rg, S); tried = false; if(!r2) break; step = 0; }
				}
			}
			vc = S.vcount;
			rc = mode < 0 ? mode : 0;
		} else {
		for(long guard = 0; guard < (1l << 40); guard++) {
			if(mode == 1 || mode < 0) break;
			if(S.nfront + W > S.eflush + R) {
				const uint32_t e1 = S.nfront - KEEP;
				for(uint32_t id = S.eflush; id < e1; id++) { const uint4_t a = ra[id & (R - 1)]; const uint2_t l = rb[id & (R - 1)]; ea[id] = EdgeA{a.x, a.y, a.z, 0}; eb[id] = EdgeB{l.x, l.y}; io.fl[id] = rf[id & (R - 1)]; }
				S.eflush = e1;
				continue;
			}
			if(mode == 3) { const uint32_t d = cta_window_host(io, rg, S, W, R); mode = 0; tried = d == 0; }
			else if(mode == 4) { cta_pop_host(io, rg, S); mode = 0; tried = false; }
			else { mode = clers_merged(io, rg, S, tried ? 1 : budget, vec && !tried, runmin, splitbits); tried = false; }
		}
		vc = S.vcount;
		rc = mode < 0 ? mode : 0;
		}
	} else rc = -99;
	for(uint32_t v = 0; v < pm.nvert; v++) for(int k = 0; k < 3; k++) prediction[v*3 + k] = pred[(size_t)v*4 + k];
	return rc;
}

}
