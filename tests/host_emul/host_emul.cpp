// tests/host_emul/host_emul.cpp — CPU exercise of the __host__ __device__ cores the kernels run
// (corto_b200/csrc/crt_device.cuh) plus the host directory walk, so their logic is checked without a GPU.
// Built and driven by tests/test_host_emul.py; compares against the oracle from Python.  Not part of the product.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include "../../corto_b200/csrc/crt_device.cuh"
#include "../../corto_b200/csrc/crt_walk.h"
#include "../../include/corto_b200.h"

using namespace crtb;

extern "C" {

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
		LeadState L; lead_init(L, io);
		FollowState F; follow_init(F);
		int lrc = 0; rc = 0;
		for(int guard = 0; guard < (1 << 30) && rc >= 0; guard++) {
			lrc = clers_lead(io, rg, L, budget);
			if(lrc < 0) { rc = lrc; break; }
			const uint32_t e1 = L.nfront > W ? L.nfront - W : 0;
			if(e1 > L.eflush) { for(uint32_t id = L.eflush; id < e1; id++) { const uint2_t l = rb[id & (R - 1)]; eb[id] = EdgeB{l.x, l.y}; io.fl[id] = rf[id & (R - 1)]; } L.eflush = e1; }
			while(F.tail < L.nlog && rc >= 0) {
				const uint32_t upto = std::min(L.nlog, F.tail + (uint32_t)budget);
				const int frc = clers_follow(io, rg, F, upto, ST, splitbits);
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
	} else rc = -99;
	for(uint32_t v = 0; v < pm.nvert; v++) for(int k = 0; k < 3; k++) prediction[v*3 + k] = pred[(size_t)v*4 + k];
	return rc;
}

}
