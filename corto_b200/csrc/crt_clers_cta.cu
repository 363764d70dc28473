// crt_clers_cta.cu — k_clers_cta: Decoder::decodeFaces (src/decoder.cpp:204-358) with ONE CTA of 256 threads per mesh.
//
// The automaton is serial per mesh, but a run of VERTEX / LEFT symbols — 99.5 % of a regular mesh, strips of hundreds of
// symbols — is a closed form of two prefix counts (crt_device.cuh, "v7"): thread = symbol, 256 symbols per step, two barriers
// per step, links and labels in the same pass.  Everything else runs on thread 0 through the scalar machine clers_merged.
//
//   shared memory (dynamic):  labels  uint4[R]   (v0, v1, v2, -) of every materialised front edge, ring over edge ids
//                             links   uint2[R]   (prev, next)
//                             flags   u8[R]      0 queued+alive | CLERS_DEL | CLERS_NQ (implicit FIFO: pop = scan for a 0 byte)
//                             symbols u8[4][1024] ring of the CLERS stream, filled by TMA 1-D bulk copies (cp.async.bulk +
//                                                mbarrier) three segments ahead of the cursor, so no step waits on global memory
//   global memory:            reach-back store for ring entries older than the window (ClersScratch slot), delayed stack,
//                             faces and predictions (written directly, one face / one prediction per thread).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "crt_device.cuh"
#include "crt_kernels.h"
#include "crt_ptx.cuh"

namespace crtb {

extern __shared__ __align__(16) uint8_t cta_smem[];

constexpr int CT = 256;                     // threads per CTA
constexpr int CTW = CT/32;
constexpr uint32_t CTA_SEG = 512;           // bytes per symbol segment
constexpr uint32_t CTA_SEG_LOG = 9;
constexpr uint32_t CTA_NSEG = 8;            // segments in the ring
constexpr uint32_t CTA_NSEG_LOG = 3;

struct CtaRings {
	uint32_t aA, aB, aX, aS;     // shared-space byte addresses
	uint32_t RM, sq;             // ring mask; sequence number of this mesh's segment 0 (slot = (sq + seg) & 7)
	__device__ __forceinline__ void ldA(uint32_t id, uint32_t &a, uint32_t &b, uint32_t &c) const {
		[[maybe_unused]] uint32_t d; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(aA + ((id & RM) << 4)));
	}
	__device__ __forceinline__ uint32_t ldA0(uint32_t id) const { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(aA + ((id & RM) << 4))); return v; }
	__device__ __forceinline__ void stA(uint32_t id, uint32_t a, uint32_t b, uint32_t c) {
		asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(aA + ((id & RM) << 4)), "r"(a), "r"(b), "r"(c), "r"(0u) : "memory");
	}
	__device__ __forceinline__ void ldB(uint32_t id, uint32_t &p, uint32_t &n) const {
		asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(p), "=r"(n) : "r"(aB + ((id & RM) << 3)));
	}
	__device__ __forceinline__ void stB(uint32_t id, uint32_t p, uint32_t n) {
		asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(aB + ((id & RM) << 3)), "r"(p), "r"(n) : "memory");
	}
	__device__ __forceinline__ void stB_prev(uint32_t id, uint32_t p) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(aB + ((id & RM) << 3)), "r"(p) : "memory"); }
	__device__ __forceinline__ void stB_next(uint32_t id, uint32_t n) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(aB + ((id & RM) << 3) + 4u), "r"(n) : "memory"); }
	__device__ __forceinline__ uint32_t ldFl(uint32_t id) const { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(aX + (id & RM))); return v; }
	__device__ __forceinline__ uint32_t ldFl4(uint32_t id) const { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(aX + (id & RM))); return v; }   // id % 4 == 0
	__device__ __forceinline__ void stFl(uint32_t id, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(aX + (id & RM)), "r"(v) : "memory"); }
	__device__ __forceinline__ uint32_t sym(uint32_t i) const {              // the segments are one contiguous 4 KB ring
		uint32_t v;
		asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(aS + (((sq << CTA_SEG_LOG) + i) & (CTA_NSEG*CTA_SEG - 1u))));
		return v;
	}
};

constexpr int NHMAX = 4;
struct CtaShared {
	MergedState S;
	int32_t mode;                 // 0 running, 1 done, < 0 error
	uint32_t tried;               // the last window attempt declined its first symbol: it goes the scalar way
	uint32_t flag, newprev;
	uint32_t windowed;            // symbols that went through the CTA-wide windows
	uint32_t pk[2][32];           // per virtual warp of a window: first stop | V count << 8 | L count << 16 (counts before the stop)
	uint32_t wA[2][CTW];          // pop: first alive id per warp
	uint32_t aL[NHMAX*CT + 1];    // prev-chain ids (slow path only), then the label v0 of the k-th edge the run's LEFTs consume
	uint32_t work;
	alignas(8) uint64_t bar[CTA_NSEG];
};

// ---- symbol segments -----------------------------------------------------------------------------------------------------
// Segment s of the current mesh (symbols [512 s, 512 s + 512)) is copy number sq + s of this CTA: slot (sq + s) & 7, and
// the ((sq + s) >> 3)-th use of that slot's mbarrier.  Segments up to (cler >> 9) + 6 are in flight (the slot of segment s + 8
// is free once the cursor has left segment s + 1).  `issued` / `ready` are uniform counters every thread keeps in registers.
__device__ __forceinline__ void cta_symbols(const ClersIO &io, const CtaRings &rg, uint64_t *bar, uint32_t cler, uint32_t nseg, uint32_t &issued, uint32_t &ready,
                                            uint32_t upto /* symbols below this index must be readable */) {
	// (one slot of slack: symbol cler - 1 — the BOUNDARY a window just retired — may still be read by a slow thread of the last step)
	const uint32_t want = min(nseg, (cler >> CTA_SEG_LOG) + CTA_NSEG - 1u);
	if(issued < want) {
		if(threadIdx.x == 0) {
			fence_proxy_async();
			for(uint32_t s = issued; s < want; s++) {
				const uint32_t q = rg.sq + s, slot = q & (CTA_NSEG - 1u);
				uint32_t bytes = io.nclers - s*CTA_SEG;
				if(bytes > CTA_SEG) bytes = CTA_SEG;
				bytes = (bytes + 15u) & ~15u;                      // the symbol arena pads every block (crt_api.cu: add_block)
				mbar_expect_tx(&bar[slot], bytes);
				tma_bulk_g2s(cta_smem + (rg.aS - smem_u32(cta_smem)) + slot*CTA_SEG, io.clers + (size_t)s*CTA_SEG, bytes, &bar[slot]);
			}
		}
		issued = want;
	}
	uint32_t need = (upto + CTA_SEG - 1u) >> CTA_SEG_LOG;
	if(need > issued) need = issued;
	while(ready < need) {
		const uint32_t q = rg.sq + ready;
		mbar_wait(&bar[q & (CTA_NSEG - 1u)], (q >> CTA_NSEG_LOG) & 1u);
		ready++;
	}
}

// ---- CTA-wide window over a run of VERTEX / LEFT symbols -------------------------------------------------------------------
// One window = NH*256 symbols; thread t takes symbols t, t + 256, ... ("virtual warp" w + 8 h holds symbols 32 (w + 8 h) ..).
// A run of VERTEX / LEFT symbols is a closed form of two prefix counts (crt_device.cuh, "v7"); the counts are one packed word per
// virtual warp, scanned with shuffles by every warp.  Runs window after window while the run lasts.  All decisions are taken
// on values every thread reads from shared memory, so every branch around a barrier is uniform.  When the symbol right after
// the run is BOUNDARY / DELAY (the end of a strip on a regular mesh) the window also retires the gate: the thread of the last
// symbol gives it a record (decoder.cpp:283, 327-331 — the edge stays in the front) and the caller goes straight to the FIFO
// pop.  Returns 1 in that case, else 0.  cler / start: uniform cursors, updated.
template <int NH>
__device__ int cta_window(const ClersIO &io, CtaRings &rg, CtaShared &sh, const uint32_t R, const uint32_t nseg, uint32_t &issued, uint32_t &ready,
                          uint32_t &cler, uint32_t &start, const uint32_t end, uint4 &popargs /* scan, ndel, nfront, eflush for the pop */) {
	constexpr uint32_t WS = NH*CT, NVW = NH*CTW;
	const uint32_t FULL = 0xffffffffu;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
	const uint32_t below = (1u << lane) - 1u;
	uint32_t done = 0, it = 0;
	bool bail = false;
	int popnext = 0;
	for(;; it ^= 1u) {
		const uint32_t lim = min(WS, min(io.nclers - cler, end - start));
		cta_symbols(io, rg, sh.bar, cler, nseg, issued, ready, min(io.nclers, cler + lim + 1u));
		// ---- phase A: classify my symbols, one packed word per virtual warp
		uint32_t bV[NH], bL[NH];
#pragma unroll
		for(int h = 0; h < NH; h++) {
			const uint32_t s = tid + (uint32_t)h*CT;
			if((uint32_t)h*CT >= lim) {                        // nothing of the stream in this quarter (uniform): an immediate stop
				bV[h] = bL[h] = 0;
				if(lane == 0) sh.pk[it][w + (uint32_t)h*CTW] = 0u;
				continue;
			}
			const uint32_t c = s < lim ? rg.sym(cler + s) : 0xffu;
			bV[h] = __ballot_sync(FULL, c == C_VERTEX); bL[h] = __ballot_sync(FULL, c == C_LEFT);
			const uint32_t stop = ~(bV[h] | bL[h]);
			const uint32_t k = stop ? (uint32_t)__ffs(stop) - 1u : 32u;
			const uint32_t pm = stop ? (1u << k) - 1u : FULL;
			if(lane == 0) sh.pk[it][w + (uint32_t)h*CTW] = k | ((uint32_t)__popc(bV[h] & pm) << 8) | ((uint32_t)__popc(bL[h] & pm) << 16) | ((bV[h] >> 31) << 24);
		}
		__syncthreads();                                   // B1 (also: state and rings written by the previous step are visible)
		// ---- phase B: ranks
		const uint32_t prev = sh.S.prev, next = sh.S.next, nfront = sh.S.nfront, vcount = sh.S.vcount, eflush = sh.S.eflush, ndel = sh.S.ndel;
		const uint32_t s_v0 = sh.S.v0, s_v1 = sh.S.v1, s_v2 = sh.S.v2, s_scan = sh.S.scan;
		const uint32_t e = lane < NVW ? sh.pk[it][lane] : 32u;
		const uint32_t stopb = __ballot_sync(FULL, (e & 0x3fu) < 32u);
		const uint32_t ws = stopb ? (uint32_t)__ffs(stopb) - 1u : 32u;            // first virtual warp with a stop
		const uint32_t cnt = lane <= ws ? (((e >> 8) & 0x3fu) | (((e >> 16) & 0x3fu) << 16)) : 0u;
		uint32_t incl = cnt;
#pragma unroll
		for(int d = 1; d < (int)NVW; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, d); if(lane >= (uint32_t)d) incl += t; }
		const uint32_t tot = __shfl_sync(FULL, incl, NVW - 1u);
		const uint32_t nV = tot & 0xffffu, nL = tot >> 16;
		const uint32_t excl = incl - cnt;
		uint32_t m = WS;
		{ const uint32_t ek = __shfl_sync(FULL, e, ws & 31u); if(ws < NVW) m = ws*32u + (ek & 0x3fu); }
		if(m > lim) m = lim;                               // (symbols past lim read 0xff: m <= lim already; belt and braces)
		if(m < 1u || nfront + nV + 1u > io.cap || vcount + nV > io.nvert) { bail = done == 0; break; }
		if(nfront + nV + 1u > eflush + R) break;           // the caller writes ring entries back first
		uint32_t nVb[NH], nLb[NH], id[NH], a[NH], pV[NH];
		// ---- prev chain, fast path: consecutive ids prev, prev + 1, ... (the queued edges of one earlier strip; on a regular mesh
		// the strip before the last one, i.e. ~2.3 strips of ids back)
		bool good = true;
#pragma unroll
		for(int h = 0; h < NH; h++) {
			const uint32_t base = __shfl_sync(FULL, excl, w + (uint32_t)h*CTW);
			pV[h] = (__shfl_sync(FULL, e, (w + (uint32_t)h*CTW - 1u) & 31u) >> 24) & 1u;     // was the last symbol of the virtual warp before mine a VERTEX
			nVb[h] = (base & 0xffffu) + (uint32_t)__popc(bV[h] & below);
			nLb[h] = (base >> 16) + (uint32_t)__popc(bL[h] & below);
			id[h] = prev + nLb[h]; a[h] = 0;
			if(tid + (uint32_t)h*CT < m && ((bL[h] >> lane) & 1u)) {
				bool g = id[h] < nfront && id[h] != next;
				if(g) {
					uint32_t pk, pn;
					if(id[h] >= eflush) { rg.ldB(id[h], pk, pn); a[h] = rg.ldA0(id[h]); }
					else { const EdgeB l_ = io.eb[id[h]]; pk = l_.prev; pn = l_.next; a[h] = io.ea[id[h]].v0; }   // reach-back store: a strip further back than the ring
					(void)pn;
					if(nLb[h] + 1u == nL) sh.newprev = pk; else g = pk == id[h] + 1u;
					sh.aL[nLb[h]] = a[h];
				}
				good = good && g;
			}
		}
		if(nL == 0 && tid == 0) sh.newprev = prev;
		if(!__syncthreads_and(good)) {                     // B2
			// slow path: thread 0 walks the chain; a walk that wraps around to the right-hand neighbour (small loop) means the
			// links change inside this window: scalar machine
			if(tid == 0) {
				uint32_t q = prev, ok = 1;
				for(uint32_t k = 0; k < nL; k++) {
					sh.aL[k] = q;
					if(q == next || q >= nfront) { ok = 0; break; }
					uint32_t pp, pq;
					if(q >= eflush) rg.ldB(q, pp, pq); else { const uint2_t t_ = lead_g_load(io.eb, q); pp = t_.x; pq = t_.y; }
					(void)pq;
					q = pp;
				}
				sh.newprev = q;
				sh.flag = ok;
			}
			__syncthreads();
			if(!sh.flag) { bail = done == 0; break; }
#pragma unroll
			for(int h = 0; h < NH; h++) {
				if(tid + (uint32_t)h*CT < m && ((bL[h] >> lane) & 1u)) {
					id[h] = sh.aL[nLb[h]];
					if(id[h] >= eflush) a[h] = rg.ldA0(id[h]); else a[h] = follow_g_load(io.ea, id[h]).x;
					sh.aL[nLb[h]] = a[h];
				}
			}
			__syncthreads();
		}
		// ---- the symbol after the run: BOUNDARY / DELAY retire the gate right here
		const uint32_t newprev = sh.newprev;
		const uint32_t nextf = nV ? nfront + nV - 1u : next;   // right-hand neighbour of the gate after the run
		const uint32_t gid = nfront + nV;                      // id of the gate's record, if it gets one
		uint32_t cm = 0xffu;
		if(cler + m < io.nclers && start + m < end) cm = rg.sym(cler + m);
		const bool gend = (cm == C_BOUNDARY || (cm == C_DELAY && ndel < io.cap)) && newprev != nextf && newprev < nfront;
		// ---- phase C: labels, outputs, ring records
#pragma unroll
		for(int h = 0; h < NH; h++) {
			const uint32_t s = tid + (uint32_t)h*CT;
			if((uint32_t)h*CT >= m) break;
			if(s < m) {
				const bool isV = (bV[h] >> lane) & 1u;
				const bool prevIsV = lane ? ((bV[h] >> (lane - 1u)) & 1u) : pV[h] != 0u;
				const uint32_t v0i = nLb[h] ? sh.aL[nLb[h] - 1u] : s_v0;
				const uint32_t v1i = nVb[h] ? vcount + nVb[h] - 1u : s_v1;
				uint32_t v2i;
				if(s == 0) v2i = s_v2;
				else if(prevIsV) v2i = nVb[h] > 1u ? vcount + nVb[h] - 2u : s_v1;
				else v2i = nLb[h] > 1u ? sh.aL[nLb[h] - 2u] : s_v0;
				const uint32_t x = vcount + nVb[h];
				const uint32_t third = isV ? x : a[h];
				const size_t at = (size_t)(start + s)*3u;
				if(io.faces16) { io.faces16[at] = (uint16_t)v1i; io.faces16[at + 1] = (uint16_t)v0i; io.faces16[at + 2] = (uint16_t)third; }
				else { io.faces32[at] = v1i; io.faces32[at + 1] = v0i; io.faces32[at + 2] = third; }
				if(isV) {
					((uint4 *)io.pred)[x] = make_uint4(v1i, v0i, v2i, 0u);
					const uint32_t b = nfront + nVb[h];
					rg.stA(b, x, v1i, v0i);
					rg.stB(b, nVb[h] + 1u < nV ? b + 1u : (gend ? gid : CLERS_NOLINK), nVb[h] ? b - 1u : next);
					rg.stFl(b, 0u);
				} else {
					if(id[h] >= eflush) rg.stFl(id[h], CLERS_DEL); else lead_g_set_flag(io.fl, id[h], CLERS_DEL);
				}
				if(s == m - 1u) {
					const uint32_t g0 = isV ? v0i : a[h], g1 = isV ? x : v1i, g2 = isV ? v1i : v0i;     // the gate after the run
					sh.S.v0 = g0; sh.S.v1 = g1; sh.S.v2 = g2;
					if(gend) {
						rg.stA(gid, g0, g1, g2); rg.stB(gid, newprev, nextf); rg.stFl(gid, CLERS_NQ);
						if(newprev >= eflush) rg.stB_next(newprev, gid); else clers_g_set_next(io.eb, newprev, gid);
						if(nV == 0) { if(next >= eflush) rg.stB_prev(next, gid); else clers_g_set_prev(io.eb, next, gid); }
						if(cm == C_DELAY) io.delayed[ndel] = gid;
					}
				}
			}
		}
		if(tid == 0) {
			if(nV) {
				if(next >= eflush) rg.stB_prev(next, nfront); else clers_g_set_prev(io.eb, next, nfront);
				sh.S.next = nextf;
			}
			sh.S.nfront = nfront + nV + (gend ? 1u : 0u); sh.S.vcount = vcount + nV;
			sh.S.prev = newprev;
			sh.S.start = start + m; sh.S.cler = cler + m + (gend ? 1u : 0u);
			sh.S.lp = sh.S.ln = 1; sh.S.cf = CLERS_NOID;
			sh.S.have = (!gend && start + m < end) ? 1u : 0u;
			if(gend && cm == C_DELAY) sh.S.ndel = ndel + 1u;
		}
		cler += m + (gend ? 1u : 0u); start += m; done += m;
		if(gend) { popnext = 1; popargs = make_uint4(s_scan, ndel + (cm == C_DELAY ? 1u : 0u), nfront + nV + 1u, eflush); break; }
		if(m < WS || start >= end || cler >= io.nclers) break;                 // the run ended (or the group / the stream did)
		if(nfront + nV + WS + 1u > eflush + R) break;                          // ring entries have to be written back first
	}
	if(tid == 0) { sh.tried = bail ? 1u : 0u; sh.windowed += done; }
	return popnext;
}

// ---- CTA-wide pop of the implicit FIFO: 1024 flag bytes per step -------------------------------------------------------------
// The popped edge becomes the gate.  When the symbol waiting for it is BOUNDARY / DELAY (decoder.cpp:283, 327-331: the edge keeps
// its record, nothing else changes) the symbol is consumed right here and the scan goes on from the next id.  Returns 1 when a
// gate is loaded and a VERTEX / LEFT symbol is next (the caller goes straight to the window), else 0.
// Nothing of the shared state is read here (scan, ndel, nfront, eflush come in registers): thread 0's writes at the end are
// ordered against every other thread's reads by the barriers of the caller's next step.
__device__ int cta_pop(const ClersIO &io, CtaRings &rg, CtaShared &sh, const uint32_t nseg, uint32_t &issued, uint32_t &ready, uint32_t &cler, const uint4 args) {
	const uint32_t FULL = 0xffffffffu;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
	uint32_t scan = args.x, ndel = args.y;
	const uint32_t nfront = args.z, eflush = args.w;
	uint32_t found = CLERS_NOID, it = 0, c = 0xffu;
	while(scan < nfront) {
		// the next queued edge is usually one of the first few ids: every thread reads the same eight flags (no barrier)
		if(scan >= eflush) {
			const uint32_t b0 = scan & ~3u;
			const uint32_t w0 = rg.ldFl4(b0), w1 = rg.ldFl4(b0 + 4u);
			uint32_t hit = CLERS_NOID;
#pragma unroll
			for(int k = 7; k >= 0; k--) {
				const uint32_t idk = b0 + (uint32_t)k, f = ((k < 4 ? w0 : w1) >> (8*(k & 3))) & 0xffu;
				if(f == 0u && idk >= scan && idk < nfront) hit = idk;
			}
			if(hit == CLERS_NOID && b0 + 8u >= nfront) { scan = nfront; break; }
			found = hit;
		}
		if(found == CLERS_NOID) {
		// thread t looks at the four ids base + 4 t .. + 3 (one 32-bit load when they are in the ring)
		const uint32_t base = scan & ~3u, id0 = base + 4u*tid;
		uint32_t word = 0x01010101u;
		if(id0 < nfront) {
			if(id0 >= eflush) word = rg.ldFl4(id0);
			else word = (uint32_t)io.fl[id0] | ((uint32_t)io.fl[id0 + 1] << 8) | ((uint32_t)io.fl[id0 + 2] << 16) | ((uint32_t)io.fl[id0 + 3] << 24);   // (store slots are padded)
		}
		uint32_t mine = CLERS_NOID;
#pragma unroll
		for(int k = 3; k >= 0; k--) { const uint32_t idk = id0 + (uint32_t)k; if(((word >> (8*k)) & 0xffu) == 0u && idk >= scan && idk < nfront) mine = idk; }
		const uint32_t bal = __ballot_sync(FULL, mine != CLERS_NOID);
		const uint32_t first = __shfl_sync(FULL, mine, bal ? (uint32_t)__ffs(bal) - 1u : 0u);
		if(lane == 0) sh.wA[it][w] = bal ? first : CLERS_NOID;
		__syncthreads();
		found = CLERS_NOID;
#pragma unroll
		for(int j = CTW - 1; j >= 0; j--) { const uint32_t x = sh.wA[it][j]; if(x != CLERS_NOID) found = x; }
		it ^= 1u;
		if(found == CLERS_NOID) { scan = base + 4u*CT; continue; }
		}
		scan = found + 1u;
		c = 0xffu;
		if(cler < io.nclers) { cta_symbols(io, rg, sh.bar, cler, nseg, issued, ready, cler + 1u); c = rg.sym(cler); }
		if(c == C_BOUNDARY || (c == C_DELAY && ndel < io.cap)) {
			if(c == C_DELAY) { if(tid == 0) io.delayed[ndel] = found; ndel++; }
			cler++;
			found = CLERS_NOID;
			continue;
		}
		break;
	}
	if(tid == 0) {
		sh.S.scan = scan < nfront ? scan : nfront;
		sh.S.cler = cler; sh.S.ndel = ndel;
		if(found != CLERS_NOID) {
			uint32_t p, q, a, b, c2;
			if(found >= eflush) { rg.ldB(found, p, q); rg.ldA(found, a, b, c2); }
			else { const uint2_t t_ = lead_g_load(io.eb, found); p = t_.x; q = t_.y; const uint4_t u_ = follow_g_load(io.ea, found); a = u_.x; b = u_.y; c2 = u_.z; }
			sh.S.prev = p; sh.S.next = q; sh.S.v0 = a; sh.S.v1 = b; sh.S.v2 = c2;
			sh.S.lp = sh.S.ln = 0; sh.S.have = 1; sh.S.cf = found;
		}
		sh.tried = 0;
	}
	return found != CLERS_NOID && c <= (uint32_t)C_LEFT;
}

template <int NHK>
__global__ void __launch_bounds__(CT) k_clers_cta(DevBatch B, const uint32_t *mesh_order, uint32_t nwork, ClersScratch scratch, uint32_t *ticket, uint32_t R, bool defer) {
	constexpr uint32_t WS = NHK*CT;
	__shared__ CtaShared sh;
	const uint32_t tid = threadIdx.x;
	CtaRings rg;
	{
		uint32_t sbase;
		asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"((uint32_t)__cvta_generic_to_shared(cta_smem)));
		rg.aA = sbase; rg.aB = sbase + R*16u; rg.aX = rg.aB + R*8u; rg.aS = rg.aX + R;
		rg.RM = R - 1u; rg.sq = 0;
	}
	if(tid == 0) { for(uint32_t k = 0; k < CTA_NSEG; k++) mbar_init(&sh.bar[k], 1); sh.flag = 0; }
	const uint32_t KEEP = R - 2u*WS;                       // ring entries kept after a write-back
	const uint32_t ROOM = WS + 4u;                         // ids one step can allocate: a window (+ the gate's record), a start triangle
	for(;;) {
		__syncthreads();
		if(tid == 0) sh.work = atomicAdd(ticket, 1u);
		__syncthreads();
		const uint32_t wk = sh.work;
		if(wk >= nwork) break;
		const uint32_t mi = mesh_order[wk];
		const MeshDesc *M = B.mesh + mi;
		const TunDesc td = B.tun[M->clers_tun];
		ClersIO io;
		io.clers = B.symbols + td.out_off; io.nclers = td.size;
		io.split = (const uint32_t *)(B.blobs + M->split_off); io.split_nwords = M->split_nwords;
		io.group_ends = B.group_ends + M->group0; io.ngroups = M->ngroups;
		io.nvert = M->nvert; io.nface = M->nface;
		const size_t slot = blockIdx.x;
		io.cap = scratch.cap;
		io.ea = scratch.ea + slot*scratch.cap; io.eb = scratch.eb + slot*scratch.cap;
		io.order = scratch.order + slot*scratch.cap; io.delayed = scratch.delayed + slot*scratch.cap;
		const uint32_t need = 3u*M->max_group_faces + 3u;
		if(need < io.cap) io.cap = need;
		io.faces32 = M->index16 ? nullptr : (uint32_t *)M->face_ptr;
		io.faces16 = M->index16 ? (uint16_t *)M->face_ptr : nullptr;
		io.pred = (uint32_t *)M->pred_ptr;
		io.fl = (uint8_t *)io.order;                       // no FIFO is stored: the `order` scratch backs the flag ring
		const int splitbits = ilog2_u32(io.nvert) + 1;
		const uint32_t nseg = (io.nclers + CTA_SEG - 1u)/CTA_SEG;
		uint32_t issued = 0, ready = 0, hint = WS;
		// Irregular meshes (real scans: runs of ~6 symbols between RIGHT / DELAY / BOUNDARY) gain nothing from CTA-wide windows and
		// pay for the barriers: a sample of the stream (how many RIGHT / DELAY symbols — BOUNDARY is common on regular meshes
		// too: every strip of a grid ends in two) decides, and such a mesh is left to the leader / follower kernel that
		// launch_clers_cta starts right after this one (B.regular[mi] bit 31 = "deferred").
		if(defer) {
			// 16 chunks of 1 KB spread over the stream (streams shorter than 32 K symbols are not worth a second kernel)
			const uint32_t ns = io.nclers >= 32768u ? 16384u : 0u;
			uint32_t other = 0;
			for(uint32_t p = 0; p < 4u && ns; p++) {
				const uint32_t chunk = p*4u + (tid >> 6);
				const uint32_t at = (((io.nclers >> 4)*chunk) & ~15u) + (tid & 63u)*16u;      // < nclers - 1024 + 1024
				const uint4 q = *(const uint4 *)(io.clers + at);
				const uint32_t ws[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
				for(int k = 0; k < 4; k++) {
#pragma unroll
					for(int j = 0; j < 4; j++) { const uint32_t c = (ws[k] >> (8*j)) & 0xffu; other += (c == (uint32_t)C_RIGHT || c == (uint32_t)C_DELAY) ? 1u : 0u; }
				}
			}
			other = __reduce_add_sync(0xffffffffu, other);
			if((tid & 31u) == 0) atomicAdd(&sh.flag, other);
			__syncthreads();
			const uint32_t irregular = ns && sh.flag*16u > ns;     // > 6 % RIGHT / DELAY (a real scan: ~14 %; grids ~0, a grid with a punched hole up to 3 %)
			__syncthreads();
			if(tid == 0) sh.flag = 0;
			if(irregular) { if(tid == 0) B.regular[mi] = 0x80000000u; continue; }
		}
		if(tid == 0) { merged_init(sh.S); sh.mode = 0; sh.tried = 0; sh.windowed = 0; }
		int mode;
		for(;;) {
			__syncthreads();                               // state, rings and outputs of the previous step are visible
			mode = sh.mode;
			if(mode) break;
			uint32_t cler = sh.S.cler, start = sh.S.start;
			const uint32_t nfront = sh.S.nfront, eflush = sh.S.eflush, end = sh.S.end;
			if(nfront + ROOM > eflush + R) {               // write ring entries leaving the window back to the reach-back store
				const uint32_t e1 = (nfront - KEEP) & ~3u;     // (a multiple of 4: the pop reads four flags per load)
				for(uint32_t id = eflush + tid; id < e1; id += CT) {
					uint32_t a, b, c, p, n;
					rg.ldA(id, a, b, c); rg.ldB(id, p, n);
					io.ea[id] = EdgeA{a, b, c, 0}; io.eb[id] = EdgeB{p, n}; io.fl[id] = (uint8_t)rg.ldFl(id);
				}
				__syncthreads();
				if(tid == 0) sh.S.eflush = e1;
				continue;
			}
			const uint32_t have = sh.S.have, tried = sh.tried;
			uint4 popargs = make_uint4(sh.S.scan, sh.S.ndel, nfront, eflush);
			const bool canpop = start < end && popargs.x < nfront;
			__syncthreads();                               // every thread has read the state: thread 0 may change it from here on
			cta_symbols(io, rg, sh.bar, cler, nseg, issued, ready, min(io.nclers, cler + 2u));
			const uint32_t c0 = cler < io.nclers ? rg.sym(cler) : 0xffu;
			int step = 2;                                  // 0 window, 1 pop, 2 one scalar symbol
			if(have && !tried && c0 <= (uint32_t)C_LEFT) step = 0;
			else if(!have && canpop) step = 1;
			if(step == 2) {
				if(tid == 0) {
					// everything else, one symbol at a time: RIGHT, END, SPLIT, BOUNDARY / DELAY on a gate the window did not retire, start
					// triangles, group changes, the delayed stack, and the symbol a window declined
					const uint32_t c1 = sh.S.cler, s1 = sh.S.start, g1 = sh.S.g, h1 = sh.S.have, n1 = sh.S.ndel, q1 = sh.S.scan;
					int rc = clers_merged(io, rg, sh.S, 1, false, 2u, splitbits);
					if(rc == 0 && c1 == sh.S.cler && s1 == sh.S.start && g1 == sh.S.g && h1 == sh.S.have && n1 == sh.S.ndel && q1 == sh.S.scan) rc = -5;   // no progress
					sh.tried = 0;
					sh.mode = rc;
				}
				continue;
			}
			// strips: window(s) -> the gate retires on BOUNDARY / DELAY -> pop -> window(s) ... without passing through the dispatch above
			for(;;) {
				if(step == 0) {
					// window width by the length of the last run: strips of a regular mesh change slowly, and a narrow window is fewer
					// instructions per step
					const uint32_t before = cler;
					int r;
					if(NHK >= 4 && hint > 3u*CT) r = cta_window<4>(io, rg, sh, R, nseg, issued, ready, cler, start, end, popargs);
					else if(NHK >= 4 && hint > 2u*CT) r = cta_window<3>(io, rg, sh, R, nseg, issued, ready, cler, start, end, popargs);
					else if(NHK >= 2 && hint > (uint32_t)CT) r = cta_window<2>(io, rg, sh, R, nseg, issued, ready, cler, start, end, popargs);
					else r = cta_window<1>(io, rg, sh, R, nseg, issued, ready, cler, start, end, popargs);
					hint = cler - before;
					if(!r) break;
					__syncthreads();                       // the flags and the state this window wrote are visible to the scan
					step = 1;
				} else {
					if(!cta_pop(io, rg, sh, nseg, issued, ready, cler, popargs)) break;
					step = 0;
				}
			}
		}
		// every copy that was issued has to land before the slots are reused by the next mesh
		while(ready < issued) { const uint32_t q = rg.sq + ready; mbar_wait(&sh.bar[q & (CTA_NSEG - 1u)], (q >> CTA_NSEG_LOG) & 1u); ready++; }
		rg.sq += issued;
		uint32_t vcount = sh.S.vcount;
		const bool bad = mode < 0;
		if(tid == 0) {
			if(bad) B.status[mi] = -5;
			B.vertex_count[mi] = vcount;
			B.regular[mi] = (uint64_t)sh.windowed*4u >= (uint64_t)io.nclers*3u ? 1u : 0u;   // the delta inverse picks its algorithm by this
		}
		// vertices the stream never created (corrupt / truncated input): neutral prediction so later passes stay in bounds
		uint4 *pred = (uint4 *)M->pred_ptr;
		if(bad) vcount = 0;
		for(uint32_t v = vcount + tid; v < M->nvert; v += CT) pred[v] = make_uint4(0, 0, 0, 0);
	}
}

int launch_clers_cta(const DevBatch &B, const uint32_t *order, uint32_t nwork, const ClersScratch &scratch, uint32_t *ticket, int sms, bool defer, cudaStream_t s) {
	// ring size: as much shared memory per mesh as leaves every mesh of the batch resident (up to 4 CTAs of 256 threads per SM);
	// windows of 1024 symbols where the ring is large enough to keep two of them plus a strip of reach-back, else 512
	uint32_t R = 8192;
	if(nwork > (uint32_t)sms) R = 4096;
	if(nwork > 2u*(uint32_t)sms) R = 2048;
	const size_t smem = (size_t)R*25u + CTA_NSEG*CTA_SEG;
	const uint32_t g = nwork < scratch.slots ? nwork : scratch.slots;
	cudaError_t e;
	if(R >= 4096) {
		e = cudaFuncSetAttribute(k_clers_cta<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if(e != cudaSuccess) return (int)e;
		k_clers_cta<4><<<g, CT, smem, s>>>(B, order, nwork, scratch, ticket, R, defer);
	} else {
		e = cudaFuncSetAttribute(k_clers_cta<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if(e != cudaSuccess) return (int)e;
		k_clers_cta<2><<<g, CT, smem, s>>>(B, order, nwork, scratch, ticket, R, defer);
	}
	e = cudaGetLastError();
	return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace crtb
