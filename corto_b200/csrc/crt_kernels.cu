// crt_kernels.cu — hand-written sm_100a kernels of the corto decode path (no tensor cores: there is no dense
// contraction anywhere in this path, it is byte / bit / index work bound by HBM and by one serial automaton).
//
// Kernel           replaces (reference file:line)                                   parallel structure
// k_tun_tables     Tunstall::createDecodingTables2   src/tunstall.cpp:125-256       one warp per entropy block
// k_tun_decode     Tunstall::decompress              src/tunstall.cpp:430-452       tiles of compressed bytes; the dictionary
//                  + InStream::decompress NONE path  src/cstream.cpp:68-73          staged in smem by TMA (cp.async.bulk +
//                                                                                   mbarrier); offsets = decoupled look-back
// k_unpack_chain   decodeArray / decodeValues        include/corto/cstream.h:294-360  meshes: tiles of 1024 vertices, all
// k_unpack_fused<MESH>                                                               components; bit offset = running total
//                                                                                   per chain, or block scan + look-back
// k_cloud_chain / k_unpack_fused<CLOUD>  ... + GenericAttr::deltaDecode (cloud) vertex_attribute.h:177-181, NormalAttr::deltaDecode
//                  (cloud) normal_attribute.cpp:202-207 + dequantize: point clouds in ONE pass; one CTA per chain with both
//                  carries in shared memory, or ticketed tiles + look-back for a few large clouds
// k_clers_cta      Decoder::decodeFaces              src/decoder.cpp:204-358        crt_clers_cta.cu: one CTA per mesh, runs of
//                                                                                   VERTEX / LEFT in closed form (regular streams)
// k_clers_lf       (the same)                                                       serial automaton, two warps per mesh: the
//                                                                                   irregular streams k_clers_cta defers
// k_clers          (the same, single-warp variant kept as A/B baseline, CORTO_CLERS=1)
// k_delta_mesh_seg GenericAttr::deltaDecode (mesh)   vertex_attribute.h:165-176     CTA per (mesh, attribute): segmented-scan
//                  NormalAttr::deltaDecode (mesh)    normal_attribute.cpp:193-201   rounds of 256 vertices (regular meshes);
// k_delta_mesh     (the same)                                                       warp per (mesh, attribute[, component])
// k_delta_mesh_cta (the same, one CTA per chain, pointer-doubling rounds; experiment switch CORTO_DELTA=cta)
// k_adj_build / k_scan_u32 / k_normal_estimate
//                  markBoundary, estimateNormals, computeNormals   normal_attribute.cpp:24-59, 281-325
// k_dequant        GenericAttr::dequantize, NormalAttr::dequantize, ColorAttr::dequantize
//                  vertex_attribute.h:184-230, normal_attribute.cpp:257-279, color_attribute.cpp:76-95
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "crt_device.cuh"
#include "crt_kernels.h"
#include "crt_ptx.cuh"

namespace crtb {

// =========================================================================================================
// small device utilities
// =========================================================================================================
// ---- decoupled look-back over a chain of tiles ------------------------------------------------------------
// state word: bits 63..62 = 0 empty | 1 aggregate | 2 inclusive prefix; bits 61..0 = value.  One 64-bit word
// carries flag and value together, so a relaxed volatile load/store pair is enough.  Tiles are taken in ticket
// order, hence every predecessor of a running tile is running or done: the spin cannot deadlock.
constexpr uint64_t LB_AGG = 1ull << 62, LB_PFX = 2ull << 62, LB_MASK = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t lb_load(const uint64_t *p) {
	uint64_t v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void lb_store(uint64_t *p, uint64_t v) {
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
// Called by ONE thread.  Returns the exclusive prefix of tile `t` (0 when `first`), publishes the inclusive one.
__device__ uint64_t lookback(uint64_t *states, uint32_t t, bool first, uint64_t aggregate) {
	aggregate &= LB_MASK;
	if(first) { lb_store(states + t, LB_PFX | aggregate); return 0; }
	lb_store(states + t, LB_AGG | aggregate);
	uint64_t excl = 0;
	uint32_t p = t - 1;
	for(;;) {
		uint64_t s = lb_load(states + p);
		uint64_t flag = s >> 62;
		if(flag == 0) continue;
		excl += s & LB_MASK;
		if(flag == 2) break;
		p--;
	}
	excl &= LB_MASK;
	lb_store(states + t, LB_PFX | ((excl + aggregate) & LB_MASK));
	return excl;
}

// Same, for chains interleaved in one state array: the state of (tile t, slot k) lives at states[t*stride + k].
__device__ uint64_t lookback_strided(uint64_t *states, uint32_t t, uint32_t stride, uint32_t slot, bool first, uint64_t aggregate) {
	aggregate &= LB_MASK;
	uint64_t *me = states + (size_t)t*stride + slot;
	if(first) { lb_store(me, LB_PFX | aggregate); return 0; }
	lb_store(me, LB_AGG | aggregate);
	uint64_t excl = 0;
	const uint64_t *p = me - stride;
	for(;;) {
		const uint64_t s = lb_load(p);
		const uint64_t flag = s >> 62;
		if(flag == 0) continue;
		excl += s & LB_MASK;
		if(flag == 2) break;
		p -= stride;
	}
	excl &= LB_MASK;
	lb_store(me, LB_PFX | ((excl + aggregate) & LB_MASK));
	return excl;
}

// exclusive scan of one u32 per thread across a 256-thread CTA; returns the exclusive prefix, total in *total.
__device__ __forceinline__ uint32_t cta_scan_excl_256(uint32_t v, uint32_t *s_warp /*[9]*/, uint32_t *total) {
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for(int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if(lane >= d) inc += o; }
	if(lane == 31) s_warp[w] = inc;
	__syncthreads();
	if(w == 0) {
		uint32_t x = lane < 8 ? s_warp[lane] : 0, xi = x;
#pragma unroll
		for(int d = 1; d < 8; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, xi, d); if(lane >= d) xi += o; }
		if(lane < 8) s_warp[lane] = xi - x;
		if(lane == 7) s_warp[8] = xi;
	}
	__syncthreads();
	uint32_t r = s_warp[w] + inc - v;
	*total = s_warp[8];
	__syncthreads();
	return r;
}

#define NEXT_TILE(ticket, ntiles, s_tile)                       \
	if(threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);        \
	__syncthreads();                                             \
	const uint32_t tile_id = s_tile;                             \
	__syncthreads();                                             \
	if(tile_id >= (uint32_t)(ntiles)) break;

// =========================================================================================================
// K1  Tunstall dictionaries
// =========================================================================================================
// Warp-cooperative build of one dictionary, step for step what tun_build_seq (crt_device.cuh) does on one thread:
//   * the most probable queue head (first strict maximum, tunstall.cpp:211-217) = warp arg-max over the n rows;
//   * its n children (slot, probability, text = parent text + one symbol) = one lane per child;
//   * the compaction of live slots into entry[256] (tunstall.cpp:242-255) = ballot + prefix count, 32 slots per step.
// The low-entropy branch (run-length words, tunstall.cpp:145-194) stays on lane 0: it is rare and tiny.
__device__ uint32_t tun_build_warp(const uint8_t *probs /* (sym,prob) pairs */, const uint32_t n, TunScratch &S, uint8_t *text, uint32_t *entry, const uint32_t lane) {
	const uint32_t FULL = 0xffffffffu;
	for(uint32_t i = lane; i < 512; i += 32) S.qprob[i] = 0;
	uint32_t pos = 0, slots = 0, nwords;
	const uint32_t p0 = (uint32_t)probs[1] << 8, p1 = (uint32_t)probs[3] << 8;
	uint32_t run = 2, pr = (p0*p0) >> 16;
	const uint32_t max_run = 255u/(n - 1);
	while(pr > p1 && run < max_run) { pr = (pr*p0) >> 16; run++; }
	__syncwarp();
	if(run >= 16) {
		if(lane == 0) {
			text[pos++] = probs[0];
			for(uint32_t k = 1; k < n; k++) {
				for(uint32_t i = 0; i + 1 < run; i++) text[pos++] = probs[0];
				text[pos++] = probs[2*k];
			}
			S.head[0] = (uint16_t)((run - 1)*n);
			for(uint32_t k = 1; k < n; k++) S.head[k] = (uint16_t)k;
			uint32_t q = pr;
			for(uint32_t c = 0; c < run; c++) {
				for(uint32_t k = 1; k < n; k++) {
					const uint32_t sl = k + c*n, pk = (uint32_t)probs[2*k + 1] << 8;
					S.qprob[sl] = (c == 0) ? pk : ((q*pk) >> 16);
					S.widx[sl] = (uint16_t)(k*run - c);
					S.wlen[sl] = (uint16_t)(c + 1);
				}
				q = (c == 0) ? p0 : ((q*p0) >> 16);
			}
			const uint32_t s0 = (run - 1)*n;
			S.qprob[s0] = q; S.widx[s0] = 0; S.wlen[s0] = (uint16_t)run;
		}
		pos = 1 + (n - 1)*run;
		nwords = 1 + run*(n - 1);
		slots = run*n;
	} else {
		for(uint32_t k = lane; k < n; k += 32) {
			S.head[k] = (uint16_t)k;
			S.qprob[k] = (uint32_t)probs[2*k + 1] << 8;
			S.widx[k] = (uint16_t)k; S.wlen[k] = 1;
			text[k] = probs[2*k];
		}
		pos = n; slots = n; nwords = n;
	}
	__syncwarp();
	while(nwords < 256) {
		// most probable head: first strict maximum in row order
		uint32_t bp = 0, bk = 0;
		for(uint32_t k = lane; k < n; k += 32) { const uint32_t p = S.qprob[S.head[k]]; if(p > bp) { bp = p; bk = k; } }
#pragma unroll
		for(int d = 16; d; d >>= 1) {
			const uint32_t op = __shfl_xor_sync(FULL, bp, d), ok = __shfl_xor_sync(FULL, bk, d);
			if(op > bp || (op == bp && ok < bk)) { bp = op; bk = ok; }
		}
		const uint32_t best = bp ? bk : 0u;
		const uint32_t parent = S.head[best], pp = S.qprob[parent], poff = S.widx[parent], plen = S.wlen[parent];
		const uint32_t room = 256 - nwords;        // the child that makes word 256 ends the round (tunstall.cpp:234-235)
		const uint32_t m = room < n ? room : n;    // children this round
		__syncwarp();
		for(uint32_t k = lane; k < m; k += 32) {
			const uint32_t sl = slots + k, at = pos + k*(plen + 1);
			if(sl < 512) {
				S.qprob[sl] = (pp*((uint32_t)probs[2*k + 1] << 8)) >> 16;
				S.widx[sl] = (uint16_t)at; S.wlen[sl] = (uint16_t)(plen + 1);
			}
			if(at + plen + 1 <= (uint32_t)TUN_TABLE_BYTES) {
				for(uint32_t j = 0; j < plen; j++) text[at + j] = text[poff + j];
				text[at + plen] = probs[2*k];
			}
		}
		slots += m; pos += m*(plen + 1);
		if(room > n && lane == 0) S.head[best] = (uint16_t)(parent + n);   // parent retires only if the loop ran to completion (:237-238)
		nwords += n - 1;
		__syncwarp();
	}
	if(slots > 512) slots = 512;
	uint32_t word = 0;
	const uint32_t below = (1u << lane) - 1u;
	for(uint32_t s0 = 0; s0 < slots && word < 256; s0 += 32) {
		const uint32_t sl = s0 + lane;
		const bool keep = sl < slots && !(S.head[sl % n] > sl);
		const uint32_t kb = __ballot_sync(FULL, keep);
		const uint32_t at = word + __popc(kb & below);
		if(keep && at < 256) entry[at] = (uint32_t)S.widx[sl] | ((uint32_t)S.wlen[sl] << 16);
		word += __popc(kb);
	}
	if(word > 256) word = 256;
	for(uint32_t i = word + lane; i < 256; i += 32) entry[i] = 0;
	__syncwarp();
	return pos < (uint32_t)TUN_TABLE_BYTES ? pos : (uint32_t)TUN_TABLE_BYTES;
}

__global__ void __launch_bounds__(32) k_tun_tables(DevBatch B, bool seq) {
	const int t = blockIdx.x;
	const TunDesc td = B.tun[t];
	if(td.raw || td.nsym <= 1) return;
	__shared__ TunScratch S;
	__shared__ __align__(16) uint8_t text[TUN_TABLE_BYTES];
	__shared__ __align__(16) uint32_t entry[256];
	__shared__ uint8_t probs[512];
	__shared__ uint32_t s_used;
	const int lane = threadIdx.x;
	for(uint32_t i = lane; i < 2*td.nsym; i += 32) probs[i] = B.blobs[td.probs_off + i];
	__syncwarp();
	uint32_t used;
	if(seq) {                                              // CORTO_TUN=seq: the one-thread build (A/B baseline, tests)
		if(lane == 0) s_used = tun_build_seq(probs, td.nsym, S, text, entry);
		__syncwarp();
		used = s_used;
	} else used = tun_build_warp(probs, td.nsym, S, text, entry, (uint32_t)lane);
	const uint32_t used16 = (used + 15u) & ~15u;
	uint8_t *rec = B.tunrec + (size_t)t*TUN_REC_BYTES;
	uint4 *dst = (uint4 *)rec;
	const uint4 *se = (const uint4 *)entry;
	for(int i = lane; i < 64; i += 32) dst[i] = se[i];
	const uint4 *st = (const uint4 *)text;
	for(uint32_t i = lane; i < used16/16; i += 32) dst[64 + i] = st[i];
	if(lane == 0) B.tun_used[t] = used16;
}

// =========================================================================================================
constexpr uint32_t TUN_STAGE = 16384;   // output bytes of one tile assembled in shared memory (mean expansion is 2.6-4.5x of 2048)

// K2  Tunstall decode:  out_off[i] = sum_{j<i} len[data[j]];  byte i copies its word;  the last byte of a block
//     copies exactly the remainder (tunstall.cpp:446-451).  Raw (NONE) and single-symbol blocks are copies/fills.
// =========================================================================================================
__global__ void __launch_bounds__(256) k_tun_decode(DevBatch B, const Tile *tiles, uint32_t ntiles, uint64_t *states, uint32_t *ticket) {
	__shared__ __align__(16) uint32_t s_entry[256];
	__shared__ __align__(16) uint8_t s_text[TUN_TABLE_BYTES];
	__shared__ __align__(16) uint8_t s_stage[TUN_STAGE + 32];
	__shared__ __align__(8) uint64_t s_bar;
	__shared__ uint32_t s_warp[9];
	__shared__ uint32_t s_tile;
	__shared__ uint64_t s_base;
	const int tid = threadIdx.x;
	if(tid == 0) mbar_init(&s_bar, 1);
	__syncthreads();
	uint32_t parity = 0;
	for(;;) {
		NEXT_TILE(ticket, ntiles, s_tile)
		const Tile tl = tiles[tile_id];
		const TunDesc td = B.tun[tl.a];
		uint8_t *out = B.symbols + td.out_off;
		if(td.raw || td.nsym <= 1) {
			// tile covers TUN_TILE*4 output bytes
			const uint32_t lo = tl.tile*(TUN_TILE*4u);
			uint32_t hi = lo + TUN_TILE*4u; if(hi > td.size) hi = td.size;
			uint32_t ssum = 0;
			if(td.raw) { const uint8_t *in = B.blobs + td.data_off; for(uint32_t i = lo + tid; i < hi; i += 256) { const uint8_t v = in[i]; out[i] = v; ssum += v; } }
			else { const uint8_t sym = td.nsym ? B.blobs[td.probs_off] : 0; for(uint32_t i = lo + tid; i < hi; i += 256) { out[i] = sym; ssum += sym; } }
#pragma unroll
			for(int d = 16; d; d >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, d);
			if((tid & 31) == 0 && ssum) atomicAdd(B.tun_bits + tl.a, (unsigned long long)ssum);
			continue;
		}
		// stage the dictionary (entries + used text) with one TMA bulk copy per part
		if(tid == 0) {
			const uint32_t used16 = B.tun_used[tl.a];
			const uint8_t *rec = B.tunrec + (size_t)tl.a*TUN_REC_BYTES;
			fence_proxy_async();
			mbar_expect_tx(&s_bar, 1024u + used16);
			tma_bulk_g2s(s_entry, rec, 1024u, &s_bar);
			if(used16) tma_bulk_g2s(s_text, rec + 1024, used16, &s_bar);
		}
		// my 8 compressed bytes
		const uint8_t *in = B.blobs + td.data_off;
		const uint32_t i0 = tl.tile*TUN_TILE + tid*8u;
		uint8_t by[8];
#pragma unroll
		for(int j = 0; j < 8; j++) by[j] = (i0 + j < td.csize) ? in[i0 + j] : 0;
		mbar_wait(&s_bar, parity); parity ^= 1;
		uint32_t mylen = 0;
#pragma unroll
		for(int j = 0; j < 8; j++) if(i0 + j < td.csize) mylen += s_entry[by[j]] >> 16;
		uint32_t total;
		uint32_t off = cta_scan_excl_256(mylen, s_warp, &total);
		if(tid == 0) s_base = lookback(states, tile_id, tl.first != 0, total);
		__syncthreads();
		const uint64_t obase = s_base;
		uint32_t ssum = 0;
		if(total <= TUN_STAGE && obase + total <= td.size && !(tl.tile*TUN_TILE + TUN_TILE >= td.csize)) {
			// common case: the tile's words are assembled in shared memory — at the byte phase (obase & 15) of their place in the
			// output, so that 16-byte chunks of the stage ARE 16-byte chunks of the output — then written with 128-bit stores
			// (the symbol arena aligns every block to 16 bytes, crt_api.cu: add_block)
			const uint32_t ph = (uint32_t)obase & 15u;
			uint32_t o = ph + off;
#pragma unroll
			for(int j = 0; j < 8; j++) {
				const uint32_t i = i0 + j;
				if(i >= td.csize) break;
				const uint32_t e = s_entry[by[j]];
				uint32_t st = e & 0xffffu, len = e >> 16;
				while(len) {                                   // eight text bytes per step: two aligned words, funnel-shifted
					const uint32_t a = st & ~3u, sh = (st & 3u)*8u;
					const uint32_t w0 = *(const uint32_t *)(s_text + (a & (TUN_TABLE_BYTES - 1))), w1 = *(const uint32_t *)(s_text + ((a + 4u) & (TUN_TABLE_BYTES - 1))),
					               w2 = *(const uint32_t *)(s_text + ((a + 8u) & (TUN_TABLE_BYTES - 1)));
					uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
					const uint32_t n = len < 8u ? len : 8u;
					if(n < 4u) { lo &= (1u << (8u*n)) - 1u; hi = 0; } else if(n < 8u) hi &= (1u << (8u*(n - 4u))) - 1u;
					ssum = __dp4a(lo, 0x01010101u, ssum); ssum = __dp4a(hi, 0x01010101u, ssum);
#pragma unroll
					for(int k = 0; k < 4; k++) if((uint32_t)k < n) s_stage[o + k] = (uint8_t)(lo >> (8*k));
#pragma unroll
					for(int k = 0; k < 4; k++) if((uint32_t)(k + 4) < n) s_stage[o + 4 + k] = (uint8_t)(hi >> (8*k));
					o += n; st += n; len -= n;
				}
			}
			__syncthreads();
			const uint32_t head = ph ? 16u - ph : 0u;          // bytes up to the first 16-byte boundary of the output
			const uint32_t nvec = total > head ? (total - head) >> 4 : 0u, tail0 = head + (nvec << 4);
			uint8_t *dst = out + obase;
			for(uint32_t i = tid; i < nvec; i += 256) *(uint4 *)(dst + head + (i << 4)) = *(const uint4 *)(s_stage + ph + head + (i << 4));
			if((uint32_t)tid < head && (uint32_t)tid < total) dst[tid] = s_stage[ph + tid];
			if(total > head) { const uint32_t t2 = tail0 + (uint32_t)tid; if(tid < 16 && t2 < total) dst[t2] = s_stage[ph + t2]; }
		} else {
			// long words, the last tile of a block (its last byte is clipped, tunstall.cpp:446-451) or a corrupt stream: direct stores
			uint64_t o = obase + off;
#pragma unroll
			for(int j = 0; j < 8; j++) {
				const uint32_t i = i0 + j;
				if(i >= td.csize) break;
				const uint32_t e = s_entry[by[j]];
				const uint32_t st = e & 0xffffu;
				uint32_t len = e >> 16;
				if(i == td.csize - 1) len = o < td.size ? (uint32_t)(td.size - o) : 0;   // last byte: the remainder
				for(uint32_t k = 0; k < len && o + k < td.size; k++) { const uint8_t v = s_text[(st + k) & (TUN_TABLE_BYTES - 1)]; out[o + k] = v; ssum += v; }
				o += len;
			}
		}
#pragma unroll
		for(int d = 16; d; d >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, d);
		if((tid & 31) == 0 && ssum) atomicAdd(B.tun_bits + tl.a, (unsigned long long)ssum);
		__syncthreads();   // everyone done with s_entry/s_text before the next TMA overwrites them
	}
}

// =========================================================================================================
// K4  CLERS automaton: one warp per mesh.  Lane 0 runs the serial machine (clers_run, crt_device.cuh) against
//     shared-memory rings; every CLERS_BUDGET symbols ALL lanes drain the staged faces / predictions to global memory
//     with coalesced stores and write ring entries leaving the window back to the reach-back store.  The machine is
//     instruction-latency bound (ncu: CPI 4.4, one warp per SM sub-partition), so the design goal is the smallest
//     number of dependent instructions per symbol, not bandwidth.
// =========================================================================================================
extern __shared__ __align__(16) uint8_t crt_smem[];

// Shared-memory ring policy for clers_run: explicit 32-bit shared-space addressing (ld/st.shared), bases precomputed
// once, so a ring access is mask + shift-add + LDS/STS.
struct SmemRings {
	uint32_t aA, aB, aQ, aF, aP;      // shared-space byte addresses of the rings
	uint32_t RM, QM, FM, PM;
	__device__ __forceinline__ void ldA(uint32_t id, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) const {
		asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(aA + ((id & RM) << 4)));
	}
	__device__ __forceinline__ void stA(uint32_t id, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
		asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(aA + ((id & RM) << 4)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
	}
	__device__ __forceinline__ void stA_del(uint32_t id) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(aA + ((id & RM) << 4) + 12u), "r"(1u) : "memory"); }
	__device__ __forceinline__ void ldB(uint32_t id, uint32_t &p, uint32_t &n) const {
		asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(p), "=r"(n) : "r"(aB + ((id & RM) << 3)));
	}
	__device__ __forceinline__ void stB(uint32_t id, uint32_t p, uint32_t n) {
		asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(aB + ((id & RM) << 3)), "r"(p), "r"(n) : "memory");
	}
	__device__ __forceinline__ void stB_prev(uint32_t id, uint32_t p) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(aB + ((id & RM) << 3)), "r"(p) : "memory"); }
	__device__ __forceinline__ void stB_next(uint32_t id, uint32_t n) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(aB + ((id & RM) << 3) + 4u), "r"(n) : "memory"); }
	__device__ __forceinline__ uint32_t ldQ(uint32_t i) const { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(aQ + ((i & QM) << 2))); return v; }
	__device__ __forceinline__ void stQ(uint32_t i, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(aQ + ((i & QM) << 2)), "r"(v) : "memory"); }
	__device__ __forceinline__ void stF(uint32_t face, uint32_t a, uint32_t b, uint32_t c) {
		asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(aF + ((face & FM) << 4)), "r"(a), "r"(b), "r"(c), "r"(0u) : "memory");
	}
	__device__ __forceinline__ void stP(uint32_t v, uint32_t a, uint32_t b, uint32_t c) {
		asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(aP + ((v & PM) << 4)), "r"(a), "r"(b), "r"(c), "r"(0u) : "memory");
	}
};

constexpr int CLERS_BUDGET = 80;       // symbols between drains; staging rings hold >= 3*budget entries
constexpr uint32_t CLERS_STAGE = 256;  // faces / predictions staged (power of two >= 3*CLERS_BUDGET)

__global__ void __launch_bounds__(32) k_clers(DevBatch B, const uint32_t *mesh_order, uint32_t nwork, ClersScratch scratch, uint32_t *ticket,
                                              uint32_t R, uint32_t Q) {
	const uint32_t lane = threadIdx.x;
	SmemRings rg;
	const uint32_t oA = 0, oB = R*16u, oQ = oB + R*8u, oF = oQ + Q*4u, oP = oF + CLERS_STAGE*16u;
	uint32_t sbase;   // opaque move: keeps the shared-window base in a register instead of re-deriving it at every access
	asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"((uint32_t)__cvta_generic_to_shared(crt_smem)));
	rg.aA = sbase + oA; rg.aB = sbase + oB; rg.aQ = sbase + oQ; rg.aF = sbase + oF; rg.aP = sbase + oP;
	rg.RM = R - 1; rg.QM = Q - 1; rg.FM = CLERS_STAGE - 1; rg.PM = CLERS_STAGE - 1;
	const uint32_t W = R - 3u*CLERS_BUDGET, QW = Q - 3u*CLERS_BUDGET;
	for(;;) {
		uint32_t w = 0;
		if(lane == 0) w = atomicAdd(ticket, 1u);
		w = __shfl_sync(0xffffffffu, w, 0);
		if(w >= nwork) break;
		const uint32_t mi = mesh_order[w];
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
		ClersState S;
		clers_state_init(S, io);
		int splitbits;
		asm volatile("mov.u32 %0, %1;" : "=r"(splitbits) : "r"(ilog2_u32(io.nvert) + 1));
		int rc = 0;
		for(;;) {
			if(lane == 0) rc = clers_run(io, rg, S, CLERS_BUDGET, splitbits);
			rc = __shfl_sync(0xffffffffu, rc, 0);
			// ---- cooperative drain (all lanes) ----
			const uint32_t f0 = __shfl_sync(0xffffffffu, S.fflush, 0), f1 = __shfl_sync(0xffffffffu, S.start, 0);
			const uint32_t p0 = __shfl_sync(0xffffffffu, S.pflush, 0), p1 = __shfl_sync(0xffffffffu, S.vertex_count, 0);
			const uint32_t e0 = __shfl_sync(0xffffffffu, S.eflush, 0), nf = __shfl_sync(0xffffffffu, S.nfront, 0);
			const uint32_t q0 = __shfl_sync(0xffffffffu, S.qflush, 0), no = __shfl_sync(0xffffffffu, S.norder, 0), cu = __shfl_sync(0xffffffffu, S.cursor, 0);
			__syncwarp();
			{   // faces: 3 index words per face, consecutive lanes write consecutive words
				const uint32_t nw = (f1 - f0)*3u;
				const uint4 *sf = (const uint4 *)(crt_smem + oF);
				for(uint32_t k = lane; k < nw; k += 32) {
					const uint32_t face = f0 + k/3u, comp = k - (k/3u)*3u;
					const uint32_t v = ((const uint32_t *)(sf + (face & rg.FM)))[comp];
					const size_t at = (size_t)f0*3u + k;
					if(io.faces16) io.faces16[at] = (uint16_t)v; else io.faces32[at] = v;
				}
			}
			{   // predictions
				const uint4 *sp = (const uint4 *)(crt_smem + oP);
				uint4 *dst = (uint4 *)io.pred;
				for(uint32_t v = p0 + lane; v < p1; v += 32) dst[v] = sp[v & rg.PM];
			}
			const uint32_t e1 = nf > W ? nf - W : 0u;
			if(e1 > e0) {   // ring entries leaving the window -> reach-back store
				for(uint32_t id = e0 + lane; id < e1; id += 32) {
					const uint4 a = ((const uint4 *)(crt_smem + oA))[id & rg.RM]; const uint2 l = ((const uint2 *)(crt_smem + oB))[id & rg.RM];
					io.ea[id] = EdgeA{a.x, a.y, a.z, a.w}; io.eb[id] = EdgeB{l.x, l.y};
				}
			}
			const uint32_t q1 = no > QW ? no - QW : 0u;
			if(q1 > q0) for(uint32_t i = max(q0, cu) + lane; i < q1; i += 32) io.order[i] = rg.ldQ(i);
			__syncwarp();
			if(lane == 0) { S.fflush = f1; S.pflush = p1; if(e1 > e0) S.eflush = e1; if(q1 > q0) S.qflush = q1; }
			if(rc != 0) break;
		}
		uint32_t vcount = __shfl_sync(0xffffffffu, S.vertex_count, 0);
		if(lane == 0) { if(rc < 0) B.status[mi] = rc; B.vertex_count[mi] = vcount; }
		// vertices the stream never created (corrupt / truncated input): neutral prediction so later passes stay in bounds
		uint4 *pred = (uint4 *)M->pred_ptr;
		if(rc < 0) vcount = 0;
		for(uint32_t v = vcount + lane; v < M->nvert; v += 32) pred[v] = make_uint4(0, 0, 0, 0);
		__syncwarp();
	}
}

// =========================================================================================================
// K4b  CLERS automaton, leader / follower (clers_lead + clers_follow, crt_device.cuh): two warps per mesh.
//      warp 0: lane 0 = link machine (the serial chain), all lanes = write-back of link-ring entries leaving the window
//      warp 1: lane 0 = label machine replaying the leader's log, all lanes = coalesced drains of faces / predictions and
//              write-back of label-ring entries.
//      The warps live on different SM sub-partitions, talk through a shared-memory log ring + head/tail words, and never
//      share global state.  Every spin is bounded; a stuck partner turns into CRT_E_TOPOLOGY, not a hang.
// =========================================================================================================
struct SmemRings4 {
	uint32_t aB, aX, aL, aA, aF, aP;      // shared-space byte addresses: links, flags, log, labels, staged faces / predictions
	uint32_t RM, LM, AM, FM, PM;
	__device__ __forceinline__ void ldB(uint32_t id, uint32_t &p, uint32_t &n) const {
		asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(p), "=r"(n) : "r"(aB + ((id & RM) << 3)));
	}
	__device__ __forceinline__ void stB(uint32_t id, uint32_t p, uint32_t n) {
		asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(aB + ((id & RM) << 3)), "r"(p), "r"(n) : "memory");
	}
	__device__ __forceinline__ void stB_prev(uint32_t id, uint32_t p) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(aB + ((id & RM) << 3)), "r"(p) : "memory"); }
	__device__ __forceinline__ void stB_next(uint32_t id, uint32_t n) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(aB + ((id & RM) << 3) + 4u), "r"(n) : "memory"); }
	__device__ __forceinline__ uint32_t ldFl(uint32_t id) const { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(aX + (id & RM))); return v; }
	__device__ __forceinline__ void stFl(uint32_t id, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(aX + (id & RM)), "r"(v) : "memory"); }
	__device__ __forceinline__ void stLog(uint32_t i, uint32_t w) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(aL + ((i & LM) << 2)), "r"(w) : "memory"); }
	__device__ __forceinline__ uint32_t ldLog(uint32_t i) const { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(aL + ((i & LM) << 2))); return v; }
	__device__ __forceinline__ void ldA(uint32_t id, uint32_t &a, uint32_t &b, uint32_t &c) const {
		[[maybe_unused]] uint32_t d; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(aA + ((id & AM) << 4)));
	}
	__device__ __forceinline__ void stA(uint32_t id, uint32_t a, uint32_t b, uint32_t c) {
		asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(aA + ((id & AM) << 4)), "r"(a), "r"(b), "r"(c), "r"(0u) : "memory");
	}
	__device__ __forceinline__ void stF(uint32_t face, uint32_t a, uint32_t b, uint32_t c) {
		asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(aF + ((face & FM) << 4)), "r"(a), "r"(b), "r"(c), "r"(0u) : "memory");
	}
	__device__ __forceinline__ void stP(uint32_t v, uint32_t a, uint32_t b, uint32_t c) {
		asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(aP + ((v & PM) << 4)), "r"(a), "r"(b), "r"(c), "r"(0u) : "memory");
	}
};

__device__ __forceinline__ uint32_t ld_vol_shared(const uint32_t *p) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v; }
__device__ __forceinline__ void st_vol_shared(uint32_t *p, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(smem_u32(p)), "r"(v) : "memory"); }

// ---- warp-wide window steps -------------------------------------------------------------------------------------------
// A run of VERTEX / LEFT symbols is a closed form for the link machine: VERTEX k pushes new id nfront+k (its links are
// its neighbours in the run), LEFT k consumes the k-th edge of the prev chain.  Only the prev-chain walk is serial (one
// dependent shared-memory load per LEFT); ids, link records, flags and log words are written by 32 lanes at once.  The
// scalar machine yields (return 3) when it sees such a run and takes over again at the first other symbol.
// One call runs window after window (32 symbols each) for as long as the run and the chunk budget last; inside the loop
// the machine state is warp-uniform (every lane holds the same copy), so nothing is shuffled between windows.  The
// symbols of the NEXT window are loaded as soon as the length of this one is known, so that the global-memory latency
// hides behind the rest of the step (`pre` carries them from call to call).
// Returns the number of symbols consumed (0: the first window must go through the scalar path — loop closure in sight,
// capacity edge, run too short).  S is valid in lane 0 only.
struct LeadPre { uint32_t sym, at; };    // per lane: clers[at + lane] (0xff past the end); `at` is warp-uniform

__device__ uint32_t lead_vector(const ClersIO &io, SmemRings4 &rg, LeadState &S, uint32_t left, uint32_t *chain, uint32_t lane, LeadPre &pre) {
	const uint32_t FULL = 0xffffffffu;
	__syncwarp();                                          // lane 0's scalar ring stores are visible to every lane from here
	if(!__shfl_sync(FULL, S.have, 0)) return 0;            // no current edge: the scalar machine has to pop one first
	uint32_t cler = __shfl_sync(FULL, S.cler, 0), start = __shfl_sync(FULL, S.start, 0);
	const uint32_t end = __shfl_sync(FULL, S.end, 0), eflush = __shfl_sync(FULL, S.eflush, 0);
	uint32_t nfront = __shfl_sync(FULL, S.nfront, 0), next = __shfl_sync(FULL, S.cnext, 0), prev = __shfl_sync(FULL, S.cprev, 0);
	uint32_t nlog = __shfl_sync(FULL, S.nlog, 0);
	const uint32_t below = (1u << lane) - 1u;
	uint32_t done = 0;
	for(;;) {
		const uint32_t lim = min(min(32u, left - done), min(io.nclers - cler, end - start));
		if(lim < 2) break;
		uint32_t sym = pre.sym;
		if(pre.at != cler) sym = cler + lane < io.nclers ? (uint32_t)io.clers[cler + lane] : 0xffu;
		// Fast path for the prev chain: the edges a strip consumes on its left are usually the queued edges of ONE earlier strip,
		// created back to back and linked in creation order (prev of id k is k+1).  Every lane checks one link; if all of them
		// hold the chain is prev, prev+1, ... and nothing is walked.  Any mismatch falls back to the serial walk.  The link is
		// loaded first thing (it only depends on `prev`), so its latency hides behind the symbol ballots.
		const uint32_t idk = prev + lane;
		uint32_t pk = 0xffffffffu, pn;
		if(idk >= eflush && idk < nfront && idk != next) rg.ldB(idk, pk, pn);
		const bool isV = lane < lim && sym == C_VERTEX, isL = lane < lim && sym == C_LEFT;
		const uint32_t bV = __ballot_sync(FULL, isV), bL = __ballot_sync(FULL, isL);
		const uint32_t stop = ~(bV | bL);
		const uint32_t m = stop ? (uint32_t)__ffs(stop) - 1u : 32u;
		if(m < 2) { pre.at = cler; pre.sym = sym; break; }
		pre.at = cler + m;                                 // the usual case: the run goes on right behind this window
		pre.sym = pre.at + lane < io.nclers ? (uint32_t)io.clers[pre.at + lane] : 0xffu;
		const uint32_t pm = m == 32 ? FULL : ((1u << m) - 1u);
		const uint32_t Vm = bV & pm, Lm = bL & pm;
		const uint32_t nV = __popc(Vm), nL = __popc(Lm);
		if(nfront + nV > io.cap) break;                    // the scalar machine flags it
		uint32_t p = prev + __popc(Lm & below), newprev = prev + nL;
		if(!__all_sync(FULL, lane >= nL || pk == idk + 1)) {
			uint32_t ok = 1;
			if(lane == 0) {
				uint32_t q = prev;
				for(uint32_t k = 0; k < nL; k++) {             // the serial part: walk the prev chain
					chain[k] = q;
					if(q == next) ok = 0;                      // the walk wraps around to the right-hand neighbour (small loop): its links
					                                           // change inside this window, so take the scalar path
					uint32_t pp, pq;
					if(q >= eflush) rg.ldB(q, pp, pq); else { const uint2_t t_ = lead_g_load(io.eb, q); pp = t_.x; pq = t_.y; }
					(void)pq;
					q = pp;
				}
				chain[nL] = q;
			}
			__syncwarp();
			ok = __shfl_sync(FULL, ok, 0);
			if(!ok) break;
			p = chain[__popc(Lm & below)]; newprev = chain[nL];
		}
		if(lane < m) {                                     // one flag byte and one log word per lane, whatever the symbol
			const uint32_t r = __popc(Vm & below), b = nfront + r;
			const uint32_t id = isV ? b : p;
			if(isV) rg.stB(b, r + 1 < nV ? b + 1 : CLERS_NOLINK, r ? b - 1 : next);
			if(id >= eflush) rg.stFl(id, isV ? 0u : CLERS_DEL); else lead_g_set_flag(io.fl, id, CLERS_DEL);
			rg.stLog(nlog + lane, ((isV ? (uint32_t)LG_V : (uint32_t)LG_L) << 28) | id);
		}
		if(nV) {
			if(lane == 0) { if(next >= eflush) rg.stB_prev(next, nfront); else clers_g_set_prev(io.eb, next, nfront); }
			nfront += nV; next = nfront - 1;
		}
		prev = newprev;
		nlog += m; start += m; cler += m; done += m;
		__syncwarp();                                      // this window's ring stores before the next window's loads (and `chain` reuse)
		if(start >= end) break;
	}
	if(done && lane == 0) {
		S.nfront = nfront; S.cprev = prev; S.cnext = next;
		S.lp = S.ln = 1; S.cf = CLERS_NOID;
		S.nlog = nlog; S.start = start;
		S.cler = cler; S.cwv = 0;                           // the scalar machine re-primes its symbol window when it next runs
		if(start >= end) S.have = 0;
	}
	return done;
}

// Implicit-FIFO pop, 32 flag bytes per step: most queued edges are dead by the time the scan reaches them (a grid deletes
// ~99 % of them), so the scalar machine would pay one dependent shared-memory load per dead edge.  On success the popped
// edge becomes the current edge (links loaded, LG_P logged); otherwise scan == nfront and the scalar machine goes on
// with the delayed stack / a start triangle.
__device__ void lead_pop_vector(const ClersIO &io, SmemRings4 &rg, LeadState &S, uint32_t lane) {
	const uint32_t FULL = 0xffffffffu;
	__syncwarp();
	uint32_t scan = __shfl_sync(FULL, S.scan, 0);
	const uint32_t nfront = __shfl_sync(FULL, S.nfront, 0), eflush = __shfl_sync(FULL, S.eflush, 0);
	uint32_t found = CLERS_NOID;
	while(scan < nfront) {
		const uint32_t id = scan + lane;
		uint32_t fl = 0xffu;
		if(id < nfront) fl = id >= eflush ? rg.ldFl(id) : lead_g_flag(io.fl, id);
		const uint32_t alive = __ballot_sync(FULL, fl == 0);
		if(alive) { found = scan + (uint32_t)__ffs(alive) - 1u; scan = found + 1; break; }
		scan += 32;
	}
	if(lane == 0) {
		S.scan = scan < nfront ? scan : nfront;
		if(found != CLERS_NOID) {
			uint32_t p, q;
			if(found >= eflush) rg.ldB(found, p, q); else { const uint2_t t_ = lead_g_load(io.eb, found); p = t_.x; q = t_.y; }
			S.cprev = p; S.cnext = q; S.lp = S.ln = 0; S.have = 1; S.cf = found;
			rg.stLog(S.nlog, ((uint32_t)LG_P << 28) | found); S.nlog++;
		}
	}
	__syncwarp();
}

// Drain of the staged faces [f0, f1) and predictions [p0, p1) to global memory (all lanes).
__device__ __forceinline__ void follow_drain(const ClersIO &io, const SmemRings4 &rg, uint32_t oF, uint32_t oP, uint32_t f0, uint32_t f1, uint32_t p0, uint32_t p1, uint32_t lane) {
	const uint32_t nw = (f1 - f0)*3u;
	const uint4 *sf = (const uint4 *)(crt_smem + oF);
	for(uint32_t k = lane; k < nw; k += 32) {
		const uint32_t face = f0 + k/3u, comp = k - (k/3u)*3u;
		const uint32_t v = ((const uint32_t *)(sf + (face & rg.FM)))[comp];
		const size_t at = (size_t)f0*3u + k;
		if(io.faces16) io.faces16[at] = (uint16_t)v; else io.faces32[at] = v;
	}
	const uint4 *sp = (const uint4 *)(crt_smem + oP);
	uint4 *dst = (uint4 *)io.pred;
	for(uint32_t v = p0 + lane; v < p1; v += 32) dst[v] = sp[v & rg.PM];
}

// Label machine over a run of VERTEX / LEFT log words: "who defined v0 / v1 last" is a bit trick on the ballot masks,
// label loads of the LEFTs go out in parallel.  Like lead_vector, one call runs window after window on warp-uniform
// state.  Faces and predictions of a window go straight to global memory (one face / one prediction per lane; whatever
// the scalar machine had staged before is drained first so that the staged range stays a contiguous suffix).
// Returns the number of log words consumed.
__device__ uint32_t follow_vector(const ClersIO &io, SmemRings4 &rg, FollowState &F, uint32_t upto, uint32_t lane, uint32_t oF, uint32_t oP) {
	const uint32_t FULL = 0xffffffffu;
	__syncwarp();
	uint32_t tail = __shfl_sync(FULL, F.tail, 0);
	if(upto - tail < 2) return 0;
	uint32_t vcount = __shfl_sync(FULL, F.vcount, 0), nf = __shfl_sync(FULL, F.nfaces, 0), amax = __shfl_sync(FULL, F.amax, 0);
	const uint32_t aflush = __shfl_sync(FULL, F.aflush, 0);
	uint32_t v0 = __shfl_sync(FULL, F.v0, 0), v1 = __shfl_sync(FULL, F.v1, 0), v2 = __shfl_sync(FULL, F.v2, 0);
	const uint32_t lt = (1u << lane) - 1u;
	uint32_t done = 0;
	for(;;) {
		const uint32_t lim = min(32u, upto - tail);
		if(lim < 2) break;
		const uint32_t w = lane < lim ? rg.ldLog(tail + lane) : 0xffffffffu;
		const uint32_t t = w >> 28, id = w & 0x0FFFFFFFu;
		const bool isV = t == LG_V, isL = t == LG_L;
		// the label a LEFT consumes is loaded at once, for every LEFT word in sight (harmless past the end of the window: ids in
		// the log are valid ring / scratch indices), so that its latency hides behind the ballots that find the window end
		uint32_t a = 0;
		if(isL) {
			uint32_t t1, t2;
			if(id >= aflush) rg.ldA(id, a, t1, t2); else { const uint4_t g_ = follow_g_load(io.ea, id); a = g_.x; }
		}
		const uint32_t vall = __ballot_sync(FULL, isV), lall = __ballot_sync(FULL, isL);
		uint32_t stop = ~(vall | lall);
		// a LEFT may consume an edge that a VERTEX of this very window creates (the prev chain wrapped around a small loop): its
		// label does not exist yet, so the window ends in front of it.  Ids grow with creation order: compare with the first V's.
		const uint32_t firstid = __shfl_sync(FULL, id, vall ? __ffs(vall) - 1 : 0);
		stop |= __ballot_sync(FULL, isL && vall && id >= firstid && (uint32_t)(__ffs(vall) - 1) < lane);
		const uint32_t m = stop ? (uint32_t)__ffs(stop) - 1u : 32u;
		if(m < 2) break;
		const uint32_t pm = m == 32 ? FULL : ((1u << m) - 1u);
		const uint32_t Vm = vall & pm, Lm = lall & pm;
		const uint32_t nV = __popc(Vm);
		if(vcount + nV > io.nvert || nf + m > io.nface) break;          // let the scalar machine flag the error
		if(!done) {                                                     // first window of this call: flush what the scalar machine staged
			const uint32_t f0 = __shfl_sync(FULL, F.fflush, 0), p0 = __shfl_sync(FULL, F.pflush, 0);
			if(f0 != nf || p0 != vcount) follow_drain(io, rg, oF, oP, f0, nf, p0, vcount, lane);
		}
		// who defined v0 / v1 last before this lane: v1 is a count (vertex ids are consecutive), v0 needs the label of the last LEFT
		const uint32_t lLT = Lm & lt, vLT = Vm & lt, vLT1 = Vm & (lt >> 1);
		const uint32_t a_lt = __shfl_sync(FULL, a, lLT ? 31 - __clz(lLT) : 0);
		const uint32_t v0_before = lLT ? a_lt : v0;
		const uint32_t x = vcount + __popc(vLT);
		const uint32_t v1_before = vLT ? x - 1u : v1, v1_before_p = vLT1 ? vcount + __popc(vLT1) - 1u : v1;   // own / lane-1's
		const uint32_t pv0b = __shfl_up_sync(FULL, v0_before, 1);
		const uint32_t v2_before = lane == 0 ? v2 : (((Vm >> (lane - 1)) & 1u) ? v1_before_p : pv0b);
		if(lane < m) {
			const size_t at = (size_t)(nf + lane)*3u;
			const uint32_t third = isV ? x : a;
			if(io.faces16) { io.faces16[at] = (uint16_t)v1_before; io.faces16[at + 1] = (uint16_t)v0_before; io.faces16[at + 2] = (uint16_t)third; }
			else { io.faces32[at] = v1_before; io.faces32[at + 1] = v0_before; io.faces32[at + 2] = third; }
			if(isV) { ((uint4 *)io.pred)[x] = make_uint4(v1_before, v0_before, v2_before, 0u); rg.stA(id, x, v1_before, v0_before); }
		}
		const uint32_t v0e = Lm ? __shfl_sync(FULL, a, 31 - __clz(Lm)) : v0;
		const uint32_t v2e = __shfl_sync(FULL, isV ? v1_before : v0_before, m - 1);
		if(nV) { amax = __shfl_sync(FULL, id, 31 - __clz(Vm)) + 1u; v1 = vcount + nV - 1u; }
		v0 = v0e; v2 = v2e;
		vcount += nV; nf += m; tail += m; done += m;
		__syncwarp();                                                   // this window's labels before the next window's loads
	}
	if(done && lane == 0) {
		F.v0 = v0; F.v1 = v1; F.v2 = v2;
		F.vcount = vcount; F.nfaces = nf; F.tail = tail; F.amax = amax;
		F.fflush = nf; F.pflush = vcount;
	}
	return done;
}

constexpr int LF_BUDGET = 160;          // symbols per leader chunk / log words per follower batch
constexpr uint32_t LF_STAGE = 512;      // staged faces / predictions (>= 3*LF_BUDGET)
constexpr uint32_t LF_LOG = 2048;       // log ring words
constexpr uint32_t LF_SPIN = 1u << 26;  // bound on every wait loop

__global__ void __launch_bounds__(64) k_clers_lf(DevBatch B, const uint32_t *mesh_order, uint32_t nwork, ClersScratch scratch, uint32_t *ticket,
                                                  uint32_t RB, uint32_t RA, bool vecmode, bool only_deferred) {
	__shared__ uint32_t ctl[8];          // 0 head, 1 tail, 2 done, 3 abort, 4 mesh, 5 lead rc
	__shared__ uint32_t chain[33];       // prev-chain of a VERTEX/LEFT window (lead_vector)
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	SmemRings4 rg;
	const uint32_t oB = 0, oX = oB + RB*8u, oL = oX + RB, oA = oL + LF_LOG*4u, oF = oA + RA*16u, oP = oF + LF_STAGE*16u;
	uint32_t sbase;
	asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"((uint32_t)__cvta_generic_to_shared(crt_smem)));
	rg.aB = sbase + oB; rg.aX = sbase + oX; rg.aL = sbase + oL; rg.aA = sbase + oA; rg.aF = sbase + oF; rg.aP = sbase + oP;
	rg.RM = RB - 1; rg.LM = LF_LOG - 1; rg.AM = RA - 1; rg.FM = LF_STAGE - 1; rg.PM = LF_STAGE - 1;
	const uint32_t WB = RB - 3u*LF_BUDGET, WA = RA - 3u*LF_BUDGET;
	for(;;) {
		__syncthreads();
		if(threadIdx.x == 0) { ctl[4] = atomicAdd(ticket, 1u); ctl[0] = 0; ctl[1] = 0; ctl[2] = 0; ctl[3] = 0; ctl[5] = 0; }
		__syncthreads();
		const uint32_t w = ctl[4];
		if(w >= nwork) break;
		const uint32_t mi = mesh_order[w];
		if(only_deferred) {                    // second pass behind k_clers_cta: only the meshes it left (uniform: one word per mesh)
			if(!(B.regular[mi] >> 31)) continue;
			__syncthreads();
			if(threadIdx.x == 0) B.regular[mi] = 0;
		}
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
		io.fl = (uint8_t *)io.order;           // v4 stores no FIFO: the `order` scratch backs the flag ring
		if(warp == 0) {
			// ------------------------------------------------ leader ------------------------------------------------
			LeadState S;
			lead_init(S, io);
			LeadPre pre{0xffu, 0xffffffffu};
			int rc = 0;
			for(;;) {
				if(lane == 0) {      // wait for log space (the follower is at most one ring behind)
					uint32_t spins = 0;
					while(S.nlog + 3u*LF_BUDGET + 8u - ld_vol_shared(&ctl[1]) > LF_LOG) {
						if(ld_vol_shared(&ctl[3]) || ++spins > LF_SPIN) { rc = -5; break; }
						__nanosleep(64);
					}
				}
				rc = __shfl_sync(0xffffffffu, rc, 0);
				// one chunk: scalar machine, interleaved with warp-wide windows over VERTEX/LEFT runs
				uint32_t left = LF_BUDGET;
				bool tried = false;                                // the last window attempt bailed: one symbol goes the scalar way
				if(vecmode && rc == 0) left -= lead_vector(io, rg, S, left, chain, lane, pre);   // a chunk usually starts in the middle of a run
				while(rc == 0 && left > 0) {
					uint32_t c0 = 0;
					if(lane == 0) { c0 = S.cler; rc = clers_lead(io, rg, S, tried ? 1 : (int)left, vecmode && !tried); c0 = S.cler - c0; }
					rc = __shfl_sync(0xffffffffu, rc, 0);
					c0 = __shfl_sync(0xffffffffu, c0, 0);
					left = left > c0 ? left - c0 : 0;
					tried = false;
					if(rc == 3) {                                  // the scalar machine saw a VERTEX/LEFT run coming
						rc = 0;
						const uint32_t m = left >= 2 ? lead_vector(io, rg, S, left, chain, lane, pre) : 0;
						if(m) left -= m; else tried = true;
					} else if(rc == 4) {                           // it needs the next queued edge
						rc = 0;
						lead_pop_vector(io, rg, S, lane);
						if(!__shfl_sync(0xffffffffu, S.have, 0)) tried = true;   // queue exhausted: scalar path (delayed stack / start triangle)
					}
				}
				if(lane == 0) { __threadfence_block(); st_vol_shared(&ctl[0], S.nlog); }
				const uint32_t e0 = __shfl_sync(0xffffffffu, S.eflush, 0), nf = __shfl_sync(0xffffffffu, S.nfront, 0);
				const uint32_t e1 = nf > WB ? nf - WB : 0u;
				if(e1 > e0) for(uint32_t id = e0 + lane; id < e1; id += 32) {
					const uint2 l = ((const uint2 *)(crt_smem + oB))[id & rg.RM];
					io.eb[id] = EdgeB{l.x, l.y}; io.fl[id] = (crt_smem + oX)[id & rg.RM];
				}
				__syncwarp();
				if(lane == 0 && e1 > e0) S.eflush = e1;
				if(rc != 0) break;
			}
			if(lane == 0) { st_vol_shared(&ctl[5], (uint32_t)rc); __threadfence_block(); st_vol_shared(&ctl[2], 1u); }
		} else {
			// ------------------------------------------------ follower ----------------------------------------------
			FollowState F;
			follow_init(F);
			int splitbits;
			asm volatile("mov.u32 %0, %1;" : "=r"(splitbits) : "r"(ilog2_u32(io.nvert) + 1));
			int rc = 0;
			for(;;) {
				uint32_t head = 0, done = 0;
				if(lane == 0) {
					uint32_t spins = 0;
					for(;;) {
						done = ld_vol_shared(&ctl[2]);
						head = ld_vol_shared(&ctl[0]);
						if(head != F.tail || done) break;
						if(++spins > LF_SPIN) { rc = -5; break; }
						__nanosleep(64);
					}
					__threadfence_block();
				}
				rc = __shfl_sync(0xffffffffu, rc, 0);
				head = __shfl_sync(0xffffffffu, head, 0); done = __shfl_sync(0xffffffffu, done, 0);
				{   // one batch of at most LF_BUDGET log words: scalar machine interleaved with warp-wide windows
					const uint32_t t0 = __shfl_sync(0xffffffffu, F.tail, 0);
					const uint32_t upto = min(head, t0 + (uint32_t)LF_BUDGET);
					bool tried = false;
					if(vecmode && rc == 0) follow_vector(io, rg, F, upto, lane, oF, oP);   // a batch usually starts in the middle of a run
					while(rc == 0 && __shfl_sync(0xffffffffu, F.tail, 0) < upto) {
						if(lane == 0) rc = clers_follow(io, rg, F, tried ? F.tail + 1 : upto, LF_STAGE, splitbits, vecmode && !tried);
						rc = __shfl_sync(0xffffffffu, rc, 0);
						tried = false;
						if(rc == 3) { rc = 0; if(follow_vector(io, rg, F, upto, lane, oF, oP) == 0) tried = true; }
					}
				}
				// ---- drains (all lanes) ----
				const uint32_t f0 = __shfl_sync(0xffffffffu, F.fflush, 0), f1 = __shfl_sync(0xffffffffu, F.nfaces, 0);
				const uint32_t p0 = __shfl_sync(0xffffffffu, F.pflush, 0), p1 = __shfl_sync(0xffffffffu, F.vcount, 0);
				const uint32_t a0 = __shfl_sync(0xffffffffu, F.aflush, 0), am = __shfl_sync(0xffffffffu, F.amax, 0);
				const uint32_t tl = __shfl_sync(0xffffffffu, F.tail, 0);
				if(f0 != f1 || p0 != p1) follow_drain(io, rg, oF, oP, f0, f1, p0, p1, lane);
				uint32_t a1 = am > WA ? am - WA : 0u;
				if(rc == 2) a1 = 0;
				if(a1 > a0) for(uint32_t id = a0 + lane; id < a1; id += 32) { const uint4 a = ((const uint4 *)(crt_smem + oA))[id & rg.AM]; io.ea[id] = EdgeA{a.x, a.y, a.z, 0}; }
				__syncwarp();
				if(lane == 0) {
					F.fflush = f1; F.pflush = p1;
					if(rc == 2) { F.nfaces = F.fflush = F.gstart; F.aflush = 0; F.amax = 0; rc = 0; }   // group restart: replay the G word
					else if(a1 > a0) F.aflush = a1;
					__threadfence_block();
					st_vol_shared(&ctl[1], F.tail);
				}
				rc = __shfl_sync(0xffffffffu, rc, 0);
				if(rc < 0) { if(lane == 0) st_vol_shared(&ctl[3], 1u); break; }
				if(done && head == tl) break;
			}
			const uint32_t lead_rc = ld_vol_shared(&ctl[5]);
			uint32_t vcount = __shfl_sync(0xffffffffu, F.vcount, 0);
			const bool bad = rc < 0 || (int)lead_rc < 0;
			if(lane == 0) { if(bad) B.status[mi] = -5; B.vertex_count[mi] = vcount; }
			uint4 *pred = (uint4 *)M->pred_ptr;
			if(bad) vcount = 0;
			for(uint32_t v = vcount + lane; v < M->nvert; v += 32) pred[v] = make_uint4(0, 0, 0, 0);
		}
	}
}

// =========================================================================================================
// K5  mesh delta inverse.  v[i] += v[a] + v[b] - v[c] (parallelogram) or v[i] += v[a], with a,b,c < i chosen by the
//     topology: a recurrence on a DAG.  A warp takes 32 consecutive vertices per round, lane = vertex:
//       1. every lane loads its prediction and residual, and gathers the operands that lie OUTSIDE the block from
//          global memory — 32 vertices' worth of independent loads in flight at once (the serial v1 kernel paid one L2
//          round trip per vertex);
//       2. operands INSIDE the block are resolved in registers: when the block is a pure chain (a = i-1, b and c outside:
//          99.7 % of a grid) the recurrence is an inclusive warp scan; otherwise a 32-step shuffle loop
//          reproduces the sequential order exactly (a not-yet-processed operand is still its residual, as in the
//          in-place reference loop, vertex_attribute.h:165-176).
// =========================================================================================================
template <typename T, int NC>
__device__ __forceinline__ void delta_mesh_rounds(T *v, const uint32_t stride, const uint4 *pred, uint32_t nvert, bool par, int lane) {
	// Two-deep software pipeline over rounds of 32 vertices:
	//   * prediction + residuals are fetched TWO rounds ahead (they are streams nobody writes before their round);
	//   * the operand gather of round r+1 is issued BEFORE round r is resolved, for every operand that is already final
	//     (vertex < base_r) or a still-untouched residual (vertex beyond round r+1: hostile streams only); operands that land
	//     inside round r are taken from round r's registers with a shuffle once it is done.
	// So the dependent chain per round is the in-register resolution only; the L2 round trip of the gather overlaps it.
	const uint32_t FULL = 0xffffffffu;
	auto ldp = [&](uint32_t i) { return i < nvert ? pred[i] : make_uint4(0, 0, 0, 0); };
	uint4 pA = ldp((uint32_t)lane), pB = ldp(32u + lane);          // round r, r+1
	uint32_t xA[NC], xB[NC], xprev[NC];
#pragma unroll
	for(int k = 0; k < NC; k++) {
		xA[k] = (uint32_t)lane < nvert ? (uint32_t)v[(size_t)lane*stride + k] : 0u;
		xB[k] = 32u + lane < nvert ? (uint32_t)v[(size_t)(32u + lane)*stride + k] : 0u;
		xprev[k] = 0;
	}
	// gathered operands of the current round + which of them wait for the previous round's registers
	uint32_t fa[NC], fb[NC], fc[NC];
	uint32_t pend = 0;                                               // bit0 a, bit1 b, bit2 c: operand lies in the previous round
	{
		// round 0 has no earlier vertices: operands are inside the block or (hostile) beyond it = residuals
		const uint32_t i = lane;
		const bool act = i < nvert && i > 0;
#pragma unroll
		for(int k = 0; k < NC; k++) {
			fa[k] = (act && pA.x >= 32u && pA.x < nvert) ? (uint32_t)v[(size_t)pA.x*stride + k] : 0u;
			fb[k] = (act && par && pA.y >= 32u && pA.y < nvert) ? (uint32_t)v[(size_t)pA.y*stride + k] : 0u;
			fc[k] = (act && par && pA.z >= 32u && pA.z < nvert) ? (uint32_t)v[(size_t)pA.z*stride + k] : 0u;
		}
	}
	for(uint32_t base = 0; base < nvert; base += 32) {
		const uint32_t i = base + lane;
		const bool in = i < nvert;
		const bool act = in && i > 0;                 // vertex 0 keeps its residual (loops start at 1)
		const uint32_t a = pA.x, b = pA.y, c = pA.z;
		// ---- operands that were inside the previous round: its final values are still in registers ----
		{
			const uint32_t la = (a - (base - 32u)) & 31u, lb = (b - (base - 32u)) & 31u, lc = (c - (base - 32u)) & 31u;
#pragma unroll
			for(int k = 0; k < NC; k++) {
				const uint32_t va = __shfl_sync(FULL, xprev[k], la), vb = __shfl_sync(FULL, xprev[k], lb), vc = __shfl_sync(FULL, xprev[k], lc);
				if(pend & 1u) fa[k] = va;
				if(pend & 2u) fb[k] = vb;
				if(pend & 4u) fc[k] = vc;
			}
		}
		// ---- issue the gather of the NEXT round (everything that does not depend on this round) + prefetch two rounds ahead ----
		uint32_t ga[NC], gb[NC], gc[NC], pend_n = 0;
		{
			const uint32_t nb = base + 32u, ni = nb + lane;
			const bool nact = ni < nvert;                               // ni > 0 always
			const uint32_t na = pB.x, nbb = pB.y, nc = pB.z;
			auto far = [&](uint32_t X, bool use, uint32_t bit, uint32_t (&g)[NC]) {
				const bool in_this = use && X - base < 32u;              // resolved from this round's registers next iteration
				const bool in_next = use && X - nb < 32u;                // inside its own round
				if(in_this) pend_n |= bit;
#pragma unroll
				for(int k = 0; k < NC; k++) g[k] = (use && !in_this && !in_next && X < nvert) ? (uint32_t)v[(size_t)X*stride + k] : 0u;
			};
			far(na, nact, 1u, ga); far(nbb, nact && par, 2u, gb); far(nc, nact && par, 4u, gc);
		}
		const uint32_t i2 = i + 64u;
		const uint4 pC = ldp(i2);
		uint32_t xC[NC];
#pragma unroll
		for(int k = 0; k < NC; k++) xC[k] = i2 < nvert ? (uint32_t)v[(size_t)i2*stride + k] : 0u;
		{
			const uint32_t farv = i + 32u*10u;                          // pull the streams into L2 well ahead
			if(farv < nvert) {
				asm volatile("prefetch.global.L2 [%0];" :: "l"(pred + farv));
				asm volatile("prefetch.global.L2 [%0];" :: "l"(v + (size_t)farv*stride));
			}
		}
		// ---- resolve this round ----
		uint32_t x[NC];
#pragma unroll
		for(int k = 0; k < NC; k++) x[k] = xA[k];
		const bool a_in = act && (a - base) < 32u, b_in = act && par && (b - base) < 32u, c_in = act && par && (c - base) < 32u;
		// fast path: b and c outside the block, a anywhere EARLIER in the block (or outside): x_i = r_i + x_parent(i) is a forest
		// whose parents have lower lane numbers -> pointer doubling, 5 shuffle rounds (a chain a = i-1, 99.7 % of a grid, and
		// the strip starts in between are both covered)
		const bool tree_ok = !act || (!b_in && !c_in && (!a_in || a < i));
		if(__all_sync(FULL, tree_ok)) {
			uint32_t parent = a_in ? (a - base) : 0xffffffffu;     // lane of the in-block parent, or none
			uint32_t r[NC];
#pragma unroll
			for(int k = 0; k < NC; k++) r[k] = act ? x[k] + (a_in ? 0u : fa[k]) + fb[k] - fc[k] : x[k];
#pragma unroll
			for(int d = 0; d < 5; d++) {
				const uint32_t src = parent == 0xffffffffu ? (uint32_t)lane : parent;
				const uint32_t pp = __shfl_sync(FULL, parent, src);
#pragma unroll
				for(int k = 0; k < NC; k++) { const uint32_t pr = __shfl_sync(FULL, r[k], src); if(parent != 0xffffffffu) r[k] += pr; }
				if(parent != 0xffffffffu) parent = pp;
			}
#pragma unroll
			for(int k = 0; k < NC; k++) x[k] = r[k];
		} else {
			// general case (operands b / c inside the block, common on irregular meshes): lanes are finalised in order; lane j's
			// final value is pushed to every later lane that names it.  A lane is final once all lower lanes were pushed, and an
			// operand naming a HIGHER lane (hostile stream) reads that lane's residual, exactly like the in-place reference loop.
			const uint32_t la = a_in ? a - base : 64u, lb = b_in ? b - base : 64u, lc = c_in ? c - base : 64u;
			uint32_t acc[NC], res[NC];
#pragma unroll
			for(int k = 0; k < NC; k++) { res[k] = x[k]; acc[k] = act ? x[k] + (a_in ? 0u : fa[k]) + (b_in ? 0u : fb[k]) - (c_in ? 0u : fc[k]) : x[k]; }
#pragma unroll
			for(int k = 0; k < NC; k++) {
				const uint32_t ra = __shfl_sync(FULL, res[k], la & 31u), rb = __shfl_sync(FULL, res[k], lb & 31u), rc = __shfl_sync(FULL, res[k], lc & 31u);
				if(la < 32u && la >= (uint32_t)lane) acc[k] += ra;
				if(lb < 32u && lb >= (uint32_t)lane) acc[k] += rb;
				if(lc < 32u && lc >= (uint32_t)lane) acc[k] -= rc;
			}
			const uint32_t named = __reduce_or_sync(FULL, (la < (uint32_t)lane ? 1u << la : 0u) | (lb < (uint32_t)lane ? 1u << lb : 0u) | (lc < (uint32_t)lane ? 1u << lc : 0u));
			for(uint32_t todo = named; todo; todo &= todo - 1) {
				const int j = __ffs(todo) - 1;             // every lane below j that anyone names has been pushed already, so acc of lane j is final
#pragma unroll
				for(int k = 0; k < NC; k++) {
					const uint32_t xj = __shfl_sync(FULL, acc[k], j);
					if(la == (uint32_t)j) acc[k] += xj;
					if(lb == (uint32_t)j) acc[k] += xj;
					if(lc == (uint32_t)j) acc[k] -= xj;
				}
			}
#pragma unroll
			for(int k = 0; k < NC; k++) x[k] = acc[k];
		}
		if(act) {
#pragma unroll
			for(int k = 0; k < NC; k++) v[(size_t)i*stride + k] = (T)x[k];
		}
		__syncwarp();
		// ---- rotate the pipeline ----
		pA = pB; pB = pC; pend = pend_n;
#pragma unroll
		for(int k = 0; k < NC; k++) { xprev[k] = x[k]; xA[k] = xB[k]; xB[k] = xC[k]; fa[k] = ga[k]; fb[k] = gb[k]; fc[k] = gc[k]; }
	}
}

// ---- block-wide rounds (experiment switch CORTO_DELTA=cta; not the default, see launch_delta_mesh) -----------------------------
// The warp kernel above walks a (mesh, attribute) chain 32 vertices at a time, ~220 dependent instructions per round, 4096
// rounds for a 128 K-vertex mesh whatever the batch size.  Here one CTA of 256 threads takes 256 vertices per round,
// thread = vertex:
//   * loads (prediction, residual, operands that are final in global memory) are issued one round ahead; operands inside the
//     PREVIOUS round are read from its finals in shared memory (two buffers);
//   * operands inside the round: the round is cut into SUB-BLOCKS [s, e) such that no vertex of a sub-block has a `b` / `c`
//     operand inside it (e = the first vertex naming something at or after s).  Inside a sub-block only `a` links remain, a
//     forest with parents at lower indices: log2(e - s) pointer-doubling steps through shared memory; `b` / `c` (and `a`) that
//     point into earlier sub-blocks read finals from the round's buffer.  A strip of a regular mesh is one sub-block (the
//     previous row lies a whole strip back), so a round costs ~12-20 barrier-separated steps instead of 8 x 220 dependent
//     instructions — but each such step (shared-memory write, barrier, dependent shared-memory read) measures ~400-500 cycles
//     with the next round's loads in flight, which is why this is not faster than the warp kernel yet (next: in-warp shuffle
//     doubling for the first five steps, a second round of prefetch);
//   * a sub-block shorter than 4 vertices (irregular meshes: the parallelogram names the vertices created just before) or an
//     operand naming its own or a later vertex (hostile streams) sends the rest of the round to the 8 warps one after the
//     other with the warp algorithm (warp_resolve), which reads the finals of the vertices before it — and the residuals of
//     the vertices after it, which is what the in-place reference loop would read — from the same buffer.
// Same in-place semantics as vertex_attribute.h:165-176 / normal_attribute.cpp:193-201 for every input.
template <int NC>
__device__ __forceinline__ void warp_resolve(uint32_t (&x)[NC], const uint32_t (&resid)[NC], const uint32_t (&fa)[NC], const uint32_t (&fb)[NC], const uint32_t (&fc)[NC],
                                             bool act, bool a_in, bool b_in, bool c_in, uint32_t wa, uint32_t wb, uint32_t wc, uint32_t lane) {
	// x: in  = residual + the contributions from outside the round (act lanes) / the final value (lanes that are done);  out = final
	// resid: the bare residual (what an operand naming this or a later lane reads);  fa / fb / fc: contributions of the in-round
	// operands that are NOT inside this warp's 32 vertices (0 when inside or unused)
	const uint32_t FULL = 0xffffffffu, NONE = 0xffffffffu;
	const bool tree_ok = !act || (!b_in && !c_in && (!a_in || wa < lane));
	if(__all_sync(FULL, tree_ok)) {
		uint32_t parent = (act && a_in) ? wa : NONE;
		uint32_t r[NC];
#pragma unroll
		for(int k = 0; k < NC; k++) r[k] = act ? x[k] + fa[k] + fb[k] - fc[k] : x[k];
#pragma unroll
		for(int d = 0; d < 5; d++) {
			const uint32_t src = parent == NONE ? lane : parent;
			const uint32_t pp = __shfl_sync(FULL, parent, src);
#pragma unroll
			for(int k = 0; k < NC; k++) { const uint32_t pr = __shfl_sync(FULL, r[k], src); if(parent != NONE) r[k] += pr; }
			if(parent != NONE) parent = pp;
		}
#pragma unroll
		for(int k = 0; k < NC; k++) x[k] = r[k];
	} else {
		const uint32_t la = (act && a_in) ? wa : 64u, lb = (act && b_in) ? wb : 64u, lc = (act && c_in) ? wc : 64u;
		uint32_t acc[NC], res[NC];
#pragma unroll
		for(int k = 0; k < NC; k++) { res[k] = resid[k]; acc[k] = act ? x[k] + fa[k] + fb[k] - fc[k] : x[k]; }
#pragma unroll
		for(int k = 0; k < NC; k++) {
			const uint32_t qa = __shfl_sync(FULL, res[k], la & 31u), qb = __shfl_sync(FULL, res[k], lb & 31u), qc = __shfl_sync(FULL, res[k], lc & 31u);
			if(la < 32u && la >= lane) acc[k] += qa;       // an operand naming this or a HIGHER lane reads that lane's residual
			if(lb < 32u && lb >= lane) acc[k] += qb;
			if(lc < 32u && lc >= lane) acc[k] -= qc;
		}
		const uint32_t named = __reduce_or_sync(FULL, (la < lane ? 1u << la : 0u) | (lb < lane ? 1u << lb : 0u) | (lc < lane ? 1u << lc : 0u));
		for(uint32_t todo = named; todo; todo &= todo - 1) {
			const uint32_t j = (uint32_t)__ffs(todo) - 1u;     // every lane below j that anyone names has been pushed already: acc of lane j is final
#pragma unroll
			for(int k = 0; k < NC; k++) {
				const uint32_t xj = __shfl_sync(FULL, acc[k], j);
				if(la == j) acc[k] += xj;
				if(lb == j) acc[k] += xj;
				if(lc == j) acc[k] -= xj;
			}
		}
#pragma unroll
		for(int k = 0; k < NC; k++) x[k] = acc[k];
	}
}

constexpr uint32_t DB = 256;           // vertices per block-wide round
constexpr uint32_t DB_MINSUB = 4;      // a shorter sub-block sends the rest of the round to the sequential warps
template <typename T, int NC>
__device__ __forceinline__ void delta_mesh_cta(T *v, const uint4 *pred, const uint32_t nvert, const bool par, const bool force_seq,
                                               uint32_t *s_fin /*[2][DB][NC]*/, uint32_t *s_r /*[2][DB][NC]*/, uint32_t *s_par /*[2][DB]*/, uint32_t *s_min /*[8]*/) {
	const uint32_t NONE = 0xffffffffu, FULL = 0xffffffffu;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const uint32_t nblk = (nvert + DB - 1u)/DB;
	// registers of the round being resolved (C) and of the next one (N)
	uint4 pC = make_uint4(0, 0, 0, 0), pN;
	uint32_t xC[NC], gaC[NC], gbC[NC], gcC[NC], xN[NC], gaN[NC], gbN[NC], gcN[NC];
	auto load_round = [&](uint32_t r, uint4 &p, uint32_t (&x)[NC], uint32_t (&ga)[NC], uint32_t (&gb)[NC], uint32_t (&gc)[NC]) {
		const uint32_t base = r*DB, i = base + tid;
		const bool in = r < nblk && i < nvert, act = in && i > 0;
		p = in ? pred[i] : make_uint4(0, 0, 0, 0);
		auto far = [&](uint32_t X, bool use, uint32_t (&g)[NC]) {
			// final in global memory (rounds < r-1), or a residual nobody has touched yet (rounds > r: hostile streams only)
			const bool prevblk = r > 0 && X - (base - DB) < DB, inblk = X - base < DB;
#pragma unroll
			for(int k = 0; k < NC; k++) g[k] = (use && !prevblk && !inblk && X < nvert) ? (uint32_t)v[(size_t)X*NC + k] : 0u;
		};
#pragma unroll
		for(int k = 0; k < NC; k++) x[k] = in ? (uint32_t)v[(size_t)i*NC + k] : 0u;
		far(p.x, act, ga); far(p.y, act && par, gb); far(p.z, act && par, gc);
	};
	load_round(0, pC, xC, gaC, gbC, gcC);
	uint32_t q = 0;                                        // pointer-doubling step counter (buffer parity), never reset
	for(uint32_t r = 0; r < nblk; r++) {
		const uint32_t cur = r & 1u, prv = cur ^ 1u;
		const uint32_t base = r*DB, i = base + tid;
		const bool in = i < nvert, act = in && i > 0;       // vertex 0 keeps its residual (loops start at 1)
		load_round(r + 1, pN, xN, gaN, gbN, gcN);             // in flight while this round resolves
		const uint32_t a = pC.x, b = pC.y, c = pC.z;
		const bool usebc = act && par;
		const bool a_prev = act && r > 0 && a - (base - DB) < DB, b_prev = usebc && r > 0 && b - (base - DB) < DB, c_prev = usebc && r > 0 && c - (base - DB) < DB;
		const bool a_in = act && a - base < DB, b_in = usebc && b - base < DB, c_in = usebc && c - base < DB;
		const uint32_t la = (a - base) & (DB - 1u), lb = (b - base) & (DB - 1u), lc = (c - base) & (DB - 1u);
		uint32_t *fin_c = s_fin + (size_t)cur*DB*NC, *fin_p = s_fin + (size_t)prv*DB*NC;
		// x: residual + every contribution from outside the round
		uint32_t x[NC];
#pragma unroll
		for(int k = 0; k < NC; k++) {
			const uint32_t fa = a_prev ? fin_p[((a - (base - DB)) & (DB - 1u))*NC + k] : gaC[k];      // (0 when inside the round or unused)
			const uint32_t fb = b_prev ? fin_p[((b - (base - DB)) & (DB - 1u))*NC + k] : gbC[k];
			const uint32_t fc = c_prev ? fin_p[((c - (base - DB)) & (DB - 1u))*NC + k] : gcC[k];
			fin_c[tid*NC + k] = xC[k];                        // until it is final a vertex shows its residual, as in the in-place loop
			x[k] = act ? xC[k] + fa + fb - fc : xC[k];
		}
		// an operand naming its own or a later vertex of the round: only the sequential order gives the reference's answer
		const bool hostile = (a_in && la >= tid) || (b_in && lb >= tid) || (c_in && lc >= tid);
		uint32_t s = __syncthreads_or(hostile || force_seq) ? 0u : DB + 1u;      // DB + 1: no fallback requested (yet)
		if(s > DB) {
			const int mbc = max(b_in ? (int)lb : -1, c_in ? (int)lc : -1);       // the last in-round vertex my b / c name
			uint32_t parent0 = a_in ? la : NONE;
			uint32_t s0 = 0;
			while(s0 < DB) {
				// e = first vertex >= s0 that names (b / c) something at or after s0
				const uint32_t cand = __ballot_sync(FULL, tid >= s0 && mbc >= (int)s0);
				if(lane == 0) s_min[warp] = cand ? warp*32u + (uint32_t)__ffs(cand) - 1u : DB;
				__syncthreads();
				uint32_t e = DB;
#pragma unroll
				for(int w8 = 0; w8 < (int)(DB/32u); w8++) e = min(e, s_min[w8]);
				if(e - s0 < DB_MINSUB && DB - s0 > 32u) { s = s0; break; }           // irregular stretch: sequential warps from s0 on
				const bool mine = tid >= s0 && tid < e;
				uint32_t parent = NONE;
				if(mine) {
#pragma unroll
					for(int k = 0; k < NC; k++) {
						if(b_in) x[k] += fin_c[lb*NC + k];            // earlier sub-blocks: final
						if(c_in) x[k] -= fin_c[lc*NC + k];
						if(a_in && la < s0) x[k] += fin_c[la*NC + k];
					}
					if(a_in && la >= s0) parent = parent0;
				}
				for(uint32_t span = 1; span < e - s0; span <<= 1, q++) {
					uint32_t *rr = s_r + (size_t)(q & 1u)*DB*NC, *pp = s_par + (q & 1u)*DB;
					if(mine) {
						pp[tid] = parent;
#pragma unroll
						for(int k = 0; k < NC; k++) rr[tid*NC + k] = x[k];
					}
					__syncthreads();
					if(parent != NONE) {
#pragma unroll
						for(int k = 0; k < NC; k++) x[k] += rr[parent*NC + k];
						parent = pp[parent];
					}
				}
				if(mine) {
#pragma unroll
					for(int k = 0; k < NC; k++) fin_c[tid*NC + k] = x[k];
				}
				__syncthreads();
				s0 = e;
			}
		}
		if(s <= DB) {
			// sequential warps over [s, DB): vertices below s are final (their x is the final value and they take no further part)
#pragma unroll 1
			for(uint32_t ws = s >> 5; ws < DB/32u; ws++) {
				if(warp == ws) {
					const bool todo = act && tid >= s;
					const bool wa_in = a_in && (la >> 5) == ws, wb_in = b_in && (lb >> 5) == ws, wc_in = c_in && (lc >> 5) == ws;
					uint32_t ga2[NC], gb2[NC], gc2[NC];
#pragma unroll
					for(int k = 0; k < NC; k++) {           // in the round but outside this warp: final (before) or residual (after)
						ga2[k] = (a_in && !wa_in) ? fin_c[la*NC + k] : 0u;
						gb2[k] = (b_in && !wb_in) ? fin_c[lb*NC + k] : 0u;
						gc2[k] = (c_in && !wc_in) ? fin_c[lc*NC + k] : 0u;
					}
					warp_resolve<NC>(x, xC, ga2, gb2, gc2, todo, wa_in, wb_in, wc_in, la & 31u, lb & 31u, lc & 31u, lane);
#pragma unroll
					for(int k = 0; k < NC; k++) fin_c[tid*NC + k] = x[k];
				}
				__syncthreads();
			}
		}
		if(act) {
#pragma unroll
			for(int k = 0; k < NC; k++) v[(size_t)i*NC + k] = (T)x[k];
		}
		__syncthreads();                                   // finals of this round (shared and global) before the next round reads them
		pC = pN;
#pragma unroll
		for(int k = 0; k < NC; k++) { xC[k] = xN[k]; gaC[k] = gaN[k]; gbC[k] = gbN[k]; gcC[k] = gcN[k]; }
	}
}

__global__ void __launch_bounds__(256) k_delta_mesh_cta(DevBatch B, const uint2 *work, uint32_t nwork, bool force_seq) {
	__shared__ uint32_t s_fin[2*DB*MAX_COMP], s_r[2*DB*MAX_COMP], s_par[2*DB], s_min[DB/32];
	const uint32_t w = blockIdx.x;
	if(w >= nwork) return;
	const MeshDesc *M = B.mesh + work[w].x;
	const AttrDesc *A = &M->attr[work[w].y & 0xffu];
	if(B.status[work[w].x]) return;
	const uint32_t nvert = M->nvert;
	const uint4 *pred = (const uint4 *)M->pred_ptr;
	const bool par = (A->strategy & S_PARALLEL) && A->codec != CODEC_NORMAL;    // normals: d += d[a] only (normal_attribute.cpp:193-201)
	if(A->codec == CODEC_COLOR) {
		uint8_t *v = (uint8_t *)A->work_ptr;
		switch(A->ncomp) {
		case 1: delta_mesh_cta<uint8_t, 1>(v, pred, nvert, par, force_seq, s_fin, s_r, s_par, s_min); break;
		case 2: delta_mesh_cta<uint8_t, 2>(v, pred, nvert, par, force_seq, s_fin, s_r, s_par, s_min); break;
		case 3: delta_mesh_cta<uint8_t, 3>(v, pred, nvert, par, force_seq, s_fin, s_r, s_par, s_min); break;
		default: delta_mesh_cta<uint8_t, 4>(v, pred, nvert, par, force_seq, s_fin, s_r, s_par, s_min); break;
		}
	} else {
		uint32_t *v = (uint32_t *)(A->codec == CODEC_NORMAL ? A->work_ptr : A->out_ptr);
		switch(A->ncomp) {
		case 1: delta_mesh_cta<uint32_t, 1>(v, pred, nvert, par, force_seq, s_fin, s_r, s_par, s_min); break;
		case 2: delta_mesh_cta<uint32_t, 2>(v, pred, nvert, par, force_seq, s_fin, s_r, s_par, s_min); break;
		case 3: delta_mesh_cta<uint32_t, 3>(v, pred, nvert, par, force_seq, s_fin, s_r, s_par, s_min); break;
		default: delta_mesh_cta<uint32_t, 4>(v, pred, nvert, par, force_seq, s_fin, s_r, s_par, s_min); break;
		}
	}
}

// ---- block-wide rounds, segmented scans (the default for meshes whose CLERS stream is mostly long VERTEX / LEFT runs) ----------
// Same skeleton as delta_mesh_cta (256 vertices per round, loads one round ahead, finals of the current / previous round in
// shared memory), different in-round resolution.  On a regular mesh the parallelogram of vertex i is (i-1, two vertices of the
// strip before): inside a round only the link a = i-1 is in-round, so the round is a SEGMENTED PREFIX SUM of
// residual + outside operands, cut wherever a vertex names something else inside the round (strip starts: 1-2 cuts per round).
//   * segment [s0, e): e = the first vertex whose in-round operands other than the link i-1 reach back to s0 or later (one
//     ballot + one barrier); those operands then lie in earlier segments and are final in shared memory;
//   * inside the segment: warp shuffle scan (5 steps, no barrier) + one shared-memory exchange of the warp totals;
//   * a segment shorter than 4 vertices (irregular stretches) or a hostile operand sends the rest of the round to the
//     sequential warps (warp_resolve), exactly as in delta_mesh_cta.
// Two barriers per segment instead of log2(length) pointer-doubling steps with a barrier each.
template <typename T, int NC>
__device__ __forceinline__ void delta_mesh_seg(T *v, const uint4 *pred, const uint32_t nvert, const bool par,
                                               uint32_t *s_fin /*[2][DB][NC]*/, uint32_t (*s_tot)[MAX_COMP] /*[8]*/, uint32_t *s_open /*[8]*/, uint32_t (*s_min)[8] /*[2]*/) {
	const uint32_t FULL = 0xffffffffu;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const uint32_t nblk = (nvert + DB - 1u)/DB;
	uint4 pC = make_uint4(0, 0, 0, 0), pN;
	uint32_t xC[NC], gaC[NC], gbC[NC], gcC[NC], xN[NC], gaN[NC], gbN[NC], gcN[NC];
	auto load_round = [&](uint32_t r, uint4 &p, uint32_t (&x)[NC], uint32_t (&ga)[NC], uint32_t (&gb)[NC], uint32_t (&gc)[NC]) {
		const uint32_t base = r*DB, i = base + tid;
		const bool in = r < nblk && i < nvert, act = in && i > 0;
		p = in ? pred[i] : make_uint4(0, 0, 0, 0);
		auto far = [&](uint32_t X, bool use, uint32_t (&g)[NC]) {
			// final in global memory (rounds < r-1), or a residual nobody has touched yet (rounds > r: hostile streams only)
			const bool prevblk = r > 0 && X - (base - DB) < DB, inblk = X - base < DB;
#pragma unroll
			for(int k = 0; k < NC; k++) g[k] = (use && !prevblk && !inblk && X < nvert) ? (uint32_t)v[(size_t)X*NC + k] : 0u;
		};
#pragma unroll
		for(int k = 0; k < NC; k++) x[k] = in ? (uint32_t)v[(size_t)i*NC + k] : 0u;
		far(p.x, act, ga); far(p.y, act && par, gb); far(p.z, act && par, gc);
	};
	load_round(0, pC, xC, gaC, gbC, gcC);
	uint32_t it = 0;
	for(uint32_t r = 0; r < nblk; r++) {
		const uint32_t cur = r & 1u, prv = cur ^ 1u;
		const uint32_t base = r*DB, i = base + tid;
		const bool in = i < nvert, act = in && i > 0;       // vertex 0 keeps its residual (loops start at 1)
		load_round(r + 1, pN, xN, gaN, gbN, gcN);             // in flight while this round resolves
		const uint32_t a = pC.x, b = pC.y, c = pC.z;
		const bool usebc = act && par;
		const bool a_prev = act && r > 0 && a - (base - DB) < DB, b_prev = usebc && r > 0 && b - (base - DB) < DB, c_prev = usebc && r > 0 && c - (base - DB) < DB;
		const bool a_in = act && a - base < DB, b_in = usebc && b - base < DB, c_in = usebc && c - base < DB;
		const uint32_t la = (a - base) & (DB - 1u), lb = (b - base) & (DB - 1u), lc = (c - base) & (DB - 1u);
		uint32_t *fin_c = s_fin + (size_t)cur*DB*NC, *fin_p = s_fin + (size_t)prv*DB*NC;
		uint32_t x[NC];
#pragma unroll
		for(int k = 0; k < NC; k++) {
			const uint32_t fa = a_prev ? fin_p[((a - (base - DB)) & (DB - 1u))*NC + k] : gaC[k];      // (0 when inside the round or unused)
			const uint32_t fb = b_prev ? fin_p[((b - (base - DB)) & (DB - 1u))*NC + k] : gbC[k];
			const uint32_t fc = c_prev ? fin_p[((c - (base - DB)) & (DB - 1u))*NC + k] : gcC[k];
			fin_c[tid*NC + k] = xC[k];                        // until it is final a vertex shows its residual, as in the in-place loop
			x[k] = act ? xC[k] + fa + fb - fc : xC[k];
		}
		const bool hostile = (a_in && la >= tid) || (b_in && lb >= tid) || (c_in && lc >= tid);
		const bool chain = a_in && la + 1u == tid;          // the link to the vertex just before me
		int mdep = max(b_in ? (int)lb : -1, c_in ? (int)lc : -1);
		if(a_in && !chain) mdep = max(mdep, (int)la);
		uint32_t s = __syncthreads_or(hostile) ? 0u : DB + 1u;      // DB + 1: no fallback requested (yet)
		if(s > DB) {
			uint32_t s0 = 0;
			while(s0 < DB) {
				// e = first vertex after s0 that names (other than through its link) something at or after s0
				const uint32_t cand = __ballot_sync(FULL, tid > s0 && mdep >= (int)s0);
				if(lane == 0) s_min[it][warp] = cand ? warp*32u + (uint32_t)__ffs(cand) - 1u : DB;
				__syncthreads();                               // (also: the finals of the previous segment are visible)
				uint32_t e = DB;
#pragma unroll
				for(int w8 = 0; w8 < (int)(DB/32u); w8++) e = min(e, s_min[it][w8]);
				it ^= 1u;
				if(e - s0 < DB_MINSUB && DB - s0 > 32u) { s = s0; break; }           // irregular stretch: sequential warps from s0 on
				const bool mine = tid >= s0 && tid < e;
				const bool link = mine && chain && tid > s0;
				if(mine) {
#pragma unroll
					for(int k = 0; k < NC; k++) {
						if(b_in) x[k] += fin_c[lb*NC + k];            // earlier segments: final
						if(c_in) x[k] -= fin_c[lc*NC + k];
						if(a_in && !link) x[k] += fin_c[la*NC + k];
					}
				}
				// segmented inclusive scan: a lane adds lane - d while everything in between links
				const uint32_t linkmask = __ballot_sync(FULL, link);
				const uint32_t heads_le = ~linkmask & (lane == 31u ? FULL : ((2u << lane) - 1u));
				const uint32_t hp = heads_le ? 31u - (uint32_t)__clz(heads_le) : 0u;
#pragma unroll
				for(int d = 1; d < 32; d <<= 1) {
#pragma unroll
					for(int k = 0; k < NC; k++) { const uint32_t t = __shfl_up_sync(FULL, x[k], d); if(lane >= (uint32_t)d && lane - (uint32_t)d >= hp) x[k] += t; }
				}
				if(lane == 31u) {
#pragma unroll
					for(int k = 0; k < NC; k++) s_tot[warp][k] = x[k];
				}
				if(lane == 0) s_open[warp] = linkmask == FULL ? 1u : 0u;
				__syncthreads();
				if(!heads_le && warp > 0) {                    // no head at or below me in this warp: continue the previous warp's last lane
					uint32_t carry[NC];
#pragma unroll
					for(int k = 0; k < NC; k++) carry[k] = 0;
					for(int w2 = (int)warp - 1; w2 >= 0; w2--) {
#pragma unroll
						for(int k = 0; k < NC; k++) carry[k] += s_tot[w2][k];
						if(!s_open[w2]) break;
					}
#pragma unroll
					for(int k = 0; k < NC; k++) x[k] += carry[k];
				}
				if(mine) {
#pragma unroll
					for(int k = 0; k < NC; k++) fin_c[tid*NC + k] = x[k];
				}
				s0 = e;
			}
			__syncthreads();                                   // the last segment's finals before anybody reads them
		}
		if(s <= DB) {
			// sequential warps over [s, DB): vertices below s are final (their x is the final value and they take no further part)
#pragma unroll 1
			for(uint32_t ws = s >> 5; ws < DB/32u; ws++) {
				if(warp == ws) {
					const bool todo = act && tid >= s;
					const bool wa_in = a_in && (la >> 5) == ws, wb_in = b_in && (lb >> 5) == ws, wc_in = c_in && (lc >> 5) == ws;
					uint32_t ga2[NC], gb2[NC], gc2[NC];
#pragma unroll
					for(int k = 0; k < NC; k++) {           // in the round but outside this warp: final (before) or residual (after)
						ga2[k] = (a_in && !wa_in) ? fin_c[la*NC + k] : 0u;
						gb2[k] = (b_in && !wb_in) ? fin_c[lb*NC + k] : 0u;
						gc2[k] = (c_in && !wc_in) ? fin_c[lc*NC + k] : 0u;
					}
					warp_resolve<NC>(x, xC, ga2, gb2, gc2, todo, wa_in, wb_in, wc_in, la & 31u, lb & 31u, lc & 31u, lane);
#pragma unroll
					for(int k = 0; k < NC; k++) fin_c[tid*NC + k] = x[k];
				}
				__syncthreads();
			}
		}
		if(act) {
#pragma unroll
			for(int k = 0; k < NC; k++) v[(size_t)i*NC + k] = (T)x[k];
		}
		__syncthreads();                                   // finals of this round (shared and global) before the next round reads them
		pC = pN;
#pragma unroll
		for(int k = 0; k < NC; k++) { xC[k] = xN[k]; gaC[k] = gaN[k]; gbC[k] = gbN[k]; gcC[k] = gcN[k]; }
	}
}

// One CTA per (mesh, attribute).  Meshes the CLERS kernel found regular (most symbols in long runs) take the segmented-scan
// rounds; the others fall back to the warp algorithm (delta_mesh_rounds) on warp 0, or on one warp per component when the
// batch is small (`split`), the remaining warps of the CTA exit at once.
__global__ void __launch_bounds__(256) k_delta_mesh_seg(DevBatch B, const uint2 *work, uint32_t nwork, bool split) {
	__shared__ uint32_t s_fin[2*DB*MAX_COMP], s_tot[DB/32][MAX_COMP], s_open[DB/32], s_min[2][8];
	const uint32_t w = blockIdx.x;
	if(w >= nwork) return;
	const MeshDesc *M = B.mesh + work[w].x;
	const AttrDesc *A = &M->attr[work[w].y & 0xffu];
	if(B.status[work[w].x]) return;
	const uint32_t nvert = M->nvert;
	const uint4 *pred = (const uint4 *)M->pred_ptr;
	const bool par = (A->strategy & S_PARALLEL) && A->codec != CODEC_NORMAL;    // normals: d += d[a] only (normal_attribute.cpp:193-201)
	const int nc = A->ncomp;
	if(!B.regular[work[w].x]) {
		const uint32_t warp = threadIdx.x >> 5;
		const int lane = threadIdx.x & 31;
		if(split) {
			if((int)warp >= nc) return;
			if(A->codec == CODEC_COLOR) delta_mesh_rounds<uint8_t, 1>((uint8_t *)A->work_ptr + warp, (uint32_t)nc, pred, nvert, par, lane);
			else delta_mesh_rounds<uint32_t, 1>((uint32_t *)(A->codec == CODEC_NORMAL ? A->work_ptr : A->out_ptr) + warp, (uint32_t)nc, pred, nvert, par, lane);
			return;
		}
		if(warp) return;
		if(A->codec == CODEC_COLOR) {
			uint8_t *v = (uint8_t *)A->work_ptr;
			switch(nc) {
			case 1: delta_mesh_rounds<uint8_t, 1>(v, 1u, pred, nvert, par, lane); break;
			case 2: delta_mesh_rounds<uint8_t, 2>(v, 2u, pred, nvert, par, lane); break;
			case 3: delta_mesh_rounds<uint8_t, 3>(v, 3u, pred, nvert, par, lane); break;
			default: delta_mesh_rounds<uint8_t, 4>(v, 4u, pred, nvert, par, lane); break;
			}
		} else {
			uint32_t *v = (uint32_t *)(A->codec == CODEC_NORMAL ? A->work_ptr : A->out_ptr);
			switch(nc) {
			case 1: delta_mesh_rounds<uint32_t, 1>(v, 1u, pred, nvert, par, lane); break;
			case 2: delta_mesh_rounds<uint32_t, 2>(v, 2u, pred, nvert, par, lane); break;
			case 3: delta_mesh_rounds<uint32_t, 3>(v, 3u, pred, nvert, par, lane); break;
			default: delta_mesh_rounds<uint32_t, 4>(v, 4u, pred, nvert, par, lane); break;
			}
		}
		return;
	}
	if(A->codec == CODEC_COLOR) {
		uint8_t *v = (uint8_t *)A->work_ptr;
		switch(nc) {
		case 1: delta_mesh_seg<uint8_t, 1>(v, pred, nvert, par, s_fin, s_tot, s_open, s_min); break;
		case 2: delta_mesh_seg<uint8_t, 2>(v, pred, nvert, par, s_fin, s_tot, s_open, s_min); break;
		case 3: delta_mesh_seg<uint8_t, 3>(v, pred, nvert, par, s_fin, s_tot, s_open, s_min); break;
		default: delta_mesh_seg<uint8_t, 4>(v, pred, nvert, par, s_fin, s_tot, s_open, s_min); break;
		}
	} else {
		uint32_t *v = (uint32_t *)(A->codec == CODEC_NORMAL ? A->work_ptr : A->out_ptr);
		switch(nc) {
		case 1: delta_mesh_seg<uint32_t, 1>(v, pred, nvert, par, s_fin, s_tot, s_open, s_min); break;
		case 2: delta_mesh_seg<uint32_t, 2>(v, pred, nvert, par, s_fin, s_tot, s_open, s_min); break;
		case 3: delta_mesh_seg<uint32_t, 3>(v, pred, nvert, par, s_fin, s_tot, s_open, s_min); break;
		default: delta_mesh_seg<uint32_t, 4>(v, pred, nvert, par, s_fin, s_tot, s_open, s_min); break;
		}
	}
}

// The default: one warp per chain.
__global__ void __launch_bounds__(32) k_delta_mesh(DevBatch B, const uint2 *work, uint32_t nwork) {
	const uint32_t w = blockIdx.x;
	if(w >= nwork) return;
	const int lane = threadIdx.x;
	const MeshDesc *M = B.mesh + work[w].x;
	const AttrDesc *A = &M->attr[work[w].y & 0xffu];
	if(B.status[work[w].x]) return;
	const uint32_t nvert = M->nvert;
	const uint4 *pred = (const uint4 *)M->pred_ptr;
	const int nc = A->ncomp;
	const bool par = (A->strategy & S_PARALLEL) && A->codec != CODEC_NORMAL;    // normals: d += d[a] only (normal_attribute.cpp:193-201)
	const uint32_t comp = (work[w].y >> 8) & 0xffu;
	if(comp != 0xffu) {
		// One warp per COMPONENT (the host splits when the batch has few chains): the recurrence is component-wise, so the nc
		// components of an attribute are independent chains; a warp walks one of them (interleaved values, stride nc).
		if(A->codec == CODEC_COLOR) delta_mesh_rounds<uint8_t, 1>((uint8_t *)A->work_ptr + comp, (uint32_t)nc, pred, nvert, par, lane);
		else delta_mesh_rounds<uint32_t, 1>((uint32_t *)(A->codec == CODEC_NORMAL ? A->work_ptr : A->out_ptr) + comp, (uint32_t)nc, pred, nvert, par, lane);
		return;
	}
	if(A->codec == CODEC_COLOR) {
		uint8_t *v = (uint8_t *)A->work_ptr;
		switch(nc) {
		case 1: delta_mesh_rounds<uint8_t, 1>(v, 1u, pred, nvert, par, lane); break;
		case 2: delta_mesh_rounds<uint8_t, 2>(v, 2u, pred, nvert, par, lane); break;
		case 3: delta_mesh_rounds<uint8_t, 3>(v, 3u, pred, nvert, par, lane); break;
		default: delta_mesh_rounds<uint8_t, 4>(v, 4u, pred, nvert, par, lane); break;
		}
	} else {
		uint32_t *v = (uint32_t *)(A->codec == CODEC_NORMAL ? A->work_ptr : A->out_ptr);
		switch(nc) {
		case 1: delta_mesh_rounds<uint32_t, 1>(v, 1u, pred, nvert, par, lane); break;
		case 2: delta_mesh_rounds<uint32_t, 2>(v, 2u, pred, nvert, par, lane); break;
		case 3: delta_mesh_rounds<uint32_t, 3>(v, 3u, pred, nvert, par, lane); break;
		default: delta_mesh_rounds<uint32_t, 4>(v, 4u, pred, nvert, par, lane); break;
		}
	}
}

// =========================================================================================================
// K6  normal estimation (ESTIMATED / BORDER).  The reference adds face normals to their three vertices in FACE
//     ORDER in fp32 (normal_attribute.cpp:44-55); fp32 addition is not associative, so instead of float atomics
//     each vertex gathers its incident faces and adds them in ascending face index: bit-identical to the sequential
//     loop for every input, no exactness guard needed.  Adjacency = 8 slots per vertex filled with one atomic per face
//     corner (one pass over the faces); a vertex of higher valence chains its further faces through an overflow list
//     (per-vertex linked lists: ohead[v] -> ovf[idx] = (face, next), so a high-valence vertex reads only its own entries).
//     zeroed scratch per mesh (u32): cnt[nvert] | ohead[nvert] | novf | (BORDER only) bnd[nvert] | cidx[nvert+1];
//     plain scratch: adj8[8*nvert] | ovf[3*nface] (uint2)
// =========================================================================================================
struct AdjView { uint32_t *cnt, *bnd, *cidx, *novf, *ohead, *adj8; uint2 *ovf; float4 *fn; };
__device__ __forceinline__ AdjView adj_view(const MeshDesc *M) {
	AdjView c;
	uint32_t *p = (uint32_t *)M->csr_ptr;
	c.cnt = p; p += M->nvert;
	c.ohead = p; p += M->nvert;
	c.novf = p; p += 1;
	c.bnd = p; p += M->nvert;                        // (present for BORDER meshes only: crt_api.cu sizes the region by the prediction)
	c.cidx = p;
	c.fn = (float4 *)M->adj_ptr;
	c.adj8 = (uint32_t *)(c.fn + M->nface);
	c.ovf = (uint2 *)(c.adj8 + (size_t)M->nvert*8);
	return c;
}
__device__ __forceinline__ void load_face(const MeshDesc *M, uint32_t f, uint32_t &a, uint32_t &b, uint32_t &c) {
	if(M->index16) { const uint16_t *x = (const uint16_t *)M->face_ptr + (size_t)f*3; a = x[0]; b = x[1]; c = x[2]; }
	else { const uint32_t *x = (const uint32_t *)M->face_ptr + (size_t)f*3; a = x[0]; b = x[1]; c = x[2]; }
}

// tiles: a = mesh, tile = block of SCAN_TILE faces.  One thread per face:
//   * its unnormalised normal (the cross product estimateNormals adds to each of its three vertices, normal_attribute.cpp:44-55)
//     is computed ONCE here from the integer positions and parked in scratch; k_normal_estimate then gathers one float4 per
//     incident face instead of three indices and nine coordinates — this kernel waits on atomics anyway;
//   * slot allocation in the per-vertex adjacency: one atomicAdd per face corner.  AGG (CORTO_ADJ_AGG=1) merges the lanes of a
//     warp that name the same vertex in the same corner (match.any) into one atomic per group; it halves the atomics of a
//     strip but the match + shuffle chain costs more than it saves on B200 (62 -> 92 us on 16 meshes), so it is off by default.
// PART: 0 both halves in one pass; 1 the adjacency only (needs the faces, not the positions: it can run BESIDE the delta inverse);
//       2 the face normals only (needs the delta-decoded positions).
template <bool AGG, int PART>
__global__ void __launch_bounds__(256) k_adj_build(DevBatch B, const Tile *tiles, uint32_t ntiles) {
	const Tile tl = tiles[blockIdx.x];
	const MeshDesc *M = B.mesh + tl.a;
	if(B.status[tl.a]) return;
	const AdjView C = adj_view(M);
	const bool border = M->attr[M->normal_attr].prediction == N_BORDER;
	const int32_t *P = (const int32_t *)M->attr[M->position_attr].out_ptr;   // still integer (decoder.cpp:191-195)
	const uint32_t FULL = 0xffffffffu, lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
	const uint32_t fend = min(M->nface, (tl.tile + 1)*SCAN_TILE);
	for(uint32_t f0 = tl.tile*SCAN_TILE; f0 < fend; f0 += 256) {          // uniform trip count: the warp collectives below need every lane
		const uint32_t f = f0 + threadIdx.x;
		uint32_t v[3] = {0, 0, 0};
		bool valid = f < fend;
		if(valid) {
			load_face(M, f, v[0], v[1], v[2]);
			valid = v[0] < M->nvert && v[1] < M->nvert && v[2] < M->nvert;
		}
		if(valid && PART != 1) {
			const float v0x = i2f(P[(size_t)v[0]*3]), v0y = i2f(P[(size_t)v[0]*3 + 1]), v0z = i2f(P[(size_t)v[0]*3 + 2]);
			const float ax = f_sub(i2f(P[(size_t)v[1]*3]), v0x), ay = f_sub(i2f(P[(size_t)v[1]*3 + 1]), v0y), az = f_sub(i2f(P[(size_t)v[1]*3 + 2]), v0z);
			const float bx = f_sub(i2f(P[(size_t)v[2]*3]), v0x), by = f_sub(i2f(P[(size_t)v[2]*3 + 1]), v0y), bz = f_sub(i2f(P[(size_t)v[2]*3 + 2]), v0z);
			C.fn[f] = make_float4(f_sub(f_mul(ay, bz), f_mul(az, by)), f_sub(f_mul(az, bx), f_mul(ax, bz)), f_sub(f_mul(ax, by), f_mul(ay, bx)), 0.f);   // point.h:113-115
		}
		if constexpr(PART != 2) {
#pragma unroll
		for(int k = 0; k < 3; k++) {
			uint32_t s = 0;
			if constexpr(AGG) {
				const uint32_t grp = __match_any_sync(FULL, valid ? v[k] : 0xffffffffu - lane);   // invalid lanes: singletons
				const uint32_t leader = (uint32_t)__ffs(grp) - 1u;
				uint32_t base = 0;
				if(valid && lane == leader) base = atomicAdd(C.cnt + v[k], (uint32_t)__popc(grp));
				s = __shfl_sync(FULL, base, leader) + (uint32_t)__popc(grp & below);
			} else if(valid) s = atomicAdd(C.cnt + v[k], 1u);
			if(valid) {
				if(s < 8) C.adj8[(size_t)v[k]*8 + s] = f;
				else {                                     // valence > 8: push onto the vertex's own overflow chain (0 ends a chain)
					const uint32_t idx = atomicAdd(C.novf, 1u);
					C.ovf[idx] = make_uint2(f, atomicExch(C.ohead + v[k], idx + 1u));
				}
			}
		}
		if(border && valid) { atomicXor(C.bnd + v[0], v[1] ^ v[2]); atomicXor(C.bnd + v[1], v[2] ^ v[0]); atomicXor(C.bnd + v[2], v[0] ^ v[1]); }   // markBoundary :24-37
		}
	}
}

// exclusive scan of the boundary flags (bnd != 0) per mesh over nvert+1 elements -> cidx: the running `count` of
// computeNormals (normal_attribute.cpp:284-293) that indexes the BORDER diffs.  tiles cover nvert+1 elements.
__global__ void __launch_bounds__(256) k_scan_u32(DevBatch B, const Tile *tiles, uint32_t ntiles, uint64_t *states, uint32_t *ticket) {
	__shared__ uint32_t s_warp[9];
	__shared__ uint32_t s_tile;
	__shared__ uint64_t s_base;
	const int tid = threadIdx.x;
	for(;;) {
		NEXT_TILE(ticket, ntiles, s_tile)
		const Tile tl = tiles[tile_id];
		const MeshDesc *M = B.mesh + tl.a;
		const AdjView C = adj_view(M);
		const uint32_t n = M->nvert;
		const uint32_t i0 = tl.tile*SCAN_TILE + tid*4u;
		uint32_t x[4];
#pragma unroll
		for(int j = 0; j < 4; j++) { const uint32_t i = i0 + j; x[j] = (i < n && C.bnd[i] != 0u) ? 1u : 0u; }
		const uint32_t s1 = x[0], s2 = s1 + x[1], s3 = s2 + x[2], s4 = s3 + x[3];
		uint32_t total;
		uint32_t off = cta_scan_excl_256(s4, s_warp, &total);
		if(tid == 0) s_base = lookback(states, tile_id, tl.first != 0, total);
		__syncthreads();
		const uint32_t base = (uint32_t)s_base + off;
		const uint32_t e[4] = { base, base + s1, base + s2, base + s3 };
#pragma unroll
		for(int j = 0; j < 4; j++) { const uint32_t i = i0 + j; if(i <= n) C.cidx[i] = e[j]; }
	}
}

// Valence > 8: the 8 slots + this vertex's overflow chain.  Up to 72 incident faces are copied to a local list and added by
// repeated "smallest face id greater than the last one" (a face naming the vertex twice is added twice, like the reference's
// corner loop); a fan pole beyond that walks the mesh's faces in order instead — O(nface) for that one vertex, but never
// quadratic in the valence.
__device__ __noinline__ void high_valence_normal(const MeshDesc *M, const AdjView &C, uint32_t i, uint32_t deg, float &ex, float &ey, float &ez) {
	auto add_face = [&](uint32_t f) { const float4 n = C.fn[f]; ex = f_add(ex, n.x); ey = f_add(ey, n.y); ez = f_add(ez, n.z); };
	constexpr uint32_t LOCAL = 72;
	if(deg <= LOCAL) {
		uint32_t fl[LOCAL];
		uint32_t n = 0;
		for(uint32_t s = 0; s < 8; s++) fl[n++] = C.adj8[(size_t)i*8 + s];
		const uint32_t novf = *C.novf;
		for(uint32_t at = C.ohead[i], guard = 0; at && at <= novf && n < deg && guard < LOCAL; guard++) { const uint2 e = C.ovf[at - 1u]; fl[n++] = e.x; at = e.y; }
		int64_t last = -1;
		uint32_t done = 0;
		while(done < n) {
			uint32_t fmin = 0xffffffffu, mult = 0;
			for(uint32_t s = 0; s < n; s++) {
				const uint32_t f = fl[s];
				if((int64_t)f > last) { if(f < fmin) { fmin = f; mult = 1; } else if(f == fmin) mult++; }
			}
			if(mult == 0) break;
			for(uint32_t m = 0; m < mult; m++) add_face(fmin);
			done += mult;
			last = (int64_t)fmin;
		}
	} else {
		for(uint32_t f = 0; f < M->nface; f++) {
			uint32_t a, b, c;
			load_face(M, f, a, b, c);
			if(a >= M->nvert || b >= M->nvert || c >= M->nvert) continue;      // (k_adj_build skipped it too)
			if(a == i) add_face(f);
			if(b == i) add_face(f);
			if(c == i) add_face(f);
		}
	}
}

// one thread per vertex; tiles: a = mesh, tile = block of SCAN_TILE vertices.  Two launches of the same body: HV = false does every
// vertex of valence <= 8 (all of a regular mesh) and nothing else — no local list, no call in its code; HV = true, right behind
// it, does the others (its CTAs leave at once when the mesh has no overflow entry).
template <bool HV>
__global__ void __launch_bounds__(256) k_normal_estimate(DevBatch B, const Tile *tiles, uint32_t ntiles) {
	const Tile tl = tiles[blockIdx.x];
	const MeshDesc *M = B.mesh + tl.a;
	if(B.status[tl.a]) return;
	const AdjView C = adj_view(M);
	if(HV && *C.novf == 0u) return;
	const AttrDesc *A = &M->attr[M->normal_attr];
	const int32_t *diffs = (const int32_t *)A->work_ptr;
	const int unit = f2i_x86(A->q);
	const bool border = A->prediction == N_BORDER;
	for(uint32_t i = tl.tile*SCAN_TILE + threadIdx.x; i < min(M->nvert, (tl.tile + 1)*SCAN_TILE); i += 256) {
		const uint32_t deg = C.cnt[i];
		if((deg > 8) != HV) continue;
		float ex = 0.f, ey = 0.f, ez = 0.f;
		auto add_face = [&](uint32_t f) {                              // estimateNormals :44-55, one corner's worth (cross product: k_adj_build)
			const float4 n = C.fn[f];
			ex = f_add(ex, n.x); ey = f_add(ey, n.y); ez = f_add(ez, n.z);
		};
		if constexpr(!HV) {
			// the usual case: sort the (at most 8) incident faces in registers, add in ascending face order
			const uint4 q0 = *(const uint4 *)(C.adj8 + (size_t)i*8), q1 = *(const uint4 *)(C.adj8 + (size_t)i*8 + 4);
			uint32_t f0 = deg > 0 ? q0.x : 0xffffffffu, f1 = deg > 1 ? q0.y : 0xffffffffu, f2 = deg > 2 ? q0.z : 0xffffffffu, f3 = deg > 3 ? q0.w : 0xffffffffu;
			uint32_t f4 = deg > 4 ? q1.x : 0xffffffffu, f5 = deg > 5 ? q1.y : 0xffffffffu, f6 = deg > 6 ? q1.z : 0xffffffffu, f7 = deg > 7 ? q1.w : 0xffffffffu;
#define CRT_CE(a, b) { const uint32_t lo_ = min(a, b), hi_ = max(a, b); a = lo_; b = hi_; }
			CRT_CE(f0, f1) CRT_CE(f2, f3) CRT_CE(f4, f5) CRT_CE(f6, f7)
			CRT_CE(f0, f2) CRT_CE(f1, f3) CRT_CE(f4, f6) CRT_CE(f5, f7)
			CRT_CE(f1, f2) CRT_CE(f5, f6) CRT_CE(f0, f4) CRT_CE(f3, f7)
			CRT_CE(f1, f5) CRT_CE(f2, f6)
			CRT_CE(f1, f4) CRT_CE(f3, f6)
			CRT_CE(f2, f4) CRT_CE(f3, f5)
			CRT_CE(f3, f4)
#undef CRT_CE
			if(deg > 0) add_face(f0);
			if(deg > 1) add_face(f1);
			if(deg > 2) add_face(f2);
			if(deg > 3) add_face(f3);
			if(deg > 4) add_face(f4);
			if(deg > 5) add_face(f5);
			if(deg > 6) add_face(f6);
			if(deg > 7) add_face(f7);
		} else {
			(void)add_face;
			high_valence_normal(M, C, i, deg, ex, ey, ez);
		}
		if(!border || C.bnd[i] != 0u) {                               // computeNormals :288-293 / :315-319
			int32_t qx, qy;
			to_octa(ex, ey, ez, unit, qx, qy);
			const uint32_t k = border ? C.cidx[i] : i;
			int32_t dx = 0, dy = 0;
			if(k < A->count) { dx = diffs[(size_t)k*2]; dy = diffs[(size_t)k*2 + 1]; }
			int32_t sx = (int32_t)((uint32_t)qx + (uint32_t)dx), sy = (int32_t)((uint32_t)qy + (uint32_t)dy);
			float nx, ny, nz;
			if(A->out_format == F_FLOAT) {
				to_sphere(sx, sy, unit, nx, ny, nz);
				float *o = (float *)A->out_ptr + (size_t)i*3;
				o[0] = nx; o[1] = ny; o[2] = nz;
			} else {
				to_sphere((int32_t)(int16_t)sx, (int32_t)(int16_t)sy, unit, nx, ny, nz);    // Point2s truncation, :293
				int16_t *o = (int16_t *)A->out_ptr + (size_t)i*3;
				o[0] = f2s_x86(f_mul(nx, 32767.0f)); o[1] = f2s_x86(f_mul(ny, 32767.0f)); o[2] = f2s_x86(f_mul(nz, 32767.0f));
			}
		} else if(A->out_format == F_FLOAT) {                         // interior vertex, no correction :320-322
			const float len = f_norm3(ex, ey, ez);
			float *o = (float *)A->out_ptr + (size_t)i*3;
			o[0] = f_div(ex, len); o[1] = f_div(ey, len); o[2] = f_div(ez, len);
		} else {                                                      // :294-302; tiny normal leaves the output untouched (H10)
			float len = f_norm3(ex, ey, ez);
			if(!(len < 0.00001f)) {
				len = f_div(32767.0f, len);
				int16_t *o = (int16_t *)A->out_ptr + (size_t)i*3;
				o[0] = f2s_x86(f_mul(ex, len)); o[1] = f2s_x86(f_mul(ey, len)); o[2] = f2s_x86(f_mul(ez, len));
			}
		}
	}
}

// =========================================================================================================
// K3c  point clouds, fused: bit unpack -> running delta -> dequantise -> output in ONE pass; only logs and bits are read,
//      only final arrays are written (the unfused path moved ~60 B/vertex of int32 intermediates for a 12 B/vertex position
//      array).  Tile = 1024 vertices of one attribute, ALL its components.  Chains (decoupled look-back, 8 state words / tile):
//        slots 0..3  bit offset per log stream (CORRELATED / normals: one stream; decodeValues: one per component, whose base
//                    inside the shared BITS block is the bit total of the earlier streams, summed up by k_tun_decode)
//        slots 4..7  running sum per component (cloud delta v[i] += v[i-N]: vertex_attribute.h:177-181, normal_attribute.cpp:202-207)
// =========================================================================================================
template <int NC> __device__ __forceinline__ void cta_scan_multi(const uint32_t (&v)[NC], uint32_t (&excl)[NC], uint32_t (&total)[NC], uint32_t (*s_w)[9]) {
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc[NC];
#pragma unroll
	for(int k = 0; k < NC; k++) {
		inc[k] = v[k];
#pragma unroll
		for(int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc[k], d); if(lane >= d) inc[k] += o; }
		if(lane == 31) s_w[k][w] = inc[k];
	}
	__syncthreads();
	if(w == 0) {
#pragma unroll
		for(int k = 0; k < NC; k++) {
			const uint32_t x = lane < 8 ? s_w[k][lane] : 0;
			uint32_t xi = x;
#pragma unroll
			for(int d = 1; d < 8; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, xi, d); if(lane >= d) xi += o; }
			if(lane < 8) s_w[k][lane] = xi - x;
			if(lane == 7) s_w[k][8] = xi;
		}
	}
	__syncthreads();
#pragma unroll
	for(int k = 0; k < NC; k++) { excl[k] = s_w[k][w] + inc[k] - v[k]; total[k] = s_w[k][8]; }
	// no trailing barrier: every caller has a __syncthreads between these reads of s_w and its next use of the scan
}

template <int NC, bool MESH>
__device__ __forceinline__ void cloud_tile(const DevBatch &B, const MeshDesc *M, const AttrDesc *A, const Tile &tl, uint32_t tile_id, uint64_t *states,
                                           uint32_t (*s_w)[9], uint64_t *s_base, uint8_t *s_out, uint64_t *s_carry = nullptr) {
	const int tid = threadIdx.x;
	const uint32_t nvert = M->nvert;
	const bool correlated = (A->codec == CODEC_NORMAL) || (A->codec == CODEC_GENERIC && (A->strategy & S_CORRELATED));
	const uint32_t i0 = tl.tile*1024u + tid*4u;
	const uint32_t *words = (const uint32_t *)(B.blobs + A->bits_off);
	const uint32_t nwords = A->bits_nwords;
	const bool first = tl.first != 0;
	// ---- logs and bit counts per stream ----
	uint32_t d[NC][4], bits[NC];
	uint64_t sbase[NC];                              // first bit of each stream inside the BITS block
	{
		uint64_t acc = 0;
#pragma unroll
		for(int k = 0; k < NC; k++) {
			if(correlated && k > 0) {
#pragma unroll
				for(int j = 0; j < 4; j++) d[k][j] = d[0][j];
				bits[k] = 0; sbase[k] = 0;
				continue;
			}
			const TunDesc td = B.tun[A->tun[k]];
			const uint8_t *logs = B.symbols + td.out_off;
			if(i0 + 3 < td.size) { const uchar4 q = *(const uchar4 *)(logs + i0); d[k][0] = q.x; d[k][1] = q.y; d[k][2] = q.z; d[k][3] = q.w; }
			else {
#pragma unroll
				for(int j = 0; j < 4; j++) d[k][j] = (i0 + j < td.size) ? logs[i0 + j] : 0;
			}
			bits[k] = d[k][0] + d[k][1] + d[k][2] + d[k][3];
			sbase[k] = acc;
			if(!correlated) acc += B.tun_bits[A->tun[k]];     // complete: k_tun_decode finished before this kernel started
		}
	}
	if(correlated) bits[0] *= (uint32_t)NC;
	uint32_t bexcl[NC], btot[NC];
	cta_scan_multi<NC>(bits, bexcl, btot, s_w);
	if(tid < (correlated ? 1 : NC)) {
		// bit offset of this tile inside each log stream: the running total when one CTA walks the whole chain (k_unpack_chain),
		// a decoupled look-back over the predecessors' totals when the tiles of a chain are spread over CTAs
		if(s_carry) { s_base[tid] = s_carry[tid]; s_carry[tid] += btot[tid]; }
		else s_base[tid] = lookback_strided(states, tile_id, 8, (uint32_t)tid, first, btot[tid]);
	}
	__syncthreads();
	// ---- unpack the residuals of my 4 vertices ----
	uint32_t r[NC][4];
	if(correlated) {
		uint64_t pos = s_base[0] + bexcl[0];
#pragma unroll
		for(int j = 0; j < 4; j++) {
			const int dd = (int)d[0][j], rd = dd > 32 ? 32 : dd;
			const uint32_t bias = array_bias(dd);
#pragma unroll
			for(int k = 0; k < NC; k++) {
				uint32_t v = 0;
				if(dd) { v = getbits(words, nwords, pos, rd) - bias; pos += (uint64_t)dd; }
				r[k][j] = v;
			}
		}
	} else {
#pragma unroll
		for(int k = 0; k < NC; k++) {
			uint64_t pos = sbase[k] + s_base[k] + bexcl[k];
#pragma unroll
			for(int j = 0; j < 4; j++) {
				const int dd = (int)d[k][j], rd = dd > 32 ? 32 : dd;
				uint32_t v = 0;
				if(dd) { v = (uint32_t)fold_value(getbits(words, nwords, pos, rd), dd); pos += (uint64_t)dd; }
				r[k][j] = v;
			}
		}
	}
	const uint32_t v_lo = tl.tile*1024u, v_hi = min(nvert, v_lo + 1024u);
	if constexpr(MESH) {
		// meshes: the residuals themselves are the product (the delta inverse needs the topology): int32 into the attribute's
		// output slice (generic; converted in place later, like the reference) or into scratch (normals int32, colours u8)
		const bool as_u8 = A->codec == CODEC_COLOR;
		const uint32_t stride = as_u8 ? (uint32_t)NC : 4u*NC;
		if(as_u8) {
			uint8_t *o = s_out + (size_t)tid*4*NC;
#pragma unroll
			for(int j = 0; j < 4; j++)
#pragma unroll
				for(int k = 0; k < NC; k++) o[j*NC + k] = (uint8_t)r[k][j];
		} else {
			uint32_t *o = (uint32_t *)s_out + (size_t)tid*4*NC;
#pragma unroll
			for(int j = 0; j < 4; j++)
#pragma unroll
				for(int k = 0; k < NC; k++) o[j*NC + k] = r[k][j];
		}
		__syncthreads();
		const uint32_t nbytes = (v_hi - v_lo)*stride;
		uint8_t *dst = (uint8_t *)((A->codec == CODEC_GENERIC) ? A->out_ptr : A->work_ptr) + (size_t)v_lo*stride;
		if((((uintptr_t)dst | nbytes) & 3u) == 0) {
			const uint32_t *src32 = (const uint32_t *)s_out; uint32_t *dst32 = (uint32_t *)dst;
			for(uint32_t i = tid; i < nbytes/4; i += 256) dst32[i] = src32[i];
		} else for(uint32_t i = tid; i < nbytes; i += 256) dst[i] = s_out[i];
		// (the next tile touches s_out / s_base only after the two barriers of NEXT_TILE)
	} else {
	__syncthreads();                                 // s_base is reused below
	// ---- running sum per component (wraps mod 2^32; colours are truncated to 8 bits at the end, which commutes) ----
	uint32_t tsum[NC], vexcl[NC], vtot[NC];
#pragma unroll
	for(int k = 0; k < NC; k++) {
#pragma unroll
		for(int j = 0; j < 4; j++) if(i0 + j >= nvert) r[k][j] = 0;
		r[k][1] += r[k][0]; r[k][2] += r[k][1]; r[k][3] += r[k][2];
		tsum[k] = r[k][3];
	}
	cta_scan_multi<NC>(tsum, vexcl, vtot, s_w);
	if(tid < NC) {
		if(s_carry) { s_base[tid] = s_carry[4 + tid]; s_carry[4 + tid] += vtot[tid]; }      // one CTA walks the chain: running sums
		else s_base[tid] = lookback_strided(states, tile_id, 8, 4u + (uint32_t)tid, first, vtot[tid]);
	}
	__syncthreads();
#pragma unroll
	for(int k = 0; k < NC; k++) {
		const uint32_t add = (uint32_t)s_base[k] + vexcl[k];
#pragma unroll
		for(int j = 0; j < 4; j++) r[k][j] += add;
	}
	__syncthreads();
	// ---- dequantise into shared memory, then one fully coalesced copy of the tile's slice (mesh slices start at arbitrary
	//      multiples of the vertex stride, so per-thread vector stores would be misaligned for most meshes) ----
	uint32_t stride;                                 // output bytes per vertex
	if(A->codec == CODEC_GENERIC) {
		stride = 4u*NC;
		uint32_t *o = (uint32_t *)s_out + (size_t)tid*4*NC;
		const float q = A->q;
#pragma unroll
		for(int j = 0; j < 4; j++)
#pragma unroll
			for(int k = 0; k < NC; k++)
				o[j*NC + k] = A->out_format == F_FLOAT ? __float_as_uint(f_mul(i2f((int32_t)r[k][j]), q)) : f2u_x86(f_mul(__uint2float_rn(r[k][j]), q));
	} else if(A->codec == CODEC_NORMAL) {
		stride = A->out_format == F_FLOAT ? 12u : 6u;
		if constexpr(NC == 2) {
			const int unit = f2i_x86(A->q);
#pragma unroll
			for(int j = 0; j < 4; j++) {
				float nx, ny, nz;
				if(A->out_format == F_FLOAT) {
					to_sphere((int32_t)r[0][j], (int32_t)r[1][j], unit, nx, ny, nz);
					float *o = (float *)s_out + (size_t)(tid*4 + j)*3;
					o[0] = nx; o[1] = ny; o[2] = nz;
				} else {
					to_sphere((int32_t)(int16_t)r[0][j], (int32_t)(int16_t)r[1][j], unit, nx, ny, nz);
					int16_t *o = (int16_t *)s_out + (size_t)(tid*4 + j)*3;
					o[0] = f2s_x86(f_mul(nx, 32767.0f)); o[1] = f2s_x86(f_mul(ny, 32767.0f)); o[2] = f2s_x86(f_mul(nz, 32767.0f));
				}
			}
		}
	} else {                                         // colour: YCC -> RGB, scale (color_attribute.cpp:76-95, point.h:214)
		const int oc = A->out_components;
		stride = (uint32_t)oc;
#pragma unroll
		for(int j = 0; j < 4; j++) {
			uint32_t c[4] = {0, 0, 0, 255};
#pragma unroll
			for(int k = 0; k < NC; k++) c[k] = r[k][j] & 255u;
			const uint32_t rgb[4] = { (c[2] + c[0]) & 255u, c[0], (c[1] + c[0]) & 255u, c[3] };
			uint8_t *o = s_out + (size_t)(tid*4 + j)*oc;
			for(int k = 0; k < oc; k++) o[k] = (uint8_t)(rgb[k]*(uint32_t)A->qc[k]);
		}
	}
	__syncthreads();
	{
		const uint32_t nbytes = (v_hi - v_lo)*stride;
		uint8_t *dst = (uint8_t *)A->out_ptr + (size_t)v_lo*stride;
		if((((uintptr_t)dst | nbytes) & 3u) == 0) {
			const uint32_t *src32 = (const uint32_t *)s_out; uint32_t *dst32 = (uint32_t *)dst;
			for(uint32_t i = tid; i < nbytes/4; i += 256) dst32[i] = src32[i];
		} else for(uint32_t i = tid; i < nbytes; i += 256) dst[i] = s_out[i];
	}
	}   // point clouds
}

// tiles: a = mesh, b = attr, tile = block of 1024 vertices, first = first tile of the attribute
template <bool MESH>
__global__ void __launch_bounds__(256) k_unpack_fused(DevBatch B, const Tile *tiles, uint32_t ntiles, uint64_t *states, uint32_t *ticket) {
	__shared__ uint32_t s_w[4][9];
	__shared__ uint32_t s_tile;
	__shared__ uint64_t s_base[4];
	__shared__ __align__(16) uint8_t s_out[1024*16];
	for(;;) {
		NEXT_TILE(ticket, ntiles, s_tile)
		const Tile tl = tiles[tile_id];
		const MeshDesc *M = B.mesh + tl.a;
		const AttrDesc *A = &M->attr[tl.b];
		switch(A->ncomp) {
		case 1: cloud_tile<1, MESH>(B, M, A, tl, tile_id, states, s_w, s_base, s_out); break;
		case 2: cloud_tile<2, MESH>(B, M, A, tl, tile_id, states, s_w, s_base, s_out); break;
		case 3: cloud_tile<3, MESH>(B, M, A, tl, tile_id, states, s_w, s_base, s_out); break;
		default: cloud_tile<4, MESH>(B, M, A, tl, tile_id, states, s_w, s_base, s_out); break;
		}
	}
}

// The same tiles, one CTA per CHAIN (all tiles of one attribute of one mesh, in order): the bit offsets are a running total in
// shared memory, so there is no ticket, no look-back and nobody spins on a predecessor (in k_unpack_fused 10.5 of every 24
// stall cycles are the other 255 threads waiting at the barrier behind the look-back thread).  Used when the batch has enough
// chains to fill the GPU; heads[c] .. heads[c+1] are the chain's tiles in the tile list.
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_unpack_chain(DevBatch B, const Tile *tiles, const uint32_t *heads, uint32_t nchains) {
	__shared__ uint32_t s_w[4][9];
	__shared__ uint64_t s_base[4], s_carry[4];
	__shared__ __align__(16) uint8_t s_out[1024*16];
	const uint32_t c = blockIdx.x;
	if(c >= nchains) return;
	if(threadIdx.x < 4) s_carry[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t t0 = heads[c], t1 = heads[c + 1];
	Tile tl = tiles[t0];
	const MeshDesc *M = B.mesh + tl.a;
	const AttrDesc *A = &M->attr[tl.b];
	const int nc = A->ncomp;
	for(uint32_t t = t0; t < t1; t++) {
		tl.tile = t - t0; tl.first = t == t0 ? 1u : 0u;
		switch(nc) {
		case 1: cloud_tile<1, true>(B, M, A, tl, t, nullptr, s_w, s_base, s_out, s_carry); break;
		case 2: cloud_tile<2, true>(B, M, A, tl, t, nullptr, s_w, s_base, s_out, s_carry); break;
		case 3: cloud_tile<3, true>(B, M, A, tl, t, nullptr, s_w, s_base, s_out, s_carry); break;
		default: cloud_tile<4, true>(B, M, A, tl, t, nullptr, s_w, s_base, s_out, s_carry); break;
		}
		__syncthreads();                             // s_w / s_base / s_out are reused by the next tile
	}
}

// Point clouds, one CTA per chain (all tiles of one attribute of one cloud, in order): both carries — the bit offset of every
// log stream and the running sum of every component — live in shared memory, so there is no look-back at all.  Used when the
// batch has enough chains to fill the GPU (launch_cloud_fused); the ticketed look-back kernel stays for a few large clouds.
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_cloud_chain(DevBatch B, const Tile *tiles, const uint32_t *heads, uint32_t nchains) {
	__shared__ uint32_t s_w[4][9];
	__shared__ uint64_t s_base[4], s_carry[8];
	__shared__ __align__(16) uint8_t s_out[1024*16];
	const uint32_t c = blockIdx.x;
	if(c >= nchains) return;
	if(threadIdx.x < 8) s_carry[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t t0 = heads[c], t1 = heads[c + 1];
	Tile tl = tiles[t0];
	const MeshDesc *M = B.mesh + tl.a;
	const AttrDesc *A = &M->attr[tl.b];
	const int nc = A->ncomp;
	for(uint32_t t = t0; t < t1; t++) {
		tl.tile = t - t0; tl.first = t == t0 ? 1u : 0u;
		switch(nc) {
		case 1: cloud_tile<1, false>(B, M, A, tl, t, nullptr, s_w, s_base, s_out, s_carry); break;
		case 2: cloud_tile<2, false>(B, M, A, tl, t, nullptr, s_w, s_base, s_out, s_carry); break;
		case 3: cloud_tile<3, false>(B, M, A, tl, t, nullptr, s_w, s_base, s_out, s_carry); break;
		default: cloud_tile<4, false>(B, M, A, tl, t, nullptr, s_w, s_base, s_out, s_carry); break;
		}
		__syncthreads();                             // s_w / s_base / s_out are reused by the next tile
	}
}

// =========================================================================================================
// K7  dequantise.  tiles: a = mesh, b = attr, tile = block of SCAN_TILE vertices
// =========================================================================================================
__global__ void __launch_bounds__(256) k_dequant(DevBatch B, const Tile *tiles, uint32_t ntiles) {
	const Tile tl = tiles[blockIdx.x];
	const MeshDesc *M = B.mesh + tl.a;
	if(B.status[tl.a]) return;
	const AttrDesc *A = &M->attr[tl.b];
	const uint32_t v0 = tl.tile*SCAN_TILE, v1 = min(M->nvert, v0 + SCAN_TILE);
	if(A->codec == CODEC_GENERIC) {
		const int nc = A->ncomp;
		const float q = A->q;
		if(A->out_format == F_FLOAT) {
			int32_t *v = (int32_t *)A->out_ptr;
			float *o = (float *)A->out_ptr;
			for(size_t i = (size_t)v0*nc + threadIdx.x; i < (size_t)v1*nc; i += 256) o[i] = f_mul(i2f(v[i]), q);      // vertex_attribute.h:190-193
		} else {
			uint32_t *v = (uint32_t *)A->out_ptr;
			for(size_t i = (size_t)v0*nc + threadIdx.x; i < (size_t)v1*nc; i += 256) v[i] = f2u_x86(f_mul(__uint2float_rn(v[i]), q));   // :200-203
		}
	} else if(A->codec == CODEC_NORMAL) {                             // DIFF only, normal_attribute.cpp:257-279
		const int32_t *d = (const int32_t *)A->work_ptr;
		const int unit = f2i_x86(A->q);
		for(uint32_t i = v0 + threadIdx.x; i < v1; i += 256) {
			float nx, ny, nz;
			if(A->out_format == F_FLOAT) {
				to_sphere(d[(size_t)i*2], d[(size_t)i*2 + 1], unit, nx, ny, nz);
				float *o = (float *)A->out_ptr + (size_t)i*3;
				o[0] = nx; o[1] = ny; o[2] = nz;
			} else {
				to_sphere((int32_t)(int16_t)d[(size_t)i*2], (int32_t)(int16_t)d[(size_t)i*2 + 1], unit, nx, ny, nz);
				int16_t *o = (int16_t *)A->out_ptr + (size_t)i*3;
				o[0] = f2s_x86(f_mul(nx, 32767.0f)); o[1] = f2s_x86(f_mul(ny, 32767.0f)); o[2] = f2s_x86(f_mul(nz, 32767.0f));
			}
		}
	} else {                                                          // colour: YCC -> RGB, scale (color_attribute.cpp:76-95, point.h:214)
		const int nc = A->ncomp, oc = A->out_components;
		const uint8_t *y = (const uint8_t *)A->work_ptr;
		uint8_t *o = (uint8_t *)A->out_ptr;
		for(uint32_t i = v0 + threadIdx.x; i < v1; i += 256) {
			uint32_t c[4] = {0, 0, 0, 255};
			for(int k = 0; k < nc; k++) c[k] = y[(size_t)i*nc + k];
			const uint32_t rgb[4] = { (c[2] + c[0]) & 255u, c[0], (c[1] + c[0]) & 255u, c[3] };
			for(int k = 0; k < oc; k++) o[(size_t)i*oc + k] = (uint8_t)(rgb[k]*(uint32_t)A->qc[k]);
		}
	}
}

// =========================================================================================================
// launchers
// =========================================================================================================
static inline uint32_t persistent_grid(uint32_t ntiles, int per_sm, int sms) {
	uint32_t g = (uint32_t)(per_sm*sms);
	return ntiles < g ? ntiles : g;
}

#define LAUNCH_CHECK() do { cudaError_t e__ = cudaGetLastError(); if(e__ != cudaSuccess) return (int)e__; } while(0)

int launch_tun_tables(const DevBatch &B, int ntun, cudaStream_t s) {
	if(ntun == 0) return 0;
	static int seq = -1;
	if(seq < 0) { const char *e = getenv("CORTO_TUN"); seq = (e && e[0] == 's') ? 1 : 0; }
	k_tun_tables<<<ntun, 32, 0, s>>>(B, seq != 0);
	LAUNCH_CHECK(); return 0;
}
int launch_tun_decode(const DevBatch &B, const Tile *tiles, uint32_t ntiles, uint64_t *states, uint32_t *ticket, int sms, cudaStream_t s) {
	if(ntiles == 0) return 0;
	k_tun_decode<<<persistent_grid(ntiles, 6, sms), 256, 0, s>>>(B, tiles, ntiles, states, ticket);
	LAUNCH_CHECK(); return 0;
}
int launch_clers(const DevBatch &B, const uint32_t *order, uint32_t nwork, const ClersScratch &scratch, uint32_t *ticket, int sms, cudaStream_t s) {
	if(nwork == 0) return 0;
	// CORTO_CLERS: 1 one warp per mesh (clers_run), 2 leader / follower warps, scalar, 3 leader / follower + 32-wide window steps,
	// 4: one CTA per mesh, CTA-wide window steps (k_clers_cta, crt_clers_cta.cu) for every mesh; default: the same for meshes whose
	// stream is regular, k_clers_lf for the others
	static int mode = -1;
	if(mode < 0) {
		const char *e = getenv("CORTO_CLERS");
		mode = (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : ((e && e[0] == '4') ? 5 : 4);
	}
	// default: k_clers_cta takes the meshes whose stream sample is regular and defers the others (B.regular bit 31) to k_clers_lf,
	// launched right behind it with its own ticket (ticket[1]); CORTO_CLERS=4 keeps everything on k_clers_cta
	const bool only_deferred = mode == 4 || mode == 5;
	if(only_deferred) {
		int rc = launch_clers_cta(B, order, nwork, scratch, ticket, sms, mode == 4, s);
		if(rc || mode == 5) return rc;
		ticket += 1;
	}
	const uint32_t g = nwork < scratch.slots ? nwork : scratch.slots;
	if(mode == 1) {
		uint32_t R = 4096, Q = 2048;
		if(nwork > (uint32_t)sms*2u) { R = 1024; Q = 1024; }
		const size_t smem = (size_t)R*24 + (size_t)Q*4 + 2*(size_t)CLERS_STAGE*16;
		cudaError_t e = cudaFuncSetAttribute(k_clers, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device: not cached
		if(e != cudaSuccess) return (int)e;
		k_clers<<<g, 32, smem, s>>>(B, order, nwork, scratch, ticket, R, Q);
	} else {
		// few meshes: big rings (2 CTAs per SM);  many meshes: small rings so that more serial chains share an SM
		uint32_t RB = 4096, RA = 2048;
		if(nwork > (uint32_t)sms*4u) { RB = 1024; RA = 1024; }     // really many meshes: more chains per SM beat bigger rings
		else if(nwork <= (uint32_t)sms) { RB = 16384; RA = 2048; }   // one mesh per SM at most: the whole shared memory for its rings
		const size_t smem = (size_t)RB*9 + (size_t)LF_LOG*4 + (size_t)RA*16 + 2*(size_t)LF_STAGE*16;
		cudaError_t e = cudaFuncSetAttribute(k_clers_lf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device: not cached
		if(e != cudaSuccess) return (int)e;
		k_clers_lf<<<g, 64, smem, s>>>(B, order, nwork, scratch, ticket, RB, RA, mode >= 3, only_deferred);
	}
	LAUNCH_CHECK(); return 0;
}
int launch_delta_mesh(const DevBatch &B, const uint2 *work, uint32_t nwork, bool split, cudaStream_t s) {
	if(nwork == 0) return 0;
	// default: one warp per chain (k_delta_mesh; the host may have split the work list per component).  CORTO_DELTA=cta: one CTA
	// per (mesh, attribute), 256 vertices per round (k_delta_mesh_cta) — measured 2.51 vs 2.40 ms on configs[1], 8.7 vs 6.5 ms on
	// configs[3], 110 vs 80 ms on 64 x tarta, so it is an experiment switch, not the default; CORTO_DELTA=seq: the same kernel with
	// every round on its sequential-warp path (tests).
	// CORTO_DELTA=warp: k_delta_mesh for every mesh; unset: k_delta_mesh_seg (segmented-scan rounds for regular meshes, the warp
	// algorithm for the others)
	static int mode = -1;
	if(mode < 0) { const char *e = getenv("CORTO_DELTA"); mode = (e && e[0] == 'c') ? 0 : ((e && e[0] == 's') ? 2 : ((e && e[0] == 'w') ? 1 : 3)); }
	if(mode == 3) k_delta_mesh_seg<<<nwork, 256, 0, s>>>(B, work, nwork, split);
	else if(mode == 1) k_delta_mesh<<<nwork, 32, 0, s>>>(B, work, nwork);
	else k_delta_mesh_cta<<<nwork, 256, 0, s>>>(B, work, nwork, mode == 2);
	LAUNCH_CHECK(); return 0;
}
int launch_adj_build(const DevBatch &B, const Tile *tiles, uint32_t ntiles, int part, cudaStream_t s) {
	if(ntiles == 0) return 0;
	static int agg = -1;
	if(agg < 0) { const char *e = getenv("CORTO_ADJ_AGG"); agg = (e && e[0] == '1') ? 1 : 0; }
	if(part == 1) { if(agg) k_adj_build<true, 1><<<ntiles, 256, 0, s>>>(B, tiles, ntiles); else k_adj_build<false, 1><<<ntiles, 256, 0, s>>>(B, tiles, ntiles); }
	else if(part == 2) k_adj_build<false, 2><<<ntiles, 256, 0, s>>>(B, tiles, ntiles);
	else if(agg) k_adj_build<true, 0><<<ntiles, 256, 0, s>>>(B, tiles, ntiles);
	else k_adj_build<false, 0><<<ntiles, 256, 0, s>>>(B, tiles, ntiles);
	LAUNCH_CHECK(); return 0;
}
int launch_scan_u32(const DevBatch &B, const Tile *tiles, uint32_t ntiles, uint64_t *states, uint32_t *ticket, int sms, cudaStream_t s) {
	if(ntiles == 0) return 0;
	k_scan_u32<<<persistent_grid(ntiles, 8, sms), 256, 0, s>>>(B, tiles, ntiles, states, ticket);
	LAUNCH_CHECK(); return 0;
}
int launch_normal_estimate(const DevBatch &B, const Tile *tiles, uint32_t ntiles, cudaStream_t s) {
	if(ntiles == 0) return 0;
	k_normal_estimate<false><<<ntiles, 256, 0, s>>>(B, tiles, ntiles);
	k_normal_estimate<true><<<ntiles, 256, 0, s>>>(B, tiles, ntiles);
	LAUNCH_CHECK(); return 0;
}
int launch_cloud_fused(const DevBatch &B, const Tile *tiles, uint32_t ntiles, const uint32_t *heads, uint32_t nchains, uint64_t *states, uint32_t *ticket, int sms, cudaStream_t s) {
	if(ntiles == 0) return 0;
	// one CTA per chain when there are enough chains to fill the GPU (CORTO_UNPACK=chain / lookback forces either, like the meshes)
	static int mode = -1;
	if(mode < 0) { const char *e = getenv("CORTO_UNPACK"); mode = (e && e[0] == 'c') ? 1 : ((e && e[0] == 'l') ? 2 : 0); }
	if(mode == 1 || (mode == 0 && nchains >= 2u*(uint32_t)sms)) {
		int o5 = 0, o6 = 0;                                    // the 6-CTA build when it saves a wave (see launch_mesh_unpack)
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o5, k_cloud_chain<5>, 256, 0);
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o6, k_cloud_chain<6>, 256, 0);
		bool six = false;
		if(o5 > 0 && o6 > o5) six = (nchains + (uint32_t)(o6*sms) - 1u)/(uint32_t)(o6*sms) < (nchains + (uint32_t)(o5*sms) - 1u)/(uint32_t)(o5*sms);
		if(six) k_cloud_chain<6><<<nchains, 256, 0, s>>>(B, tiles, heads, nchains);
		else k_cloud_chain<5><<<nchains, 256, 0, s>>>(B, tiles, heads, nchains);
	}
	else k_unpack_fused<false><<<persistent_grid(ntiles, 6, sms), 256, 0, s>>>(B, tiles, ntiles, states, ticket);
	LAUNCH_CHECK(); return 0;
}
int launch_mesh_unpack(const DevBatch &B, const Tile *tiles, uint32_t ntiles, const uint32_t *heads, uint32_t nchains, uint64_t *states, uint32_t *ticket, int sms, cudaStream_t s) {
	if(ntiles == 0) return 0;
	// one CTA per chain when there are enough chains to fill the GPU (CORTO_UNPACK=chain / lookback forces either: tests, A/B runs)
	static int mode = -1;
	if(mode < 0) { const char *e = getenv("CORTO_UNPACK"); mode = (e && e[0] == 'c') ? 1 : ((e && e[0] == 'l') ? 2 : 0); }
	const bool chain = mode == 1 || (mode == 0 && nchains >= 2u*(uint32_t)sms);
	// Two builds of the chain kernel: 48 registers (5 CTAs per SM by the occupancy calculator) and capped at 40 (a few spills, 6
	// CTAs).  A chain is one CTA from start to end, so what matters is the number of WAVES the batch needs: the 6-CTA build is
	// taken when it saves a wave (e.g. 768 chains on 148 SMs: 2 waves -> 1), never otherwise.  CORTO_UNPACK_OCC=5|6 forces either.
	static int occ = -1;
	if(occ < 0) { const char *e = getenv("CORTO_UNPACK_OCC"); occ = (e && (e[0] == '5' || e[0] == '6')) ? e[0] - '0' : 0; }
	bool six = occ == 6;
	if(occ == 0 && chain) {
		int o5 = 0, o6 = 0;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o5, k_unpack_chain<5>, 256, 0);
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o6, k_unpack_chain<6>, 256, 0);
		if(o5 > 0 && o6 > o5) {
			const uint32_t w5 = (nchains + (uint32_t)(o5*sms) - 1u)/(uint32_t)(o5*sms), w6 = (nchains + (uint32_t)(o6*sms) - 1u)/(uint32_t)(o6*sms);
			six = w6 < w5;
		}
	}
	if(chain && six) k_unpack_chain<6><<<nchains, 256, 0, s>>>(B, tiles, heads, nchains);
	else if(chain) k_unpack_chain<5><<<nchains, 256, 0, s>>>(B, tiles, heads, nchains);
	else k_unpack_fused<true><<<persistent_grid(ntiles, 6, sms), 256, 0, s>>>(B, tiles, ntiles, states, ticket);
	LAUNCH_CHECK(); return 0;
}
int launch_dequant(const DevBatch &B, const Tile *tiles, uint32_t ntiles, cudaStream_t s) {
	if(ntiles == 0) return 0;
	k_dequant<<<ntiles, 256, 0, s>>>(B, tiles, ntiles);
	LAUNCH_CHECK(); return 0;
}

}  // namespace crtb
