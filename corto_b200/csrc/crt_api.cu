// crt_api.cu — the C ABI of include/corto_b200.h: batch object (device-resident decode), single-decoder API with
// host buffers, and the reference shims' own symbol names.  Host orchestration only; every byte of decode work
// happens in crt_kernels.cu.  There is no CPU decode path in this library.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/corto_b200.h"
#include "crt_common.h"
#include "crt_kernels.h"
#include "crt_walk.h"

using namespace crtb;

// ---------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local int g_code = 0;     // code of the last failure on this thread (crt_batch_create returns a pointer, not a code)
static int fail(int code, const std::string &msg) { g_err = msg; g_code = code; return code; }
static int cuda_fail(cudaError_t e, const char *what) {
	g_err = std::string(what) + ": " + cudaGetErrorString(e);
	return CRT_E_CUDA;
}
#define CU(x) do { cudaError_t e_ = (x); if(e_ != cudaSuccess) return cuda_fail(e_, #x); } while(0)

extern "C" const char *crt_last_error(void) { return g_err.c_str(); }

extern "C" int crt_device_available(void) {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if(e != cudaSuccess || n == 0) { g_err = std::string("no CUDA device: ") + cudaGetErrorString(e); cudaGetLastError(); return 0; }
	return 1;
}

static inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1)/a*a; }

// Small device blocks are recycled instead of going back to the driver: a crt::Decoder-shaped call (one mesh, host buffers)
// would otherwise spend more time in cudaMalloc / cudaFree (implicit device syncs) than in its kernels.  Blocks up to 64 MB,
// at most 512 MB held per process, keyed by device; anything larger is a plain cudaMalloc / cudaFree.  A block returns here only
// after the work that used it has completed (crt_decode synchronises; batch_free_device synchronises the device first).
namespace {
struct DevCache {
	std::mutex m;
	std::multimap<std::pair<int, size_t>, void *> free_blocks;
	std::map<void *, std::pair<int, size_t>> live;
	size_t held = 0;
};
DevCache g_dev_cache;
constexpr size_t DEV_CACHE_BLOCK = 64ull << 20, DEV_CACHE_TOTAL = 512ull << 20;
cudaError_t dev_alloc(void **p, size_t n) {
	int dev = 0;
	cudaGetDevice(&dev);
	if(n <= DEV_CACHE_BLOCK) {
		std::lock_guard<std::mutex> g(g_dev_cache.m);
		auto it = g_dev_cache.free_blocks.lower_bound({dev, n});
		if(it != g_dev_cache.free_blocks.end() && it->first.first == dev && it->first.second <= 2*n + 65536) {
			*p = it->second;
			g_dev_cache.live[*p] = it->first;
			g_dev_cache.held -= it->first.second;
			g_dev_cache.free_blocks.erase(it);
			return cudaSuccess;
		}
	}
	cudaError_t e = cudaMalloc(p, n);
	if(e == cudaSuccess && n <= DEV_CACHE_BLOCK) { std::lock_guard<std::mutex> g(g_dev_cache.m); g_dev_cache.live[*p] = {dev, n}; }
	return e;
}
void dev_free(void *p) {
	if(!p) return;
	{
		std::lock_guard<std::mutex> g(g_dev_cache.m);
		auto it = g_dev_cache.live.find(p);
		if(it != g_dev_cache.live.end()) {
			const std::pair<int, size_t> key = it->second;
			g_dev_cache.live.erase(it);
			if(g_dev_cache.held + key.second <= DEV_CACHE_TOTAL) {
				cudaDeviceSynchronize();                   // what cudaFree would have implied: nothing in flight may still touch the block
				g_dev_cache.free_blocks.insert({key, p}); g_dev_cache.held += key.second;
				return;
			}
		}
	}
	cudaFree(p);
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------
struct Binding { void *ptr; int format; int components; };

struct Stage { const char *name; cudaEvent_t ev; };

struct crt_batch {
	std::vector<ParsedMesh> meshes;
	std::vector<uint64_t> vert_base, face_base;
	uint64_t total_bytes = 0;
	std::map<std::string, Binding> binds;
	std::map<std::string, int> comps;  // components per bound attribute name (must agree across the batch)

	// host images of the device tables
	std::vector<MeshDesc> h_mesh;
	std::vector<TunDesc> h_tun;
	std::vector<uint32_t> h_groups;
	std::vector<Tile> t_tun, t_bits, t_dequant, t_faces, t_verts, t_vscan, t_cfused;
	std::vector<uint2> w_delta;        // delta work items; the first n_delta_crit are the positions an ESTIMATED / BORDER normal waits for
	size_t n_delta_crit = 0;
	std::vector<uint32_t> c_bits;      // per unpack chain (mesh, attribute): index of its first tile in t_bits, + one sentinel
	std::vector<uint32_t> c_cfused;    // the same for the point-cloud chains (t_cfused)
	std::vector<uint32_t> clers_order;
	bool any_border = false;

	// device
	int device = 0, sms = 148;
	uint8_t *d_blobs = nullptr;        size_t blobs_bytes = 0;
	std::vector<std::vector<uint8_t>> tapes;   // crt_batch_create_device: the walk tapes (the directory is re-walked from them)
	bool blobs_external = false;       // crt_batch_create_device: d_blobs is the caller's arena (not owned, never copied into)
	uint8_t *d_tables = nullptr;       size_t tables_bytes = 0;     // MeshDesc | TunDesc | groups | tiles ... (one H2D)
	std::vector<uint8_t> h_tables;
	uint8_t *d_scratch = nullptr;      size_t scratch_bytes = 0;    // symbols, per-mesh work, dictionaries, CLERS slots
	uint8_t *d_zero = nullptr;         size_t zero_bytes = 0;       // region cleared at every decode: tickets, states, status, csr counters
	// offsets inside d_tables
	size_t o_mesh = 0, o_tun = 0, o_groups = 0, o_t_tun = 0, o_t_bits = 0, o_t_dequant = 0, o_t_faces = 0, o_t_verts = 0,
	       o_t_vscan = 0, o_w_delta = 0, o_order = 0, o_t_cfused = 0, o_c_bits = 0, o_c_cfused = 0;
	// offsets inside d_zero
	size_t z_ticket = 0, z_status = 0, z_vcount = 0, z_regular = 0, z_states = 0, z_csr = 0, z_tunbits = 0;
	bool delta_split = false;          // irregular meshes: one warp per component (small batches)
	size_t n_states = 0;
	// scratch pieces
	uint8_t *d_symbols = nullptr, *d_tunrec = nullptr; uint32_t *d_tun_used = nullptr;
	ClersScratch clers{};
	bool uploaded = false;
	uint64_t signature = 0;            // of the directory + bindings the device tables were built from
	size_t dir_bytes = 0;              // prefix of the table image that holds the directory
	int launches = 0;
	bool profiling = false;
	std::vector<Stage> stages;
	std::vector<int> h_status;
	// optional intra-batch overlap (CORTO_OVERLAP=1): the attribute unpack runs beside the CLERS automaton on a side stream
	cudaStream_t side = nullptr;
	cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
	int overlap = -1;                  // -1: read CORTO_OVERLAP on first use
};

extern "C" crt_batch *crt_batch_create(int n, const unsigned char *const *blobs, const int *lens) {
	if(n < 0 || (n > 0 && (!blobs || !lens))) { fail(CRT_E_ARG, "bad arguments"); return nullptr; }
	crt_batch *b = new crt_batch();
	b->meshes.resize(n);
	b->vert_base.assign(n + 1, 0); b->face_base.assign(n + 1, 0);
	for(int i = 0; i < n; i++) {
		std::string err;
		int rc = parse_header(blobs[i], lens[i], b->meshes[i], err);
		if(rc == CRT_OK) rc = walk_directory(b->meshes[i], err);
		if(rc != CRT_OK) { fail(rc, "blob " + std::to_string(i) + ": " + err); delete b; return nullptr; }
		b->vert_base[i + 1] = b->vert_base[i] + b->meshes[i].nvert;
		b->face_base[i + 1] = b->face_base[i] + b->meshes[i].nface;
		b->total_bytes += (uint64_t)lens[i];
	}
	return b;
}

// The bytes parse_header + walk_directory read from one blob (header, group table, block headers — a few hundred bytes, no
// payload), in reading order.  Returns the tape length (the tape is written only when it fits `cap`), or a CRT_E_* code.
extern "C" int crt_walk_tape(const unsigned char *blob, int len, unsigned char *tape, int cap, uint32_t *nvert, uint32_t *nface, uint32_t *nattr) {
	ParsedMesh pm;
	std::vector<uint8_t> rec;
	pm.record = &rec;
	std::string err;
	int rc = parse_header(blob, len, pm, err);
	if(rc == CRT_OK) rc = walk_directory(pm, err);
	if(rc != CRT_OK) return fail(rc, err);
	if(nvert) *nvert = pm.nvert;
	if(nface) *nface = pm.nface;
	if(nattr) *nattr = (uint32_t)pm.attrs.size();
	if(tape && (int)rec.size() <= cap && !rec.empty()) memcpy(tape, rec.data(), rec.size());
	return (int)rec.size();
}

// A batch whose blobs are ALREADY in device memory (e.g. received over NVLink from the rank that ingested them): `arena` holds
// them back to back, blob i at the sum of the 16-byte-rounded lengths before it; the directory of each comes from its walk
// tape (crt_walk_tape on the rank that had the host copy).  The arena is borrowed for the lifetime of the batch; nothing is
// copied to or from the host but the tables.
extern "C" crt_batch *crt_batch_create_device(int n, const unsigned char *const *tapes, const int *tape_lens, const int *blob_lens, const void *arena) {
	if(n < 0 || (n > 0 && (!tapes || !tape_lens || !blob_lens || !arena)) || ((uintptr_t)arena & 15)) { fail(CRT_E_ARG, "bad arguments"); return nullptr; }
	crt_batch *b = new crt_batch();
	b->meshes.resize(n);
	b->tapes.reserve(n);
	b->vert_base.assign(n + 1, 0); b->face_base.assign(n + 1, 0);
	for(int i = 0; i < n; i++) {
		std::string err;
		ParsedMesh &pm = b->meshes[i];
		b->tapes.emplace_back(tapes[i], tapes[i] + (tape_lens[i] > 0 ? tape_lens[i] : 0));
		pm.tape = b->tapes.back().data(); pm.tape_len = (uint32_t)b->tapes.back().size();
		int rc = parse_header(nullptr, blob_lens[i], pm, err);
		if(rc == CRT_OK) rc = walk_directory(pm, err);
		if(rc != CRT_OK) { fail(rc, "blob " + std::to_string(i) + ": " + err); delete b; return nullptr; }
		b->vert_base[i + 1] = b->vert_base[i] + pm.nvert;
		b->face_base[i + 1] = b->face_base[i] + pm.nface;
		b->total_bytes += (uint64_t)blob_lens[i];
	}
	b->d_blobs = (uint8_t *)arena;
	b->blobs_external = true;
	return b;
}

static void batch_free_device(crt_batch *b) {
	if(b->d_tables || b->d_scratch || b->d_zero) cudaDeviceSynchronize();   // (cudaFree used to imply it; recycled blocks must be idle)
	if(b->d_blobs && !b->blobs_external) dev_free(b->d_blobs);
	if(b->d_tables) dev_free(b->d_tables);
	if(b->d_scratch) dev_free(b->d_scratch);
	if(b->d_zero) dev_free(b->d_zero);
	if(!b->blobs_external) b->d_blobs = nullptr;
	b->d_tables = b->d_scratch = b->d_zero = nullptr;
	for(auto &s: b->stages) cudaEventDestroy(s.ev);
	b->stages.clear();
	for(int k = 0; k < 2; k++) {
		if(b->ev_fork[k]) cudaEventDestroy(b->ev_fork[k]);
		if(b->ev_join[k]) cudaEventDestroy(b->ev_join[k]);
		b->ev_fork[k] = b->ev_join[k] = nullptr;
	}
	if(b->side) { cudaStreamDestroy(b->side); b->side = nullptr; }
	b->uploaded = false;
}

extern "C" void crt_batch_destroy(crt_batch *b) {
	if(!b) return;
	batch_free_device(b);
	delete b;
}

extern "C" int crt_batch_count(const crt_batch *b) { return (int)b->meshes.size(); }
extern "C" uint64_t crt_batch_total_verts(const crt_batch *b) { return b->vert_base.back(); }
extern "C" uint64_t crt_batch_total_faces(const crt_batch *b) { return b->face_base.back(); }
extern "C" uint64_t crt_batch_total_bytes(const crt_batch *b) { return b->total_bytes; }
extern "C" const uint64_t *crt_batch_vert_base(const crt_batch *b) { return b->vert_base.data(); }
extern "C" const uint64_t *crt_batch_face_base(const crt_batch *b) { return b->face_base.data(); }

extern "C" int crt_batch_mesh_info(const crt_batch *b, int i, uint32_t *nvert, uint32_t *nface, uint32_t *attr_mask) {
	if(i < 0 || i >= (int)b->meshes.size()) return fail(CRT_E_ARG, "mesh index out of range");
	const ParsedMesh &m = b->meshes[i];
	if(nvert) *nvert = m.nvert;
	if(nface) *nface = m.nface;
	if(attr_mask) {
		uint32_t k = m.nface ? CRT_HAS_INDEX : 0;
		for(auto &a: m.attrs) {
			if(a.name == "position") k |= CRT_HAS_POSITION;
			else if(a.name == "normal") k |= CRT_HAS_NORMAL;
			else if(a.name == "color") k |= CRT_HAS_COLOR;
			else if(a.name == "uv") k |= CRT_HAS_UV;
			else k |= CRT_HAS_OTHER;
		}
		*attr_mask = k;
	}
	return CRT_OK;
}

extern "C" int crt_batch_attr_components(const crt_batch *b, const char *name) {
	if(!name) return 0;
	for(const ParsedMesh &m: b->meshes) { const int a = m.find(name); if(a >= 0) return m.attrs[a].N; }
	return 0;
}

extern "C" int crt_batch_bind(crt_batch *b, const char *name, void *device_ptr, int format, int components) {
	if(!name) return fail(CRT_E_ARG, "null attribute name");
	if(!device_ptr) { b->binds.erase(name); return CRT_OK; }      // unbind: takes effect at the next crt_batch_upload / rewalk
	b->binds[name] = Binding{device_ptr, format, components};
	return CRT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Build every host table from the parsed directories + the current bindings.
static int build_tables(crt_batch *b, uint64_t &symbols_bytes, uint64_t &work_bytes, std::vector<uint64_t> &work_off /*per mesh*/,
                        uint64_t &zero_csr_bytes, std::vector<uint64_t> &csr_off, uint64_t &adj_bytes, std::vector<uint64_t> &adj_off) {
	const int n = (int)b->meshes.size();
	b->h_mesh.assign(n, MeshDesc{});
	b->h_tun.clear(); b->h_groups.clear();
	b->t_tun.clear(); b->t_bits.clear(); b->t_dequant.clear(); b->t_faces.clear(); b->t_verts.clear(); b->t_vscan.clear(); b->t_cfused.clear();
	b->w_delta.clear(); b->clers_order.clear();
	b->any_border = false;
	b->comps.clear();
	symbols_bytes = 0; work_bytes = 0; zero_csr_bytes = 0; adj_bytes = 0;
	work_off.assign(n, 0); csr_off.assign(n, 0); adj_off.assign(n, 0);
	uint64_t blob_off = 0;

	auto add_block = [&](const ParsedMesh &pm, const Block &blk, uint64_t mesh_blob_off) -> int {
		TunDesc td{};
		td.probs_off = mesh_blob_off + blk.probs_off;
		td.data_off = mesh_blob_off + blk.data_off;
		td.out_off = symbols_bytes;
		td.nsym = blk.nsym; td.size = blk.size; td.csize = blk.csize; td.raw = blk.raw ? 1 : 0;
		td.tile0 = (uint32_t)b->t_tun.size();
		symbols_bytes += align_up((uint64_t)blk.size + 32, 16);   // k_clers reads its symbols through 8-byte windows, 16 bytes ahead
		const int id = (int)b->h_tun.size();
		if(blk.size) {
			if(blk.raw || blk.nsym <= 1) {
				uint32_t nt = (blk.size + TUN_TILE*4 - 1)/(TUN_TILE*4);
				for(uint32_t t = 0; t < nt; t++) b->t_tun.push_back(Tile{(uint32_t)id, 0, t, 1});
			} else {
				uint32_t nt = (blk.csize + TUN_TILE - 1)/TUN_TILE;
				for(uint32_t t = 0; t < nt; t++) b->t_tun.push_back(Tile{(uint32_t)id, 0, t, t == 0 ? 1u : 0u});
			}
		}
		b->h_tun.push_back(td);
		(void)pm;
		return id;
	};

	for(int i = 0; i < n; i++) {
		const ParsedMesh &pm = b->meshes[i];
		MeshDesc &M = b->h_mesh[i];
		M.blob_off = blob_off; M.blob_len = pm.len;
		M.nvert = pm.nvert; M.nface = pm.nface; M.nattr = (uint32_t)pm.attrs.size();
		M.ngroups = (uint32_t)pm.group_ends.size(); M.group0 = (uint32_t)b->h_groups.size();
		uint32_t prev = 0, maxg = 0;
		for(uint32_t e: pm.group_ends) {
			b->h_groups.push_back(e);
			uint32_t ee = std::min(e, pm.nface);
			if(ee > prev) maxg = std::max(maxg, ee - prev);
			prev = e;
		}
		M.max_group_faces = maxg;
		M.clers_tun = -1; M.position_attr = pm.find("position"); M.normal_attr = -1;
		M.max_front = pm.max_front;
		uint64_t wb = 0;                       // per-mesh work bytes (16-aligned pieces)
		auto take = [&](uint64_t bytes) { uint64_t o = wb; wb += align_up(bytes, 16); return o; };
		uint64_t pred_o = 0, face_o = 0;
		if(pm.nface) {
			M.clers_tun = add_block(pm, pm.clers, blob_off);
			M.nclers = pm.clers.size;
			M.split_off = blob_off + pm.split_off; M.split_nwords = pm.split_nwords;
			pred_o = take((uint64_t)pm.nvert*16);
			b->clers_order.push_back((uint32_t)i);
		}
		// index binding
		auto ib = b->binds.find("index");
		bool index_bound = pm.nface && ib != b->binds.end();
		if(index_bound) {
			if(ib->second.format != CRT_UINT32 && ib->second.format != CRT_UINT16) return fail(CRT_E_FORMAT, "index must be CRT_UINT32 or CRT_UINT16");
			M.index16 = ib->second.format == CRT_UINT16;
			M.index_ptr = (uint64_t)ib->second.ptr + b->face_base[i]*3*(M.index16 ? 2 : 4);
		} else if(pm.nface) face_o = take((uint64_t)pm.nface*12);

		std::vector<uint64_t> attr_work(pm.attrs.size(), 0);
		for(size_t a = 0; a < pm.attrs.size(); a++) {
			const ParsedAttr &pa = pm.attrs[a];
			const AttrStreams &st = pm.streams[a];
			AttrDesc &A = M.attr[a];
			A.codec = pa.codec; A.N = pa.N; A.ncomp = pa.codec == CODEC_NORMAL ? 2 : pa.N;
			A.strategy = pa.strategy; A.q = pa.q;
			A.out_format = -1; A.out_components = pa.N; A.prediction = st.prediction;
			for(int k = 0; k < 4; k++) A.qc[k] = st.qc[k];
			A.bits_off = blob_off + st.bits_off; A.bits_nwords = st.bits_nwords;
			A.ntun = (int)st.blocks.size();
			A.count = st.blocks.empty() ? 0 : st.blocks[0].size;
			auto it = b->binds.find(pa.name);
			const bool bound = it != b->binds.end();
			if(bound) {
				const Binding &bd = it->second;
				// every mesh's slice of an arena is vert_base * stride: the component count behind a name has to be the same for all
				// meshes of the batch (and is what the caller sized the arena by: crt_batch_attr_components)
				// (colours are written with the caller's component count, normals always as triples: their N may differ per mesh)
				auto nc = b->comps.find(pa.name);
				if(nc == b->comps.end()) b->comps[pa.name] = pa.N;
				else if(nc->second != pa.N && pa.codec == CODEC_GENERIC) return fail(CRT_E_LIMIT, "attribute '" + pa.name + "' has different component counts in the meshes of this batch");
				uint64_t stride;
				if(pa.codec == CODEC_NORMAL) {
					if(bd.format != CRT_FLOAT && bd.format != CRT_INT16) return fail(CRT_E_FORMAT, "Format not supported for normal attribute (float, int16 only)");
					stride = bd.format == CRT_FLOAT ? 12 : 6;
				} else if(pa.codec == CODEC_COLOR) {
					if(bd.format != CRT_UINT8) return fail(CRT_E_FORMAT, "Unsupported color output format.");
					if(bd.components < 1 || bd.components > 4) return fail(CRT_E_ARG, "colour components must be 1..4");
					A.out_components = bd.components;
					stride = (uint64_t)bd.components;
				} else {
					if(bd.format != CRT_FLOAT && bd.format != CRT_INT32 && bd.format != CRT_UINT32) return fail(CRT_E_FORMAT, "generic attributes decode to CRT_FLOAT, CRT_INT32 or CRT_UINT32");
					stride = (uint64_t)pa.N*4;
				}
				A.out_format = bd.format;
				A.out_ptr = (uint64_t)bd.ptr + b->vert_base[i]*stride;
			}
			// entropy blocks are registered only for attributes somebody consumes (unbound ones are skipped)
			for(int k = 0; k < A.ntun; k++) A.tun[k] = bound ? add_block(pm, st.blocks[k], blob_off) : -1;
			if(!bound) continue;
			if(pa.codec == CODEC_NORMAL) attr_work[a] = take((uint64_t)pm.nvert*8 + 16);
			if(pa.codec == CODEC_COLOR) attr_work[a] = take((uint64_t)pm.nvert*pa.N + 16);
			const bool normal_diff = pa.codec == CODEC_NORMAL && st.prediction == N_DIFF;
			if(pm.nface == 0) {
				// point cloud: one fused kernel (bit unpack -> running delta -> dequantise) per attribute; normals other than DIFF are
				// parsed only (the reference's cloud path runs no postDelta, decoder.cpp:133-147)
				if(pa.codec != CODEC_NORMAL || normal_diff) {
					uint32_t nt = (pm.nvert + 1023)/1024;
					for(uint32_t t = 0; t < nt; t++) b->t_cfused.push_back(Tile{(uint32_t)i, (uint32_t)a, t, t == 0 ? 1u : 0u});
				}
				continue;
			}
			// bit-unpack tiles (all components of 1024 vertices per tile)
			{
				uint32_t nt = (pm.nvert + 1023)/1024;
				for(uint32_t t = 0; t < nt; t++) b->t_bits.push_back(Tile{(uint32_t)i, (uint32_t)a, t, t == 0 ? 1u : 0u});
			}
			// delta inverse
			if(pa.codec != CODEC_NORMAL || normal_diff) {
				b->w_delta.push_back(make_uint2((unsigned)i, (unsigned)a | 0xff00u | ((unsigned)A.ncomp << 16)));   // all components; meshes only: clouds went through the fused kernel
			}
			// dequantise (normals: only DIFF goes through k_dequant; ESTIMATED/BORDER are finished by k_normal_estimate)
			if(pa.codec != CODEC_NORMAL || normal_diff) {
				uint32_t nt = (pm.nvert + SCAN_TILE - 1)/SCAN_TILE;
				for(uint32_t t = 0; t < nt; t++) b->t_dequant.push_back(Tile{(uint32_t)i, (uint32_t)a, t, 0});
			}
			if(pa.codec == CODEC_NORMAL && !normal_diff && pm.nface) {
				if(st.prediction != N_ESTIMATED && st.prediction != N_BORDER) return fail(CRT_E_FORMAT, "unknown normal prediction in stream");
				M.normal_attr = (int)a;
				if(st.prediction == N_BORDER) b->any_border = true;
			}
		}
		if(M.normal_attr >= 0) {
			if(M.position_attr < 0) return fail(CRT_E_NOPOSITION, "No position attribute found. Use DIFF normal strategy instead.");   // normal_attribute.cpp:219-221
			if(M.attr[M.position_attr].out_format != CRT_FLOAT || !M.attr[M.position_attr].out_ptr)
				return fail(CRT_E_NOPOSITION, "ESTIMATED/BORDER normals need the position attribute bound as float (normal_attribute.cpp:230)");
			if(M.attr[M.position_attr].N != 3 || M.attr[M.position_attr].codec != CODEC_GENERIC)     // estimateNormals reads Point3i (normal_attribute.cpp:230)
				return fail(CRT_E_LIMIT, "ESTIMATED/BORDER normals need a 3-component generic position attribute");
			csr_off[i] = zero_csr_bytes;
			// cnt | ohead | novf, and for BORDER prediction bnd | cidx[+1] (crt_kernels.cu: adj_view)
			const bool border_pred = pm.streams[M.normal_attr].prediction == N_BORDER;
			zero_csr_bytes += align_up(((uint64_t)pm.nvert*(border_pred ? 4 : 2) + 3)*4, 16);
			adj_off[i] = adj_bytes;
			adj_bytes += align_up((uint64_t)pm.nface*16 + (uint64_t)pm.nvert*32 + (uint64_t)pm.nface*24, 16);   // face normals + 8 slots per vertex + overflow pairs
			uint32_t nf = (pm.nface + SCAN_TILE - 1)/SCAN_TILE, nv = (pm.nvert + SCAN_TILE - 1)/SCAN_TILE, ns = (pm.nvert + 1 + SCAN_TILE - 1)/SCAN_TILE;
			for(uint32_t t = 0; t < nf; t++) b->t_faces.push_back(Tile{(uint32_t)i, 0, t, 0});
			for(uint32_t t = 0; t < nv; t++) b->t_verts.push_back(Tile{(uint32_t)i, 0, t, 0});
			if(border_pred) for(uint32_t t = 0; t < ns; t++) b->t_vscan.push_back(Tile{(uint32_t)i, 0, t, t == 0 ? 1u : 0u});   // the boundary scan: BORDER meshes only (the others have no bnd / cidx)
		}
		// stash per-mesh work offsets (resolved to pointers once the arena address is known)
		M.pred_ptr = pred_o; M.face_ptr = face_o;
		for(size_t a = 0; a < pm.attrs.size(); a++) M.attr[a].work_ptr = attr_work[a];
		work_off[i] = work_bytes;
		work_bytes += align_up(wb, 256);
		blob_off += align_up(pm.len, 16);
	}
	// biggest meshes first: the serial automaton is the long pole (LPT order)
	std::stable_sort(b->clers_order.begin(), b->clers_order.end(), [&](uint32_t x, uint32_t y) { return b->meshes[x].nface > b->meshes[y].nface; });
	return CRT_OK;
}

template <class T> static size_t put(std::vector<uint8_t> &img, const std::vector<T> &v) {
	size_t o = align_up(img.size(), 256);
	img.resize(o + v.size()*sizeof(T) + 16);
	if(!v.empty()) memcpy(img.data() + o, v.data(), v.size()*sizeof(T));
	return o;
}

static uint64_t directory_signature(const crt_batch *b);

static int batch_prepare(crt_batch *b, cudaStream_t stream, bool copy_blobs) {
	CU(cudaGetDevice(&b->device));
	CU(cudaDeviceGetAttribute(&b->sms, cudaDevAttrMultiProcessorCount, b->device));
	uint64_t symbols_bytes, work_bytes, csr_bytes, adj_bytes;
	std::vector<uint64_t> work_off, csr_off, adj_off;
	int rc = build_tables(b, symbols_bytes, work_bytes, work_off, csr_bytes, csr_off, adj_bytes, adj_off);
	if(rc) return rc;
	const int n = (int)b->meshes.size();

	// ---- blob arena ----
	uint64_t blobs_bytes = 16;
	for(auto &m: b->meshes) blobs_bytes += align_up(m.len, 16);
	if(b->blobs_external) copy_blobs = false;
	else if(!b->d_blobs || b->blobs_bytes < blobs_bytes) {
		if(b->d_blobs) { dev_free(b->d_blobs); b->d_blobs = nullptr; }
		CU(dev_alloc((void **)&b->d_blobs, blobs_bytes));
		b->blobs_bytes = blobs_bytes;
		copy_blobs = true;
	}
	if(copy_blobs && n) {
		// host blobs that already sit in one buffer with the arena's layout (16-byte-rounded, back to back) travel as ONE copy
		bool contiguous = true;
		for(int i = 1; i < n && contiguous; i++)
			contiguous = b->meshes[i].blob == b->meshes[0].blob + b->h_mesh[i].blob_off;
		if(contiguous) CU(cudaMemcpyAsync(b->d_blobs, b->meshes[0].blob, b->h_mesh[n - 1].blob_off + b->meshes[n - 1].len, cudaMemcpyHostToDevice, stream));
		else for(int i = 0; i < n; i++)
			CU(cudaMemcpyAsync(b->d_blobs + b->h_mesh[i].blob_off, b->meshes[i].blob, b->meshes[i].len, cudaMemcpyHostToDevice, stream));
	}

	// ---- scratch arena: symbols | per-mesh work | adj | dictionaries | CLERS slots ----
	const uint64_t ntun = b->h_tun.size();
	uint32_t cap = 16, nmesh_faces = (uint32_t)b->clers_order.size();
	for(auto &M: b->h_mesh) if(M.nface) cap = std::max(cap, 3u*M.max_group_faces + 16u);
	uint32_t slots = std::min<uint32_t>(nmesh_faces, (uint32_t)b->sms*8u);
	const uint64_t per_slot = (uint64_t)cap*(sizeof(EdgeA) + sizeof(EdgeB) + 8);
	while(slots > 1 && per_slot*slots > (48ull << 30)) slots /= 2;
	uint64_t o_sym = 0, o_work = align_up(o_sym + symbols_bytes + 64, 256), o_adj = align_up(o_work + work_bytes, 256),
	         o_rec = align_up(o_adj + adj_bytes, 256), o_used = align_up(o_rec + ntun*TUN_REC_BYTES, 256),
	         o_clers = align_up(o_used + ntun*4 + 16, 256), total = o_clers + per_slot*slots + 256;
	if(!b->d_scratch || b->scratch_bytes < total) {
		if(b->d_scratch) { dev_free(b->d_scratch); b->d_scratch = nullptr; }
		CU(dev_alloc((void **)&b->d_scratch, total));
		b->scratch_bytes = total;
	}
	b->d_symbols = b->d_scratch + o_sym;
	b->d_tunrec = b->d_scratch + o_rec;
	b->d_tun_used = (uint32_t *)(b->d_scratch + o_used);
	b->clers.cap = cap; b->clers.slots = slots;
	uint8_t *cs = b->d_scratch + o_clers;
	b->clers.ea = (EdgeA *)cs; cs += (uint64_t)cap*slots*sizeof(EdgeA);
	b->clers.eb = (EdgeB *)cs; cs += (uint64_t)cap*slots*sizeof(EdgeB);
	b->clers.order = (uint32_t *)cs; cs += (uint64_t)cap*slots*4;
	b->clers.delayed = (uint32_t *)cs;

	// ---- zeroed control region: tickets | status | vertex_count | look-back states | csr counters ----
	b->n_states = b->t_tun.size() + 8*b->t_bits.size() + b->t_vscan.size() + 8*b->t_cfused.size();
	b->z_ticket = 0;
	b->z_status = 256;
	b->z_vcount = align_up(b->z_status + (uint64_t)n*4, 256);
	b->z_regular = align_up(b->z_vcount + (uint64_t)n*4, 256);
	b->z_states = align_up(b->z_regular + (uint64_t)n*4, 256);
	b->z_tunbits = align_up(b->z_states + b->n_states*8, 256);
	b->z_csr = align_up(b->z_tunbits + ntun*8, 256);
	uint64_t zero_total = b->z_csr + csr_bytes + 256;
	if(!b->d_zero || b->zero_bytes < zero_total) {
		if(b->d_zero) { dev_free(b->d_zero); b->d_zero = nullptr; }
		CU(dev_alloc((void **)&b->d_zero, zero_total));
	}
	b->zero_bytes = zero_total;

	// ---- resolve scratch pointers inside the descriptors ----
	for(int i = 0; i < n; i++) {
		MeshDesc &M = b->h_mesh[i];
		uint8_t *w = b->d_scratch + o_work + work_off[i];
		if(M.nface) {
			M.pred_ptr = (uint64_t)(w + M.pred_ptr);
			M.face_ptr = M.index_ptr ? M.index_ptr : (uint64_t)(w + M.face_ptr);
			if(!M.index_ptr) M.index16 = 0;
		} else { M.pred_ptr = 0; M.face_ptr = 0; }
		for(uint32_t a = 0; a < M.nattr; a++) {
			AttrDesc &A = M.attr[a];
			if(A.out_format >= 0 && (A.codec == CODEC_NORMAL || A.codec == CODEC_COLOR)) A.work_ptr = (uint64_t)(w + A.work_ptr);
			else A.work_ptr = 0;
		}
		if(M.normal_attr >= 0) {
			M.csr_ptr = (uint64_t)(b->d_zero + b->z_csr + csr_off[i]);
			M.adj_ptr = (uint64_t)(b->d_scratch + o_adj + adj_off[i]);
		}
	}

	// ---- one table image, one H2D ----
	std::vector<uint8_t> &img = b->h_tables;
	img.clear();
	b->o_mesh = put(img, b->h_mesh);
	b->o_tun = put(img, b->h_tun);
	b->o_groups = put(img, b->h_groups);
	b->dir_bytes = img.size();              // MeshDesc | TunDesc | groups: the directory proper
	b->o_t_tun = put(img, b->t_tun);
	b->o_t_bits = put(img, b->t_bits);
	b->c_bits.clear();
	for(size_t t = 0; t < b->t_bits.size(); t++) if(b->t_bits[t].first) b->c_bits.push_back((uint32_t)t);
	b->c_bits.push_back((uint32_t)b->t_bits.size());
	b->o_c_bits = put(img, b->c_bits);
	b->o_t_dequant = put(img, b->t_dequant);
	b->o_t_faces = put(img, b->t_faces);
	b->o_t_verts = put(img, b->t_verts);
	b->o_t_vscan = put(img, b->t_vscan);
	// Delta work items: one warp per (mesh, attribute) walks all components together (they share the prediction loads and
	// interleave in the pipeline).  When that leaves most SMs without a warp (few, large meshes) the components — independent
	// chains — get a warp each instead: measured 94 -> 80 ms on 64 x 1.7 M-vertex meshes, no gain (c2) or a loss (c4) on big batches.
	// (only the default warp kernel takes per-component items; the block-wide kernel, CORTO_DELTA=cta|seq, walks all components)
	const char *force = getenv("CORTO_DELTA_SPLIT");                // 0 / 1 overrides the heuristic (tests, A/B runs)
	const char *dmode = getenv("CORTO_DELTA");
	const bool warp_kernel = dmode && dmode[0] == 'w';               // k_delta_mesh takes per-component items from the host
	const bool want_split = force ? force[0] == '1' : b->w_delta.size() < 2u*(size_t)b->sms;
	b->delta_split = want_split;                                      // k_delta_mesh_seg splits inside the CTA (irregular meshes only)
	if(warp_kernel && want_split) {
		std::vector<uint2> split;
		for(const uint2 &w: b->w_delta)
			for(unsigned c = 0; c < (w.y >> 16); c++) split.push_back(make_uint2(w.x, (w.y & 0xffu) | (c << 8)));
		b->w_delta.swap(split);
	}
	// positions that a normal estimation waits for go first: CORTO_OVERLAP=2 runs the rest beside the estimation (crt_batch_decode)
	{
		auto crit = [&](const uint2 &w) { const MeshDesc &M = b->h_mesh[w.x]; return M.normal_attr >= 0 && (int)(w.y & 0xffu) == M.position_attr; };
		std::stable_partition(b->w_delta.begin(), b->w_delta.end(), crit);
		b->n_delta_crit = 0;
		for(const uint2 &w: b->w_delta) if(crit(w)) b->n_delta_crit++;
	}
	b->o_w_delta = put(img, b->w_delta);
	b->o_order = put(img, b->clers_order);
	b->o_t_cfused = put(img, b->t_cfused);
	b->c_cfused.clear();
	for(size_t t = 0; t < b->t_cfused.size(); t++) if(b->t_cfused[t].first) b->c_cfused.push_back((uint32_t)t);
	b->c_cfused.push_back((uint32_t)b->t_cfused.size());
	b->o_c_cfused = put(img, b->c_cfused);
	if(!b->d_tables || b->tables_bytes < img.size()) {
		if(b->d_tables) { dev_free(b->d_tables); b->d_tables = nullptr; }
		CU(dev_alloc((void **)&b->d_tables, img.size() + 256));
		b->tables_bytes = img.size() + 256;
	}
	CU(cudaMemcpyAsync(b->d_tables, img.data(), img.size(), cudaMemcpyHostToDevice, stream));
	// the table image is pageable host memory owned by the batch; wait so it may be rebuilt safely
	CU(cudaStreamSynchronize(stream));
	b->uploaded = true;
	b->signature = directory_signature(b);
	return CRT_OK;
}

extern "C" int crt_batch_upload(crt_batch *b, void *stream) {
	if(!crt_device_available()) return CRT_E_CUDA;
	return batch_prepare(b, (cudaStream_t)stream, true);
}

// Signature of a walked directory: every offset / size the device tables are derived from.
static uint64_t directory_signature(const crt_batch *b) {
	uint64_t h = 0xcbf29ce484222325ull;
	auto mix = [&](uint64_t v) { h ^= v; h *= 0x100000001b3ull; };
	auto blk = [&](const Block &k) { mix(k.probs_off); mix(k.nsym); mix(k.size); mix(k.csize); mix(k.data_off); mix(k.raw); };
	for(const ParsedMesh &m: b->meshes) {
		mix(m.nvert); mix(m.nface); mix(m.max_front); mix(m.split_off); mix(m.split_nwords);
		for(uint32_t e: m.group_ends) mix(e);
		blk(m.clers);
		for(const AttrStreams &s: m.streams) {
			mix(s.bits_off); mix(s.bits_nwords); mix((uint64_t)s.prediction);
			for(int k = 0; k < 4; k++) mix((uint64_t)s.qc[k]);
			for(const Block &k: s.blocks) blk(k);
		}
	}
	for(auto &kv: b->binds) { for(char c: kv.first) mix((uint64_t)c); mix((uint64_t)kv.second.ptr); mix((uint64_t)kv.second.format); mix((uint64_t)kv.second.components); }
	return h;
}

// (tests: a batch built from walk tapes must carry the same directory as one built from the host blobs)
extern "C" uint64_t crt_batch_directory_signature(const crt_batch *b) { return directory_signature(b); }

extern "C" int crt_batch_rewalk(crt_batch *b, void *stream) {
	if(!b->uploaded) return fail(CRT_E_ARG, "crt_batch_rewalk before crt_batch_upload");
	for(size_t i = 0; i < b->meshes.size(); i++) {
		std::string err;
		int rc = walk_directory(b->meshes[i], err);
		if(rc) return fail(rc, "blob " + std::to_string(i) + ": " + err);
	}
	// The directory (MeshDesc / TunDesc / group table) goes to the device again.  The launch geometry derived from it (tile
	// lists, work orders) is a pure function of the directory: it is rebuilt only when the walk found something different.
	if(directory_signature(b) == b->signature && b->dir_bytes) {
		CU(cudaMemcpyAsync(b->d_tables, b->h_tables.data(), b->dir_bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
		return CRT_OK;
	}
	return batch_prepare(b, (cudaStream_t)stream, false);
}

extern "C" int crt_batch_set_profiling(crt_batch *b, int on) { b->profiling = on != 0; return CRT_OK; }

static int mark(crt_batch *b, const char *name, size_t &k, cudaStream_t s) {
	if(!b->profiling) return CRT_OK;
	if(k >= b->stages.size()) { Stage st{name, nullptr}; CU(cudaEventCreate(&st.ev)); b->stages.push_back(st); }
	b->stages[k].name = name;
	CU(cudaEventRecord(b->stages[k].ev, s));
	k++;
	return CRT_OK;
}

extern "C" int crt_batch_decode(crt_batch *b, void *stream_) {
	if(!b->uploaded) return fail(CRT_E_ARG, "crt_batch_decode before crt_batch_upload");
	cudaStream_t s = (cudaStream_t)stream_;
	DevBatch B{};
	B.blobs = b->d_blobs; B.symbols = b->d_symbols;
	B.mesh = (const MeshDesc *)(b->d_tables + b->o_mesh);
	B.tun = (const TunDesc *)(b->d_tables + b->o_tun);
	B.tunrec = b->d_tunrec; B.tun_used = b->d_tun_used;
	B.group_ends = (const uint32_t *)(b->d_tables + b->o_groups);
	B.status = (int32_t *)(b->d_zero + b->z_status);
	B.vertex_count = (uint32_t *)(b->d_zero + b->z_vcount);
	B.regular = (uint32_t *)(b->d_zero + b->z_regular);
	B.tun_bits = (unsigned long long *)(b->d_zero + b->z_tunbits);
	uint32_t *tickets = (uint32_t *)(b->d_zero + b->z_ticket);
	uint64_t *states = (uint64_t *)(b->d_zero + b->z_states);
	const Tile *t_tun = (const Tile *)(b->d_tables + b->o_t_tun), *t_bits = (const Tile *)(b->d_tables + b->o_t_bits),
	           *t_dequant = (const Tile *)(b->d_tables + b->o_t_dequant),
	           *t_faces = (const Tile *)(b->d_tables + b->o_t_faces), *t_verts = (const Tile *)(b->d_tables + b->o_t_verts),
	           *t_vscan = (const Tile *)(b->d_tables + b->o_t_vscan);
	int launches = 0;
	size_t k = 0;
	int rc;
#define RUN(call, cond) do { if(cond) { int e_ = (call); if(e_) return cuda_fail((cudaError_t)e_, #call); launches++; } } while(0)
	if((rc = mark(b, "begin", k, s))) return rc;
	CU(cudaMemsetAsync(b->d_zero, 0, b->zero_bytes, s));
	RUN(launch_tun_tables(B, (int)b->h_tun.size(), s), !b->h_tun.empty());
	if((rc = mark(b, "tun_tables", k, s))) return rc;
	uint64_t *st = states;
	RUN(launch_tun_decode(B, t_tun, (uint32_t)b->t_tun.size(), st, tickets + 0, b->sms, s), !b->t_tun.empty());
	st += b->t_tun.size();
	if((rc = mark(b, "tun_decode", k, s))) return rc;
	// Stage order inside one batch (CORTO_OVERLAP, three bits, default 3; stage timers — profiling — always run serial, 0 = one stream):
	//   bit 0  the attribute unpack beside the CLERS automaton on a side stream (the automaton leaves most issue slots idle).
	//   bit 1  what an ESTIMATED / BORDER normal does not wait for leaves the critical path CLERS -> position delta -> face normals
	//          -> estimation: the adjacency (it needs the faces, not the positions), the boundary scan and the delta inverse of the
	//          other attributes run on the side stream beside the position delta.
	//   bit 2  the dequantisation beside the estimation (which reads face normals + adjacency, not the positions).  Both are
	//          HBM-bound: no gain, noisier (6.73 / 7.19 ms in two runs) — off by default.
	//   configs[1] on one box: 7.57 (0) -> 6.90 (2) -> 6.74 ms (3); configs[3]: 27.0 -> 26.4 -> 25.2 ms.
	if(b->overlap < 0) { const char *e = getenv("CORTO_OVERLAP"); b->overlap = e ? atoi(e) & 7 : 3; }
	const bool ovl = (b->overlap & 1) && !b->profiling && !b->clers_order.empty() && !b->t_bits.empty();
	// bit 1: the delta inverse of everything a normal estimation does NOT wait for (uv, colours, ...) runs beside the estimation
	const bool ovl2 = (b->overlap & 2) && !b->profiling && !b->t_faces.empty() && b->n_delta_crit > 0;
	cudaStream_t s2 = s;
	if(ovl || ovl2) {
		if(!b->side) {
			CU(cudaStreamCreateWithFlags(&b->side, cudaStreamNonBlocking));
			for(int k2 = 0; k2 < 2; k2++) {
				CU(cudaEventCreateWithFlags(&b->ev_fork[k2], cudaEventDisableTiming));
				CU(cudaEventCreateWithFlags(&b->ev_join[k2], cudaEventDisableTiming));
			}
		}
	}
	if(ovl) s2 = b->side;
	uint64_t *st_bits = st, *st_cfused = st + 8*b->t_bits.size();
	st = st_cfused + 8*b->t_cfused.size();
	if(ovl) {
		CU(cudaEventRecord(b->ev_fork[0], s));
		RUN(launch_clers(B, (const uint32_t *)(b->d_tables + b->o_order), (uint32_t)b->clers_order.size(), b->clers, tickets + 2, b->sms, s), true);
		CU(cudaStreamWaitEvent(s2, b->ev_fork[0], 0));
	}
	RUN(launch_mesh_unpack(B, t_bits, (uint32_t)b->t_bits.size(), (const uint32_t *)(b->d_tables + b->o_c_bits), (uint32_t)b->c_bits.size() - 1u, st_bits, tickets + 1, b->sms, s2), !b->t_bits.empty());
	if((rc = mark(b, "bit_unpack", k, s))) return rc;
	RUN(launch_cloud_fused(B, (const Tile *)(b->d_tables + b->o_t_cfused), (uint32_t)b->t_cfused.size(), (const uint32_t *)(b->d_tables + b->o_c_cfused), (uint32_t)b->c_cfused.size() - 1u, st_cfused, tickets + 6, b->sms, s2), !b->t_cfused.empty());
	if((rc = mark(b, "cloud_fused", k, s))) return rc;
	if(ovl) {
		CU(cudaEventRecord(b->ev_join[0], s2));
		CU(cudaStreamWaitEvent(s, b->ev_join[0], 0));
	} else {
		RUN(launch_clers(B, (const uint32_t *)(b->d_tables + b->o_order), (uint32_t)b->clers_order.size(), b->clers, tickets + 2, b->sms, s), !b->clers_order.empty());
	}
	if((rc = mark(b, "clers", k, s))) return rc;
	if(ovl2) {
		const uint2 *wd = (const uint2 *)(b->d_tables + b->o_w_delta);
		CU(cudaEventRecord(b->ev_fork[1], s));
		CU(cudaStreamWaitEvent(b->side, b->ev_fork[1], 0));
		RUN(launch_delta_mesh(B, wd, (uint32_t)b->n_delta_crit, b->delta_split, s), true);
		// side stream: the adjacency (faces only — no position is read) and the boundary scan, then the rest of the delta inverse
		RUN(launch_adj_build(B, t_faces, (uint32_t)b->t_faces.size(), 1, b->side), true);
		RUN(launch_scan_u32(B, t_vscan, (uint32_t)b->t_vscan.size(), st, tickets + 4, b->sms, b->side), b->any_border);
		RUN(launch_delta_mesh(B, wd + b->n_delta_crit, (uint32_t)(b->w_delta.size() - b->n_delta_crit), b->delta_split, b->side), true);
	} else
	RUN(launch_delta_mesh(B, (const uint2 *)(b->d_tables + b->o_w_delta), (uint32_t)b->w_delta.size(), b->delta_split, s), !b->w_delta.empty());
	if((rc = mark(b, "delta", k, s))) return rc;
	if(!b->t_faces.empty()) {                      // (the adjacency build reads the delta-decoded positions: it cannot move in front of the delta inverse)
		if(ovl2) {
			RUN(launch_adj_build(B, t_faces, (uint32_t)b->t_faces.size(), 2, s), true);      // face normals: the positions are final now
			CU(cudaEventRecord(b->ev_join[1], b->side));
			CU(cudaStreamWaitEvent(s, b->ev_join[1], 0));                                    // adjacency + boundary scan (+ the other attributes' delta)
			// the estimation reads face normals + adjacency, not the positions: the dequantisation (which turns the integer positions
			// into floats in place) runs beside it on the side stream, after the face normals have read them
			if(b->overlap & 4) {
				CU(cudaEventRecord(b->ev_fork[0], s));
				CU(cudaStreamWaitEvent(b->side, b->ev_fork[0], 0));
				RUN(launch_dequant(B, t_dequant, (uint32_t)b->t_dequant.size(), b->side), !b->t_dequant.empty());
			}
		} else {
			RUN(launch_adj_build(B, t_faces, (uint32_t)b->t_faces.size(), 0, s), true);
			RUN(launch_scan_u32(B, t_vscan, (uint32_t)b->t_vscan.size(), st, tickets + 4, b->sms, s), b->any_border);
		}
		st += b->t_vscan.size();
		RUN(launch_normal_estimate(B, t_verts, (uint32_t)b->t_verts.size(), s), true);
	}
	if((rc = mark(b, "normals", k, s))) return rc;
	if(ovl2 && (b->overlap & 4)) {
		CU(cudaEventRecord(b->ev_join[0], b->side));
		CU(cudaStreamWaitEvent(s, b->ev_join[0], 0));
	} else
	RUN(launch_dequant(B, t_dequant, (uint32_t)b->t_dequant.size(), s), !b->t_dequant.empty());
	if((rc = mark(b, "dequant", k, s))) return rc;
#undef RUN
	b->launches = launches;
	return CRT_OK;
}

extern "C" int crt_batch_launches(const crt_batch *b) { return b->launches; }

extern "C" int crt_batch_stage_times(crt_batch *b, const char **names, float *ms, int cap) {
	int n = 0;
	for(size_t i = 1; i < b->stages.size() && n < cap; i++) {
		float t = 0;
		if(cudaEventElapsedTime(&t, b->stages[i - 1].ev, b->stages[i].ev) != cudaSuccess) { cudaGetLastError(); t = -1; }
		names[n] = b->stages[i].name; ms[n] = t; n++;
	}
	return n;
}

extern "C" int crt_batch_status(crt_batch *b, int *per_mesh) {
	const size_t n = b->meshes.size();
	b->h_status.assign(n, 0);
	if(n) CU(cudaMemcpy(b->h_status.data(), b->d_zero + b->z_status, n*4, cudaMemcpyDeviceToHost));
	int first = CRT_OK;
	for(size_t i = 0; i < n; i++) {
		if(per_mesh) per_mesh[i] = b->h_status[i];
		if(first == CRT_OK && b->h_status[i]) { first = b->h_status[i]; fail(first, "mesh " + std::to_string(i) + ": Decoding topology failed"); }
	}
	return first;
}

extern "C" int crt_batch_debug_clers(crt_batch *b, int i, unsigned char *out, uint32_t cap, uint32_t *nout) {
	if(!b->uploaded || i < 0 || i >= (int)b->meshes.size()) return fail(CRT_E_ARG, "bad mesh index");
	const MeshDesc &M = b->h_mesh[i];
	if(M.clers_tun < 0) { *nout = 0; return CRT_OK; }
	const TunDesc &td = b->h_tun[M.clers_tun];
	*nout = td.size;
	CU(cudaMemcpy(out, b->d_symbols + td.out_off, std::min(cap, td.size), cudaMemcpyDeviceToHost));
	return CRT_OK;
}

extern "C" int crt_batch_debug_prediction(crt_batch *b, int i, uint32_t *out) {
	if(!b->uploaded || i < 0 || i >= (int)b->meshes.size()) return fail(CRT_E_ARG, "bad mesh index");
	const MeshDesc &M = b->h_mesh[i];
	if(!M.nface) return CRT_OK;
	std::vector<uint32_t> tmp((size_t)M.nvert*4);
	CU(cudaMemcpy(tmp.data(), (const void *)M.pred_ptr, tmp.size()*4, cudaMemcpyDeviceToHost));
	for(uint32_t v = 0; v < M.nvert; v++) { out[v*3] = tmp[v*4]; out[v*3 + 1] = tmp[v*4 + 1]; out[v*3 + 2] = tmp[v*4 + 2]; }
	return CRT_OK;
}

extern "C" int crt_shard_lpt(int n, const uint32_t *nvert, const uint32_t *nface, const uint32_t *nattr, int world, int *rank_of) {
	if(n < 0 || world < 1 || !rank_of) return fail(CRT_E_ARG, "bad arguments");
	std::vector<int> order(n);
	std::vector<double> cost(n);
	for(int i = 0; i < n; i++) { order[i] = i; cost[i] = 4.0*(nface ? nface[i] : 0) + 1.0*(double)nvert[i]*(nattr ? nattr[i] : 1); }
	std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
	std::vector<double> load(world, 0.0);
	for(int i: order) {
		int best = 0;
		for(int r = 1; r < world; r++) if(load[r] < load[best]) best = r;
		rank_of[i] = best; load[best] += cost[i];
	}
	return CRT_OK;
}

// =========================================================================================================
// single decoder with host buffers
// =========================================================================================================
struct HostBind { void *ptr; int format; int components; bool typed = false; };

struct crt_decoder {
	ParsedMesh pm;
	std::map<std::string, HostBind> binds;
	void *index = nullptr; int index16 = 0;
	std::vector<uint32_t> group_ends; std::vector<Props> group_props;
	bool groups_known = false;
	int normal_prediction = 0;
	int qc[4] = {4, 4, 4, 8};
};

static void ensure_groups(crt_decoder *d) {
	// the reference fills index.groups inside decode() (decoder.cpp:165); the shims read them after decode().
	// Here the group table is parsed on first use so ngroups()/groups() also work before decode().
	if(d->groups_known) return;
	std::string err;
	if(walk_directory(d->pm, err) == CRT_OK) { d->group_ends = d->pm.group_ends; d->group_props = d->pm.group_props; }
	d->groups_known = true;
}

extern "C" crt_decoder *crt_new_decoder(int len, const unsigned char *buffer) {
	crt_decoder *d = new crt_decoder();
	std::string err;
	int rc = parse_header(buffer, len, d->pm, err);
	if(rc) { fail(rc, err); delete d; return nullptr; }
	return d;
}
extern "C" void crt_delete_decoder(crt_decoder *d) { delete d; }
extern "C" uint32_t crt_nvert(const crt_decoder *d) { return d->pm.nvert; }
extern "C" uint32_t crt_nface(const crt_decoder *d) { return d->pm.nface; }
extern "C" int crt_ngroups(const crt_decoder *d) { ensure_groups((crt_decoder *)d); return (int)d->group_ends.size(); }
extern "C" void crt_groups(const crt_decoder *d, int *ends) { ensure_groups((crt_decoder *)d); for(size_t i = 0; i < d->group_ends.size(); i++) ends[i] = (int)d->group_ends[i]; }
extern "C" int crt_group_nprops(const crt_decoder *d, int g) { ensure_groups((crt_decoder *)d); return (g < 0 || g >= (int)d->group_props.size()) ? 0 : (int)d->group_props[g].size(); }
extern "C" const char *crt_group_prop(const crt_decoder *d, int g, int i, const char **value) {
	ensure_groups((crt_decoder *)d);
	if(g < 0 || g >= (int)d->group_props.size() || i < 0 || i >= (int)d->group_props[g].size()) return nullptr;
	if(value) *value = d->group_props[g][i].second.c_str();
	return d->group_props[g][i].first.c_str();
}
extern "C" int crt_nexif(const crt_decoder *d) { return (int)d->pm.exif.size(); }
extern "C" const char *crt_exif(const crt_decoder *d, int i, const char **value) {
	if(i < 0 || i >= (int)d->pm.exif.size()) return nullptr;
	if(value) *value = d->pm.exif[i].second.c_str();
	return d->pm.exif[i].first.c_str();
}
extern "C" int crt_has_attr(const crt_decoder *d, const char *name) { return d->pm.find(name) >= 0; }
extern "C" int crt_nattr(const crt_decoder *d) { return (int)d->pm.attrs.size(); }
extern "C" const char *crt_attr_info(const crt_decoder *d, int i, int *codec, float *q, int *components, int *format, int *strategy) {
	if(i < 0 || i >= (int)d->pm.attrs.size()) return nullptr;
	const ParsedAttr &a = d->pm.attrs[i];
	if(codec) *codec = a.codec; if(q) *q = a.q; if(components) *components = a.N; if(format) *format = a.format; if(strategy) *strategy = a.strategy;
	return a.name.c_str();
}

extern "C" int crt_set_attribute(crt_decoder *d, const char *name, char *buffer, int format) {   // decoder.cpp:96-102
	if(d->pm.find(name) < 0) return 0;
	int comps = d->pm.attrs[d->pm.find(name)].N;
	auto it = d->binds.find(name);
	if(it != d->binds.end()) comps = it->second.components;
	HostBind hb; hb.ptr = buffer; hb.format = format; hb.components = comps;
	d->binds[name] = hb;
	return 1;
}
static int set_typed(crt_decoder *d, const char *name, void *b, int format) {
	const int rc = crt_set_attribute(d, name, (char *)b, format);
	if(rc) d->binds[name].typed = true;
	return rc;
}
extern "C" int crt_set_positions(crt_decoder *d, float *b) { return set_typed(d, "position", b, CRT_FLOAT); }
extern "C" int crt_set_normals32(crt_decoder *d, float *b) { return set_typed(d, "normal", b, CRT_FLOAT); }
extern "C" int crt_set_normals16(crt_decoder *d, int16_t *b) { return set_typed(d, "normal", b, CRT_INT16); }
extern "C" int crt_set_uvs(crt_decoder *d, float *b) { return set_typed(d, "uv", b, CRT_FLOAT); }
extern "C" int crt_set_colors(crt_decoder *d, unsigned char *b, int components) {               // decoder.cpp:116-123
	if(d->pm.find("color") < 0) return 0;
	HostBind hb; hb.ptr = b; hb.format = CRT_UINT8; hb.components = components;
	d->binds["color"] = hb;
	return 1;
}
extern "C" void crt_set_index32(crt_decoder *d, uint32_t *b) { d->index = b; d->index16 = 0; }
extern "C" void crt_set_index16(crt_decoder *d, uint16_t *b) { d->index = b; d->index16 = 1; }
extern "C" int crt_normal_prediction(const crt_decoder *d) { return d->normal_prediction; }
extern "C" void crt_color_q(const crt_decoder *d, int qc[4]) { for(int k = 0; k < 4; k++) qc[k] = d->qc[k]; }

extern "C" int crt_decode(crt_decoder *d) {
	if(!crt_device_available()) return CRT_E_CUDA;
	// setPositions / setUvs / setNormals take float[3 nvert] / float[2 nvert] / [3 nvert] arrays (decoder.h:51-55): a header that
	// announces another component count would write past them (the reference does, it has no checks)
	for(const ParsedAttr &pa: d->pm.attrs) {
		auto it = d->binds.find(pa.name);
		if(it == d->binds.end() || !it->second.ptr || !it->second.typed) continue;
		const int want = pa.name == "uv" ? 2 : 3;
		if(pa.N != want) return fail(CRT_E_LIMIT, "attribute '" + pa.name + "' has " + std::to_string(pa.N) + " components, the typed setter binds " + std::to_string(want));
	}
	const unsigned char *blob = d->pm.blob;
	int len = (int)d->pm.len;
	crt_batch *b = crt_batch_create(1, &blob, &len);
	if(!b) return g_code ? g_code : CRT_E_TRUNCATED;
	d->group_ends = b->meshes[0].group_ends; d->group_props = b->meshes[0].group_props; d->groups_known = true;
	const ParsedMesh &pm = b->meshes[0];
	struct Out { void *dev; void *host; size_t bytes; };
	std::vector<Out> outs;
	int rc = CRT_OK;
	auto cleanup = [&]() { for(auto &o: outs) dev_free(o.dev); crt_batch_destroy(b); };
	for(size_t a = 0; a < pm.attrs.size() && rc == CRT_OK; a++) {
		const ParsedAttr &pa = pm.attrs[a];
		if(pa.codec == CODEC_NORMAL) d->normal_prediction = pm.streams[a].prediction;
		if(pa.codec == CODEC_COLOR) for(int k = 0; k < 4; k++) d->qc[k] = pm.streams[a].qc[k];
		auto it = d->binds.find(pa.name);
		if(it == d->binds.end() || !it->second.ptr) continue;
		const HostBind &hb = it->second;
		size_t stride;
		if(pa.codec == CODEC_NORMAL) stride = hb.format == CRT_INT16 ? 6 : 12;
		else if(pa.codec == CODEC_COLOR) stride = (size_t)hb.components;
		else stride = (size_t)pa.N*4;
		Out o{nullptr, hb.ptr, stride*pm.nvert};
		cudaError_t e = dev_alloc((void **)&o.dev, o.bytes + 16);
		if(e != cudaSuccess) { rc = cuda_fail(e, "cudaMalloc(output)"); break; }
		outs.push_back(o);
		// the caller's buffer content is preserved where the reference leaves elements untouched
		if(pa.codec == CODEC_NORMAL && hb.format == CRT_INT16) cudaMemcpy(o.dev, o.host, o.bytes, cudaMemcpyHostToDevice);
		else cudaMemsetAsync(o.dev, 0, o.bytes, nullptr);   // elements no kernel writes (a stream shorter than nvert) come back as 0, never as stale device memory
		rc = crt_batch_bind(b, pa.name.c_str(), o.dev, hb.format, hb.components);
	}
	if(rc == CRT_OK && pm.nface && d->index) {
		Out o{nullptr, d->index, (size_t)pm.nface*3*(d->index16 ? 2 : 4)};
		cudaError_t e = dev_alloc((void **)&o.dev, o.bytes + 16);
		if(e != cudaSuccess) rc = cuda_fail(e, "cudaMalloc(index)");
		else { outs.push_back(o); cudaMemsetAsync(o.dev, 0, o.bytes, nullptr); rc = crt_batch_bind(b, "index", o.dev, d->index16 ? CRT_UINT16 : CRT_UINT32, 0); }   // (faces past the last group end: 0)
	}
	if(rc == CRT_OK) rc = crt_batch_upload(b, nullptr);
	if(rc == CRT_OK) rc = crt_batch_decode(b, nullptr);
	if(rc == CRT_OK) { cudaError_t e = cudaStreamSynchronize(nullptr); if(e != cudaSuccess) rc = cuda_fail(e, "decode kernels"); }
	if(rc == CRT_OK) rc = crt_batch_status(b, nullptr);
	if(rc == CRT_OK) for(auto &o: outs) {
		cudaError_t e = cudaMemcpy(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost);
		if(e != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpy(D2H)"); break; }
	}
	std::string keep = g_err;
	cleanup();
	if(rc != CRT_OK) g_err = keep;
	return rc;
}

// =========================================================================================================
// the reference shims' own names (emcorto.cpp:14-89, corto_codec.h:41-43)
// =========================================================================================================
extern "C" crt_decoder *newDecoder(int n, const unsigned char *buffer) { return crt_new_decoder(n, buffer); }
extern "C" void deleteDecoder(crt_decoder *d) { crt_delete_decoder(d); }
extern "C" int ngroups(crt_decoder *d) { return crt_ngroups(d); }
extern "C" void groups(crt_decoder *d, int *ends) { crt_groups(d, ends); }
extern "C" int nvert(crt_decoder *d) { return (int)crt_nvert(d); }
extern "C" int nface(crt_decoder *d) { return (int)crt_nface(d); }
extern "C" int hasAttr(crt_decoder *d, const char *attr) { return crt_has_attr(d, attr); }
extern "C" int hasNormal(crt_decoder *d) { return crt_has_attr(d, "normal"); }
extern "C" int hasColor(crt_decoder *d) { return crt_has_attr(d, "color"); }
extern "C" int hasUv(crt_decoder *d) { return crt_has_attr(d, "uv"); }
extern "C" void setPositions(crt_decoder *d, float *b) { crt_set_positions(d, b); }
extern "C" void setNormals32(crt_decoder *d, float *b) { crt_set_normals32(d, b); }
extern "C" void setNormals16(crt_decoder *d, int16_t *b) { crt_set_normals16(d, b); }
extern "C" void setColors(crt_decoder *d, unsigned char *b, int components) { crt_set_colors(d, b, components); }
extern "C" void setUvs(crt_decoder *d, float *b) { crt_set_uvs(d, b); }
extern "C" void setIndex16(crt_decoder *d, uint16_t *b) { crt_set_index16(d, b); }
extern "C" void setIndex32(crt_decoder *d, uint32_t *b) { crt_set_index32(d, b); }
extern "C" void decode(crt_decoder *d) { crt_decode(d); }

extern "C" crt_decoder *CreateDecoder(int length, unsigned char *data, crt_Vector2 *info) {
	crt_decoder *d = crt_new_decoder(length, data);
	if(d && info) { info->x = (float)d->pm.nface; info->y = (float)d->pm.nvert; }   // corto_codec.cpp:11-14
	return d;
}
extern "C" void DestroyDecoder(crt_decoder *d) { crt_delete_decoder(d); }
extern "C" int DecodeMesh(crt_decoder *d, crt_Vector3 *vertices, int *indices, crt_Vector3 *normals, crt_Color *colors, crt_Vector2 *texcoord) {
	if(d->pm.nface == 0) return -1;                                                  // corto_codec.cpp:27-30
	crt_set_index32(d, (uint32_t *)indices);
	if(d->pm.nvert > 0) crt_set_positions(d, (float *)vertices);
	if(crt_has_attr(d, "normal")) crt_set_normals32(d, (float *)normals);
	std::vector<unsigned char> rgba;
	if(crt_has_attr(d, "color") && colors) { rgba.resize((size_t)d->pm.nvert*4); crt_set_colors(d, rgba.data(), 4); }
	if(crt_has_attr(d, "uv")) crt_set_uvs(d, (float *)texcoord);
	if(crt_decode(d) != CRT_OK) return -1;
	for(size_t i = 0; i < rgba.size()/4; i++) {
		colors[i].r = rgba[i*4]/255.0f; colors[i].g = rgba[i*4 + 1]/255.0f; colors[i].b = rgba[i*4 + 2]/255.0f; colors[i].a = rgba[i*4 + 3]/255.0f;
	}
	return (int)d->pm.nface;
}
