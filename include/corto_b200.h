/* corto_b200.h — C ABI of the B200-native corto DECODE path (libcorto_b200.so).
 *
 * Plain C: pointers and sizes only, no C++ / torch types.  Three groups of entry points:
 *
 *  (1) crt_*            single-decoder API with HOST buffers, a superset of the two C shims the reference
 *                       ships (WASM: html/js/emscripten/emcorto.cpp:14-89, Unity: src/corto_codec.h:41-43).
 *                       crt_decode() = H2D of the blob, the CUDA kernels, D2H into the bound host arrays,
 *                       synchronous — the semantics of crt::Decoder::decode() (src/decoder.cpp:126-131).
 *  (2) reference-named  newDecoder/.../decode and CreateDecoder/DestroyDecoder/DecodeMesh: the exact
 *                       symbols the reference bindings (post.js ccall, unity/CortoMeshLoader.cs P/Invoke)
 *                       resolve, so those bindings load this library unchanged.  See INTEGRATION.md.
 *  (3) crt_batch_*      NEW: batched, device-resident decode of many independent .crt blobs — what
 *                       bench.py times.  Blobs go H2D once, outputs stay in HBM in flat arenas.
 *
 * Errors: functions returning int give CRT_OK (0) or a negative CRT_E_* code; crt_last_error() returns a
 * thread-local message.  The reference throws `const char*` (src/decoder.cpp:44,51,274) — the C++ facade
 * (include/corto_b200/decoder.h) rethrows these messages to stay source compatible.
 * There is NO CPU fallback: without a CUDA device every decode call fails with CRT_E_CUDA.
 */
#ifndef CORTO_B200_H
#define CORTO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* crt::VertexAttribute::Format — include/corto/vertex_attribute.h:32 */
enum { CRT_UINT32 = 0, CRT_INT32, CRT_UINT16, CRT_INT16, CRT_UINT8, CRT_INT8, CRT_FLOAT, CRT_DOUBLE };
/* crt::NormalAttr::Prediction — include/corto/normal_attribute.h:42-44 */
enum { CRT_NORMAL_DIFF = 0, CRT_NORMAL_ESTIMATED = 1, CRT_NORMAL_BORDER = 2 };
/* crt::VertexAttribute::CODEC — include/corto/vertex_attribute.h:34 */
enum { CRT_GENERIC_CODEC = 1, CRT_NORMAL_CODEC = 2, CRT_COLOR_CODEC = 3 };

enum {
	CRT_OK = 0,
	CRT_E_ALIGN = -1,       /* "Memory must be alignegned on 4 bytes."  decoder.cpp:43-44 */
	CRT_E_MAGIC = -2,       /* "Not a crt file."                         decoder.cpp:50-51 */
	CRT_E_TRUNCATED = -3,   /* a block runs past the end of the blob (the reference has no bounds checks) */
	CRT_E_ENTROPY = -4,     /* "Unknown entropy"                         cstream.cpp:81-85 */
	CRT_E_TOPOLOGY = -5,    /* "Decoding topology failed"                decoder.cpp:272-275 */
	CRT_E_FORMAT = -6,      /* unsupported output format                 normal_attribute.cpp:249-253, color_attribute.cpp:112 */
	CRT_E_NOPOSITION = -7,  /* "No position attribute found..."          normal_attribute.cpp:219-221 */
	CRT_E_LIMIT = -8,       /* more attributes / components than this build handles */
	CRT_E_CUDA = -9,        /* CUDA runtime error or no device */
	CRT_E_ARG = -10
};

const char *crt_last_error(void);
/* 1 if a CUDA device is usable by this library, else 0 (message in crt_last_error). */
int crt_device_available(void);

/* ------------------------------------------------------------------------------------------------ */
/* (1) single decoder, host buffers                                                                  */
typedef struct crt_decoder crt_decoder;

/* Header parse only, like crt::Decoder::Decoder (src/decoder.cpp:41-89).  `buffer` is BORROWED (cstream.h:226)
 * and must stay valid and 4-byte aligned until crt_decode returns.  NULL on error. */
crt_decoder *crt_new_decoder(int len, const unsigned char *buffer);
void crt_delete_decoder(crt_decoder *d);

uint32_t crt_nvert(const crt_decoder *d);
uint32_t crt_nface(const crt_decoder *d);
int crt_ngroups(const crt_decoder *d);
void crt_groups(const crt_decoder *d, int *ends);                         /* emcorto.cpp:22-27 */
int crt_group_nprops(const crt_decoder *d, int group);
const char *crt_group_prop(const crt_decoder *d, int group, int i, const char **value);
int crt_nexif(const crt_decoder *d);
const char *crt_exif(const crt_decoder *d, int i, const char **value);
int crt_has_attr(const crt_decoder *d, const char *name);                 /* decoder.h:47 */
int crt_nattr(const crt_decoder *d);
/* attribute i in wire (= std::map name) order: name, codec, q, N, header format, strategy (decoder.cpp:62-70) */
const char *crt_attr_info(const crt_decoder *d, int i, int *codec, float *q, int *components, int *format, int *strategy);

/* Bind outputs (decoder.h:49-61).  Return 1 if the attribute exists, else 0 — like the reference's bool.  The typed setters
 * promise float[3 nvert] (positions, normals) / float[2 nvert] (uvs) arrays: crt_decode fails with CRT_E_LIMIT instead of
 * writing past them when the header announces another component count (use crt_set_attribute + crt_attr_info for those). */
int crt_set_positions(crt_decoder *d, float *buffer);
int crt_set_normals32(crt_decoder *d, float *buffer);
int crt_set_normals16(crt_decoder *d, int16_t *buffer);
int crt_set_uvs(crt_decoder *d, float *buffer);
int crt_set_colors(crt_decoder *d, unsigned char *buffer, int components);
int crt_set_attribute(crt_decoder *d, const char *name, char *buffer, int format);
void crt_set_index32(crt_decoder *d, uint32_t *buffer);
void crt_set_index16(crt_decoder *d, uint16_t *buffer);

/* Decode into the bound host arrays.  After success crt_normal_prediction / crt_color_q report what the
 * stream carried (NormalAttr::prediction normal_attribute.cpp:179, ColorAttr::qc color_attribute.h:56-57). */
int crt_decode(crt_decoder *d);
int crt_normal_prediction(const crt_decoder *d);
void crt_color_q(const crt_decoder *d, int qc[4]);

/* ------------------------------------------------------------------------------------------------ */
/* (2) the reference shims' own symbol names                                                         */
crt_decoder *newDecoder(int n, const unsigned char *buffer);               /* emcorto.cpp:14 */
void deleteDecoder(crt_decoder *d);                                        /* emcorto.cpp:87 */
int ngroups(crt_decoder *d);                                               /* emcorto.cpp:18 */
void groups(crt_decoder *d, int *ends);                                    /* emcorto.cpp:22 */
int nvert(crt_decoder *d);                                                 /* emcorto.cpp:29 */
int nface(crt_decoder *d);                                                 /* emcorto.cpp:33 */
int hasAttr(crt_decoder *d, const char *attr);                             /* emcorto.cpp:37 */
int hasNormal(crt_decoder *d);                                             /* emcorto.cpp:41 */
int hasColor(crt_decoder *d);                                              /* emcorto.cpp:45 */
int hasUv(crt_decoder *d);                                                 /* emcorto.cpp:49 */
void setPositions(crt_decoder *d, float *buffer);                          /* emcorto.cpp:55 */
void setNormals32(crt_decoder *d, float *buffer);                          /* emcorto.cpp:59 */
void setNormals16(crt_decoder *d, int16_t *buffer);                        /* emcorto.cpp:63 */
void setColors(crt_decoder *d, unsigned char *buffer, int components);     /* emcorto.cpp:67 */
void setUvs(crt_decoder *d, float *buffer);                                /* emcorto.cpp:71 */
void setIndex16(crt_decoder *d, uint16_t *buffer);                         /* emcorto.cpp:75 */
void setIndex32(crt_decoder *d, uint32_t *buffer);                         /* emcorto.cpp:79 */
void decode(crt_decoder *d);                                               /* emcorto.cpp:83 */

typedef struct { float r, g, b, a; } crt_Color;                            /* corto_codec.h:19-25 */
typedef struct { float x, y; } crt_Vector2;                                /* corto_codec.h:26-30 */
typedef struct { float x, y, z; } crt_Vector3;                             /* corto_codec.h:31-36 */
crt_decoder *CreateDecoder(int length, unsigned char *data, crt_Vector2 *decoderInfo);   /* corto_codec.h:41 */
void DestroyDecoder(crt_decoder *decoder);                                              /* corto_codec.h:42 */
/* returns nface, or -1 for a point cloud (corto_codec.cpp:27-30).  The reference's Color(FLOAT) output is
 * broken upstream; here colours come out as r,g,b,a in [0,1] = u8/255. */
int DecodeMesh(crt_decoder *decoder, crt_Vector3 *vertices, int *indices, crt_Vector3 *normals, crt_Color *colors,
               crt_Vector2 *texcoord);                                                  /* corto_codec.h:43 */

/* ------------------------------------------------------------------------------------------------ */
/* (3) batched, device-resident decode                                                               */
typedef struct crt_batch crt_batch;

/* Attribute presence mask bits reported by crt_batch_mesh_info. */
enum { CRT_HAS_POSITION = 1, CRT_HAS_NORMAL = 2, CRT_HAS_COLOR = 4, CRT_HAS_UV = 8, CRT_HAS_OTHER = 16, CRT_HAS_INDEX = 32 };

/* Parse the headers and walk the stream directory of `n` HOST blobs (each 4-byte aligned; borrowed until
 * crt_batch_upload returns).  No GPU work.  NULL on error. */
crt_batch *crt_batch_create(int n, const unsigned char *const *blobs, const int *lens);
void crt_batch_destroy(crt_batch *b);

/* Multi-GPU ingest: the rank that holds a blob in HOST memory records the few hundred bytes the header
 * parse and the directory walk read from it (crt_walk_tape: no payload byte, returns the tape length or a CRT_E_* code; the
 * tape is written when it fits `cap`; nvert / nface / nattr feed crt_shard_lpt); the rank that receives the blob in DEVICE
 * memory builds its batch from the tapes and the arena the blobs arrived in (blob i at the sum of the 16-byte-rounded lengths
 * before it, arena 16-byte aligned, borrowed while the batch lives).  No payload ever returns to a host. */
int crt_walk_tape(const unsigned char *blob, int len, unsigned char *tape, int cap, uint32_t *nvert, uint32_t *nface, uint32_t *nattr);
crt_batch *crt_batch_create_device(int n, const unsigned char *const *tapes, const int *tape_lens, const int *blob_lens, const void *arena);
/* FNV-1a over every offset / size of the walked directory (and the bindings): equal for a batch built from host blobs and one
 * built from their tapes. */
uint64_t crt_batch_directory_signature(const crt_batch *b);

int crt_batch_count(const crt_batch *b);
int crt_batch_mesh_info(const crt_batch *b, int i, uint32_t *nvert, uint32_t *nface, uint32_t *attr_mask);
/* Component count the headers give attribute `name` (0 = no mesh of the batch carries it): an arena bound under that name
 * holds components * total_verts elements.  Meshes of one batch must agree on it (crt_batch_upload fails with CRT_E_LIMIT
 * otherwise); ESTIMATED / BORDER normals additionally need a 3-component "position". */
int crt_batch_attr_components(const crt_batch *b, const char *name);
uint64_t crt_batch_total_verts(const crt_batch *b);
uint64_t crt_batch_total_faces(const crt_batch *b);
uint64_t crt_batch_total_bytes(const crt_batch *b);     /* sum of blob lengths */
/* Prefix arrays (n+1 entries) locating mesh i inside every output arena: vertices [vert_base[i], vert_base[i+1]),
 * faces [face_base[i], face_base[i+1]). */
const uint64_t *crt_batch_vert_base(const crt_batch *b);
const uint64_t *crt_batch_face_base(const crt_batch *b);

/* Output arenas are flat DEVICE arrays, meshes concatenated in batch order:
 *   "position" float[3*V]  "uv" float[2*V]  "normal" float[3*V] (or int16[3*V])  "color" u8[components*V]
 *   "index" u32[3*F] (or u16[3*F]; vertex ids are per mesh, not rebased)  custom generic attrs: float[N*V]
 * Meshes lacking an attribute leave their slice untouched.  format: CRT_FLOAT / CRT_INT16 (normal) /
 * CRT_UINT8 (color) / CRT_UINT32|CRT_UINT16 (index) / CRT_INT32|CRT_UINT32 (generic, integer dequantise).
 * `components` is used by "color" only (3 or 4, crt::Decoder::setColors decoder.cpp:116-123).
 * Unbound attributes are parsed and skipped. */
int crt_batch_bind(crt_batch *b, const char *name, void *device_ptr, int format, int components);

/* Copy blobs + directory to the device and allocate scratch (async on `stream`, a cudaStream_t passed as void*). */
int crt_batch_upload(crt_batch *b, void *stream);
/* Re-run the host directory walk and re-send the (small) directory; blobs stay resident.  bench.py calls this
 * inside the timed step so that no part of Decoder::decode's work is hoisted out of the measurement. */
int crt_batch_rewalk(crt_batch *b, void *stream);
/* Launch the decode kernels (async on `stream`).  Outputs are complete when the stream reaches this point. */
int crt_batch_decode(crt_batch *b, void *stream);
/* After the stream has been synchronised: per-mesh status (CRT_OK or CRT_E_TOPOLOGY ...). Returns the first
 * failing code or CRT_OK. */
int crt_batch_status(crt_batch *b, int *per_mesh /* n entries or NULL */);
/* Number of kernel launches issued by the last crt_batch_decode (for bench.py's gpu_launches). */
int crt_batch_launches(const crt_batch *b);
/* CUDA-event timing of the last decode, per stage, milliseconds (call after a stream sync).  names/ms arrays of
 * `cap` entries; returns the number of stages.  Enabled with crt_batch_set_profiling(b, 1). */
int crt_batch_set_profiling(crt_batch *b, int on);
int crt_batch_stage_times(crt_batch *b, const char **names, float *ms, int cap);
/* Debug pins (device -> host copies of intermediates of mesh i; synchronous): clers bytes, prediction triples. */
int crt_batch_debug_clers(crt_batch *b, int i, unsigned char *out, uint32_t cap, uint32_t *n);
int crt_batch_debug_prediction(crt_batch *b, int i, uint32_t *out /* nvert*3 */);

/* Sharding helper for multi-GPU runs: longest-processing-time greedy over the cost model
 * cost = nface*alpha + nvert*nattr*beta.  Writes the rank (0..world-1) of each of the n blobs. No GPU work. */
int crt_shard_lpt(int n, const uint32_t *nvert, const uint32_t *nface, const uint32_t *nattr, int world, int *rank_of);

#ifdef __cplusplus
}
#endif
#endif /* CORTO_B200_H */
