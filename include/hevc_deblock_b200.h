/*
 * hevc_deblock_b200.h -- C ABI of the B200 (sm_100a) HEVC-style deblocking filter.
 *
 * Drop-in boundary for the deblocking path of RomanKazantsev/gpu_video_codec
 * (directory hevc_deblocking_filter/).  Reference files cited below:
 *   cpu.h   = hevc_deblocking_filter_cpu.h   (class ReadYuvFrame: the CPU filter and the data contract)
 *   gpu.cu  = hevc_deblocking_filter_gpu.cu  (ExecuteGpu: the GPU entry point being replaced)
 *   main.cu = main.cu                        (the only caller of both)
 *
 * Data contract (identical to the reference):
 *   - frames are 8-bit YV12: a luma plane of width*height samples followed by two
 *     chroma planes of (width/2)*(height/2) samples, rows tightly packed (cpu.h:66-82);
 *   - width and height must be multiples of 8 (cpu.h:46-48);
 *   - one quantisation parameter per frame (cpu.h:35-37), values above 51 act as 51
 *     (cpu.h:1064-1072);
 *   - luma boundary strengths in the reference's two arrays (cpu.h:86-99):
 *       vert_bs[(height/8) rows][(width/8 + 1) edges]   num_vert_bs = (width/8 + 1) * height / 8
 *       hor_bs [(height/8 + 1) edge rows][(width/8)]     num_hor_bs  = (height/8 + 1) * width / 8
 *     an edge is filtered when its strength is > 0 (cpu.h:164).  NULL selects the
 *     reference constructor's defaults (cpu.h:92-99).  Sizes must match exactly
 *     (cpu.h:122-123).  Chroma strengths are not caller-settable in the reference
 *     and are not here (always the defaults of cpu.h:110-117);
 *   - results are bit-identical to ReadYuvFrame::DeblockingFilter(1) (cpu.h:134)
 *     run with zero plane padding and zero-extended chroma strength arrays.
 *
 * Errors: the reference throws string literals (cpu.h:44, 47, 123) or prints and
 * ignores CUDA errors (gpu.cu:1104-1169, 1271-1288).  Here every entry point
 * returns an int status and never throws; hdb_last_error() gives the text, which
 * for the three reference conditions is the reference's own message.
 *
 * There is no CPU fallback: every entry point that filters requires a CUDA
 * device of compute capability 10.x and fails with HDB_ERR_CUDA otherwise.
 */
#ifndef HEVC_DEBLOCK_B200_H
#define HEVC_DEBLOCK_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HDB_OK               0
#define HDB_ERR_INVALID_ARG  1  /* NULL where a pointer is required, n == 0, ... */
#define HDB_ERR_FILE_SIZE    2  /* "Incorrect file size"                                   cpu.h:43-45 */
#define HDB_ERR_DIMENSIONS   3  /* "Width and height of image must be multiplier of ..."   cpu.h:46-48 */
#define HDB_ERR_BS_SIZE      4  /* "Incorrect size of input boundary strenght array"       cpu.h:122-123 */
#define HDB_ERR_CUDA         5  /* CUDA runtime / driver failure, or no sm_100 device */
#define HDB_ERR_IO           6  /* cannot open / read / write a file */
#define HDB_ERR_NO_MEMORY    7

typedef struct hdb_context hdb_context;   /* one GPU: device memory, pinned staging, streams */
typedef struct hdb_multi hdb_multi;       /* several GPUs of one box */

/*
 * Timing of one call, the three figures the reference prints (gpu.cu:1292, 1302-1303)
 * plus byte counts.  Device times come from CUDA events.
 */
/*
 * A call that fits one small slice (<= 4 MB) on page-locked, mapped host buffers is served without
 * the copy engines: the kernel reads and writes host memory itself.  copy_ms is then 0 and kernel_ms
 * covers the transfers as well, because they happen inside the kernel.
 */
typedef struct hdb_timing {
	double kernel_ms;          /* "Execution Time without copy on GPU"  gpu.cu:1266-1292 */
	double copy_ms;            /* "Copy Operation Time"  (H2D + D2H busy time)  gpu.cu:1246-1258, 1294-1303 */
	double total_ms;           /* "Execution Time with copy on GPU": wall time of the whole call */
	unsigned long long h2d_bytes, d2h_bytes;
	int kernel_launches;       /* filter-kernel launches made by the call */
	int path;                  /* 1 = tiled kernel, TMA loads; 3 = tiled kernel, cp.async loads (width not a multiple of 32);
	                              2 = padded-scratch kernels (width below 32 or unaligned buffers) */
} hdb_timing;

typedef struct hdb_device_info {  /* what GetGpuDeviceInfo prints, main.cu:92-107 */
	char name[256];
	int cc_major, cc_minor, sm_count, warp_size, max_threads_per_block;
	unsigned long long total_global_mem, total_const_mem, shared_mem_per_block;
} hdb_device_info;

/* ---- context ---------------------------------------------------------- */
int hdb_create(hdb_context **ctx, int device);
void hdb_destroy(hdb_context *ctx);
const char *hdb_last_error(const hdb_context *ctx);      /* ctx may be NULL: last creation error */
const char *hdb_status_string(int status);
int hdb_get_device_info(const hdb_context *ctx, hdb_device_info *info);  /* main.cu:92-107 */

/* ---- the data contract, host side -------------------------------------- */
unsigned hdb_num_vert_bs(unsigned width, unsigned height);   /* cpu.h:86 */
unsigned hdb_num_hor_bs(unsigned width, unsigned height);    /* cpu.h:87 */
/* constructor defaults, cpu.h:92-99; either pointer may be NULL */
int hdb_default_boundary_strength(unsigned width, unsigned height, unsigned char *vert_bs, unsigned char *hor_bs);
int hdb_beta(unsigned qp);   /* cpu.h:1064-1067 */
int hdb_tc(unsigned qp);     /* cpu.h:1069-1072 */

/* pinned host memory, as the reference allocates for its planes (gpu.cu:1103-1111) */
void *hdb_host_alloc(size_t bytes);
void hdb_host_free(void *p);

/* ---- one frame, host buffers ------------------------------------------- */
/*
 * Replaces  ReadYuvFrame(file, w, h, Qp) [cpu.h:35] -> SetBoundaryStrenght(...) [cpu.h:120]
 *           -> DeblockingFilter(n) [cpu.h:134] -> Save(file) [cpu.h:995]
 * and the GPU twin ExecuteGpu(...) [gpu.cu:1230], with buffers instead of files.
 * yuv_in / yuv_out: one tight YV12 frame, 3*width*height/2 bytes; may alias.
 */
int hdb_deblock_frame(hdb_context *ctx, const unsigned char *yuv_in, unsigned char *yuv_out,
	unsigned width, unsigned height, unsigned qp,
	const unsigned char *vert_bs, unsigned num_vert_bs,
	const unsigned char *hor_bs, unsigned num_hor_bs, hdb_timing *timing);

/* Same with the three planes given separately (the reference keeps ext_Y / ext_U / ext_V, cpu.h:1041-1048). */
int hdb_deblock_planes(hdb_context *ctx,
	const unsigned char *y_in, const unsigned char *u_in, const unsigned char *v_in,
	unsigned char *y_out, unsigned char *u_out, unsigned char *v_out,
	unsigned width, unsigned height, unsigned qp,
	const unsigned char *vert_bs, unsigned num_vert_bs,
	const unsigned char *hor_bs, unsigned num_hor_bs, hdb_timing *timing);

/*
 * File in, file out: ExecuteGpu(input_file_name, output_file_name, width, height, Qp, ...)
 * of main.cu:87-90 / gpu.cu:1230-1306 (the four block-dimension arguments have no
 * meaning for this implementation and are dropped).
 */
int hdb_execute_gpu(hdb_context *ctx, const char *input_file_name, const char *output_file_name,
	unsigned width, unsigned height, unsigned qp, hdb_timing *timing);

/* ---- a batch of frames, host buffers (pinned multi-stream pipeline) ----- */
/*
 * n_frames frames of one size back to back (frame f at frames + f*3*width*height/2).
 * qp[n_frames]; vert_bs[n_frames][num_vert_bs], hor_bs[n_frames][num_hor_bs] or both NULL
 * for the defaults.  H2D copy, filter kernel and D2H copy of successive slices of
 * the batch overlap on three streams.  Host buffers that are not page-locked are
 * page-locked for the duration of the call when large, staged otherwise.
 */
int hdb_deblock_batch(hdb_context *ctx, const unsigned char *frames_in, unsigned char *frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *qp,
	const unsigned char *vert_bs, const unsigned char *hor_bs, hdb_timing *timing);

/* ---- a batch of frames already in device memory ------------------------ */
/*
 * All pointers are device pointers on ctx's device (d_vert_bs / d_hor_bs may be NULL
 * for the defaults; d_out may equal d_in).  The kernel is enqueued on `stream`
 * (a cudaStream_t; NULL = the legacy default stream) and the call returns
 * without synchronising.
 */
int hdb_deblock_batch_device(hdb_context *ctx, const void *d_frames_in, void *d_frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *d_qp,
	const unsigned char *d_vert_bs, const unsigned char *d_hor_bs, void *stream);

/* ---- decoder-style surfaces in device memory ---------------------------- */
/*
 * Three separately addressed planes with row pitches, as a hardware decoder or a frame pool hands
 * them over (the reference only knows tight planes, cpu.h:55-82; this is the first of the data
 * formats on the caller's side of the path).  Plane p of frame f starts at
 * base_p + f * frame_stride_{y|c}; row r of it at + r * pitch_{y|c}.  Pitches, frame strides and
 * base addresses must be multiples of 16 bytes, pitches at least the plane width, and the width a
 * multiple of 32 (HDB_ERR_DIMENSIONS otherwise).  Bytes between the plane width and the pitch are
 * neither read into the result nor written.  out may equal in (in place).  Asynchronous on `stream`
 * like hdb_deblock_batch_device; d_qp, d_vert_bs, d_hor_bs as there.
 */
typedef struct hdb_surfaces {
	void *y, *u, *v;                        /* device pointers to frame 0 of each plane */
	size_t pitch_y, pitch_c;                /* bytes between rows of the luma plane / of each chroma plane */
	size_t frame_stride_y, frame_stride_c;  /* bytes between frames; ignored when n_frames == 1 */
} hdb_surfaces;

int hdb_deblock_surfaces_device(hdb_context *ctx, const hdb_surfaces *in, const hdb_surfaces *out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *d_qp,
	const unsigned char *d_vert_bs, const unsigned char *d_hor_bs, void *stream);

/* ---- edge maps: per-edge QP and boundary-strength derivation ------------------------------ */
/*
 * The first of the components upstream of the filter.  The reference stubs both: one QP per frame
 * (cpu.h:35-37, 136-137) and constant boundary strengths (cpu.h:91-99); its setter takes strengths
 * only (cpu.h:120-132).  An EDGE BYTE carries both for one 8-sample edge:
 *     e = bs | qp << 2          bs 0..3 (0 = off, luma filters when > 0 as cpu.h:164), qp 0..51
 * in the reference's own two array layouts (vert[(H/8)][W/8 + 1], hor[(H/8 + 1)][W/8], sizes
 * hdb_num_vert_bs / hdb_num_hor_bs, exact-size check as cpu.h:122-123).  With edge maps every luma
 * edge segment takes beta / tc (cpu.h:1064-1072) from its own entry's qp; geometry, order, decisions
 * and filters are unchanged, and with one QP in every entry the result equals hdb_deblock_frame's.
 * The chroma planes follow the reference as before: the frame's qp and the constructor's chroma
 * strengths (cpu.h:110-117), which are not caller-settable there either.
 */
#define HDB_EDGE(bs, qp) ((unsigned char)(((bs) & 3u) | ((qp) << 2)))

int hdb_deblock_frame_edges(hdb_context *ctx, const unsigned char *yuv_in, unsigned char *yuv_out,
	unsigned width, unsigned height, unsigned qp,
	const unsigned char *vert_edges, unsigned num_vert_edges,
	const unsigned char *hor_edges, unsigned num_hor_edges, hdb_timing *timing);

/* batch forms of the same: hdb_deblock_batch / hdb_deblock_batch_device with edge maps (required) */
int hdb_deblock_batch_edges(hdb_context *ctx, const unsigned char *frames_in, unsigned char *frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *qp,
	const unsigned char *vert_edges, const unsigned char *hor_edges, hdb_timing *timing);
int hdb_deblock_batch_edges_device(hdb_context *ctx, const void *d_frames_in, void *d_frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *d_qp,
	const unsigned char *d_vert_edges, const unsigned char *d_hor_edges, void *stream);

/*
 * What a decoder knows about one 8x8 luma block.  hdb_derive_edges turns info[n_frames][H/8][W/8]
 * into edge maps by the published H.265 rule (8.7.2.4), which the reference leaves out: an edge that
 * is neither a transform nor a prediction edge is off; intra on either side -> 2; transform edge
 * with coded coefficients on either side -> 1; different reference pictures or motion-vector count,
 * or a motion-vector difference of >= 4 quarter samples in a component -> 1; else 0; the edge's qp
 * is (QpP + QpQ + 1) >> 1, above 51 acting as 51; the four sides of the picture are off.  All-intra
 * blocks with every edge a transform edge give the reference's "2 everywhere" on the inner edges.
 */
typedef struct hdb_block_info {
	short mv[2][2];          /* [list][x, y], quarter samples */
	signed char ref[2];      /* reference picture per list, < 0 = list unused */
	unsigned char flags;     /* HDB_BLK_* */
	unsigned char qp;        /* QpY */
} hdb_block_info;            /* 12 bytes */
#define HDB_BLK_INTRA   1u   /* intra-predicted */
#define HDB_BLK_CBF     2u   /* the transform block covering it has non-zero coefficients */
#define HDB_BLK_TU_LEFT 4u   /* its left edge is a transform-block edge */
#define HDB_BLK_TU_TOP  8u
#define HDB_BLK_PU_LEFT 16u  /* its left edge is a prediction-block edge */
#define HDB_BLK_PU_TOP  32u

/* device pointers; asynchronous on `stream`; the maps of frame f start at + f * hdb_num_*_bs */
int hdb_derive_edges_device(hdb_context *ctx, const hdb_block_info *d_info, unsigned n_frames,
	unsigned width, unsigned height, unsigned char *d_vert_edges, unsigned char *d_hor_edges, void *stream);
/* host pointers (runs on the GPU all the same: copy in, kernel, copy out) */
int hdb_derive_edges(hdb_context *ctx, const hdb_block_info *info, unsigned n_frames,
	unsigned width, unsigned height, unsigned char *vert_edges, unsigned char *hor_edges);

/* ---- NV12 (a hardware decoder's surface format) and the flagged general forms -------------- */
/*
 * The second of the data formats on the caller's side: Y plane followed by one plane of interleaved
 * chroma, U0 V0 U1 V1 ..., width bytes per row, height / 2 rows.  The reference reads planar YV12
 * only (cpu.h:66-82) and filters its two chroma planes identically (cpu.h:453-992); an NV12 frame
 * gives, sample for sample, what the reference gives on the same frame stored as YV12.  A tight NV12
 * frame has the size of a YV12 one (3 * width * height / 2), so the batch entry points take it as
 * is.  NV12 needs a width that is a multiple of 32 (HDB_ERR_DIMENSIONS otherwise: there is no
 * second kernel path for it).
 */
#define HDB_FLAG_EDGE_MAPS 1u   /* vert / hor hold edge bytes (HDB_EDGE) and are required */
#define HDB_FLAG_NV12      2u   /* frames are NV12 instead of YV12 */

/*
 * Samples deeper than 8 bits (the third of the caller-side data formats): planar frames like the 8-bit ones -- Y,
 * then the two chroma planes, tight rows -- of 16-bit little-endian containers holding `depth`-bit samples in their
 * low bits (what software decoders call yuv420p10le), 3 * width * height bytes a frame.  The reference is 8-bit
 * only: unsigned char planes (cpu.h:1041-1048) and the sample maximum 255 in its clips (cpu.h:1202, 1228, 1474).
 * Deeper samples follow the one rule the published standard adds (H.265 8.7.2.5.3): beta and tc from the
 * reference's tables times 1 << (depth - 8), and (1 << depth) - 1 in place of 255; geometry, order, decisions and
 * formulas are the reference's.  HDB_FLAG_DEPTH(8) -- 8-bit samples in 16-bit containers -- gives the reference's
 * results exactly; 9 and 10 have nothing in the reference to be compared with.  Width a multiple of 16, at least
 * 32; combines with HDB_FLAG_EDGE_MAPS.
 *   HDB_FLAG_NV12 | HDB_FLAG_DEPTH(10) is P010, the layout a hardware decoder writes for 10-bit video: Y plane, then
 * one plane of interleaved chroma (U0 V0 U1 V1 ...), every sample in the HIGH 10 bits of its 16-bit container (the
 * low 6 bits are written as 0).  Sample for sample the result of the planar form.
 */
#define HDB_FLAG_SAMPLES16 4u
#define HDB_FLAG_DEPTH(d)  (HDB_FLAG_SAMPLES16 | ((unsigned)(d) << 8))   /* d = 8, 9 or 10 */

/* hdb_deblock_batch / hdb_deblock_batch_device with flags (0 = exactly those two) */
int hdb_deblock_batch_ex(hdb_context *ctx, const unsigned char *frames_in, unsigned char *frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *qp,
	const unsigned char *vert, const unsigned char *hor, unsigned flags, hdb_timing *timing);
int hdb_deblock_batch_device_ex(hdb_context *ctx, const void *d_frames_in, void *d_frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *d_qp,
	const unsigned char *d_vert, const unsigned char *d_hor, unsigned flags, void *stream);

/*
 * The decoder-loop form from host memory: frames and their block data in, deblocked frames out.  The block
 * data of every slice travels with its frames (12 bytes per 8x8 block), the edge maps are derived on the
 * device (hdb_derive_edges_device) and never leave it.  flags: HDB_FLAG_NV12 or 0.
 */
int hdb_deblock_batch_blocks(hdb_context *ctx, const unsigned char *frames_in, unsigned char *frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *qp,
	const hdb_block_info *info, unsigned flags, hdb_timing *timing);

/* pitched NV12 surfaces in device memory, as hdb_deblock_surfaces_device; flags: HDB_FLAG_EDGE_MAPS or 0 */
typedef struct hdb_nv12_surfaces {
	void *y, *uv;                            /* device pointers to frame 0 of each plane */
	size_t pitch_y, pitch_uv;                /* bytes between rows (>= width for both) */
	size_t frame_stride_y, frame_stride_uv;  /* bytes between frames; ignored when n_frames == 1 */
} hdb_nv12_surfaces;

int hdb_deblock_nv12_surfaces_device(hdb_context *ctx, const hdb_nv12_surfaces *in, const hdb_nv12_surfaces *out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *d_qp,
	const unsigned char *d_vert, const unsigned char *d_hor, unsigned flags, void *stream);

/*
 * File in, file out for a raw sequence: any number of frames back to back in one file (the reference's
 * reader takes exactly one, cpu.h:43; its driver names one file per run, main.cu:111-133).  Same
 * checks and messages as hdb_execute_gpu, the length check becoming "a whole number of frames".
 * One qp for all frames, default strengths; flags: HDB_FLAG_NV12 or 0.  The file is processed a
 * bounded chunk at a time through the batch pipeline; timings are summed over the chunks.
 */
int hdb_execute_gpu_sequence(hdb_context *ctx, const char *input_file_name, const char *output_file_name,
	unsigned width, unsigned height, unsigned qp, unsigned flags, unsigned *n_frames, hdb_timing *timing);

/* ---- sample adaptive offset: the in-loop stage after deblocking --------------------------------- */
/*
 * The third of the components next to the path: absent from the reference, whose in-loop
 * chain ends with the deblocking filter (cpu.h:134-993), so this follows the published process (H.265 8.7.3) and
 * has nothing in the reference to be compared with.  Per coding tree block (ctb_size x ctb_size luma samples,
 * half that in each direction for the chroma planes; ctb_size 16, 32 or 64) and colour component:
 *   type 0  off
 *   type 1  band offset: band = sample >> 3; the four bands from `position` on (wrapping at 32) get offset[0..3]
 *   type 2  edge offset: class `position` -- 0 horizontal, 1 vertical, 2 diagonal (up-left / down-right),
 *           3 diagonal (up-right / down-left); a sample below both neighbours gets offset[0], below one and equal
 *           to the other offset[1], above one and equal to the other offset[2], above both offset[3]; a sample
 *           with a neighbour outside the picture is left alone
 * results clipped to 0..255; neighbours are always input samples, so the stage runs out of place.
 * params[n_frames][3][ceil(height / ctb_size)][ceil(width / ctb_size)], component order Y, U, V.  8-bit YV12;
 * deeper samples through the _ex entry points.
 */
typedef struct hdb_sao_params {
	unsigned char type;        /* 0, 1, 2 */
	unsigned char position;    /* band position 0..31, or edge class 0..3 */
	signed char offset[4];
} hdb_sao_params;             /* 6 bytes */

/* device pointers (d_out != d_in); asynchronous on `stream` */
int hdb_sao_batch_device(hdb_context *ctx, const void *d_frames_in, void *d_frames_out, unsigned n_frames,
	unsigned width, unsigned height, const hdb_sao_params *d_params, unsigned ctb_size, void *stream);
/* host pointers (runs on the GPU: copy in, kernel, copy out) */
int hdb_sao_batch(hdb_context *ctx, const unsigned char *frames_in, unsigned char *frames_out, unsigned n_frames,
	unsigned width, unsigned height, const hdb_sao_params *params, unsigned ctb_size);
/*
 * The same with flags: 0, or HDB_FLAG_DEPTH(d), d = 8 .. 12 -- planar frames of d-bit samples in 16-bit containers
 * (the layout of hdb_deblock_batch_device_ex with that flag; device frames 8-byte aligned).  The band of a sample
 * is then sample >> (d - 5) (8.7.3: bandShift), results are clipped to (1 << d) - 1, and the offsets are final
 * values as before (the standard's SaoOffsetVal: |offset| <= 31 at ten bits).
 */
int hdb_sao_batch_device_ex(hdb_context *ctx, const void *d_frames_in, void *d_frames_out, unsigned n_frames,
	unsigned width, unsigned height, const hdb_sao_params *d_params, unsigned ctb_size, unsigned flags, void *stream);
int hdb_sao_batch_ex(hdb_context *ctx, const void *frames_in, void *frames_out, unsigned n_frames,
	unsigned width, unsigned height, const hdb_sao_params *params, unsigned ctb_size, unsigned flags);
/*
 * The in-loop pair on device-resident frames: the deblocking filter (arguments as hdb_deblock_batch_device_ex,
 * flags 0, HDB_FLAG_EDGE_MAPS, HDB_FLAG_DEPTH(8 .. 10) or both) into scratch memory of the call's own, then sample
 * adaptive offset into d_frames_out; both on `stream`, asynchronous.  d_frames_out may equal d_frames_in.
 */
int hdb_deblock_sao_batch_device(hdb_context *ctx, const void *d_frames_in, void *d_frames_out, unsigned n_frames,
	unsigned width, unsigned height, const unsigned *d_qp, const unsigned char *d_vert, const unsigned char *d_hor,
	unsigned flags, const hdb_sao_params *d_params, unsigned ctb_size, void *stream);

/* number of filter-kernel launches made through ctx since creation */
unsigned long long hdb_kernel_launch_count(const hdb_context *ctx);

/* ---- several GPUs of one box (no reference analogue: the reference is single-GPU, main.cu:93-95) */
/* devices == NULL: the first n_devices visible devices; n_devices <= 0: all of them */
int hdb_multi_create(hdb_multi **m, const int *devices, int n_devices);
void hdb_multi_destroy(hdb_multi *m);
int hdb_multi_device_count(const hdb_multi *m);
const char *hdb_multi_last_error(const hdb_multi *m);

/* Frame-sharded batch: contiguous frame ranges per GPU, each with its own pipeline; no collective. */
int hdb_multi_deblock_batch(hdb_multi *m, const unsigned char *frames_in, unsigned char *frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *qp,
	const unsigned char *vert_bs, const unsigned char *hor_bs, hdb_timing *timing);

/* the same with HDB_FLAG_EDGE_MAPS / HDB_FLAG_NV12 (declared further up) */
int hdb_multi_deblock_batch_ex(hdb_multi *m, const unsigned char *frames_in, unsigned char *frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *qp,
	const unsigned char *vert, const unsigned char *hor, unsigned flags, hdb_timing *timing);

/*
 * One large frame split into strips of CTU rows (64 luma lines), one strip per GPU.
 * Each GPU uploads only its own rows; the 4-row input halos come from the
 * neighbouring GPUs' memory by peer-to-peer copies (NVLink where present).
 */
int hdb_multi_deblock_frame_strips(hdb_multi *m, const unsigned char *yuv_in, unsigned char *yuv_out,
	unsigned width, unsigned height, unsigned qp,
	const unsigned char *vert_bs, unsigned num_vert_bs,
	const unsigned char *hor_bs, unsigned num_hor_bs, hdb_timing *timing);
/* the same with edge maps: flags = HDB_FLAG_EDGE_MAPS (or 0); the maps keep their frame-global indexing */
int hdb_multi_deblock_frame_strips_ex(hdb_multi *m, const unsigned char *yuv_in, unsigned char *yuv_out,
	unsigned width, unsigned height, unsigned qp,
	const unsigned char *vert, unsigned num_vert,
	const unsigned char *hor, unsigned num_hor, unsigned flags, hdb_timing *timing);

/*
 * A STREAM of such frames (n_frames frames back to back in host memory, qp[n_frames], strengths per frame or
 * NULL), every frame split into the same strips.  The stream is pipelined: on every GPU the upload of frame
 * f + 1, the kernel of frame f and the download of frame f - 1 overlap (three streams and three buffers per
 * GPU, one host thread per GPU), and the halo rows are not copied at all -- each GPU's kernel reads the 4
 * input rows beyond a cut straight from the neighbouring GPU's buffer over NVLink.  hdb_multi_deblock_frame_strips
 * is this call with n_frames = 1.  Blocking; results are identical to the unsplit filter.
 */
int hdb_multi_deblock_stream_strips(hdb_multi *m, const unsigned char *frames_in, unsigned char *frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *qp,
	const unsigned char *vert, const unsigned char *hor, unsigned flags, hdb_timing *timing);

/*
 * The split itself: fills first_row[d] / n_rows[d] (luma rows; either may be NULL) for the devices in use and
 * returns their number -- CTU rows (64 luma lines) dealt out evenly, earlier devices first; fewer strips than
 * devices when the frame has fewer CTU rows.  0 on invalid arguments.
 */
int hdb_multi_strip_layout(const hdb_multi *m, unsigned height, int *first_row, int *n_rows);

/*
 * Strips that already sit in each GPU's memory (a decoder per GPU, an earlier stage): nothing crosses the
 * host link.  in[d] / out[d] are the three planes of strip d on device d, tight rows (pitch = width for Y,
 * width / 2 for U and V), holding exactly the rows hdb_multi_strip_layout assigns to d; 16-byte aligned;
 * out of place (a strip's input rows are read by its neighbours' kernels too).  d_vert[d] / d_hor[d]: frame-sized
 * strength (or edge-map) arrays on device d, or both NULL for the defaults.  streams[d]: the cudaStream_t of
 * device d to enqueue on (NULL array: the library's own); a strip's kernel starts after the work already
 * enqueued on its own and on both neighbours' streams.  With timing the call waits for completion and reports
 * kernel_ms = the slowest device's kernel (halo traffic included: it happens inside the kernel); without, it
 * returns once everything is enqueued.  width must be a multiple of 32.
 */
typedef struct hdb_strip_planes { void *y, *u, *v; } hdb_strip_planes;
int hdb_multi_deblock_strips_device(hdb_multi *m, const hdb_strip_planes *in, const hdb_strip_planes *out,
	unsigned width, unsigned height, unsigned qp,
	const unsigned char *const *d_vert, const unsigned char *const *d_hor,
	unsigned flags, void *const *streams, hdb_timing *timing);
/* bytes that cross NVLink per frame of that size: 4 rows of each plane per cut and direction */
unsigned long long hdb_multi_strip_halo_bytes(const hdb_multi *m, unsigned width, unsigned height);

#ifdef __cplusplus
}
#endif
#endif
