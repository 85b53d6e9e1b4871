/*
 * deblock_kernels.cuh -- launch-side interface of the sm_100a deblocking kernels
 * (internal to the library; the public boundary is include/hevc_deblock_b200.h).
 */
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hdb {

/* exact n / d for 0 <= n < 2^31 with one multiply-high */
struct FastDiv {
	uint32_t mul, shr, d;
	__host__ void Init(uint32_t div)
	{
		d = div;
		if (div <= 1) { mul = 0; shr = 0; return; }
		uint32_t lg = 0;
		while ((1ull << lg) < div) lg++;
		uint32_t p = 31 + lg;
		mul = (uint32_t)(((1ull << p) + div - 1) / div);
		shr = p - 32;
	}
	__device__ __forceinline__ uint32_t Div(uint32_t n) const { return d <= 1 ? n : __umulhi(n, mul) >> shr; }
};

/*
 * One batch of frames of one size, as the tiled kernel sees it.  Input planes are
 * read through three tensor maps (Y, U, V), output planes are written with plain
 * vector stores; everything else is here.  "row0" describe a CTU-row strip of a larger frame (multi-GPU strip
 * split): block rows are numbered frame-globally, the buffer holds rows starting
 * at global block row row0 (padded-block units), and every boundary-strength
 * index and border test uses the global numbers.  For a whole frame row0 = 0.
 */
struct TiledBatch {
	int n_frames;
	int nv12;                       /* 1: two planes, Y and interleaved UV (maps.in[1], out_plane[1]); else Y, U, V */
	/* frame-global geometry */
	int width, height;              /* luma */
	int luma_nbx, luma_nby;         /* (W+8)/8, (H+8)/8 */
	int chroma_nbx, chroma_nby;     /* (W/2+8)/8, (H/2+8)/8 */
	/* rows held by the buffers (whole frame: 0 .. nby) */
	int luma_row0, luma_rows;       /* first global block row, number of block rows to filter */
	int chroma_row0, chroma_rows;
	int luma_y_org, chroma_y_org;   /* buffer row of the first block row's first sample row:
	                                   -4 when the buffer starts at image row 0, else 0 */
	/* work decomposition: item = (frame, plane, row group, 32-block chunk) */
	int luma_chunks, chroma_chunks;
	int luma_groups, chroma_groups;
	int items_luma, items_chroma, items_per_frame;
	long long total_items;
	FastDiv div_items_per_frame, div_luma_chunks, div_chroma_chunks, div_items_chroma;
	/* chroma width = 4 mod 8: the reference's blocks tile a virtual plane of pitch chroma_vpitch laid over its
	 * padded buffer of pitch chroma_rpitch (deblock_kernels.cu: SkewToImage); cp.async load mode only */
	int chroma_skew, chroma_vpitch, chroma_rpitch;
	FastDiv div_chroma_rpitch;
	/* per-frame inputs */
	const unsigned *qp;             /* [n_frames], or NULL: every frame uses qp_value */
	unsigned qp_value;
	const unsigned char *vert_bs;   /* luma BS, reference layout */
	const unsigned char *hor_bs;
	long long vert_bs_stride, hor_bs_stride; /* bytes between frames; 0 = one array for all */
	const uint32_t *chroma_bs;      /* [chroma_nby][chroma_nbx] resolved, packed 4 x u8 */
	/* how the tiles are fetched (deblock_kernels.cu): 0 tensor maps, 1 tensor maps + peer halo rows, 2 cp.async */
	int load_mode;
	const unsigned char *in_plane[3];       /* input plane base pointers of frame 0 (load mode 2 reads through them) */
	/* load mode 1, a strip of a frame split across GPUs: the 4 rows above the buffer's first row / below its
	 * last row in the neighbouring strips' buffers (peer pointers, same pitch), NULL at the frame's own top / bottom */
	const unsigned char *halo_top[3], *halo_bot[3];
	/* output: plane base pointers of frame 0 (may equal the input's), rows present, frame pitch */
	unsigned char *out_plane[3];
	int luma_buf_rows, chroma_buf_rows;
	int luma_pitch, chroma_pitch;                      /* bytes between rows (>= plane width, multiple of 16) */
	long long frame_stride, chroma_frame_stride;       /* bytes between frames, luma plane / chroma planes */
};

struct TensorMaps3 { CUtensorMap in[3]; };   /* Y, U, V input planes, 32-bit elements */

/* kernel configuration the library was built with */
struct TiledConfig { int block_rows_per_item, stages, warps_per_cta, smem_bytes, max_ctas_per_sm; };
TiledConfig GetTiledConfig();

/* Fill the work-decomposition fields from the geometry fields. */
void FinalizeTiledBatch(TiledBatch &b);

/* Once per device (the current one): function attributes; returns the resident CTAs per SM the grid is sized
 * with, per kernel variant: index = (edge maps ? 1 : 0) | (NV12 ? 2 : 0) | load mode << 2. */
constexpr int kTiledVariants = 12;
cudaError_t PrepareTiled(int ctas_per_sm[kTiledVariants]);
/* Persistent tiled kernel: one launch filters Y, U and V of every frame of the batch. */
cudaError_t LaunchTiled(const TensorMaps3 &maps, const TiledBatch &b, void *sched, int sm_count, const int ctas_per_sm[kTiledVariants],
	cudaStream_t stream, bool edge_qp = false);
/* 16-bit samples (planar; bit depth 8 .. 10 in 16-bit containers): the tensor maps describe rows of 2 * width
 * bytes with a 132-element box; pitches in TiledBatch are bytes as always.  Variant index = (edge maps ? 1 : 0)
 * | (depth - 8) << 1; 6 and 7 are P010 (TiledBatch.nv12: Y + interleaved UV, 10 bits in the high bits of their
 * containers). */
constexpr int kTiled16Variants = 8;
cudaError_t PrepareTiled16(int ctas_per_sm[kTiled16Variants]);
cudaError_t LaunchTiled16(const TensorMaps3 &maps, const TiledBatch &b, void *sched, int sm_count,
	const int ctas_per_sm[kTiled16Variants], cudaStream_t stream, int bit_depth, bool edge_qp);
/* bytes of zero-initialised device memory the scheduler needs (one per context) */
size_t TiledSchedBytes();

/* Resolved chroma boundary strengths for one frame size -> table[chroma_nby * chroma_nbx]. */
cudaError_t LaunchChromaBsTable(uint32_t *table, int width, int height, cudaStream_t stream);

/* Reference-default luma boundary strengths (cpu.h:92-99) generated on the device. */
cudaError_t LaunchDefaultLumaBs(unsigned char *vert_bs, unsigned char *hor_bs, int width, int height,
	cudaStream_t stream);

/*
 * Generic path for frame sizes the tensor maps cannot describe (W % 32 != 0):
 * explicit zero-padded planes in scratch memory, the reference's own addressing
 * pitch (blocks_x * 8), one thread per padded 8x8 block.
 */
struct GenericBatch {
	int n_frames, width, height;
	const unsigned char *in; unsigned char *out;   /* tight YV12 frames, frame_stride apart */
	long long frame_stride;
	unsigned char *scratch;                        /* GenericScratchBytes(width, height) * n_frames */
	const unsigned *qp;                            /* or NULL: qp_value */
	unsigned qp_value;
	const unsigned char *vert_bs, *hor_bs;
	long long vert_bs_stride, hor_bs_stride;
	bool edge_qp;                                  /* vert_bs / hor_bs hold edge bytes (bs | qp << 2) */
};
size_t GenericScratchBytes(int width, int height);
cudaError_t LaunchGeneric(const GenericBatch &b, cudaStream_t stream);

/*
 * Edge maps (bs | qp << 2 per entry, reference array layouts) from per-8x8-block decoder data:
 * info = hdb_block_info[n_frames][H/8][W/8] (12 bytes each, 4-byte aligned) in device memory.
 */
cudaError_t LaunchDeriveEdges(const void *info, int width, int height, int n_frames,
	unsigned char *vert_edges, unsigned char *hor_edges, cudaStream_t stream);

/*
 * Sample adaptive offset (sao_kernels.cu; H.265 8.7.3, absent from the reference): tight planar 4:2:0 frames in
 * device memory -- a byte a sample (sample_bytes 1, bit_depth 8) or 16-bit containers (sample_bytes 2, bit_depth
 * 8 .. 12; 8-byte aligned) -- out of place, d_params = hdb_sao_params[n_frames][3][ceil(H / ctb)][ceil(W / ctb)].
 * d_tickets: kSaoTicketWords words of device memory that are the launch's own until it has run (the counters the
 * blocks are handed out through; zeroed here, in stream order).
 */
constexpr int kSaoTicketWords = 4096;
cudaError_t LaunchSao(const unsigned char *d_in, unsigned char *d_out, int n_frames, int width, int height,
	const void *d_params, int ctb, int bit_depth, int sample_bytes, unsigned *d_tickets, int sm_count, cudaStream_t stream);

} // namespace hdb
