/*
 * deblock_kernels.cu -- sm_100a kernels of the deblocking path.
 *
 * DeblockTiledKernel: persistent, warp-autonomous, one warp per CTA.  Every warp
 * runs its own TMA -> registers -> 128-bit-store pipeline over "items" = 32
 * horizontally adjacent padded 8x8 blocks x kRB (= 2) block rows of one plane of
 * one frame; the block rows of an item go through one rolled copy of the block
 * code, their boundary strengths are fetched one item ahead.
 *
 * Geometry.  The reference pads every plane by 4 samples (cpu.h:55-82), so padded
 * block (bx, by) covers image columns 8bx-4 .. 8bx+3 and rows 8by-4 .. 8by+3: it
 * is centred on an 8x8 grid crossing.  Blocks are independent (cpu.h:146-447
 * touches only the block's own 64 samples).  A TMA box must start on a 16-byte
 * boundary in the innermost dimension (measured: anything else faults), and the
 * blocks of an item start 4 bytes before one, so the item's box is fetched from
 * the 16-byte boundary 12 bytes further left: a 272 x 8*kRB byte box of 32-bit
 * elements at x = 256*chunk - 16, y = 8*by0 - 4 -- the item's 32 blocks plus a
 * 12-byte lead-in and a 4-byte tail that belong to the neighbouring items and
 * are never used.  TMA's out-of-bounds zero fill IS the reference's padding
 * (the oracle defines it as zero).  Lane l owns block l: it reads its 8x8
 * samples from shared memory (bytes 12+8l .. 19+8l of each row), runs the
 * vertical-edge and horizontal-edge passes fused in registers
 * (deblock_core.cuh) and writes them back to the same place.  The warp then
 * streams the item's own columns out of shared memory: fifteen 16-byte vectors
 * per row (image columns 256*chunk .. +239, two rows per instruction) plus one
 * instruction of 32-bit words for the 4 + 12 fringe bytes per row that do not
 * fill a vector on their own; writes the reference would make into its padding
 * are dropped by the bounds predicate.  Each output byte is written by exactly
 * one item and only bytes an item owns are consumed, so filtering in place
 * (out == in) is safe.
 *
 * Work is handed out dynamically in runs of consecutive items from one global
 * counter (luma items cost ~3x chroma items, so a static split would be
 * unbalanced); the runs shrink as the work runs out, and the last CTA to leave
 * resets the counter for the next launch.
 *
 * Two template parameters widen the same kernel: kEdgeQp -- the luma strength
 * arrays hold edge bytes (bs | qp << 2) and every luma segment takes beta / tc from its own entry -- and
 * kNv12 -- two planes, Y and interleaved UV, a UV item holding 16 block columns of U and of V that lane
 * pairs split, filter and re-interleave (UvBlockRow).  <false, false> is the kernel described above.
 *
 * DeriveEdgesKernel builds edge maps from per-8x8-block decoder data (H.265 8.7.2.4).
 * The generic kernels at the end cover frame sizes TMA cannot describe.
 */
#include "deblock_kernels.cuh"
#include "deblock_core.cuh"
#include "deblock_core16.cuh"
#include "tma_ptx.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#ifndef HDB_RB
#define HDB_RB 2        /* block rows per item */
#endif
#ifndef HDB_STAGES
#define HDB_STAGES 2
#endif
/*
 * One warp per CTA.  Warps share nothing -- each runs its own pipeline -- so the CTA is only the
 * unit of residency, and with a single warp in it the warp's shared-memory addresses are
 * compile-time constants: with one block row per item the kernel then fitted 72 registers without
 * spilling (96 with four warps per CTA; 80 spills) and 28 warps were resident per SM instead of 20:
 * 6.47 -> 6.04 ms for 1024 4K frames (profiles/r01i_cta_shapes.txt).  With two block rows per item:
 * 80 registers, no spills, 23 resident warps (shared memory is the limit).
 */
#ifndef HDB_WARPS
#define HDB_WARPS 1
#endif
#ifndef HDB_MIN_CTAS
#define HDB_MIN_CTAS 21
#endif
#ifndef HDB_NV12_MIN_CTAS
#define HDB_NV12_MIN_CTAS HDB_MIN_CTAS          /* the two-plane variants (measurement knob) */
#endif
#ifndef HDB_LDGSTS_MIN_CTAS
#define HDB_LDGSTS_MIN_CTAS HDB_MIN_CTAS         /* the cp.async loader variants: 80 registers as well, no spills */
#endif
#ifndef HDB_SMEM_ALIGN
#define HDB_SMEM_ALIGN 128                      /* a TMA destination without swizzle needs 128 bytes */
#endif
#ifndef HDB_TRACE
#define HDB_TRACE 0     /* 1: development build that records per-warp timestamps (tools/trace_tail.py) */
#endif

namespace hdb {

constexpr int kRB = HDB_RB;
constexpr int kStages = HDB_STAGES;
constexpr int kWarps = HDB_WARPS;
constexpr int kLead = 12;                       /* bytes from the 16-byte-aligned box origin to block 0 */
constexpr int kTileW = 272;                     /* bytes: lead-in + 32 blocks of 8 columns + 4 tail */
constexpr int kTileWords = kTileW / 4;          /* 68 */
constexpr int kTileH = 8 * kRB;
constexpr int kTileBytes = kTileW * kTileH;     /* 2176 * kRB: a multiple of 128 */
constexpr int kSmemBytes = kWarps * kStages * (kTileBytes + 8 + 8);
constexpr uint32_t kNoItem = 0xffffffffu;
/*
 * How a tile gets into shared memory (third template parameter of the kernel):
 *   kLoadTma      one TMA box per item (UTMALDG.3D), issued by lane 0 -- the kernel described above;
 *   kLoadTmaHalo  the same for one CTU-row strip of a frame that is split across GPUs (deblock_multi.cu):
 *                 the buffer holds the strip's own rows only, and the 4 input rows on the far side of a
 *                 cut are read by the warp straight from the neighbouring GPU's buffer (peer loads over
 *                 NVLink) into the tile rows TMA has zero-filled -- the halo exchange is fused into the
 *                 filter kernel, there is no copy step;
 *   kLoadLdgsts   planes a tensor map cannot describe (row pitch not a multiple of 16 bytes: every width
 *                 that is a multiple of 8 but not of 32, cpu.h:46-48): all lanes fetch the box with 4-byte
 *                 cp.async (LDGSTS, zero fill outside the plane) one item ahead, and rows are stored as
 *                 32-bit words because their global addresses are only 4-byte aligned.
 */
enum { kLoadTma = 0, kLoadTmaHalo = 1, kLoadLdgsts = 2 };
static_assert(kStages >= 2, "the stage after the current one holds the next item (boundary-strength prefetch)");

#if HDB_TRACE
/* development builds only (tools/trace_tail.py): per warp {kernel entry, first tile ready, loop exit} in ns and the item count */
__device__ unsigned long long g_trace[8192][4];
__device__ __forceinline__ unsigned long long GlobalTimerNs()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}
#endif

/* scheduler words in global memory: [0] next item, [1] CTAs done */
struct SchedWords { unsigned int next; unsigned int done; };

/* ------------------------------------------------------------------ */
/* item bookkeeping                                                    */
/* ------------------------------------------------------------------ */

/* kNv12: plane 1 is the interleaved UV plane; one of its items holds 16 block columns of U and of V */
template <bool kNv12>
__device__ __forceinline__ void ItemCoords(uint32_t desc, int &plane, int &bx0, int &lby0)
{
	plane = (int)(desc & 3u);
	bx0 = (int)((desc >> 2) & 0x3ffu) * ((kNv12 && plane) ? 16 : 32);
	lby0 = (int)(desc >> 12) * kRB;
}

/*
 * Producer-side cursor over the item space.  Items are numbered frame-major, then
 * plane (Y, U, V), then row group, then 32-block chunk.  A run of consecutive item
 * ids is decoded once (three multiply-high divisions) and then advanced by one.
 */
struct ItemCursor {
	uint32_t frame, plane, grp, chunk, chunks, groups;

	__device__ __forceinline__ void Seek(const TiledBatch &b, uint32_t id)
	{
		frame = b.div_items_per_frame.Div(id);
		uint32_t r = id - frame * (uint32_t)b.items_per_frame;
		if (r < (uint32_t)b.items_luma) {
			plane = 0; chunks = (uint32_t)b.luma_chunks; groups = (uint32_t)b.luma_groups;
			grp = b.div_luma_chunks.Div(r);
		} else {
			r -= (uint32_t)b.items_luma;
			plane = 1 + (r >= (uint32_t)b.items_chroma ? 1u : 0u);
			r -= (plane - 1) * (uint32_t)b.items_chroma;
			chunks = (uint32_t)b.chroma_chunks; groups = (uint32_t)b.chroma_groups;
			grp = b.div_chroma_chunks.Div(r);
		}
		chunk = r - grp * chunks;
	}
	template <bool kNv12>
	__device__ __forceinline__ void Next(const TiledBatch &b)
	{
		if (++chunk < chunks) return;
		chunk = 0;
		if (++grp < groups) return;
		grp = 0;
		if (++plane == (kNv12 ? 2 : 3)) { plane = 0; frame++; }
		chunks = (uint32_t)(plane ? b.chroma_chunks : b.luma_chunks);
		groups = (uint32_t)(plane ? b.chroma_groups : b.luma_groups);
	}
	__device__ __forceinline__ uint32_t Desc() const { return plane | (chunk << 2) | (grp << 12); }
};

/* what the consumer needs to know about the plane an item belongs to */
struct PlaneInfo {
	int nbx, rows, row0, pw, pitch, buf_rows, y_org;
	long long frame_stride;
	unsigned char *out;
};

template <bool kNv12>
__device__ __forceinline__ PlaneInfo MakePlaneInfo(const TiledBatch &b, int plane)
{
	PlaneInfo p;
	const bool c = plane != 0;
	p.nbx = c ? b.chroma_nbx : b.luma_nbx;
	p.rows = c ? b.chroma_rows : b.luma_rows;
	p.row0 = c ? b.chroma_row0 : b.luma_row0;
	p.pw = (c && !kNv12) ? (b.width >> 1) : b.width;      /* bytes per row of the plane: the UV plane is as wide as luma */
	p.pitch = c ? b.chroma_pitch : b.luma_pitch;
	p.frame_stride = c ? b.chroma_frame_stride : b.frame_stride;
	p.buf_rows = c ? b.chroma_buf_rows : b.luma_buf_rows;
	p.y_org = c ? b.chroma_y_org : b.luma_y_org;
	p.out = plane == 0 ? b.out_plane[0] : (plane == 1 ? b.out_plane[1] : b.out_plane[2]);
	return p;
}

/*
 * The boundary strengths of the block lane `lane` owns in block row `rb` of item {frame, desc}, as
 * loaded and not yet looked at: four bytes for a luma block (cpu.h:159-163, 223-227, 287-291,
 * 368-372), one packed word of the per-size table for a chroma block; all 0 for lanes and rows
 * outside the plane (a block whose four strengths are 0 is left untouched).  Loads and use are
 * taken apart on purpose: the loads of item i+1 are issued while item i is filtered and nothing
 * waits on them until PackBs() one item later (+1 % on long launches, +1.7 % on short ones).
 */
struct BsRaw { u32 a, b, c, d; };

template <bool kNv12>
__device__ __forceinline__ BsRaw ItemBsRaw(const TiledBatch &b, uint32_t frame, uint32_t desc, int lane, int rb)
{
	BsRaw r = { 0u, 0u, 0u, 0u };
	if (desc == kNoItem) return r;
	int plane, bx0, lby0;
	ItemCoords<kNv12>(desc, plane, bx0, lby0);
	const bool c = plane != 0;
	const int bx = bx0 + ((kNv12 && c) ? (lane & 15) : lane), lby = lby0 + rb;
	if (bx >= (c ? b.chroma_nbx : b.luma_nbx) || lby >= (c ? b.chroma_rows : b.luma_rows)) return r;
	const int by = (c ? b.chroma_row0 : b.luma_row0) + lby;
	if (c) { r.a = __ldg(b.chroma_bs + by * b.chroma_nbx + bx); return r; }
	const int w8 = b.width >> 3;
	const unsigned char *vert_bs = b.vert_bs + (long long)frame * b.vert_bs_stride;
	const unsigned char *hor_bs = b.hor_bs + (long long)frame * b.hor_bs_stride;
	if (by > 0) r.a = __ldg(vert_bs + (by - 1) * (w8 + 1) + bx);
	if (by < b.luma_nby - 1) r.b = __ldg(vert_bs + by * (w8 + 1) + bx);
	if (bx > 0) r.c = __ldg(hor_bs + by * w8 + bx - 1);
	if (bx < b.luma_nbx - 1) r.d = __ldg(hor_bs + by * w8 + bx);
	return r;
}

/* BS_ver1 | BS_ver2 << 8 | BS_hor1 << 16 | BS_hor2 << 24 */
__device__ __forceinline__ u32 PackBs(const BsRaw &r, bool chroma)
{
	return chroma ? r.a : (r.a | (r.b << 8) | (r.c << 16) | (r.d << 24));
}

/* ------------------------------------------------------------------ */
/* the kernel                                                          */
/* ------------------------------------------------------------------ */
/*
 * One block row of an interleaved-UV item (NV12), all 32 lanes: lane l & 15 names the block pair, lanes
 * 0-15 filter its U samples, lanes 16-31 its V samples, then the two swap one word per row and each
 * writes two of the row's four words.  Every lane runs the whole sequence, also those whose block lies
 * outside the plane (their strengths are 0, nothing is filtered, nothing is written back): the exchange
 * then sits outside divergent control flow with a full mask.  Under `if (live)` with a ballot mask the
 * two-plane kernels needed 96 registers (80 spilled 424 bytes) and lost three resident warps per SM.
 */
__device__ __forceinline__ void UvBlockRow(u32 *base, u32 s4, const FilterParams &fp, bool live, bool v_plane)
{
	u32 lo[8], hi[8];
#pragma unroll
	for (int r = 0; r < 8; r++) {
		/* the pair's 16 bytes of a row start 8 bytes into a 16-byte unit: two 64-bit accesses */
		const uint2 *w = reinterpret_cast<const uint2 *>(base + r * kTileWords);
		const uint2 a = w[0], c = w[1];
		SplitUv(a.x, a.y, c.x, c.y, v_plane, lo[r], hi[r]);
	}
	ChromaBlock(lo, hi, s4 & 0xff, (s4 >> 8) & 0xff, (s4 >> 16) & 0xff, s4 >> 24, fp);
#pragma unroll
	for (int r = 0; r < 8; r++) {
		/* the U lane needs the other plane's lo, the V lane the other plane's hi */
		const u32 other = __shfl_xor_sync(0xffffffffu, v_plane ? lo[r] : hi[r], 16);
		uint2 o;
		if (v_plane) JoinUv(other, hi[r], o.x, o.y);
		else JoinUv(lo[r], other, o.x, o.y);
		if (live) *reinterpret_cast<uint2 *>(base + r * kTileWords + (v_plane ? 2 : 0)) = o;
	}
}

/*
 * Skewed chroma plane (see fetch_tile in the kernel): byte (vr, vc) of the virtual padded plane -- pitch
 * b.chroma_vpitch = 8 * chroma_nbx -- is byte o = vr * vpitch + vc of the reference's padded buffer, whose real
 * pitch is b.chroma_rpitch = cw + 8 with the image at (4, 4): buffer row r = o / rpitch, column c = o - r * rpitch.
 * A SkewWalk follows one virtual column down the rows of an item: one division where it starts, then
 * c += vpitch with at most one wrap per row (vpitch = rpitch - 4).
 */
struct SkewWalk {
	int r, c;          /* row / column in the padded buffer of the unit's first byte */
	int img;           /* byte offset in the image plane that position would have: (r - 4) * cw + (c - 4) */
	bool lo_col, hi_col;   /* the unit's first / second four virtual columns exist (0 <= column < vpitch) */

	__device__ __forceinline__ void Start(const TiledBatch &b, int vr, int vc)
	{
		lo_col = (unsigned)vc < (unsigned)b.chroma_vpitch;
		hi_col = (unsigned)(vc + 4) < (unsigned)b.chroma_vpitch;
		/* vc can be a little negative (the box's lead-in): count from 8 buffer rows further up to stay positive */
		const uint32_t o = (uint32_t)(vr * b.chroma_vpitch + vc + 8 * b.chroma_rpitch);
		const uint32_t q = b.div_chroma_rpitch.Div(o);
		r = (int)q - 8; c = (int)(o - q * (uint32_t)b.chroma_rpitch);
		img = (r - 4) * (b.width >> 1) + c - 4;
	}
	/* one virtual row down: the same column is vpitch bytes further on in the buffer, i.e. vpitch further in the
	 * image too unless the buffer row wraps, which takes the 8 bytes of padding out */
	__device__ __forceinline__ void NextRow(const TiledBatch &b)
	{
		c += b.chroma_vpitch; img += b.chroma_vpitch;
		if (c >= b.chroma_rpitch) { c -= b.chroma_rpitch; r++; img -= 8; }
	}
	/* which halves of the 8-byte unit at this position are image samples (a unit can straddle the 4-sample
	 * padding at either end of a buffer row, or an end of the virtual row) */
	__device__ __forceinline__ void Halves(const TiledBatch &b, bool &lo, bool &hi) const
	{
		const int cw = b.width >> 1;
		const bool row_ok = (unsigned)(r - 4) < (unsigned)b.chroma_buf_rows;
		lo = lo_col && row_ok && (unsigned)(c - 4) <= (unsigned)(cw - 4);      /* 4 <= c <= cw */
		hi = hi_col && row_ok && (unsigned)c <= (unsigned)(cw - 4);            /* 0 <= c <= cw - 4 */
	}
};

/*
 * The same walk in closed form for planes of at least 64 chroma columns (tools: the recurrence above is the
 * definition, tests/test_gpu_parity.py compares both against the oracle).  Going one virtual row down moves a
 * column one buffer row down and 4 bytes to the left (c -= 4, image offset += vpitch - 8); the only exception is
 * the row where c reaches 0: from there the column jumps to the end of the SAME buffer row (c = cw + 4, image
 * offset += vpitch).  So with t0 = c / 4 of the first row, the rows t0, t0 + 1, t0 + 2 of the walk are the ones
 * that straddle the padding (first half, both halves, second half of the unit are padding), the image offset
 * is linear with one bump of 8 after row t0, and the next such row is rpitch / 4 >= 18 rows away: at most one
 * per item.  A lane keeps two 16-bit masks (which rows' halves are image samples) instead of walking.
 */
struct SkewRun {
	int img0;                  /* image offset of the unit in the item's first row */
	int bump;                  /* rows after this one are 8 bytes further on (>= kTileH: none) */
	unsigned lo_mask, hi_mask; /* bit t: the unit's first / second four bytes in row t are image samples */

	__device__ __forceinline__ static unsigned RowRange(int x, int y)   /* bits x .. y - 1 of 0 .. 15 */
	{
		x = min(max(x, 0), 16); y = min(max(y, 0), 16);
		return y > x ? ((1u << y) - 1u) & ~((1u << x) - 1u) : 0u;
	}
	__device__ __forceinline__ void Start(int vpitch, int rpitch, int cw, int buf_rows, const FastDiv &div_rpitch, int vr, int vc)
	{
		const bool lo_col = (unsigned)vc < (unsigned)vpitch, hi_col = (unsigned)(vc + 4) < (unsigned)vpitch;
		const uint32_t o = (uint32_t)(vr * vpitch + vc + 8 * rpitch);
		const uint32_t q = div_rpitch.Div(o);
		const int r0 = (int)q - 8, c0 = (int)(o - q * (uint32_t)rpitch);
		img0 = (r0 - 4) * cw + c0 - 4;
		const int t0 = c0 == cw + 4 ? -1 : (c0 == cw ? -2 : c0 >> 2);
		const int sh = min(t0 + 2, 24);
		const unsigned lo_inv = (3u << sh) >> 2, hi_inv = (6u << sh) >> 2;
		bump = t0 >= 0 ? t0 : 99;
		/* buffer rows 4 .. buf_rows + 3 hold the image: row r0 + t before the bump, r0 + t - 1 after it */
		const int a = 4 - r0, e = buf_rows + 4 - r0;
		const unsigned before = RowRange(0, bump + 1);
		const unsigned rows = (RowRange(a, e) & before) | (RowRange(a + 1, e + 1) & ~before);
		lo_mask = lo_col ? rows & ~lo_inv : 0u;
		hi_mask = hi_col ? rows & ~hi_inv : 0u;
	}
};

/* planes narrower than that (a column can meet the padding twice in one item): the walk itself, out of line so that
 * the kernels' hot loops stay small */
__device__ __noinline__ void SkewFetchNarrow(const TiledBatch &b, const unsigned char *src, uint32_t d, int vr, int vc)
{
	SkewWalk wk;
	wk.Start(b, vr, vc);
#pragma unroll 1
	for (int t = 0; t < kTileH; t++) {
		bool lo, hi;
		wk.Halves(b, lo, hi);
		const unsigned char *g = src + wk.img;
		if (lo == hi) CpAsync8(d, lo ? g : src, lo ? 8u : 0u);
		else { CpAsync4(d, lo ? g : src, lo ? 4u : 0u); CpAsync4(d + 4u, hi ? g + 4 : src, hi ? 4u : 0u); }
		wk.NextRow(b);
		d += kTileW;
	}
}

__device__ __noinline__ void SkewStoreNarrow(const TiledBatch &b, unsigned char *dst, const unsigned char *sp, int vr, int vc, bool live)
{
	SkewWalk wk;
	wk.Start(b, vr, vc);
	if (!live) wk.lo_col = wk.hi_col = false;
#pragma unroll 1
	for (int rr = 0; rr < kTileH; rr++) {
		bool lo, hi;
		wk.Halves(b, lo, hi);
		const uint2 v = *reinterpret_cast<const uint2 *>(sp);
		if (lo && hi) *reinterpret_cast<uint2 *>(dst + wk.img) = v;
		else {
			if (lo) *reinterpret_cast<u32 *>(dst + wk.img) = v.x;
			if (hi) *reinterpret_cast<u32 *>(dst + wk.img + 4) = v.y;
		}
		wk.NextRow(b);
		sp += kTileW;
	}
}

template <bool kNv12>
__device__ __noinline__ void FetchTileLdgsts(const TiledBatch &b, const uint2 it, const uint32_t dst, const int lane)
{
	if (it.y != kNoItem) {
		int plane, bx0, lby0;
		ItemCoords<kNv12>(it.y, plane, bx0, lby0);
		const bool c = plane != 0;
		const int pitch = c ? b.chroma_pitch : b.luma_pitch;
		const unsigned char *src = b.in_plane[plane] + (long long)it.x * (c ? b.chroma_frame_stride : b.frame_stride);
		const int xb0 = 256 * (int)((it.y >> 2) & 0x3ffu) - 16;        /* plane byte of tile byte 0 */
		/* loops kept rolled: the cp.async variants must not outgrow the instruction cache (ncu: "no instruction"
		 * was their first stall reason when these were unrolled) */
		if (!kNv12 && c && b.chroma_skew) {
			/* 8-byte units like the straight planes: tile byte 0 is virtual column 4 mod 8, which is where the image's
			 * 8-byte boundaries fall (SkewWalk); lane l takes unit l of every row, and units 32 and 33 of the rows go
			 * one to a lane afterwards.  A unit half in the padding goes as two 4-byte copies. */
			const int vpitch = b.chroma_vpitch, rpitch = b.chroma_rpitch, cw = b.width >> 1, buf_rows = b.chroma_buf_rows;
			if (rpitch >= 72) {
				SkewRun k;
				k.Start(vpitch, rpitch, cw, buf_rows, b.div_chroma_rpitch, 8 * lby0, xb0 + 4 + 8 * lane);
				const unsigned char *g = src + k.img0;
				uint32_t d = dst + 8u * (uint32_t)lane;
				const int step = vpitch - 8;
#pragma unroll 1
				for (int t = 0; t < kTileH; t++) {
					const bool lo = (k.lo_mask >> t) & 1u, hi = (k.hi_mask >> t) & 1u;
					if (lo == hi) CpAsync8(d, lo ? g : src, lo ? 8u : 0u);
					else { CpAsync4(d, lo ? g : src, lo ? 4u : 0u); CpAsync4(d + 4u, hi ? g + 4 : src, hi ? 4u : 0u); }
					g += t == k.bump ? step + 8 : step;
					d += kTileW;
				}
			} else SkewFetchNarrow(b, src, dst + 8u * (uint32_t)lane, 8 * lby0, xb0 + 4 + 8 * lane);
			/* units 32 and 33: row lane / 2 (and 16 rows further on, for taller items), unit 32 + lane % 2 */
#pragma unroll 1
			for (int t = lane >> 1; t < kTileH; t += 16) {
				SkewWalk wk;
				wk.Start(b, 8 * lby0 + t, xb0 + 4 + 8 * (32 + (lane & 1)));
				bool lo, hi;
				wk.Halves(b, lo, hi);
				const unsigned char *g = src + wk.img;
				const uint32_t d = dst + (uint32_t)(t * kTileW) + 8u * (uint32_t)(32 + (lane & 1));
				if (lo == hi) CpAsync8(d, lo ? g : src, lo ? 8u : 0u);
				else { CpAsync4(d, lo ? g : src, lo ? 4u : 0u); CpAsync4(d + 4u, hi ? g + 4 : src, hi ? 4u : 0u); }
			}
		} else {
			const int pw = (c && !kNv12) ? (b.width >> 1) : b.width;
			const int rows = c ? b.chroma_buf_rows : b.luma_buf_rows;
			const int y0 = 8 * lby0 + (c ? b.chroma_y_org : b.luma_y_org);
			/* 34 units of 8 bytes per row: lane l takes unit l of every row; units 32 and 33 go one to a lane
			 * afterwards (row lane / 2, unit 32 + lane % 2) */
			const int xb = xb0 + 8 * lane;
			const bool col_ok = xb >= 0 && xb < pw;
			const unsigned char *g = src + (long long)y0 * pitch + xb;
			uint32_t d = dst + 8u * (uint32_t)lane;
#pragma unroll 1
			for (int t = 0; t < kTileH; t++) {
				const bool ok = col_ok && (unsigned)(y0 + t) < (unsigned)rows;
				CpAsync8(d, ok ? g : src, ok ? 8u : 0u);
				g += pitch; d += kTileW;
			}
			const int xe = xb0 + 8 * (32 + (lane & 1));
#pragma unroll 1
			for (int t = lane >> 1; t < kTileH; t += 16) {
				const bool ok = xe < pw && (unsigned)(y0 + t) < (unsigned)rows;
				CpAsync8(dst + (uint32_t)(t * kTileW) + 8u * (uint32_t)(32 + (lane & 1)), ok ? src + (long long)(y0 + t) * pitch + xe : src, ok ? 8u : 0u);
			}
		}
	}
CpAsyncCommit();
}

/* kEdgeQp: the luma boundary-strength arrays hold edge bytes (bs | qp << 2, deblock_core.cuh) and every
 * luma segment takes beta / tc from its own entry; chroma is untouched by it (frame QP, cpu.h:110-117) */
/* kNv12: two planes, Y and interleaved UV (lanes 0-15 filter the U blocks of a UV item, lanes 16-31 the V blocks) */
template <bool kEdgeQp, bool kNv12, int kLoad>
__global__ void __launch_bounds__(kWarps * 32, kLoad == kLoadLdgsts ? HDB_LDGSTS_MIN_CTAS : (kNv12 ? HDB_NV12_MIN_CTAS : HDB_MIN_CTAS))
DeblockTiledKernel(const __grid_constant__ TensorMaps3 maps, const __grid_constant__ TiledBatch b,
	SchedWords *sched, const uint32_t run_len)
{
	extern __shared__ __align__(HDB_SMEM_ALIGN) unsigned char smem[];
	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	unsigned char *tiles = smem + warp * (kStages * kTileBytes);
	uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kWarps * kStages * kTileBytes) + warp * kStages;
	/* what each stage holds, written by lane 0 when it issues the load: {frame, desc} */
	uint2 *held = reinterpret_cast<uint2 *>(smem + kWarps * kStages * (kTileBytes + 8)) + warp * kStages;
	const uint32_t tiles_s = SmemAddr(tiles);
	const uint32_t bars_s = SmemAddr(bars);
	const uint32_t total = (uint32_t)b.total_items;

	/* lane 0 is the warp's producer: scheduler state, TMA issue */
	uint32_t run_next = 0, run_end = 0, run_pending = 0, pending_len = 0;
	/*
	 * Guided self-scheduling: a request takes (items still unassigned) / (2 x warps in the grid)
	 * consecutive items, at most run_len and at least 1.  Long runs keep the scheduler atomic and
	 * the cursor decode cheap while there is plenty of work; towards the end the runs shrink to
	 * single items, so the last warp finishes one item after the others instead of run_len items.
	 */
	const uint32_t two_grid_warps = 2u * gridDim.x * kWarps;
	auto next_len = [&](uint32_t seen) -> uint32_t {
		const uint32_t left = total > seen ? total - seen : 0u;
		return min(max(left / two_grid_warps, 1u), run_len);
	};
	ItemCursor cur;
	cur.frame = cur.plane = cur.grp = cur.chunk = 0; cur.chunks = cur.groups = 1;

	auto issue_load = [&](int s) {         /* lane 0 only */
		if (run_next == run_end) {
			run_next = run_pending;
			run_end = run_pending + pending_len;
			if (run_next < total) {
				/* ask for the run after this one now; it is not needed before this one is used up */
				pending_len = next_len(run_end);
				run_pending = atomicAdd(&sched->next, pending_len);
				cur.Seek(b, run_next);
			}
		} else {
			cur.template Next<kNv12>(b);
		}
		if (run_next >= total) { held[s] = make_uint2(0u, kNoItem); return; }
		run_next++;
		held[s] = make_uint2(cur.frame, cur.Desc());
		if (kLoad == kLoadLdgsts) return;      /* every lane fetches its share of the box: fetch_tile */
		const uint32_t bar = bars_s + 8u * s;
		MbarArriveExpectTx(bar, kTileBytes);
		/* x in 32-bit elements: (8*bx0 - 4 - kLead) / 4, a multiple of 4 elements = 16 bytes */
		TmaLoad3D(tiles_s + s * kTileBytes, &maps.in[cur.plane], 64 * (int)cur.chunk - 4,
			8 * kRB * (int)cur.grp + (cur.plane ? b.chroma_y_org : b.luma_y_org), (int)cur.frame, bar);
	};
	/*
	 * kLoadLdgsts: the box of the item lane 0 has just put into held[s], fetched by all lanes with cp.async;
	 * where the box leaves the plane the copy writes zeros instead (src-size 0: the zero fill is the reference's
	 * padding, as TMA's out-of-bounds fill is in the other variants).  One commit group per item, empty for "no
	 * item", so that wait_group counts items.
	 *   Straight planes (luma always; chroma when its width is a multiple of 8): 8-byte copies -- rows of such
	 * planes start on 8-byte boundaries.  34 units per tile row: lane l takes unit l of every row, and the two
	 * left-over units of row l / 2.
	 *   Skewed chroma planes (chroma width = 4 mod 8, i.e. frame width = 8 mod 16): the reference pads such a
	 * plane to cw + 8 columns but addresses it with a pitch of 8 * (number of blocks) = cw + 4 (cpu.h:450-451,
	 * 469-471), so its blocks tile a VIRTUAL plane -- the same linear memory read with the shorter pitch.
	 * SkewToImage maps a word of the virtual plane to the image sample it really is (or to padding); 4-byte
	 * copies, one division per word.  The generic kernels reproduce the same thing with an explicit padded copy.
	 */
	/* kLoadLdgsts: the box of stage s, fetched by all lanes after a __syncwarp() that follows issue_load; out of
	 * line, so that its two call sites share one copy of the code (instruction cache) */
	auto fetch_tile = [&](int s) { FetchTileLdgsts<kNv12>(b, held[s], tiles_s + s * kTileBytes, lane); };

#if HDB_TRACE
	const unsigned long long t_entry = GlobalTimerNs();
	unsigned long long t_first = 0, n_done = 0;
#endif
	if (lane == 0) {
		for (int s = 0; s < kStages; s++) MbarInit(bars_s + 8u * s, 1);
		FenceMbarInit();
		pending_len = next_len(0u);
		run_pending = atomicAdd(&sched->next, pending_len);
		for (int s = 0; s < kStages; s++) issue_load(s);
	}
	__syncwarp();
	if (kLoad == kLoadLdgsts) {
		for (int s = 0; s < kStages; s++) fetch_tile(s);
	}

	int s = 0;
	uint32_t parity = 0;
	uint32_t cur_frame = 0xffffffffu;
	int cur_plane = -1;
	FilterParams fp;
	PlaneInfo pl;
	MakeFilterParams(fp, 0, 0);
	pl = MakePlaneInfo<kNv12>(b, 0);
	/* boundary strengths travel one item ahead of the samples */
	BsRaw bs_raw[kRB];
	{
		const uint2 it0 = held[0];
#pragma unroll
		for (int rb = 0; rb < kRB; rb++) bs_raw[rb] = ItemBsRaw<kNv12>(b, it0.x, it0.y, lane, rb);
	}
	for (;;) {
		const uint2 it = held[s];
		const uint32_t frame = it.x, desc = it.y;
		if (desc == kNoItem) break;
		int plane, bx0, lby0;
		ItemCoords<kNv12>(desc, plane, bx0, lby0);
		const bool chroma = plane != 0;
		/* beta / tc once per frame (cpu.h:136-137); plane constants once per plane.  Runs of
		 * consecutive items share both, so these two branches are rarely taken. */
		if (frame != cur_frame) {
			cur_frame = frame;
			const unsigned qp = b.qp ? __ldg(b.qp + frame) : b.qp_value;
			MakeFilterParams(fp, BetaOf(qp), TcOf(qp));
		}
		if (plane != cur_plane) { cur_plane = plane; pl = MakePlaneInfo<kNv12>(b, plane); }
		const int bx = bx0 + ((kNv12 && chroma) ? (lane & 15) : lane);

		/* this item's boundary strengths were loaded during the previous item; start the next item's loads */
		u32 bs[kRB];
		{
			const uint2 it_next = held[s + 1 == kStages ? 0 : s + 1];
#pragma unroll
			for (int rb = 0; rb < kRB; rb++) bs[rb] = PackBs(bs_raw[rb], chroma);
#pragma unroll
			for (int rb = 0; rb < kRB; rb++) bs_raw[rb] = ItemBsRaw<kNv12>(b, it_next.x, it_next.y, lane, rb);
		}

		if (kLoad == kLoadLdgsts) {
			/* this item's group is the older of the kStages in flight; the copies of the other lanes become
			 * visible through the warp barrier */
			CpAsyncWait<kStages - 1>();
			__syncwarp();
		} else {
			MbarWait(bars_s + 8u * s, parity);
		}
#if HDB_TRACE
		if (n_done++ == 0) t_first = GlobalTimerNs();
#endif
		u32 *tile = reinterpret_cast<u32 *>(tiles + s * kTileBytes);
		unsigned char *dst = pl.out + (long long)frame * pl.frame_stride;
		if (kLoad == kLoadTmaHalo) {
			/*
			 * A strip of a frame that is split across GPUs: the buffer holds the strip's own rows, so the tile
			 * rows above its first and below its last row came out of TMA as zeros.  Where a neighbouring strip
			 * exists they are its last / first 4 input rows: fetch them from the neighbour's buffer (a peer
			 * pointer: NVLink) into the tile.  Rare: the first and the last row group of a plane only.
			 */
			const int y_first = 8 * lby0 + pl.y_org;
			const unsigned char *top = b.halo_top[plane], *bot = b.halo_bot[plane];
			const bool need_top = y_first < 0 && top != nullptr;
			const bool need_bot = y_first + kTileH > pl.buf_rows && bot != nullptr;
			if (need_top || need_bot) {
				const int xw0 = (int)((desc >> 2) & 0x3ffu) * 64 - 4;       /* plane word of tile word 0 */
				for (int t = 0; t < kTileH; t++) {
					const int y = y_first + t;
					const unsigned char *row = nullptr;
					if (need_top && y < 0 && y >= -4) row = top + (long long)(y + 4) * pl.pitch;
					if (need_bot && y >= pl.buf_rows && y < pl.buf_rows + 4) row = bot + (long long)(y - pl.buf_rows) * pl.pitch;
					if (row == nullptr) continue;
					for (int w = lane; w < kTileWords; w += 32) {
						const int xb = 4 * (xw0 + w);
						tile[t * kTileWords + w] = (xb >= 0 && xb < pl.pw) ? LdPeer32(row + xb) : 0u;
					}
				}
				__syncwarp();
			}
		}
		/* not unrolled: one copy of the block code, or the kernel outgrows the instruction cache
		 * (unrolled, two block rows per item run 25 % slower than one) */
		/* tc = 0 (QP <= 17, cpu.h:1028-1033): every correction is clipped to +-0 (cpu.h:1192-1211, 1254-1276,
		 * 1464-1476) -- the frame passes through unchanged, so its items are only moved.  Chroma always takes the
		 * frame's QP; luma does unless the strengths are edge bytes. */
		const bool pass_through = fp.tc == 0 && (chroma || !kEdgeQp);
#pragma unroll 1
		for (int rb = 0; rb < (pass_through ? 0 : kRB); rb++) {
			if (kNv12 && chroma) {
				u32 s4 = bs[0];
#pragma unroll
				for (int k = 1; k < kRB; k++) if (rb == k) s4 = bs[k];
				/* interleaved UV: the block pair of column l = lane & 15 is 16 bytes per row, 8 bytes after the
				 * box origin; lane l filters its U samples, lane l + 16 its V samples, then the two swap
				 * halves and each writes two of the row's four words */
				UvBlockRow(tile + rb * 8 * kTileWords + 2 + 4 * (lane & 15), s4, fp, bx < pl.nbx && lby0 + rb < pl.rows, lane >= 16);
			} else if (bx < pl.nbx && lby0 + rb < pl.rows) {
				u32 *base = tile + rb * 8 * kTileWords + kLead / 4 + 2 * lane;
				u32 lo[8], hi[8];
#pragma unroll
				for (int r = 0; r < 8; r++) { lo[r] = base[r * kTileWords]; hi[r] = base[r * kTileWords + 1]; }
				u32 s4 = bs[0];          /* select instead of indexing: bs[] stays in registers */
#pragma unroll
				for (int k = 1; k < kRB; k++) if (rb == k) s4 = bs[k];
				if (!kNv12 && chroma) ChromaBlock(lo, hi, s4 & 0xff, (s4 >> 8) & 0xff, (s4 >> 16) & 0xff, s4 >> 24, fp);
				else LumaBlockT<kEdgeQp>(lo, hi, s4 & 0xff, (s4 >> 8) & 0xff, (s4 >> 16) & 0xff, s4 >> 24, fp);
#pragma unroll
				for (int r = 0; r < 8; r++) { base[r * kTileWords] = lo[r]; base[r * kTileWords + 1] = hi[r]; }
			}
		}
		__syncwarp();

		/* stream the item's own columns out: image columns x_img .. x_img + 251 start at tile byte kLead */
		{
			const int y0 = 8 * lby0 + pl.y_org;
			/* plane byte of tile byte kLead; a UV item starts 8 bytes into the box instead of 12 but covers the
			 * same 256 plane bytes per row relative to the box, four bytes further left */
			const bool uv = kNv12 && chroma;
			const int x_img = uv ? 16 * bx0 - 4 : 8 * bx0 - 4;
			const unsigned char *tile8 = reinterpret_cast<const unsigned char *>(tile);
			/* 16-byte vectors: tile bytes 16 .. 255 <-> image columns x_img + 4 .. x_img + 243; two rows per pass */
			const int j = lane & 15, half = lane >> 4;
			const int vcol = x_img + 4 + 16 * j;
			const bool vec_lane = j < 15 && vcol < pl.pw;
			const unsigned char *vsrc = tile8 + half * kTileW + 16 + 16 * j;
			unsigned char *vdst = dst + (long long)(y0 + half) * pl.pitch + vcol;
			const long long pitch2 = 2ll * pl.pitch;
			/* fringe words: tile word 3 (image column x_img) and words 64..66 (x_img + 244 ..): 4 per row */
			const int ftr = lane >> 2, k = lane & 3;
			const int fword = uv ? (k < 2 ? 2 + k : 62 + k) : (k == 0 ? 3 : 63 + k);
			const int fcol = x_img + 4 * (fword - 3);
			const bool f_lane = fcol >= 0 && fcol < pl.pw;
			if (kLoad == kLoadLdgsts && !kNv12 && chroma && b.chroma_skew) {
				/* the item's 256 bytes per row of the VIRTUAL plane (tile bytes 12 .. 267) to the image samples they are:
				 * 31 units of 8 bytes from tile byte 16 on (image-aligned, see fetch) and one word at each end */
				const int vpitch = b.chroma_vpitch, rpitch = b.chroma_rpitch, cw = b.width >> 1, buf_rows = b.chroma_buf_rows;
				const unsigned char *sp = tile8 + 16 + 8 * lane;
				if (rpitch >= 72) {
					SkewRun k;
					k.Start(vpitch, rpitch, cw, buf_rows, b.div_chroma_rpitch, 8 * lby0, 8 * bx0 + 4 + 8 * lane);
					if (lane == 31) k.lo_mask = k.hi_mask = 0u;
					unsigned char *g = dst + k.img0;
					const int step = vpitch - 8;
#pragma unroll 1
					for (int rr = 0; rr < kTileH; rr++) {
						const bool lo = (k.lo_mask >> rr) & 1u, hi = (k.hi_mask >> rr) & 1u;
						const uint2 v = *reinterpret_cast<const uint2 *>(sp);
						if (lo && hi) *reinterpret_cast<uint2 *>(g) = v;
						else {
							if (lo) *reinterpret_cast<u32 *>(g) = v.x;
							if (hi) *reinterpret_cast<u32 *>(g + 4) = v.y;
						}
						g += rr == k.bump ? step + 8 : step;
						sp += kTileW;
					}
				} else SkewStoreNarrow(b, dst, sp, 8 * lby0, 8 * bx0 + 4 + 8 * lane, lane < 31);
				/* the word at virtual column 8 bx0 / 8 bx0 + 252 of every row: row lane / 2, end lane % 2 */
#pragma unroll 1
				for (int rr = lane >> 1; rr < kTileH; rr += 16) {
					SkewWalk fr;
					fr.Start(b, 8 * lby0 + rr, 8 * bx0 + (lane & 1 ? 252 : 0));
					if (fr.lo_col && (unsigned)(fr.r - 4) < (unsigned)buf_rows && (unsigned)(fr.c - 4) < (unsigned)cw)
						*reinterpret_cast<u32 *>(dst + fr.img) = tile[rr * kTileWords + ((lane & 1) ? 66 : 3)];
				}
			} else if (kLoad == kLoadLdgsts) {
				/* rows start on 8-byte boundaries: 64-bit stores.  The item's columns begin 4 bytes before such a
				 * boundary (at one, for a UV item): 31 aligned units per row (tile byte 16 + 8 l <-> plane byte
				 * 256 chunk + 8 l) and one word at each end, or 32 units (tile byte 8 + 8 l) */
				const int xb0 = 256 * (int)((desc >> 2) & 0x3ffu);             /* plane byte of tile byte 16 */
				const int col = uv ? xb0 - 8 + 8 * lane : xb0 + 8 * lane;
				const bool col_ok = (uv || lane < 31) && col >= 0 && col < pl.pw;
				const unsigned char *sp = tile8 + (uv ? 8 : 16) + 8 * lane;
				unsigned char *gp = dst + (long long)y0 * pl.pitch + col;
#pragma unroll 1
				for (int rr = 0; rr < kTileH; rr++) {
					const int y = y0 + rr;
					if (col_ok && y >= 0 && y < pl.buf_rows) *reinterpret_cast<uint2 *>(gp) = *reinterpret_cast<const uint2 *>(sp);
					sp += kTileW; gp += pl.pitch;
				}
				if (!uv) {
#pragma unroll 1
					for (int r0 = 0; r0 < kTileH; r0 += 16) {
						const int rr = r0 + (lane >> 1), y = y0 + rr;
						const int fw = (lane & 1) ? 66 : 3;                      /* tile word (byte 264 or 12); plane byte xb0 + 248 or xb0 - 4 */
						const int fc = (lane & 1) ? xb0 + 248 : xb0 - 4;
						if (rr < kTileH && fc >= 0 && fc < pl.pw && y >= 0 && y < pl.buf_rows)
							*reinterpret_cast<u32 *>(dst + (long long)y * pl.pitch + fc) = tile[rr * kTileWords + fw];
					}
				}
			} else if (y0 >= 0 && y0 + kTileH <= pl.buf_rows) {
				/* every row of the tile is inside the plane */
#pragma unroll
				for (int rr = 0; rr < kTileH; rr += 2) {
					if (vec_lane) *reinterpret_cast<uint4 *>(vdst) = *reinterpret_cast<const uint4 *>(vsrc + rr * kTileW);
					vdst += pitch2;
				}
#pragma unroll
				for (int rr = 0; rr < kTileH; rr += 8)
					if (f_lane) *reinterpret_cast<u32 *>(dst + (long long)(y0 + rr + ftr) * pl.pitch + fcol) = tile[(rr + ftr) * kTileWords + fword];
			} else {
				/* first / last rows of a plane: the reference writes these into its padding; drop them */
#pragma unroll
				for (int rr = 0; rr < kTileH; rr += 2) {
					const int y = y0 + rr + half;
					if (vec_lane && y >= 0 && y < pl.buf_rows) *reinterpret_cast<uint4 *>(vdst) = *reinterpret_cast<const uint4 *>(vsrc + rr * kTileW);
					vdst += pitch2;
				}
#pragma unroll
				for (int rr = 0; rr < kTileH; rr += 8) {
					const int y = y0 + rr + ftr;
					if (f_lane && y >= 0 && y < pl.buf_rows) *reinterpret_cast<u32 *>(dst + (long long)y * pl.pitch + fcol) = tile[(rr + ftr) * kTileWords + fword];
				}
			}
		}
		/* the stage is free again: refill it with the item kStages ahead */
		if (kLoad != kLoadLdgsts) FenceProxyAsyncShared();
		__syncwarp();
		if (lane == 0) issue_load(s);
		__syncwarp();
		if (kLoad == kLoadLdgsts) fetch_tile(s);
		if (++s == kStages) { s = 0; parity ^= 1u; }
	}
	if (lane == 0) {
		/* the last run request must have returned before the counter may be reset */
		asm volatile("" :: "r"(run_pending) : "memory");
#if HDB_TRACE
		const unsigned w = blockIdx.x * kWarps + warp;
		if (w < 8192) { g_trace[w][0] = t_entry; g_trace[w][1] = t_first; g_trace[w][2] = GlobalTimerNs(); g_trace[w][3] = n_done; }
#endif
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(&sched->done, 1u) == gridDim.x - 1) {
			sched->next = 0u;
			sched->done = 0u;
			__threadfence();
		}
	}
}

#if HDB_TRACE
extern "C" int hdb_debug_read_trace(unsigned long long *dst, int n_warps)
{
	return (int)cudaMemcpyFromSymbol(dst, g_trace, sizeof(unsigned long long) * 4 * (size_t)std::min(n_warps, 8192));
}
#endif

/* ------------------------------------------------------------------ */
/* the same kernel for 16-bit samples   */
/* ------------------------------------------------------------------ */
/*
 * Planar frames of 16-bit containers holding kDepth-bit samples (deblock_core16.cuh; depth 8 .. 10).  Items,
 * scheduling, boundary-strength prefetch and the per-warp two-stage TMA pipeline are those of DeblockTiledKernel;
 * what differs is the geometry -- a block row is 16 bytes per block, so an item's box is 528 x 16 bytes (8 bytes
 * of lead-in to the 16-byte boundary, 32 blocks of 16 bytes, 8 of tail: 132 32-bit elements) at byte
 * 512 * chunk - 16 of the row -- and the arithmetic.  Per pixel it moves twice the bytes for the same number of
 * instructions, which puts it nearer the HBM roofline than the 8-bit kernel.  TMA loads only: the chroma pitch
 * (width bytes) must be a multiple of 16, i.e. the width a multiple of 16.
 */
constexpr int kTileW16 = 528;
constexpr int kTileWords16 = kTileW16 / 4;        /* 132 */
constexpr int kTileBytes16 = kTileW16 * kTileH;   /* 8448: a multiple of 128 */
constexpr int kSmemBytes16 = kWarps * kStages * (kTileBytes16 + 8 + 8);
#ifndef HDB_MIN_CTAS16
#define HDB_MIN_CTAS16 12
#endif

/*
 * One block row of an interleaved-UV item of a P010 frame (16-bit samples in the HIGH bits, chroma interleaved:
 * what a hardware decoder writes for 10-bit video): as UvBlockRow, lanes 0-15 filter the U samples of block pair
 * lane & 15 and lanes 16-31 its V samples.  The pair's 32 bytes of a row are eight words (U_k | V_k << 16); a lane
 * picks its plane's halves, brings them down to the low bits, filters, swaps two words per row with its partner
 * and re-interleaves four of the eight words.  Every lane runs the whole sequence (full-mask shuffles); `live`
 * only gates the write-back.
 */
template <int kDepth>
__device__ __forceinline__ void UvBlockRow16(u32 *base, u32 s4, const FilterParams16 &fp, bool live, bool v_plane)
{
	constexpr int kShift = 16 - kDepth;
	const u32 sel = v_plane ? 0x7632u : 0x5410u;
	u32 w[8][4];
#pragma unroll
	for (int r = 0; r < 8; r++) {
		const uint4 a = *reinterpret_cast<const uint4 *>(base + r * kTileWords16);
		const uint4 c = *reinterpret_cast<const uint4 *>(base + r * kTileWords16 + 4);
		w[r][0] = (Prmt(a.x, a.y, sel) >> kShift) & Depth<kDepth>::kMaxRep;
		w[r][1] = (Prmt(a.z, a.w, sel) >> kShift) & Depth<kDepth>::kMaxRep;
		w[r][2] = (Prmt(c.x, c.y, sel) >> kShift) & Depth<kDepth>::kMaxRep;
		w[r][3] = (Prmt(c.z, c.w, sel) >> kShift) & Depth<kDepth>::kMaxRep;
	}
	ChromaBlock16<kDepth>(w, s4 & 0xff, (s4 >> 8) & 0xff, (s4 >> 16) & 0xff, s4 >> 24, fp);
#pragma unroll
	for (int r = 0; r < 8; r++) {
		/* the U lane writes words 0-3 (samples 0-3 of both planes), the V lane words 4-7 */
		const u32 g0 = __shfl_xor_sync(0xffffffffu, v_plane ? w[r][0] : w[r][2], 16);
		const u32 g1 = __shfl_xor_sync(0xffffffffu, v_plane ? w[r][1] : w[r][3], 16);
		const u32 u0 = v_plane ? g0 : w[r][0], u1 = v_plane ? g1 : w[r][1];
		const u32 v0 = v_plane ? w[r][2] : g0, v1 = v_plane ? w[r][3] : g1;
		uint4 o;
		o.x = PairLo(u0, v0) << kShift; o.y = PairHi(u0, v0) << kShift;
		o.z = PairLo(u1, v1) << kShift; o.w = PairHi(u1, v1) << kShift;
		if (live) *reinterpret_cast<uint4 *>(base + r * kTileWords16 + (v_plane ? 4 : 0)) = o;
	}
}

/* kP010: two planes, Y and interleaved UV, samples in the high bits of their containers (P010 / P016) */
template <int kDepth, bool kEdgeQp, bool kP010>
__global__ void __launch_bounds__(kWarps * 32, HDB_MIN_CTAS16)
DeblockTiled16Kernel(const __grid_constant__ TensorMaps3 maps, const __grid_constant__ TiledBatch b,
	SchedWords *sched, const uint32_t run_len)
{
	extern __shared__ __align__(HDB_SMEM_ALIGN) unsigned char smem[];
	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	unsigned char *tiles = smem + warp * (kStages * kTileBytes16);
	uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kWarps * kStages * kTileBytes16) + warp * kStages;
	uint2 *held = reinterpret_cast<uint2 *>(smem + kWarps * kStages * (kTileBytes16 + 8)) + warp * kStages;
	const uint32_t tiles_s = SmemAddr(tiles);
	const uint32_t bars_s = SmemAddr(bars);
	const uint32_t total = (uint32_t)b.total_items;

	uint32_t run_next = 0, run_end = 0, run_pending = 0, pending_len = 0;
	const uint32_t two_grid_warps = 2u * gridDim.x * kWarps;
	auto next_len = [&](uint32_t seen) -> uint32_t {
		const uint32_t left = total > seen ? total - seen : 0u;
		return min(max(left / two_grid_warps, 1u), run_len);
	};
	ItemCursor cur;
	cur.frame = cur.plane = cur.grp = cur.chunk = 0; cur.chunks = cur.groups = 1;
	auto issue_load = [&](int s) {         /* lane 0 only */
		if (run_next == run_end) {
			run_next = run_pending;
			run_end = run_pending + pending_len;
			if (run_next < total) {
				pending_len = next_len(run_end);
				run_pending = atomicAdd(&sched->next, pending_len);
				cur.Seek(b, run_next);
			}
		} else {
			cur.template Next<kP010>(b);
		}
		if (run_next >= total) { held[s] = make_uint2(0u, kNoItem); return; }
		run_next++;
		held[s] = make_uint2(cur.frame, cur.Desc());
		const uint32_t bar = bars_s + 8u * s;
		MbarArriveExpectTx(bar, kTileBytes16);
		/* x in 32-bit elements: (512 * chunk - 16) / 4 */
		TmaLoad3D(tiles_s + s * kTileBytes16, &maps.in[cur.plane], 128 * (int)cur.chunk - 4,
			8 * kRB * (int)cur.grp + (cur.plane ? b.chroma_y_org : b.luma_y_org), (int)cur.frame, bar);
	};
	if (lane == 0) {
		for (int s = 0; s < kStages; s++) MbarInit(bars_s + 8u * s, 1);
		FenceMbarInit();
		pending_len = next_len(0u);
		run_pending = atomicAdd(&sched->next, pending_len);
		for (int s = 0; s < kStages; s++) issue_load(s);
	}
	__syncwarp();

	int s = 0;
	uint32_t parity = 0;
	uint32_t cur_frame = 0xffffffffu;
	int cur_plane = -1;
	FilterParams16 fp;
	PlaneInfo pl;
	MakeFilterParams16<kDepth>(fp, 0, 0);
	pl = MakePlaneInfo<kP010>(b, 0);
	BsRaw bs_raw[kRB];
	{
		const uint2 it0 = held[0];
#pragma unroll
		for (int rb = 0; rb < kRB; rb++) bs_raw[rb] = ItemBsRaw<kP010>(b, it0.x, it0.y, lane, rb);
	}
	for (;;) {
		const uint2 it = held[s];
		const uint32_t frame = it.x, desc = it.y;
		if (desc == kNoItem) break;
		int plane, bx0, lby0;
		ItemCoords<kP010>(desc, plane, bx0, lby0);
		const bool chroma = plane != 0;
		if (frame != cur_frame) {
			cur_frame = frame;
			const unsigned qp = b.qp ? __ldg(b.qp + frame) : b.qp_value;
			MakeFilterParams16<kDepth>(fp, BetaOf(qp), TcOf(qp));      /* cpu.h:136-137, scaled to the depth */
		}
		if (plane != cur_plane) { cur_plane = plane; pl = MakePlaneInfo<kP010>(b, plane); }
		const int bx = bx0 + ((kP010 && chroma) ? (lane & 15) : lane);
		u32 bs[kRB];
		{
			const uint2 it_next = held[s + 1 == kStages ? 0 : s + 1];
#pragma unroll
			for (int rb = 0; rb < kRB; rb++) bs[rb] = PackBs(bs_raw[rb], chroma);
#pragma unroll
			for (int rb = 0; rb < kRB; rb++) bs_raw[rb] = ItemBsRaw<kP010>(b, it_next.x, it_next.y, lane, rb);
		}
		MbarWait(bars_s + 8u * s, parity);
		u32 *tile = reinterpret_cast<u32 *>(tiles + s * kTileBytes16);
		unsigned char *dst = pl.out + (long long)frame * pl.frame_stride;
#pragma unroll 1
		for (int rb = 0; rb < kRB; rb++) {
			if (kP010 && chroma) {
				u32 s4 = bs[0];
#pragma unroll
				for (int k = 1; k < kRB; k++) if (rb == k) s4 = bs[k];
				/* block pair lane & 15: 32 bytes per row from the box origin on (the pair of block column 16 chunk
				 * starts exactly there) */
				UvBlockRow16<kDepth>(tile + rb * 8 * kTileWords16 + 8 * (lane & 15), s4, fp, bx < pl.nbx && lby0 + rb < pl.rows, lane >= 16);
			} else if (bx < pl.nbx && lby0 + rb < pl.rows) {
				/* lane l owns bytes 8 + 16 l .. 23 + 16 l of each of the block row's 8 tile rows */
				u32 *base = tile + rb * 8 * kTileWords16 + 2 + 4 * lane;
				u32 w[8][4];
#pragma unroll
				for (int r = 0; r < 8; r++) {
					const uint2 a = *reinterpret_cast<const uint2 *>(base + r * kTileWords16);
					const uint2 c = *reinterpret_cast<const uint2 *>(base + r * kTileWords16 + 2);
					w[r][0] = a.x; w[r][1] = a.y; w[r][2] = c.x; w[r][3] = c.y;
					if (kP010) {
#pragma unroll
						for (int j = 0; j < 4; j++) w[r][j] = (w[r][j] >> (16 - kDepth)) & Depth<kDepth>::kMaxRep;
					}
				}
				u32 s4 = bs[0];
#pragma unroll
				for (int k = 1; k < kRB; k++) if (rb == k) s4 = bs[k];
				if (chroma) ChromaBlock16<kDepth>(w, s4 & 0xff, (s4 >> 8) & 0xff, (s4 >> 16) & 0xff, s4 >> 24, fp);
				else LumaBlock16<kDepth, kEdgeQp>(w, s4 & 0xff, (s4 >> 8) & 0xff, (s4 >> 16) & 0xff, s4 >> 24, fp);
#pragma unroll
				for (int r = 0; r < 8; r++) {
					if (kP010) {
#pragma unroll
						for (int j = 0; j < 4; j++) w[r][j] <<= 16 - kDepth;
					}
					*reinterpret_cast<uint2 *>(base + r * kTileWords16) = make_uint2(w[r][0], w[r][1]);
					*reinterpret_cast<uint2 *>(base + r * kTileWords16 + 2) = make_uint2(w[r][2], w[r][3]);
				}
			}
		}
		__syncwarp();
		/* the item's own bytes of every row: tile bytes 8 .. 519 <-> plane bytes 512 chunk - 8 .. + 503: thirty-one
		 * 16-byte vectors (tile byte 16 + 16 j <-> plane byte 512 chunk + 16 j) and two words at each end */
		{
			const int y0 = 8 * lby0 + pl.y_org;
			const int pwb = 2 * pl.pw;                                  /* plane width in bytes */
			const int xb0 = 512 * (int)((desc >> 2) & 0x3ffu);
			const unsigned char *tile8 = reinterpret_cast<const unsigned char *>(tile);
			const bool uv = kP010 && chroma;                           /* a UV item owns tile bytes 0 .. 511: 32 vectors, no fringe */
			const int vcol = uv ? xb0 - 16 + 16 * lane : xb0 + 16 * lane;
			const bool vec_lane = (uv || lane < 31) && vcol >= 0 && vcol < pwb;
			const unsigned char *vsrc = tile8 + (uv ? 0 : 16) + 16 * lane;
			unsigned char *vdst = dst + (long long)y0 * pl.pitch + vcol;
			const int k = lane & 3, ftr = lane >> 2;
			const int fword = k < 2 ? 2 + k : 126 + k;                  /* tile words 2, 3 and 128, 129 */
			const int fcol = k < 2 ? xb0 - 8 + 4 * k : xb0 + 496 + 4 * (k - 2);
			const bool f_lane = !uv && fcol >= 0 && fcol < pwb;
#pragma unroll
			for (int rr = 0; rr < kTileH; rr++) {
				const int y = y0 + rr;
				if (vec_lane && y >= 0 && y < pl.buf_rows) *reinterpret_cast<uint4 *>(vdst) = *reinterpret_cast<const uint4 *>(vsrc + rr * kTileW16);
				vdst += pl.pitch;
			}
#pragma unroll
			for (int rr = 0; rr < kTileH; rr += 8) {
				const int y = y0 + rr + ftr;
				if (f_lane && y >= 0 && y < pl.buf_rows) *reinterpret_cast<u32 *>(dst + (long long)y * pl.pitch + fcol) = tile[(rr + ftr) * kTileWords16 + fword];
			}
		}
		FenceProxyAsyncShared();
		__syncwarp();
		if (lane == 0) issue_load(s);
		__syncwarp();
		if (++s == kStages) { s = 0; parity ^= 1u; }
	}
	if (lane == 0) asm volatile("" :: "r"(run_pending) : "memory");
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(&sched->done, 1u) == gridDim.x - 1) {
			sched->next = 0u;
			sched->done = 0u;
			__threadfence();
		}
	}
}

typedef void (*TiledKernelFn16)(const TensorMaps3, const TiledBatch, SchedWords *, const uint32_t);

/* variant index = (edge maps ? 1 : 0) | (depth - 8) << 1; 6, 7: P010 (depth 10, semi-planar, high-aligned) */
static TiledKernelFn16 TiledKernel16Of(int v)
{
	switch (v) {
	case 0: return DeblockTiled16Kernel<8, false, false>;
	case 1: return DeblockTiled16Kernel<8, true, false>;
	case 2: return DeblockTiled16Kernel<9, false, false>;
	case 3: return DeblockTiled16Kernel<9, true, false>;
	case 4: return DeblockTiled16Kernel<10, false, false>;
	case 5: return DeblockTiled16Kernel<10, true, false>;
	case 6: return DeblockTiled16Kernel<10, false, true>;
	case 7: return DeblockTiled16Kernel<10, true, true>;
	}
	return nullptr;
}

cudaError_t PrepareTiled16(int ctas_per_sm[kTiled16Variants])
{
	for (int v = 0; v < kTiled16Variants; v++) {
		const TiledKernelFn16 fn = TiledKernel16Of(v);
		cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes16);
		if (err != cudaSuccess) return err;
		int n = 0;
		err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, kWarps * 32, kSmemBytes16);
		if (err != cudaSuccess) return err;
		ctas_per_sm[v] = std::max(n, 1);
	}
	return cudaSuccess;
}

cudaError_t LaunchTiled16(const TensorMaps3 &maps, const TiledBatch &b, void *sched, int sm_count,
	const int ctas_per_sm[kTiled16Variants], cudaStream_t stream, int bit_depth, bool edge_qp)
{
	if (b.total_items <= 0) return cudaSuccess;
	if (b.total_items >= (1ll << 31) || bit_depth < 8 || bit_depth > 10) return cudaErrorInvalidValue;
	if (b.nv12 && bit_depth != 10) return cudaErrorInvalidValue;
	const int variant = b.nv12 ? (6 | (edge_qp ? 1 : 0)) : ((edge_qp ? 1 : 0) | ((bit_depth - 8) << 1));
	const int resident = std::max(ctas_per_sm[variant], 1);
	long long ctas = std::min<long long>((long long)sm_count * resident, (b.total_items + kWarps - 1) / kWarps);
	ctas = std::max<long long>(ctas, 1);
	const uint32_t run_len = b.total_items / (ctas * kWarps) >= 512 ? 8 : 4;
	TiledKernel16Of(variant)<<<(unsigned)ctas, kWarps * 32, kSmemBytes16, stream>>>(maps, b, reinterpret_cast<SchedWords *>(sched), run_len);
	return cudaGetLastError();
}

int HostBeta(unsigned qp) { return BetaOf(qp); }
int HostTc(unsigned qp) { return TcOf(qp); }

TiledConfig GetTiledConfig()
{
	TiledConfig c;
	c.block_rows_per_item = kRB; c.stages = kStages; c.warps_per_cta = kWarps;
	c.smem_bytes = kSmemBytes; c.max_ctas_per_sm = HDB_MIN_CTAS;
	return c;
}

void FinalizeTiledBatch(TiledBatch &b)
{
	b.luma_chunks = (b.luma_nbx + 31) / 32;
	b.chroma_chunks = b.nv12 ? (b.chroma_nbx + 15) / 16 : (b.chroma_nbx + 31) / 32;
	b.luma_groups = (b.luma_rows + kRB - 1) / kRB;
	b.chroma_groups = (b.chroma_rows + kRB - 1) / kRB;
	b.items_luma = b.luma_chunks * b.luma_groups;
	b.items_chroma = b.chroma_chunks * b.chroma_groups;
	b.items_per_frame = b.items_luma + (b.nv12 ? 1 : 2) * b.items_chroma;
	b.total_items = (long long)b.items_per_frame * b.n_frames;
	b.div_items_per_frame.Init((uint32_t)b.items_per_frame);
	b.div_luma_chunks.Init((uint32_t)b.luma_chunks);
	b.div_chroma_chunks.Init((uint32_t)b.chroma_chunks);
	b.div_items_chroma.Init((uint32_t)std::max(b.items_chroma, 1));
	/* the reference addresses a padded chroma plane with a pitch of 8 * (number of blocks) (cpu.h:469-471), which
	 * is 4 short of its real pitch when the chroma width is 4 mod 8 */
	const int cw = b.width / 2;
	b.chroma_skew = (!b.nv12 && cw % 8 != 0) ? 1 : 0;
	b.chroma_vpitch = 8 * b.chroma_nbx;
	b.chroma_rpitch = cw + 8;
	b.div_chroma_rpitch.Init((uint32_t)b.chroma_rpitch);
	if (b.chroma_skew && b.chroma_row0 == 0) {
		/* The virtual rows run 4 bytes short of the buffer rows, so the reference's block rows end before the
		 * buffer does: the last image rows lie beyond them and it never touches them (Save copies them out as they
		 * came in).  Out of place they still have to be moved: row groups to the end of the buffer, the ones past
		 * the last block row without arithmetic. */
		const long long buffer_bytes = (long long)(b.height / 2 + 8) * b.chroma_rpitch;
		const int virtual_rows = (int)((buffer_bytes + b.chroma_vpitch - 1) / b.chroma_vpitch);
		b.chroma_groups = std::max(b.chroma_groups, (virtual_rows + 8 * kRB - 1) / (8 * kRB));
		b.items_chroma = b.chroma_chunks * b.chroma_groups;
		b.items_per_frame = b.items_luma + 2 * b.items_chroma;
		b.total_items = (long long)b.items_per_frame * b.n_frames;
		b.div_items_per_frame.Init((uint32_t)b.items_per_frame);
		b.div_items_chroma.Init((uint32_t)std::max(b.items_chroma, 1));
	}
}

typedef void (*TiledKernelFn)(const TensorMaps3, const TiledBatch, SchedWords *, const uint32_t);

/* variant index = (edge maps ? 1 : 0) | (NV12 ? 2 : 0) | load mode << 2; strips are planar only */
static TiledKernelFn TiledKernelOf(int v)
{
	switch (v) {
	case 0: return DeblockTiledKernel<false, false, kLoadTma>;
	case 1: return DeblockTiledKernel<true, false, kLoadTma>;
	case 2: return DeblockTiledKernel<false, true, kLoadTma>;
	case 3: return DeblockTiledKernel<true, true, kLoadTma>;
	case 4: return DeblockTiledKernel<false, false, kLoadTmaHalo>;
	case 5: return DeblockTiledKernel<true, false, kLoadTmaHalo>;
	case 8: return DeblockTiledKernel<false, false, kLoadLdgsts>;
	case 9: return DeblockTiledKernel<true, false, kLoadLdgsts>;
	case 10: return DeblockTiledKernel<false, true, kLoadLdgsts>;
	case 11: return DeblockTiledKernel<true, true, kLoadLdgsts>;
	}
	return nullptr;
}

static int VariantOf(const TiledBatch &b, bool edge_qp)
{
	return (edge_qp ? 1 : 0) | (b.nv12 ? 2 : 0) | (b.load_mode << 2);
}

cudaError_t PrepareTiled(int ctas_per_sm[kTiledVariants])
{
	for (int v = 0; v < kTiledVariants; v++) {
		ctas_per_sm[v] = 1;
		const TiledKernelFn fn = TiledKernelOf(v);
		if (!fn) continue;
		cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
		if (err != cudaSuccess) return err;
		int n = 0;
		err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, kWarps * 32, kSmemBytes);
		if (err != cudaSuccess) return err;
		ctas_per_sm[v] = std::max(n, 1);
	}
	return cudaSuccess;
}

cudaError_t LaunchTiled(const TensorMaps3 &maps, const TiledBatch &b, void *sched, int sm_count, const int ctas_per_sm[kTiledVariants],
	cudaStream_t stream, bool edge_qp)
{
	if (b.total_items <= 0) return cudaSuccess;
	if (b.total_items >= (1ll << 31)) return cudaErrorInvalidValue;
	const int variant = VariantOf(b, edge_qp);
	const TiledKernelFn fn = TiledKernelOf(variant);
	if (!fn) return cudaErrorInvalidValue;
	const long long warps_wanted = b.total_items;
	const int resident = std::max(ctas_per_sm[variant], 1);
	long long ctas = std::min<long long>((long long)sm_count * resident, (warps_wanted + kWarps - 1) / kWarps);
	ctas = std::max<long long>(ctas, 1);
	/* consecutive items per scheduler request: an upper bound, the kernel shortens the runs as the work runs
	 * out.  Longer runs re-read fewer tile-edge sectors (neighbouring items share one per row); shorter runs
	 * level the end of a short launch out better. */
	uint32_t run_len = b.total_items / (ctas * kWarps) >= 512 ? 8 : 4;
	static const int run_len_env = getenv("HDB_RUN_LEN") ? atoi(getenv("HDB_RUN_LEN")) : 0;   /* measurement knob */
	if (run_len_env > 0) run_len = (uint32_t)run_len_env;
	fn<<<(unsigned)ctas, kWarps * 32, kSmemBytes, stream>>>(maps, b, reinterpret_cast<SchedWords *>(sched), run_len);
	return cudaGetLastError();
}

size_t TiledSchedBytes() { return sizeof(SchedWords); }

/* ------------------------------------------------------------------ */
/* small helper kernels                                                */
/* ------------------------------------------------------------------ */
__global__ void ChromaBsTableKernel(uint32_t *table, int width, int height)
{
	const int cnbx = (width / 2 + 8) / 8, cnby = (height / 2 + 8) / 8;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= cnbx * cnby) return;
	const int by = i / cnbx, bx = i - by * cnbx;
	table[i] = ChromaBsOfBlock(bx, by, (unsigned)width / 2, (unsigned)height / 2, (width + 8) / 8, (height + 8) / 8);
}

cudaError_t LaunchChromaBsTable(uint32_t *table, int width, int height, cudaStream_t stream)
{
	const int n = ((width / 2 + 8) / 8) * ((height / 2 + 8) / 8);
	ChromaBsTableKernel<<<(n + 255) / 256, 256, 0, stream>>>(table, width, height);
	return cudaGetLastError();
}

/* cpu.h:92-99 */
__global__ void DefaultLumaBsKernel(unsigned char *vert_bs, unsigned char *hor_bs, int width, int height)
{
	const unsigned nv = NumVertBs(width, height), nh = NumHorBs(width, height);
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < nv) vert_bs[i] = (i % (width / 8 + 1) == 0) ? 0 : 2;
	if (i < nh) hor_bs[i] = (i % (height / 8 + 1) == 0) ? 0 : 2;
}

cudaError_t LaunchDefaultLumaBs(unsigned char *vert_bs, unsigned char *hor_bs, int width, int height,
	cudaStream_t stream)
{
	const unsigned n = std::max(NumVertBs(width, height), NumHorBs(width, height));
	if (n == 0) return cudaSuccess;
	DefaultLumaBsKernel<<<(n + 255) / 256, 256, 0, stream>>>(vert_bs, hor_bs, width, height);
	return cudaGetLastError();
}

/*
 * Edge maps from block data (deblock_core.cuh: DeriveEdge).  Point (r, x) of the (H/8 + 1) x (W/8 + 1)
 * grid of block corners owns the vertical entry [r][x] -- the edge between blocks (r, x-1) and (r, x) --
 * and the horizontal entry [r][x] -- between (r-1, x) and (r, x).  Entries on the picture's four sides
 * are 0 (H.265 does not filter them).  One thread walks one column of corners down kDeriveRows rows: it
 * fetches and normalises its own block's record once per row (three 32-bit loads; the 12-byte records of
 * neighbouring lanes share their sectors), keeps it for the next row, where it is the block above, and gets the
 * block to the left from the neighbouring lane by shuffle (a warp covers 31 columns; its lane 0 repeats the column
 * before them and writes nothing).  No shared memory and no barrier:
 * the first version staged raw records of 128 x 8 corners in shared memory and normalised them in every edge that
 * used them -- 200 instructions per corner and a barrier stall on top (profiles/r01o_ncu_variants_summary.json).
 * Per 8x8 block 12 bytes are read (+ 1/kDeriveRows for the row above) and 2 written, against the 192 the filter
 * moves.
 */
constexpr int kDeriveWarps = 4, kDeriveRows = 32;
constexpr int kDeriveCols = 31 * kDeriveWarps;     /* corner columns per CTA: lane 0 of every warp only provides the block to the left of lane 1 */

__global__ void __launch_bounds__(32 * kDeriveWarps)
DeriveEdgesKernel(const uint32_t *info, int w8, int h8, int n_frames, unsigned char *vert_edges, unsigned char *hor_edges)
{
	const int lane = threadIdx.x & 31;
	/* corner column of this lane: a warp covers 31 of them, and its lane 0 repeats the last column of the warp
	 * before it (column -1 for the first: no block there) so that every lane finds its left neighbour one lane down */
	const int x = blockIdx.x * kDeriveCols + 31 * (int)(threadIdx.x >> 5) + lane - 1;
	const int r0 = blockIdx.y * kDeriveRows, r1 = min(r0 + kDeriveRows, h8 + 1);
	const long long nv = (long long)(w8 + 1) * h8, nh = (long long)(h8 + 1) * w8;
	const bool col_has_block = x >= 0 && x < w8;
	const bool writes = lane > 0;
	for (int f = blockIdx.z; f < n_frames; f += gridDim.z) {
		const uint32_t *rec = info + 3ll * ((long long)f * w8 * h8 + (long long)(r0 - 1) * w8 + x);     /* record (r0 - 1, x) */
		auto fetch = [&](int r) -> BlockWords {
			BlockWords raw = { 0u, 0u, 0u };
			if (col_has_block && r >= 0 && r < h8) { raw.mv0 = __ldg(rec); raw.mv1 = __ldg(rec + 1); raw.meta = __ldg(rec + 2); raw = NormalizeBlock(raw); }
			rec += 3 * w8;
			return raw;
		};
		BlockWords above = fetch(r0 - 1);
		unsigned char *vf = vert_edges + f * nv + (long long)r0 * (w8 + 1) + x;
		unsigned char *hf = hor_edges + f * nh + (long long)r0 * w8 + x;
#pragma unroll 2
		for (int r = r0; r < r1; r++) {
			const BlockWords q = fetch(r);
			BlockWords left;
			left.mv0 = __shfl_up_sync(0xffffffffu, q.mv0, 1);
			left.mv1 = __shfl_up_sync(0xffffffffu, q.mv1, 1);
			left.meta = __shfl_up_sync(0xffffffffu, q.meta, 1);
			const u32 fl = q.meta >> 16;
			const bool has_q = r < h8 && col_has_block;
			u32 ev = 0, eh = 0;
			if (has_q && x > 0) ev = DeriveEdge(left, q, fl & kBlkTuLeft, fl & kBlkPuLeft);
			if (has_q && r > 0) eh = DeriveEdge(above, q, fl & kBlkTuTop, fl & kBlkPuTop);
			if (writes && r < h8 && x <= w8) *vf = (unsigned char)ev;
			if (writes && x < w8) *hf = (unsigned char)eh;
			vf += w8 + 1; hf += w8;
			above = q;
		}
	}
}

cudaError_t LaunchDeriveEdges(const void *info, int width, int height, int n_frames,
	unsigned char *vert_edges, unsigned char *hor_edges, cudaStream_t stream)
{
	const int w8 = width / 8, h8 = height / 8;
	if (n_frames <= 0 || w8 <= 0 || h8 <= 0) return cudaSuccess;
	const dim3 grid((w8 + kDeriveCols) / kDeriveCols, (h8 + kDeriveRows) / kDeriveRows, (unsigned)std::min(n_frames, 32768));
	DeriveEdgesKernel<<<grid, 32 * kDeriveWarps, 0, stream>>>(static_cast<const uint32_t *>(info), w8, h8, n_frames, vert_edges, hor_edges);
	return cudaGetLastError();
}

/* ------------------------------------------------------------------ */
/* generic path                                                        */
/* ------------------------------------------------------------------ */
struct GenericPlane {
	int w, h;            /* tight plane size */
	int pw, ph;          /* padded size (w + 8, h + 8) */
	int nbx, nby;        /* pw / 8, ph / 8 */
	long long plane_off; /* offset of the plane inside a tight frame */
	long long pad_off;   /* offset of the padded plane inside a frame's scratch */
};

__host__ __device__ inline GenericPlane MakeGenericPlane(int width, int height, int plane)
{
	GenericPlane g;
	const int cw = width / 2, ch = height / 2;
	g.w = plane ? cw : width; g.h = plane ? ch : height;
	g.pw = g.w + 8; g.ph = g.h + 8;
	g.nbx = g.pw / 8; g.nby = g.ph / 8;
	g.plane_off = plane == 0 ? 0 : (long long)width * height + (long long)(plane - 1) * cw * ch;
	const long long ysz = (((long long)(width + 8) * (height + 8)) + 15) & ~15ll;
	const long long csz = (((long long)(cw + 8) * (ch + 8)) + 15) & ~15ll;
	g.pad_off = plane == 0 ? 0 : ysz + (plane - 1) * csz;
	return g;
}

size_t GenericScratchBytes(int width, int height)
{
	const long long ysz = (((long long)(width + 8) * (height + 8)) + 15) & ~15ll;
	const long long csz = (((long long)(width / 2 + 8) * (height / 2 + 8)) + 15) & ~15ll;
	return (size_t)(ysz + 2 * csz);
}

/* tight -> padded at (4,4), zero border (cpu.h:55-82 with the padding defined as zero) */
__global__ void GenericPadKernel(GenericBatch b, long long scratch_per_frame)
{
	const int plane = blockIdx.y, frame = blockIdx.z;
	const GenericPlane g = MakeGenericPlane(b.width, b.height, plane);
	const unsigned char *src = b.in + frame * b.frame_stride + g.plane_off;
	unsigned char *dst = b.scratch + frame * scratch_per_frame + g.pad_off;
	const long long n = (long long)g.pw * g.ph;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const int y = (int)(i / g.pw) - 4, x = (int)(i % g.pw) - 4;
		dst[i] = (x >= 0 && x < g.w && y >= 0 && y < g.h) ? src[(long long)y * g.w + x] : 0;
	}
}

/* one thread per padded block, addressing pitch nbx * 8 as in cpu.h:169-172 / 469-471 */
template <bool kEdgeQp>
__global__ void GenericFilterKernel(GenericBatch b, long long scratch_per_frame)
{
	const int plane = blockIdx.y, frame = blockIdx.z;
	const GenericPlane g = MakeGenericPlane(b.width, b.height, plane);
	const int luma_nbx = (b.width + 8) / 8, luma_nby = (b.height + 8) / 8;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= g.nbx * g.nby) return;
	const int by = i / g.nbx, bx = i - by * g.nbx;
	const long long pitch = (long long)g.nbx * 8;
	unsigned char *blk = b.scratch + frame * scratch_per_frame + g.pad_off + (long long)by * 8 * pitch + bx * 8;
	FilterParams fp;
	const unsigned qp = b.qp ? b.qp[frame] : b.qp_value;
	MakeFilterParams(fp, BetaOf(qp), TcOf(qp));
	u32 lo[8], hi[8];
#pragma unroll
	for (int r = 0; r < 8; r++) {
		/* scratch planes are 16-byte aligned and pitch is a multiple of 8 */
		const uint2 v = *reinterpret_cast<const uint2 *>(blk + r * pitch);
		lo[r] = v.x; hi[r] = v.y;
	}
	if (plane == 0) {
		const u32 s4 = LumaBsOfBlock(bx, by, b.width >> 3, g.nbx, g.nby,
			b.vert_bs + frame * b.vert_bs_stride, b.hor_bs + frame * b.hor_bs_stride);
		LumaBlockT<kEdgeQp>(lo, hi, s4 & 0xff, (s4 >> 8) & 0xff, (s4 >> 16) & 0xff, s4 >> 24, fp);
	} else {
		const u32 s4 = ChromaBsOfBlock(bx, by, (unsigned)g.w, (unsigned)g.h, luma_nbx, luma_nby);
		ChromaBlock(lo, hi, s4 & 0xff, (s4 >> 8) & 0xff, (s4 >> 16) & 0xff, s4 >> 24, fp);
	}
#pragma unroll
	for (int r = 0; r < 8; r++) *reinterpret_cast<uint2 *>(blk + r * pitch) = make_uint2(lo[r], hi[r]);
}

/* padded -> tight (cpu.h:995-1018) */
__global__ void GenericCropKernel(GenericBatch b, long long scratch_per_frame)
{
	const int plane = blockIdx.y, frame = blockIdx.z;
	const GenericPlane g = MakeGenericPlane(b.width, b.height, plane);
	const unsigned char *src = b.scratch + frame * scratch_per_frame + g.pad_off;
	unsigned char *dst = b.out + frame * b.frame_stride + g.plane_off;
	const long long n = (long long)g.w * g.h;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const int y = (int)(i / g.w), x = (int)(i % g.w);
		dst[i] = src[(long long)(y + 4) * g.pw + x + 4];
	}
}

cudaError_t LaunchGeneric(const GenericBatch &b, cudaStream_t stream)
{
	if (b.n_frames <= 0) return cudaSuccess;
	const long long spf = (long long)GenericScratchBytes(b.width, b.height);
	const long long ypix = (long long)(b.width + 8) * (b.height + 8);
	const int nblk = ((b.width + 8) / 8) * ((b.height + 8) / 8);
	/* blockIdx.z carries the frame: at most 65535 per launch */
	for (int f0 = 0; f0 < b.n_frames; f0 += 65535) {
		GenericBatch part = b;
		part.n_frames = std::min(65535, b.n_frames - f0);
		part.in = b.in + (long long)f0 * b.frame_stride;
		part.out = b.out + (long long)f0 * b.frame_stride;
		part.scratch = b.scratch + (long long)f0 * spf;
		part.qp = b.qp ? b.qp + f0 : nullptr;
		part.vert_bs = b.vert_bs + (long long)f0 * b.vert_bs_stride;
		part.hor_bs = b.hor_bs + (long long)f0 * b.hor_bs_stride;
		const dim3 gpix((unsigned)std::min<long long>((ypix + 255) / 256, 1024), 3, (unsigned)part.n_frames);
		GenericPadKernel<<<gpix, 256, 0, stream>>>(part, spf);
		const dim3 gblk((nblk + 127) / 128, 3, (unsigned)part.n_frames);
		if (b.edge_qp) GenericFilterKernel<true><<<gblk, 128, 0, stream>>>(part, spf);
		else GenericFilterKernel<false><<<gblk, 128, 0, stream>>>(part, spf);
		GenericCropKernel<<<gpix, 256, 0, stream>>>(part, spf);
	}
	return cudaGetLastError();
}

} // namespace hdb
