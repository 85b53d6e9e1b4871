/*
 * sao_kernels.cu -- sample adaptive offset, the in-loop stage that follows deblocking in HEVC (SURVEY.md section
 * 8(f) rank 3).  The reference stops after the deblocking filter (hevc_deblocking_filter_cpu.h:134-993 is the end
 * of its in-loop chain), so this stage restates the published process (H.265 8.7.3) and is checked against
 * oracle/deblock_oracle.c::hdbo_sao_frame, which in turn agrees with an independent second statement
 * (oracle/sao_spec.py).  Parity is unpinned by nature.
 *
 * A separate pass, out of place: an edge offset needs the DEBLOCKED values of a sample's neighbours, and the
 * blocks of the deblocking kernel are independent only up to their own 64 samples -- a tile would have to be
 * deblocked with a one-block apron to offset its rim, which costs more than the second pass moves.
 *
 * One warp takes one coding tree block of one plane at a time, so type, class and offsets are the same for all its
 * lanes and the kernel branches on them: a block that is off is moved, a band-offset block never looks at a
 * neighbour, a horizontal edge class needs no other row, a vertical one no other column.  (The first version
 * walked 256-byte strips across blocks and had to compute the edge category AND the band index of every sample
 * and select: 182 instructions per 8 samples, 0.27 of the HBM peak.)  A lane owns eight samples of a row (two
 * words); ctb / 8 lanes cover the block's width, and the warp's 32 / (ctb / 8) lane groups each take a run of
 * the block's rows, four rows at a time with all their loads issued before the first is used; the classes that
 * look up and down keep six rows in registers and fetch the next four while they work.  Left / right neighbours
 * come from the adjacent lanes by shuffle; the block's first and last lane load theirs from the neighbouring
 * block.
 *
 * Blocks are handed out IN ORDER through a ticket counter (chunks of kSaoChunk consecutive blocks per warp, the
 * next ticket drawn while the current chunk is worked on).  With a static round-robin the warps drifted apart --
 * a block that is off costs a tenth of an edge-offset block -- and a block's rim rows and columns, which its
 * neighbours read too, had left the L2 by the time the neighbour came by: 1.6x the algorithmic DRAM reads.  In
 * order, every warp of the device works inside one window of a few thousand consecutive blocks.
 *
 * Four samples at a time as two 16x2 pairs, with the packed-halfword instructions the deblocking core uses:
 *   edge offset  1 + sign(c - a) per half = relu(min(256 + c - a - 255, 2)) (one VIADDMNMX.RELU); category
 *                2 + sign + sign in 0..4
 *   band offset  min(((c >> 3) - band_position) & 31, 4) in 0..4
 *   either way the index picks the offset out of an 8-byte pool with one byte permute (selector nibble | 8 for the
 *   sign-replicating upper byte), and VIADDMNMX.RELU adds it and clips to 0..255.
 * Samples with a neighbour outside the picture are left alone through a per-byte keep mask.
 */
#include "deblock_kernels.cuh"
#include "sao_core.cuh"

#include <algorithm>

namespace hdb {

/* shape knobs (tools/build_variants.sh) */
#ifndef HDB_SAO_WARPS
#define HDB_SAO_WARPS 4
#endif
#ifndef HDB_SAO_CHUNK
#define HDB_SAO_CHUNK 1
#endif
#ifndef HDB_SAO_BAND
#define HDB_SAO_BAND 0
#endif
#ifndef HDB_SAO_STATIC
#define HDB_SAO_STATIC 0
#endif
#ifndef HDB_SAO_COUNTERS
#define HDB_SAO_COUNTERS 32
#endif
constexpr int kSaoWarps = HDB_SAO_WARPS;
constexpr int kSaoCounters = HDB_SAO_COUNTERS;   /* ticket counters, a 128-byte line each */
constexpr int kSaoChunk = HDB_SAO_CHUNK;     /* consecutive blocks per ticket */
constexpr int kSaoBand = HDB_SAO_BAND ? kSaoChunk : 1;   /* block rows a chunk runs down before the next column */

struct SaoBatch {
	const unsigned char *in; unsigned char *out;    /* tight YV12 frames, frame_stride apart */
	long long frame_stride;
	int n_frames, width, height;
	int ctb;                                         /* luma CTB size: 16, 32 or 64 */
	int bit_depth;                                   /* 8 with a byte a sample; 8 .. 12 in 16-bit containers */
	int ctb_cols, ctb_rows;
	const unsigned char *params;                     /* hdb_sao_params[n][3][ctb_rows][ctb_cols], 6 bytes each */
	/* task space: frame-major, then plane, band of kSaoChunk CTB rows, CTB column, CTB row within the band */
	int tasks_per_plane, tasks_per_frame;
	long long total_tasks;
	FastDiv div_tasks_per_frame, div_tasks_per_plane, div_band;
	unsigned *ticket;                                /* zero at launch: kSaoCounters counters, 32 words apart */
};

/* what a lane knows about the block it works on */
struct SaoLane {
	const u32 *src; u32 *dst;          /* the plane, as words */
	unsigned pww;                      /* words per row */
	int ph;                            /* rows of the plane */
	/* word columns: own two (clamped into the row), the neighbouring block's (first lane: left, others: right),
	 * and the right one again for a block one lane wide */
	unsigned xc0, xc1, xe, xr;
	int ya, yb, rpg;                   /* this lane group's rows [ya, yb), rows per group */
	bool live0, live1, first_col, last_col, one_lane;
	u32 pool_lo, pool_hi, band_add, keep0, keep1;
};

/* a row as loaded (own words, the neighbouring block's) and as worked on (own words, the words left and right) */
struct RawRow { u32 w0, w1, e, e2; };
struct RowWin { u32 L, w0, w1, R; };

/* Rows outside the picture are read from the nearest one inside, words outside from the nearest inside -- a sample
 * whose neighbour lies outside is kept whatever the comparison says -- so the loads are branch-free. */
template <bool kSides, bool kPair>
__device__ __forceinline__ RawRow LoadRow(const SaoLane &t, int y)
{
	const unsigned row = (unsigned)min(max(y, 0), t.ph - 1) * t.pww;
	RawRow w{0u, 0u, 0u, 0u};
	if (kPair) {      /* rows of an even number of words: the lane's two words are one aligned load */
		const uint2 v = __ldg(reinterpret_cast<const uint2 *>(t.src + (row + t.xc0)));
		w.w0 = v.x; w.w1 = v.y;
	} else { w.w0 = __ldg(t.src + (row + t.xc0)); w.w1 = __ldg(t.src + (row + t.xc1)); }
	if (kSides) {
		w.e = __ldg(t.src + (row + t.xe));
		if (t.one_lane) w.e2 = __ldg(t.src + (row + t.xr));
	}
	return w;
}

/* the lanes hand each other their neighbours' words; every lane of the warp must be here */
template <bool kSides>
__device__ __forceinline__ RowWin Neighbours(const SaoLane &t, const RawRow &w)
{
	RowWin o{0u, w.w0, w.w1, 0u};
	if (kSides) {
		o.L = __shfl_up_sync(0xffffffffu, w.w1, 1);
		o.R = __shfl_down_sync(0xffffffffu, w.w0, 1);
		if (t.first_col) o.L = w.e;
		if (t.last_col) o.R = t.one_lane ? w.e2 : w.e;
	}
	return o;
}

template <bool kPair>
__device__ __forceinline__ void StoreRow(const SaoLane &t, u32 *orow, bool row_live, u32 v0, u32 v1)
{
	if (kPair) { if (row_live && t.live0) *reinterpret_cast<uint2 *>(orow) = make_uint2(v0, v1); }
	else {
		if (row_live && t.live0) orow[0] = v0;
		if (row_live && t.live1) orow[1] = v1;
	}
}

/* kMode 0: the block is off -- moved; 1: band offset.  No sample needs another: four rows' loads at a time. */
template <int kMode, bool kPair>
__device__ __forceinline__ void FlatRun(const SaoLane &t)
{
	u32 *orow = t.dst + ((unsigned)t.ya * t.pww + t.xc0);
#pragma unroll 1
	for (int k0 = 0; k0 < t.rpg; k0 += 4) {
		RawRow r[4];
#pragma unroll
		for (int j = 0; j < 4; j++) r[j] = k0 + j < t.rpg ? LoadRow<false, kPair>(t, t.ya + k0 + j) : RawRow{0u, 0u, 0u, 0u};
#pragma unroll
		for (int j = 0; j < 4; j++) {
			if (k0 + j < t.rpg) {       /* the same for all lanes */
				u32 v0 = r[j].w0, v1 = r[j].w1;
				if (kMode == 1) { v0 = BandOffset4(v0, t.band_add, t.pool_lo, t.pool_hi); v1 = BandOffset4(v1, t.band_add, t.pool_lo, t.pool_hi); }
				StoreRow<kPair>(t, orow, t.ya + k0 + j < t.yb, v0, v1);
			}
			orow += t.pww;
		}
	}
}

/*
 * Edge offset (8.7.3.2).  Class 0 compares a sample with its left and right neighbours, 1 with the ones above and
 * below, 2 with up-left and down-right, 3 with up-right and down-left.  The two comparisons of a sample are one
 * comparison each of two samples: with S(y, i) = 1 + sign(c(y, i) - a(y, i)), a = the FIRST neighbour (left,
 * above, up-left, up-right), the second comparison of a sample is the negated first comparison of the sample
 * that has it as its first neighbour, so
 *   category(y, i) = 2 + S(y, i) - S(y + vy, i - hx)          (vy, hx) = (0, -1), (1, 0), (1, -1), (1, 1)
 * and the walk computes one S per sample, a byte each, and reuses it a row (and / or a sample) later.  The S of
 * the one sample beyond the lane's eight comes from the neighbouring lane by shuffle, or, for the block's last /
 * first lane, from the neighbouring block's word.  The classes that look up and down run one row ahead of their
 * output: iteration y computes S(y) and writes row y - 1.
 * Rows are fetched two iterations before the shuffles that first touch them.
 */
template <int kClass, bool kPair>
__device__ __forceinline__ void EdgeRun(const SaoLane &t)
{
	constexpr bool kSides = kClass != 1, kTall = kClass != 0;
	const int n_iter = t.rpg + (kTall ? 1 : 0);
	const RawRow none{0u, 0u, 0u, 0u};
	u32 *orow = t.dst + ((long long)(t.ya - (kTall ? 1 : 0)) * (long long)t.pww + t.xc0);       /* row ya - 1 is never written */
	/* iteration i: row ya + i, fetched into `slot` two iterations ago, becomes `cur`; `above` is the row before it
	 * with its S in s_above; S of the current row goes to s_cur */
	auto step = [&](int i, RawRow &slot, const RowWin &above, const u32 (&s_above)[2], RowWin &cur, u32 (&s_cur)[2]) {
		cur = Neighbours<kSides>(t, slot);
		slot = i + 2 < n_iter ? LoadRow<kSides, kPair>(t, t.ya + i + 2) : none;
		const RowWin &ar = kTall ? above : cur;        /* the row the first neighbour lies in, and the row written */
		/* first neighbours, as words aligned with the current row's */
		u32 a0 = ar.w0, a1 = ar.w1;
		if (kClass == 3) { a0 = __funnelshift_r(ar.w0, ar.w1, 8); a1 = __funnelshift_r(ar.w1, ar.R, 8); }           /* one sample to the right */
		else if (kClass != 1) { a0 = __funnelshift_r(ar.L, ar.w0, 24); a1 = __funnelshift_r(ar.w0, ar.w1, 24); }    /* one to the left */
		const u32 s0 = Signs4(cur.w0, a0), s1 = Signs4(cur.w1, a1);
		/* S of the sample it is the first neighbour of, aligned with the lane's samples */
		u32 t0 = s0, t1 = s1;
		if (kClass == 0 || kClass == 2) {
			/* S(y, 8): the neighbouring lane's first, or sample 0 of the word to the right against sample 7 of ar */
			const u32 mine = AddMinRelu2((cur.R & 0xffu) + 0x100u - (ar.w1 >> 24), 0xff01ff01u, 0x00020002u);
			const u32 theirs = __shfl_down_sync(0xffffffffu, s0, 1);
			const u32 sr = t.last_col ? mine : theirs;
			t0 = __funnelshift_r(s0, s1, 8); t1 = __funnelshift_r(s1, sr, 8);
		}
		if (kClass == 3) {
			/* S(y, -1): the neighbouring lane's last, or sample 3 of the word to the left against sample 0 of ar */
			const u32 mine = AddMinRelu2((cur.L >> 24) + 0x100u - (ar.w0 & 0xffu), 0xff01ff01u, 0x00020002u) << 24;
			const u32 theirs = __shfl_up_sync(0xffffffffu, s1, 1);
			const u32 sl = t.first_col ? mine : theirs;
			t0 = __funnelshift_r(sl, s0, 24); t1 = __funnelshift_r(s0, s1, 24);
		}
		const u32 e0 = (kTall ? s_above[0] : s0) + 0x02020202u - t0, e1 = (kTall ? s_above[1] : s1) + 0x02020202u - t1;
		const u32 v0 = ApplyOffsets4(ar.w0, e0, t.pool_lo, t.pool_hi), v1 = ApplyOffsets4(ar.w1, e1, t.pool_lo, t.pool_hi);
		/* the row written; the picture's first and last row are left alone by the classes that look up and down */
		const int yo = t.ya + i - (kTall ? 1 : 0);
		const bool keep_row = kTall && (yo == 0 || yo == t.ph - 1);
		const u32 k0m = keep_row ? 0xffffffffu : t.keep0, k1m = keep_row ? 0xffffffffu : t.keep1;
		StoreRow<kPair>(t, orow, yo >= t.ya && yo < t.yb, (ar.w0 & k0m) | (v0 & ~k0m), (ar.w1 & k1m) | (v1 & ~k1m));
		orow += t.pww;
		s_cur[0] = s0; s_cur[1] = s1;
	};
	/* two rows a round, so that the registers swap roles instead of being copied */
	RowWin wa{0u, 0u, 0u, 0u}, wb{0u, 0u, 0u, 0u};
	u32 sa[2] = {0u, 0u}, sb[2] = {0u, 0u};
	if (kTall) wb = Neighbours<kSides>(t, LoadRow<kSides, kPair>(t, t.ya - 1));
	RawRow p = LoadRow<kSides, kPair>(t, t.ya), q = n_iter > 1 ? LoadRow<kSides, kPair>(t, t.ya + 1) : none;
#pragma unroll 1
	for (int i = 0; i < n_iter; i += 2) {
		step(i, p, wb, sb, wa, sa);
		if (i + 1 < n_iter) step(i + 1, q, wa, sa, wb, sb);
	}
}

template <bool kPair>
__global__ void __launch_bounds__(kSaoWarps * 32)
SaoKernel(const __grid_constant__ SaoBatch b)
{
	const int lane = threadIdx.x & 31;
	const unsigned total = (unsigned)b.total_tasks;
	/* Chunk v * kSaoCounters + c is ticket v of counter c.  One counter handed out 0.4 G tickets a second -- atomics
	 * on one address are serial in the L2 -- which bounded the blocks that are off or band offset; the warps are
	 * spread over kSaoCounters counters, which advance at the same pace and so still hand the blocks out in order. */
	const unsigned n_chunks = (total + kSaoChunk - 1) / kSaoChunk;
	const unsigned warp_id = blockIdx.x * kSaoWarps + (threadIdx.x >> 5);
#if HDB_SAO_STATIC
	const unsigned stride = gridDim.x * kSaoWarps;
	unsigned next = warp_id;
#else
	const unsigned counter = warp_id % kSaoCounters;
	unsigned *ticket = b.ticket + 32u * counter;
	unsigned next = 0;
	if (lane == 0) next = atomicAdd(ticket, 1u);
	next = __shfl_sync(0xffffffffu, next, 0) * kSaoCounters + counter;
#endif
	while (next < n_chunks) {
		const unsigned first = next * kSaoChunk, last = min(first + (unsigned)kSaoChunk, total);
#if HDB_SAO_STATIC
		next += stride;
#else
		if (lane == 0) next = atomicAdd(ticket, 1u);      /* in flight while this chunk is worked on */
#endif
		for (unsigned task = first; task < last; task++) {
			const uint32_t frame = b.div_tasks_per_frame.Div(task);
			uint32_t r = task - frame * (uint32_t)b.tasks_per_frame;
			const int plane = (int)b.div_tasks_per_plane.Div(r);
			r -= (uint32_t)plane * (uint32_t)b.tasks_per_plane;
			/* kSaoChunk blocks one above the other, then the next column: a chunk's blocks share a row of samples
			 * each, its neighbours in the row -- which share a column of 32-byte sectors per row with it -- are other
			 * warps' at the same time */
			const int band = (int)b.div_band.Div(r), in_band = (int)r - band * (kSaoBand * b.ctb_cols);
			const int crow = band * kSaoBand + in_band % kSaoBand, ccol = in_band / kSaoBand;
			if (crow >= b.ctb_rows) continue;             /* the last band may be short */
			const int pw = plane ? b.width >> 1 : b.width, ph = plane ? b.height >> 1 : b.height;
			const int ctb = plane ? b.ctb >> 1 : b.ctb;                 /* 64 .. 8 samples */
			const long long plane_off = plane == 0 ? 0 : (long long)b.width * b.height + (plane == 2 ? (long long)(b.width >> 1) * (b.height >> 1) : 0);
			SaoLane t;
			t.src = reinterpret_cast<const u32 *>(b.in + (long long)frame * b.frame_stride + plane_off);
			t.dst = reinterpret_cast<u32 *>(b.out + (long long)frame * b.frame_stride + plane_off);
			t.ph = ph;
			const int pww = pw >> 2;
			t.pww = (unsigned)pww;
			/* lanes: ctb / 8 of them side by side cover the block's width, eight samples each; the 32 / (ctb / 8)
			 * groups they form split the block's rows into runs */
			const int lpr = ctb >> 3, groups = 32 / lpr;
			const int col = lane & (lpr - 1), group = lane / lpr;
			t.rpg = (ctb + groups - 1) / groups;                        /* rows per group: 16, 4, 1, 1 */
			const int x = ccol * (ctb >> 2) + 2 * col;                  /* this lane's first word */
			t.live0 = x < pww; t.live1 = x + 1 < pww;
			t.ya = crow * ctb + group * t.rpg;
			t.yb = min(min(t.ya + t.rpg, (crow + 1) * ctb), ph);
			t.first_col = col == 0; t.last_col = col == lpr - 1; t.one_lane = lpr == 1;
			t.xc0 = (unsigned)min(x, pww - (kPair ? 2 : 1)); t.xc1 = (unsigned)min(x + 1, pww - 1); t.xr = (unsigned)min(x + 2, pww - 1);
			t.xe = t.first_col ? (unsigned)min(max(x - 1, 0), pww - 1) : t.xr;
			/* the block's parameters: the same for every lane */
			const unsigned char *p = b.params + 6ll * ((((long long)frame * 3 + plane) * b.ctb_rows + crow) * b.ctb_cols + ccol);
			const u32 type = __ldg(p);
			u32 pos = __ldg(p + 1);
			const u32 o0 = __ldg(p + 2), o1 = __ldg(p + 3), o2 = __ldg(p + 4), o3 = __ldg(p + 5);
			OffsetPool(type, o0, o1, o2, o3, t.pool_lo, t.pool_hi);
			if (type == 2u) pos &= 3u;
			t.band_add = (32u - pos) * 0x01010101u;
			/* bytes that keep their value whatever the row: the picture's first / last column (the vertical class
			 * does not look at them) */
			t.keep0 = (x == 0 ? 0x000000ffu : 0u) | (x == pww - 1 ? 0xff000000u : 0u);
			t.keep1 = x + 1 == pww - 1 ? 0xff000000u : 0u;
			if (type == 2u && pos == 1u) t.keep0 = t.keep1 = 0u;
			/* a dead lane's first word lies outside the row: its stores are off, its output pointer is formed from the
			 * clamped column and never used */
			if (type == 0u) FlatRun<0, kPair>(t);
			else if (type == 1u) FlatRun<1, kPair>(t);
			else if (pos == 0u) EdgeRun<0, kPair>(t);
			else if (pos == 1u) EdgeRun<1, kPair>(t);
			else if (pos == 2u) EdgeRun<2, kPair>(t);
			else EdgeRun<3, kPair>(t);
		}
#if !HDB_SAO_STATIC
		next = __shfl_sync(0xffffffffu, next, 0) * kSaoCounters + counter;
#endif
	}
}

/* ------------------------------------------------------------------ */
/* samples of 9 .. 12 bits in 16-bit containers                        */
/* ------------------------------------------------------------------ */
/*
 * The same walk with a sample per half: a lane's eight samples are four words (one 16-byte load when the rows are
 * a whole number of them), nothing is unpacked or packed, S and the categories stay 16x2.  Band shift, sample
 * maximum and the bias of the comparison follow the bit depth (8.7.3: bandShift = bitDepth - 5; offsets arrive
 * as final values, |offset| <= 31 at ten bits).
 */
struct Sao16Lane {
	const u32 *src; u32 *dst;          /* the plane, as words of two samples */
	unsigned pww;                      /* words per row (even) */
	int ph;
	unsigned xa, xb, xe, xr;           /* word columns: own first / third (clamped), the neighbouring block's, the right one */
	unsigned x;                        /* own first word, unclamped */
	int ya, yb, rpg;
	bool live_a, live_b, first_col, last_col, one_lane;
	u32 pool_lo, pool_hi, band_add, band_shift;
	u32 bias2, unbias2, max2;          /* 1 << depth, -((1 << depth) - 1), (1 << depth) - 1 in both halves */
};
struct RawRow16 { u32 w[4], e, e2; };
struct RowWin16 { u32 L, w[4], R; };

template <bool kSides, bool kQuad>
__device__ __forceinline__ RawRow16 LoadRow16(const Sao16Lane &t, int y)
{
	const unsigned row = (unsigned)min(max(y, 0), t.ph - 1) * t.pww;
	RawRow16 r;
	if (kQuad) {
		const uint4 v = __ldg(reinterpret_cast<const uint4 *>(t.src + (row + t.xa)));
		r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
	} else {
		const uint2 a = __ldg(reinterpret_cast<const uint2 *>(t.src + (row + t.xa))), c = __ldg(reinterpret_cast<const uint2 *>(t.src + (row + t.xb)));
		r.w[0] = a.x; r.w[1] = a.y; r.w[2] = c.x; r.w[3] = c.y;
	}
	r.e = r.e2 = 0u;
	if (kSides) {
		r.e = __ldg(t.src + (row + t.xe));
		if (t.one_lane) r.e2 = __ldg(t.src + (row + t.xr));
	}
	return r;
}

template <bool kSides>
__device__ __forceinline__ RowWin16 Neighbours16(const Sao16Lane &t, const RawRow16 &r)
{
	RowWin16 o;
#pragma unroll
	for (int j = 0; j < 4; j++) o.w[j] = r.w[j];
	o.L = o.R = 0u;
	if (kSides) {
		o.L = __shfl_up_sync(0xffffffffu, r.w[3], 1);
		o.R = __shfl_down_sync(0xffffffffu, r.w[0], 1);
		if (t.first_col) o.L = r.e;
		if (t.last_col) o.R = t.one_lane ? r.e2 : r.e;
	}
	return o;
}

template <bool kQuad>
__device__ __forceinline__ void StoreRow16(const Sao16Lane &t, u32 *orow, bool row_live, const u32 (&v)[4])
{
	if (kQuad) { if (row_live && t.live_a) *reinterpret_cast<uint4 *>(orow) = make_uint4(v[0], v[1], v[2], v[3]); }
	else {
		if (row_live && t.live_a) *reinterpret_cast<uint2 *>(orow) = make_uint2(v[0], v[1]);
		if (row_live && t.live_b) *reinterpret_cast<uint2 *>(orow + 2) = make_uint2(v[2], v[3]);
	}
}

template <int kMode, bool kQuad>
__device__ __forceinline__ void FlatRun16(const Sao16Lane &t)
{
	u32 *orow = t.dst + ((unsigned)t.ya * t.pww + t.xa);
#pragma unroll 1
	for (int k0 = 0; k0 < t.rpg; k0 += 2) {      /* four rows at a time cost the edge classes their registers */
		RawRow16 r[2];
#pragma unroll
		for (int j = 0; j < 2; j++) if (k0 + j < t.rpg) r[j] = LoadRow16<false, kQuad>(t, t.ya + k0 + j);
#pragma unroll
		for (int j = 0; j < 2; j++) {
			if (k0 + j < t.rpg) {
				u32 v[4];
#pragma unroll
				for (int i = 0; i < 4; i++) {
					v[i] = r[j].w[i];
					if (kMode == 1) {
						v[i] = BandOffset2(v[i], t.band_shift, t.band_add, t.pool_lo, t.pool_hi, t.max2);
					}
				}
				StoreRow16<kQuad>(t, orow, t.ya + k0 + j < t.yb, v);
			}
			orow += t.pww;
		}
	}
}

/* see EdgeRun: category(y, i) = 2 + S(y, i) - S(y + vy, i - hx), one S per sample */
template <int kClass, bool kQuad>
__device__ __forceinline__ void EdgeRun16(const Sao16Lane &t)
{
	constexpr bool kSides = kClass != 1, kTall = kClass != 0;
	const int n_iter = t.rpg + (kTall ? 1 : 0);
	u32 *orow = t.dst + ((long long)(t.ya - (kTall ? 1 : 0)) * (long long)t.pww + t.xa);
	auto step = [&](int i, RawRow16 &slot, const RowWin16 &above, const u32 (&s_above)[4], RowWin16 &cur, u32 (&s_cur)[4]) {
		cur = Neighbours16<kSides>(t, slot);
		if (i + 2 < n_iter) slot = LoadRow16<kSides, kQuad>(t, t.ya + i + 2);
		const RowWin16 &ar = kTall ? above : cur;
		u32 s[5], tt[4];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			u32 a = ar.w[j];
			if (kClass == 3) a = Prmt(ar.w[j], j == 3 ? ar.R : ar.w[j + 1], 0x5432);          /* one sample to the right */
			else if (kClass != 1) a = Prmt(j == 0 ? ar.L : ar.w[j - 1], ar.w[j], 0x5432);      /* one to the left */
			s[j] = Signs2(cur.w[j], a, t.bias2, t.unbias2);
		}
		if (kClass == 0 || kClass == 2) {
			const u32 mine = AddMinRelu2((cur.R & 0xffffu) + (t.bias2 & 0xffffu) - (ar.w[3] >> 16), t.unbias2, 0x00020002u);
			const u32 theirs = __shfl_down_sync(0xffffffffu, s[0], 1);
			s[4] = t.last_col ? mine : theirs;
#pragma unroll
			for (int j = 0; j < 4; j++) tt[j] = Prmt(s[j], s[j + 1], 0x5432);
		} else if (kClass == 3) {
			const u32 mine = AddMinRelu2((cur.L >> 16) + (t.bias2 & 0xffffu) - (ar.w[0] & 0xffffu), t.unbias2, 0x00020002u) << 16;
			const u32 theirs = __shfl_up_sync(0xffffffffu, s[3], 1);
			const u32 sl = t.first_col ? mine : theirs;
#pragma unroll
			for (int j = 0; j < 4; j++) tt[j] = Prmt(j == 0 ? sl : s[j - 1], s[j], 0x5432);
		} else {
#pragma unroll
			for (int j = 0; j < 4; j++) tt[j] = s[j];
		}
		const int yo = t.ya + i - (kTall ? 1 : 0);
		const bool keep_row = kTall && (yo == 0 || yo == t.ph - 1);
		u32 v[4];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const u32 r = ApplyOffsets2(ar.w[j], (kTall ? s_above[j] : s[j]) + 0x00020002u - tt[j], t.pool_lo, t.pool_hi, t.max2);
			/* the picture's first / last column (not for the vertical class), first / last row (not for the horizontal one) */
			u32 keep = 0u;
			if (kClass != 1) keep = (t.x + j == 0u ? 0x0000ffffu : 0u) | (t.x + j == t.pww - 1u ? 0xffff0000u : 0u);
			if (keep_row) keep = 0xffffffffu;
			v[j] = (ar.w[j] & keep) | (r & ~keep);
			s_cur[j] = s[j];
		}
		StoreRow16<kQuad>(t, orow, yo >= t.ya && yo < t.yb, v);
		orow += t.pww;
	};
	RowWin16 wa, wb;
	u32 sa[4] = {0u, 0u, 0u, 0u}, sb[4] = {0u, 0u, 0u, 0u};
	wb.L = wb.R = 0u; wb.w[0] = wb.w[1] = wb.w[2] = wb.w[3] = 0u;
	wa = wb;
	if (kTall) wb = Neighbours16<kSides>(t, LoadRow16<kSides, kQuad>(t, t.ya - 1));
	RawRow16 p = LoadRow16<kSides, kQuad>(t, t.ya), q = p;
	if (n_iter > 1) q = LoadRow16<kSides, kQuad>(t, t.ya + 1);
#pragma unroll 1
	for (int i = 0; i < n_iter; i += 2) {
		step(i, p, wb, sb, wa, sa);
		if (i + 1 < n_iter) step(i + 1, q, wa, sa, wb, sb);
	}
}

#ifndef HDB_SAO16_MIN_CTAS
#define HDB_SAO16_MIN_CTAS 6
#endif
template <bool kQuad>
__global__ void __launch_bounds__(kSaoWarps * 32, HDB_SAO16_MIN_CTAS)
Sao16Kernel(const __grid_constant__ SaoBatch b)
{
	const int lane = threadIdx.x & 31;
	const unsigned total = (unsigned)b.total_tasks;
	const unsigned n_chunks = (total + kSaoChunk - 1) / kSaoChunk;
	const unsigned warp_id = blockIdx.x * kSaoWarps + (threadIdx.x >> 5);
	const unsigned counter = warp_id % kSaoCounters;
	unsigned *ticket = b.ticket + 32u * counter;
	unsigned next = 0;
	if (lane == 0) next = atomicAdd(ticket, 1u);
	next = __shfl_sync(0xffffffffu, next, 0) * kSaoCounters + counter;
	while (next < n_chunks) {
		const unsigned first = next * kSaoChunk, last = min(first + (unsigned)kSaoChunk, total);
		if (lane == 0) next = atomicAdd(ticket, 1u);
		for (unsigned task = first; task < last; task++) {
			const uint32_t frame = b.div_tasks_per_frame.Div(task);
			uint32_t r = task - frame * (uint32_t)b.tasks_per_frame;
			const int plane = (int)b.div_tasks_per_plane.Div(r);
			r -= (uint32_t)plane * (uint32_t)b.tasks_per_plane;
			const int band = (int)b.div_band.Div(r), in_band = (int)r - band * (kSaoBand * b.ctb_cols);
			const int crow = band * kSaoBand + in_band % kSaoBand, ccol = in_band / kSaoBand;
			if (crow >= b.ctb_rows) continue;
			const int pw = plane ? b.width >> 1 : b.width, ph = plane ? b.height >> 1 : b.height;
			const int ctb = plane ? b.ctb >> 1 : b.ctb;
			const long long plane_off = plane == 0 ? 0 : (long long)b.width * b.height + (plane == 2 ? (long long)(b.width >> 1) * (b.height >> 1) : 0);
			Sao16Lane t;
			t.src = reinterpret_cast<const u32 *>(b.in + (long long)frame * b.frame_stride + 2 * plane_off);
			t.dst = reinterpret_cast<u32 *>(b.out + (long long)frame * b.frame_stride + 2 * plane_off);
			t.ph = ph;
			const int pww = pw >> 1;
			t.pww = (unsigned)pww;
			const int lpr = ctb >> 3, groups = 32 / lpr;
			const int col = lane & (lpr - 1), group = lane / lpr;
			t.rpg = (ctb + groups - 1) / groups;
			const int x = ccol * (ctb >> 1) + 4 * col;                  /* this lane's first word */
			t.x = (unsigned)x;
			t.live_a = x < pww; t.live_b = x + 2 < pww;
			t.ya = crow * ctb + group * t.rpg;
			t.yb = min(min(t.ya + t.rpg, (crow + 1) * ctb), ph);
			t.first_col = col == 0; t.last_col = col == lpr - 1; t.one_lane = lpr == 1;
			t.xa = (unsigned)min(x, pww - (kQuad ? 4 : 2)); t.xb = (unsigned)min(x + 2, pww - 2); t.xr = (unsigned)min(x + 4, pww - 1);
			t.xe = t.first_col ? (unsigned)min(max(x - 1, 0), pww - 1) : t.xr;
			const unsigned char *p = b.params + 6ll * ((((long long)frame * 3 + plane) * b.ctb_rows + crow) * b.ctb_cols + ccol);
			const u32 type = __ldg(p);
			u32 pos = __ldg(p + 1);
			const u32 o0 = __ldg(p + 2), o1 = __ldg(p + 3), o2 = __ldg(p + 4), o3 = __ldg(p + 5);
			OffsetPool(type, o0, o1, o2, o3, t.pool_lo, t.pool_hi);
			if (type == 2u) pos &= 3u;
			t.band_add = (32u - pos) * 0x10001u;
			t.band_shift = (u32)b.bit_depth - 5u;
			const u32 one = 1u << b.bit_depth;
			t.bias2 = one * 0x10001u; t.max2 = (one - 1u) * 0x10001u; t.unbias2 = ((0x10000u - (one - 1u)) & 0xffffu) * 0x10001u;
			if (type == 0u) FlatRun16<0, kQuad>(t);
			else if (type == 1u) FlatRun16<1, kQuad>(t);
			else if (pos == 0u) EdgeRun16<0, kQuad>(t);
			else if (pos == 1u) EdgeRun16<1, kQuad>(t);
			else if (pos == 2u) EdgeRun16<2, kQuad>(t);
			else EdgeRun16<3, kQuad>(t);
		}
		next = __shfl_sync(0xffffffffu, next, 0) * kSaoCounters + counter;
	}
}

cudaError_t LaunchSao(const unsigned char *d_in, unsigned char *d_out, int n_frames, int width, int height,
	const void *d_params, int ctb, int bit_depth, int sample_bytes, unsigned *d_tickets, int sm_count, cudaStream_t stream)
{
	if (n_frames <= 0) return cudaSuccess;
	/* CTAs an SM holds of each of the four kernels */
	static const struct Resident {
		int n[4];
		Resident()
		{
			const void *k[4] = { (const void *)SaoKernel<false>, (const void *)SaoKernel<true>, (const void *)Sao16Kernel<false>, (const void *)Sao16Kernel<true> };
			for (int i = 0; i < 4; i++)
				if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n[i], k[i], kSaoWarps * 32, 0) != cudaSuccess || n[i] <= 0) n[i] = 6;
		}
	} resident;
	SaoBatch b{};
	b.in = d_in; b.out = d_out;
	b.frame_stride = (long long)width * height * 3 / 2 * sample_bytes;
	b.width = width; b.height = height; b.ctb = ctb; b.bit_depth = bit_depth;
	b.ctb_cols = (width + ctb - 1) / ctb; b.ctb_rows = (height + ctb - 1) / ctb;
	b.params = static_cast<const unsigned char *>(d_params);
	/* one task per coding tree block and plane (the chroma planes have the same block grid as luma), the block rows
	 * rounded up to whole bands */
	b.tasks_per_plane = (b.ctb_rows + kSaoBand - 1) / kSaoBand * kSaoBand * b.ctb_cols;
	b.tasks_per_frame = 3 * b.tasks_per_plane;
	b.div_tasks_per_frame.Init((uint32_t)b.tasks_per_frame);
	b.div_tasks_per_plane.Init((uint32_t)b.tasks_per_plane);
	b.div_band.Init((uint32_t)(kSaoBand * b.ctb_cols));
	/* frames per launch bounded by the 31-bit task index */
	const long long max_frames = std::max<long long>(1, ((1ll << 31) - 1) / b.tasks_per_frame);
	const long long parts = (n_frames + max_frames - 1) / max_frames;
	constexpr int kWordsPerPart = 32 * kSaoCounters;
	static_assert(kWordsPerPart <= kSaoTicketWords, "ticket buffer too small");
	if (parts > kSaoTicketWords / kWordsPerPart) return cudaErrorInvalidValue;
	cudaError_t e = cudaMemsetAsync(d_tickets, 0, sizeof(unsigned) * kWordsPerPart * (size_t)parts, stream);
	if (e != cudaSuccess) return e;
	for (long long f0 = 0; f0 < n_frames; f0 += max_frames) {
		SaoBatch part = b;
		part.n_frames = (int)std::min<long long>(max_frames, n_frames - f0);
		part.in = d_in + f0 * b.frame_stride; part.out = d_out + f0 * b.frame_stride;
		part.params = b.params + 6ll * 3 * b.ctb_rows * b.ctb_cols * f0;
		part.total_tasks = (long long)part.n_frames * b.tasks_per_frame;
		part.ticket = d_tickets + f0 / max_frames * kWordsPerPart;
		/* as many warps as the device holds at once, every one of them drawing tickets until none are left */
		const long long per_cta = (long long)kSaoWarps * kSaoChunk;
		const bool pair = width % 16 == 0 && ((uintptr_t)part.in | (uintptr_t)part.out) % 8 == 0;
		const bool quad = width % 16 == 0 && ((uintptr_t)part.in | (uintptr_t)part.out) % 16 == 0;
		const int per_sm = resident.n[sample_bytes == 2 ? 2 + quad : pair];
		const long long ctas = std::min<long long>((part.total_tasks + per_cta - 1) / per_cta, (long long)sm_count * per_sm);
		/* both planes' rows an even number of words and the frames 8-byte aligned: the lanes move their two words at once */
		if (sample_bytes == 2) {
			if (quad) Sao16Kernel<true><<<(unsigned)ctas, kSaoWarps * 32, 0, stream>>>(part);
			else Sao16Kernel<false><<<(unsigned)ctas, kSaoWarps * 32, 0, stream>>>(part);
		} else if (pair) SaoKernel<true><<<(unsigned)ctas, kSaoWarps * 32, 0, stream>>>(part);
		else SaoKernel<false><<<(unsigned)ctas, kSaoWarps * 32, 0, stream>>>(part);
	}
	return cudaGetLastError();
}

} // namespace hdb
