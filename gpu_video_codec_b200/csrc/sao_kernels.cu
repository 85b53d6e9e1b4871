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
#include "deblock_core.cuh"

#include <algorithm>

namespace hdb {

constexpr int kSaoWarps = 4;
constexpr int kSaoChunk = 4;      /* consecutive blocks per ticket */

struct SaoBatch {
	const unsigned char *in; unsigned char *out;    /* tight YV12 frames, frame_stride apart */
	long long frame_stride;
	int n_frames, width, height;
	int ctb;                                         /* luma CTB size: 16, 32 or 64 */
	int ctb_cols, ctb_rows;
	const unsigned char *params;                     /* hdb_sao_params[n][3][ctb_rows][ctb_cols], 6 bytes each */
	/* task space: frame-major, then plane, CTB row, CTB column */
	int tasks_per_plane, tasks_per_frame;
	long long total_tasks;
	FastDiv div_tasks_per_frame, div_tasks_per_plane, div_ctb_cols;
	unsigned *ticket;                                /* zero at launch: the next block to hand out */
};

/* bytes (b0, b1) / (b2, b3) of a word as a 16x2 pair */
__device__ __forceinline__ u32 PairA(u32 w) { return Prmt(w, 0u, 0x4140); }
__device__ __forceinline__ u32 PairB(u32 w) { return Prmt(w, 0u, 0x4342); }

/* 1 + sign(c - n) per half; c256 = c + 256 per half */
__device__ __forceinline__ u32 SignPlus1(u32 c256, u32 n) { return AddMinRelu2(c256 - n, 0xff01ff01u, 0x00020002u); }

/* offsets picked by the index halves (0..4) out of the pool, as signed halves */
__device__ __forceinline__ u32 PickOffsets(u32 idx, u32 pool_lo, u32 pool_hi)
{
	const u32 t = idx * 0x11u;                                   /* nibble | nibble << 4 in bytes 0 and 2 */
	const u32 sel = Prmt(t, 0u, 0x4420) | 0x8080u;               /* nibbles: i0, i0 | 8, i1, i1 | 8 */
	/* prmt.b32 itself: bit 3 of a selector nibble replicates the sign of the selected byte (__byte_perm documents
	 * the low three bits of each nibble only) */
	u32 d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(pool_lo), "r"(pool_hi), "r"(sel));
	return d;
}

/* four samples of word c against the neighbour words a, n: edge offset */
__device__ __forceinline__ u32 EdgeOffset4(u32 c, u32 a, u32 n, u32 pool_lo, u32 pool_hi)
{
	const u32 cA = PairA(c), cB = PairB(c), cA256 = cA + 0x01000100u, cB256 = cB + 0x01000100u;
	const u32 eA = SignPlus1(cA256, PairA(a)) + SignPlus1(cA256, PairA(n));
	const u32 eB = SignPlus1(cB256, PairB(a)) + SignPlus1(cB256, PairB(n));
	const u32 rA = AddMinRelu2(cA, PickOffsets(eA, pool_lo, pool_hi), 0x00ff00ffu);
	const u32 rB = AddMinRelu2(cB, PickOffsets(eB, pool_lo, pool_hi), 0x00ff00ffu);
	return Prmt(rA, rB, 0x6420);
}

/* four samples of word c: band offset */
__device__ __forceinline__ u32 BandOffset4(u32 c, u32 band_add, u32 pool_lo, u32 pool_hi)
{
	const u32 cA = PairA(c), cB = PairB(c);
	const u32 kA = MinU2((((cA >> 3) & 0x001f001fu) + band_add) & 0x001f001fu, 0x00040004u);
	const u32 kB = MinU2((((cB >> 3) & 0x001f001fu) + band_add) & 0x001f001fu, 0x00040004u);
	const u32 rA = AddMinRelu2(cA, PickOffsets(kA, pool_lo, pool_hi), 0x00ff00ffu);
	const u32 rB = AddMinRelu2(cB, PickOffsets(kB, pool_lo, pool_hi), 0x00ff00ffu);
	return Prmt(rA, rB, 0x6420);
}

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
template <bool kSides>
__device__ __forceinline__ RawRow LoadRow(const SaoLane &t, int y)
{
	const unsigned row = (unsigned)min(max(y, 0), t.ph - 1) * t.pww;
	RawRow w{0u, 0u, 0u, 0u};
	w.w0 = __ldg(t.src + (row + t.xc0)); w.w1 = __ldg(t.src + (row + t.xc1));
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

__device__ __forceinline__ void StoreRow(const SaoLane &t, u32 *orow, bool row_live, u32 v0, u32 v1)
{
	if (row_live && t.live0) orow[0] = v0;
	if (row_live && t.live1) orow[1] = v1;
}

/* kMode 0: the block is off -- moved; 1: band offset; 2: edge offset, horizontal class.  No row needs another. */
template <int kMode>
__device__ __forceinline__ void FlatRun(const SaoLane &t)
{
	u32 *orow = t.dst + ((unsigned)t.ya * t.pww + 2u * (t.xc0 >> 1));
	for (int k0 = 0; k0 < t.rpg; k0 += 4) {
		RawRow r[4];
#pragma unroll
		for (int j = 0; j < 4; j++) r[j] = k0 + j < t.rpg ? LoadRow<kMode == 2>(t, t.ya + k0 + j) : RawRow{0u, 0u, 0u, 0u};
#pragma unroll
		for (int j = 0; j < 4; j++) {
			if (k0 + j < t.rpg) {       /* the same for all lanes */
				u32 v0 = r[j].w0, v1 = r[j].w1;
				if (kMode == 1) { v0 = BandOffset4(v0, t.band_add, t.pool_lo, t.pool_hi); v1 = BandOffset4(v1, t.band_add, t.pool_lo, t.pool_hi); }
				if (kMode == 2) {
					const RowWin w = Neighbours<true>(t, r[j]);
					const u32 a0 = __funnelshift_r(w.L, w.w0, 24), a1 = __funnelshift_r(w.w0, w.w1, 24);
					const u32 n0 = __funnelshift_r(w.w0, w.w1, 8), n1 = __funnelshift_r(w.w1, w.R, 8);
					const u32 e0 = EdgeOffset4(w.w0, a0, n0, t.pool_lo, t.pool_hi), e1 = EdgeOffset4(w.w1, a1, n1, t.pool_lo, t.pool_hi);
					v0 = (v0 & t.keep0) | (e0 & ~t.keep0); v1 = (v1 & t.keep1) | (e1 & ~t.keep1);
				}
				StoreRow(t, orow, t.ya + k0 + j < t.yb, v0, v1);
			}
			orow += t.pww;
		}
	}
}

/* edge offset, the classes that look up and down (8.7.3.2, hPos / vPos): 1 rows y - 1 / y + 1; 2 left of row
 * y - 1 / right of row y + 1; 3 right of row y - 1 / left of row y + 1.  Row k of the run is picture row
 * ya - 1 + k; rows 0 .. rpg + 1 are needed.  Six rows in registers give four rows of output; the four rows the
 * next round adds are loaded before this round's work and handed round (the shuffles are the first use of a
 * load) after it. */
template <int kClass>
__device__ __forceinline__ void TallRun(const SaoLane &t)
{
	constexpr bool kSides = kClass != 1;
	const RawRow none{0u, 0u, 0u, 0u};
	RowWin w[6];
	{
		RawRow r[6];
#pragma unroll
		for (int j = 0; j < 6; j++) r[j] = j <= t.rpg + 1 ? LoadRow<kSides>(t, t.ya - 1 + j) : none;
#pragma unroll
		for (int j = 0; j < 6; j++) w[j] = Neighbours<kSides>(t, r[j]);
	}
	u32 *orow = t.dst + ((unsigned)t.ya * t.pww + 2u * (t.xc0 >> 1));
#pragma unroll 1
	for (int k0 = 0; k0 < t.rpg; k0 += 4) {
		const bool more = k0 + 4 < t.rpg;
		RawRow nx[4] = {none, none, none, none};
		if (more) {
#pragma unroll
			for (int j = 0; j < 4; j++) nx[j] = LoadRow<kSides>(t, t.ya + k0 + 5 + j);
		}
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int y = t.ya + k0 + j;
			const RowWin &up = w[j], &c = w[j + 1], &dn = w[j + 2];
			u32 a0, a1, n0, n1;
			if (kClass == 1) { a0 = up.w0; a1 = up.w1; n0 = dn.w0; n1 = dn.w1; }
			else if (kClass == 2) {
				a0 = __funnelshift_r(up.L, up.w0, 24); a1 = __funnelshift_r(up.w0, up.w1, 24);
				n0 = __funnelshift_r(dn.w0, dn.w1, 8); n1 = __funnelshift_r(dn.w1, dn.R, 8);
			} else {
				a0 = __funnelshift_r(up.w0, up.w1, 8); a1 = __funnelshift_r(up.w1, up.R, 8);
				n0 = __funnelshift_r(dn.L, dn.w0, 24); n1 = __funnelshift_r(dn.w0, dn.w1, 24);
			}
			const u32 e0 = EdgeOffset4(c.w0, a0, n0, t.pool_lo, t.pool_hi), e1 = EdgeOffset4(c.w1, a1, n1, t.pool_lo, t.pool_hi);
			/* the picture's first and last row are left alone */
			const bool keep_row = y == 0 || y == t.ph - 1;
			const u32 k0m = keep_row ? 0xffffffffu : t.keep0, k1m = keep_row ? 0xffffffffu : t.keep1;
			StoreRow(t, orow, k0 + j < t.rpg && y < t.yb, (c.w0 & k0m) | (e0 & ~k0m), (c.w1 & k1m) | (e1 & ~k1m));
			orow += t.pww;
		}
		if (more) {
			w[0] = w[4]; w[1] = w[5];
#pragma unroll
			for (int j = 0; j < 4; j++) w[2 + j] = Neighbours<kSides>(t, nx[j]);
		}
	}
}

__global__ void __launch_bounds__(kSaoWarps * 32)
SaoKernel(const __grid_constant__ SaoBatch b)
{
	const int lane = threadIdx.x & 31;
	const unsigned total = (unsigned)b.total_tasks;
	unsigned next = 0;
	if (lane == 0) next = atomicAdd(b.ticket, (unsigned)kSaoChunk);
	next = __shfl_sync(0xffffffffu, next, 0);
	while (next < total) {
		const unsigned first = next, last = min(first + (unsigned)kSaoChunk, total);
		if (lane == 0) next = atomicAdd(b.ticket, (unsigned)kSaoChunk);      /* in flight while this chunk is worked on */
		for (unsigned task = first; task < last; task++) {
			const uint32_t frame = b.div_tasks_per_frame.Div(task);
			uint32_t r = task - frame * (uint32_t)b.tasks_per_frame;
			const int plane = (int)b.div_tasks_per_plane.Div(r);
			r -= (uint32_t)plane * (uint32_t)b.tasks_per_plane;
			const int crow = (int)b.div_ctb_cols.Div(r), ccol = (int)r - crow * b.ctb_cols;
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
			t.xc0 = (unsigned)min(x, pww - 1); t.xc1 = (unsigned)min(x + 1, pww - 1); t.xr = (unsigned)min(x + 2, pww - 1);
			t.xe = t.first_col ? (unsigned)min(max(x - 1, 0), pww - 1) : t.xr;
			/* the block's parameters: the same for every lane */
			const unsigned char *p = b.params + 6ll * ((((long long)frame * 3 + plane) * b.ctb_rows + crow) * b.ctb_cols + ccol);
			const u32 type = __ldg(p);
			u32 pos = __ldg(p + 1);
			const u32 o0 = __ldg(p + 2), o1 = __ldg(p + 3), o2 = __ldg(p + 4), o3 = __ldg(p + 5);
			if (type == 2u) { t.pool_lo = o0 | (o1 << 8) | (o2 << 24); t.pool_hi = o3; pos &= 3u; }       /* categories 0, 1, (2: none), 3, 4 */
			else { t.pool_lo = o0 | (o1 << 8) | (o2 << 16) | (o3 << 24); t.pool_hi = 0u; }              /* bands k = 0..3, (4: none) */
			t.band_add = (32u - pos) * 0x10001u;
			/* bytes that keep their value whatever the row: the picture's first / last column (the vertical class
			 * does not look at them) */
			t.keep0 = (x == 0 ? 0x000000ffu : 0u) | (x == pww - 1 ? 0xff000000u : 0u);
			t.keep1 = x + 1 == pww - 1 ? 0xff000000u : 0u;
			if (type == 2u && pos == 1u) t.keep0 = t.keep1 = 0u;
			/* a dead lane's first word lies outside the row: its stores are off, its output pointer is formed from the
			 * clamped column and never used */
			if (type == 0u) FlatRun<0>(t);
			else if (type == 1u) FlatRun<1>(t);
			else if (pos == 0u) FlatRun<2>(t);
			else if (pos == 1u) TallRun<1>(t);
			else if (pos == 2u) TallRun<2>(t);
			else TallRun<3>(t);
		}
		next = __shfl_sync(0xffffffffu, next, 0);
	}
}

cudaError_t LaunchSao(const unsigned char *d_in, unsigned char *d_out, int n_frames, int width, int height,
	const void *d_params, int ctb, unsigned *d_tickets, int sm_count, cudaStream_t stream)
{
	if (n_frames <= 0) return cudaSuccess;
	static const int per_sm = [] {
		int n = 0;
		return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, SaoKernel, kSaoWarps * 32, 0) == cudaSuccess && n > 0 ? n : 8;
	}();
	SaoBatch b{};
	b.in = d_in; b.out = d_out;
	b.frame_stride = (long long)width * height * 3 / 2;
	b.width = width; b.height = height; b.ctb = ctb;
	b.ctb_cols = (width + ctb - 1) / ctb; b.ctb_rows = (height + ctb - 1) / ctb;
	b.params = static_cast<const unsigned char *>(d_params);
	/* one task per coding tree block and plane; the chroma planes have the same block grid as luma */
	b.tasks_per_plane = b.ctb_rows * b.ctb_cols;
	b.tasks_per_frame = 3 * b.tasks_per_plane;
	b.div_tasks_per_frame.Init((uint32_t)b.tasks_per_frame);
	b.div_tasks_per_plane.Init((uint32_t)b.tasks_per_plane);
	b.div_ctb_cols.Init((uint32_t)b.ctb_cols);
	/* frames per launch bounded by the 31-bit task index */
	const long long max_frames = std::max<long long>(1, ((1ll << 31) - 1) / b.tasks_per_frame);
	const long long parts = (n_frames + max_frames - 1) / max_frames;
	if (parts > kSaoTicketWords) return cudaErrorInvalidValue;
	cudaError_t e = cudaMemsetAsync(d_tickets, 0, sizeof(unsigned) * (size_t)parts, stream);
	if (e != cudaSuccess) return e;
	for (long long f0 = 0; f0 < n_frames; f0 += max_frames) {
		SaoBatch part = b;
		part.n_frames = (int)std::min<long long>(max_frames, n_frames - f0);
		part.in = d_in + f0 * b.frame_stride; part.out = d_out + f0 * b.frame_stride;
		part.params = b.params + 6ll * 3 * b.ctb_rows * b.ctb_cols * f0;
		part.total_tasks = (long long)part.n_frames * b.tasks_per_frame;
		part.ticket = d_tickets + f0 / max_frames;
		/* as many warps as the device holds at once, every one of them drawing tickets until none are left */
		const long long per_cta = (long long)kSaoWarps * kSaoChunk;
		const long long ctas = std::min<long long>((part.total_tasks + per_cta - 1) / per_cta, (long long)sm_count * per_sm);
		SaoKernel<<<(unsigned)ctas, kSaoWarps * 32, 0, stream>>>(part);
	}
	return cudaGetLastError();
}

} // namespace hdb
