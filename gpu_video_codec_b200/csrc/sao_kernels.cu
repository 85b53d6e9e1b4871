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
 * One warp walks one column strip of 256 bytes (lane l: the eight samples of words 64 * strip + 2 l and + 1) down the
 * rows of one coding-tree-block row of one plane, so the parameters a lane needs are constant for the walk.  A lane
 * keeps the words of rows y - 1, y, y + 1 and their left / right neighbours (from the adjacent lanes by shuffle;
 * the strip's first and last lane load theirs) in registers.  Neighbouring CTBs of a strip differ in type and
 * class, so the row loop is branch-free: which rows and which side the two neighbours come from is a per-lane
 * selection followed by one funnel shift, and the edge category and the band index are both computed and one is
 * selected.  Four samples at a time as two 16x2 pairs, with the packed-halfword instructions the deblocking core
 * uses:
 *   edge offset  256 + sign(c - a) per half = max(min(256 + c - a, 257), 255); category 2 + sign + sign in 0..4
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

struct SaoBatch {
	const unsigned char *in; unsigned char *out;    /* tight YV12 frames, frame_stride apart */
	long long frame_stride;
	int n_frames, width, height;
	int ctb;                                         /* luma CTB size: 16, 32 or 64 */
	int ctb_cols, ctb_rows;
	const unsigned char *params;                     /* hdb_sao_params[n][3][ctb_rows][ctb_cols], 6 bytes each */
	/* task space: frame-major, then plane, CTB row, strip */
	int strips_luma, strips_chroma, tasks_luma, tasks_chroma, tasks_per_frame;
	long long total_tasks;
	FastDiv div_tasks_per_frame, div_strips_luma, div_strips_chroma;
};

/* bytes (b0, b1) / (b2, b3) of a word as a 16x2 pair */
__device__ __forceinline__ u32 PairA(u32 w) { return Prmt(w, 0u, 0x4140); }
__device__ __forceinline__ u32 PairB(u32 w) { return Prmt(w, 0u, 0x4342); }

/* 256 + sign(c - n) per half */
__device__ __forceinline__ u32 SignBiased(u32 c, u32 n) { return Max2(Min2(c + 0x01000100u - n, 0x01010101u), 0x00ff00ffu); }

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

/* the word of a row's (L, w0, w1, R) window that lies `shift` samples to the left (shift = -1), at (0) or to the
 * right (+1) of w0 / w1, as one funnel shift of a selected register pair */
struct RowWin { u32 L, w0, w1, R; };

__device__ __forceinline__ void Shifted(const RowWin &r, bool left, u32 amount, u32 &o0, u32 &o1)
{
	/* left: (w0 << 8 | L >> 24, w1 << 8 | w0 >> 24) = funnel-right by 24 of (L, w0), (w0, w1);
	 * at / right: funnel-right by 0 / 8 of (w0, w1), (w1, R) */
	const u32 a0 = left ? r.L : r.w0, a1 = left ? r.w0 : r.w1, a2 = left ? r.w1 : r.R;
	o0 = __funnelshift_r(a0, a1, amount);
	o1 = __funnelshift_r(a1, a2, amount);
}

__global__ void __launch_bounds__(kSaoWarps * 32)
SaoKernel(const __grid_constant__ SaoBatch b)
{
	const int lane = threadIdx.x & 31;
	const long long warp0 = (long long)blockIdx.x * kSaoWarps + (threadIdx.x >> 5);
	const long long n_warps = (long long)gridDim.x * kSaoWarps;
	for (long long task = warp0; task < b.total_tasks; task += n_warps) {
		const uint32_t frame = b.div_tasks_per_frame.Div((uint32_t)task);
		uint32_t r = (uint32_t)task - frame * (uint32_t)b.tasks_per_frame;
		int plane = 0;
		if (r >= (uint32_t)b.tasks_luma) { r -= (uint32_t)b.tasks_luma; plane = 1; if (r >= (uint32_t)b.tasks_chroma) { r -= (uint32_t)b.tasks_chroma; plane = 2; } }
		const int strips = plane ? b.strips_chroma : b.strips_luma;
		const uint32_t crow = plane ? b.div_strips_chroma.Div(r) : b.div_strips_luma.Div(r);
		const int strip = (int)(r - crow * (uint32_t)strips);
		const int pw = plane ? b.width >> 1 : b.width, ph = plane ? b.height >> 1 : b.height;
		const int ctb = plane ? b.ctb >> 1 : b.ctb;
		const long long plane_off = plane == 0 ? 0 : (long long)b.width * b.height + (plane == 2 ? (long long)(b.width >> 1) * (b.height >> 1) : 0);
		const u32 *src = reinterpret_cast<const u32 *>(b.in + (long long)frame * b.frame_stride + plane_off);
		u32 *dst = reinterpret_cast<u32 *>(b.out + (long long)frame * b.frame_stride + plane_off);
		const int pww = pw >> 2;                                 /* words per row */
		const int x = 2 * (32 * strip + lane);                   /* this lane's first word: eight samples, words x and x + 1 */
		const bool live0 = x < pww, live1 = x + 1 < pww;
		const int y0 = (int)crow * ctb, y1 = min(y0 + ctb, ph);
		/* parameters of the CTB this lane's samples lie in (eight samples never straddle CTBs: their width is a
		 * multiple of 8 and so is the lane's first sample) */
		u32 type = 0, pos = 0, pool_lo = 0, pool_hi = 0;
		if (live0) {
			const unsigned char *p = b.params + 6ll * ((((long long)frame * 3 + plane) * b.ctb_rows + crow) * b.ctb_cols + (4 * x) / ctb);
			type = __ldg(p); pos = __ldg(p + 1);
			const u32 o0 = __ldg(p + 2), o1 = __ldg(p + 3), o2 = __ldg(p + 4), o3 = __ldg(p + 5);
			if (type == 2) { pool_lo = o0 | (o1 << 8) | (o2 << 24); pool_hi = o3; pos &= 3u; }     /* categories 0, 1, (2: none), 3, 4 */
			else { pool_lo = o0 | (o1 << 8) | (o2 << 16) | (o3 << 24); pool_hi = 0u; }            /* bands k = 0..3, (4: none) */
		}
		const bool edge = type == 2u;
		/* which rows and which side the two neighbours of the class come from (8.7.3.2: hPos / vPos):
		 * class 0: left / right of row y; 1: rows y - 1 / y + 1; 2: left of y - 1 / right of y + 1; 3: right of y - 1 / left of y + 1 */
		const bool a_same_row = pos == 0u, a_left = pos == 0u || pos == 2u, b_left = pos == 3u;
		const u32 a_amount = pos == 1u ? 0u : (a_left ? 24u : 8u), b_amount = pos == 1u ? 0u : (b_left ? 24u : 8u);
		/* bytes that keep their value whatever the row: nothing to do, or the picture's first / last column for every
		 * class but the vertical one */
		u32 keep0 = 0u, keep1 = 0u;
		if (type == 0u) keep0 = keep1 = 0xffffffffu;
		else if (edge && pos != 1u) { keep0 = x == 0 ? 0x000000ffu : 0u; keep0 |= x == pww - 1 ? 0xff000000u : 0u; keep1 = x + 1 == pww - 1 ? 0xff000000u : 0u; }
		const u32 band_add = (32u - pos) * 0x10001u;

		/* (L, w0, w1, R) of a row for this lane.  Branch-free: rows and words outside the picture are read from the
		 * nearest ones inside instead -- a sample whose neighbour lies outside is kept whatever the comparison says. */
		const int xc0 = min(x, pww - 1), xc1 = min(x + 1, pww - 1), xl = min(max(x - 1, 0), pww - 1), xr = min(x + 2, pww - 1);
		auto load_row = [&](int y) -> RowWin {
			const u32 *p = src + (long long)min(max(y, 0), ph - 1) * pww;
			RowWin w;
			w.w0 = __ldg(p + xc0); w.w1 = __ldg(p + xc1);
			const u32 el = __ldg(p + xl), er = __ldg(p + xr);
			w.L = __shfl_up_sync(0xffffffffu, w.w1, 1);
			w.R = __shfl_down_sync(0xffffffffu, w.w0, 1);
			if (lane == 0) w.L = el;
			if (lane == 31) w.R = er;
			return w;
		};
		auto offset4 = [&](u32 c, u32 a, u32 n) -> u32 {          /* four samples of word c against neighbour words a, n */
			const u32 cA = PairA(c), cB = PairB(c);
			const u32 eA = SignBiased(cA, PairA(a)) + SignBiased(cA, PairA(n)) - 0x01fe01feu;
			const u32 eB = SignBiased(cB, PairB(a)) + SignBiased(cB, PairB(n)) - 0x01fe01feu;
			const u32 kA = MinU2((((cA >> 3) & 0x001f001fu) + band_add) & 0x001f001fu, 0x00040004u);
			const u32 kB = MinU2((((cB >> 3) & 0x001f001fu) + band_add) & 0x001f001fu, 0x00040004u);
			const u32 rA = AddMinRelu2(cA, PickOffsets(edge ? eA : kA, pool_lo, pool_hi), 0x00ff00ffu);
			const u32 rB = AddMinRelu2(cB, PickOffsets(edge ? eB : kB, pool_lo, pool_hi), 0x00ff00ffu);
			return Prmt(rA, rB, 0x6420);
		};
		/* the row after next is fetched while this one is worked on: with the loads of row y + 1 issued in the
		 * iteration that needs them the walk ran at the latency of one load per row (0.25 of the HBM peak) */
		RowWin r0 = load_row(y0 - 1), r1 = load_row(y0), r2 = load_row(y0 + 1);
		u32 *orow = dst + (long long)y0 * pww + x;
#pragma unroll 1
		for (int y = y0; y < y1; y++) {
			const RowWin r3 = load_row(y + 2);
			const RowWin ra = a_same_row ? r1 : r0, rb = a_same_row ? r1 : r2;
			u32 a0, a1, n0, n1;
			Shifted(ra, a_left, a_amount, a0, a1);
			Shifted(rb, b_left, b_amount, n0, n1);
			const u32 res0 = offset4(r1.w0, a0, n0), res1 = offset4(r1.w1, a1, n1);
			/* the picture's first / last row for every class but the horizontal one */
			const bool keep_row = edge && pos != 0u && (y == 0 || y == ph - 1);
			const u32 k0 = keep_row ? 0xffffffffu : keep0, k1 = keep_row ? 0xffffffffu : keep1;
			if (live0) orow[0] = (r1.w0 & k0) | (res0 & ~k0);
			if (live1) orow[1] = (r1.w1 & k1) | (res1 & ~k1);
			r0 = r1; r1 = r2; r2 = r3;
			orow += pww;
		}
	}
}

cudaError_t LaunchSao(const unsigned char *d_in, unsigned char *d_out, int n_frames, int width, int height,
	const void *d_params, int ctb, int sm_count, cudaStream_t stream)
{
	if (n_frames <= 0) return cudaSuccess;
	SaoBatch b{};
	b.in = d_in; b.out = d_out;
	b.frame_stride = (long long)width * height * 3 / 2;
	b.width = width; b.height = height; b.ctb = ctb;
	b.ctb_cols = (width + ctb - 1) / ctb; b.ctb_rows = (height + ctb - 1) / ctb;
	b.params = static_cast<const unsigned char *>(d_params);
	b.strips_luma = (width / 4 + 63) / 64; b.strips_chroma = (width / 8 + 63) / 64;     /* 64 words = 256 bytes a strip */
	/* the chroma planes have as many CTB rows as luma (a chroma CTB is half as high as the plane's share of rows) */
	b.tasks_luma = b.strips_luma * b.ctb_rows; b.tasks_chroma = b.strips_chroma * b.ctb_rows;
	b.tasks_per_frame = b.tasks_luma + 2 * b.tasks_chroma;
	b.div_tasks_per_frame.Init((uint32_t)b.tasks_per_frame);
	b.div_strips_luma.Init((uint32_t)b.strips_luma); b.div_strips_chroma.Init((uint32_t)b.strips_chroma);
	/* frames per launch bounded by the 31-bit task index */
	const long long max_frames = std::max<long long>(1, ((1ll << 31) - 1) / b.tasks_per_frame);
	for (long long f0 = 0; f0 < n_frames; f0 += max_frames) {
		SaoBatch part = b;
		part.n_frames = (int)std::min<long long>(max_frames, n_frames - f0);
		part.in = d_in + f0 * b.frame_stride; part.out = d_out + f0 * b.frame_stride;
		part.params = b.params + 6ll * 3 * b.ctb_rows * b.ctb_cols * f0;
		part.total_tasks = (long long)part.n_frames * b.tasks_per_frame;
		const long long ctas = std::min<long long>((part.total_tasks + kSaoWarps - 1) / kSaoWarps, (long long)sm_count * 16);
		SaoKernel<<<(unsigned)ctas, kSaoWarps * 32, 0, stream>>>(part);
	}
	return cudaGetLastError();
}

} // namespace hdb
