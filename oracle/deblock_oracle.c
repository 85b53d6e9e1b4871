/*
 * deblock_oracle.c -- TEST INFRASTRUCTURE ONLY (see deblock_oracle.h).
 *
 * Plain-C restatement of the reference CPU deblocking filter.  Every function
 * cites the lines of hevc_deblocking_filter/hevc_deblocking_filter_cpu.h
 * ("cpu.h") it follows.  Written from the behaviour of that file, organised
 * differently: the reference spells out 32 pointers per edge segment, this
 * file describes a segment by (p base, q base, tap step, line step).
 */
#include "deblock_oracle.h"

#include <stdlib.h>
#include <string.h>

#define S 8 /* cpu.h:1035 sample_block_size */

/* cpu.h:1021-1026 */
static const unsigned char k_beta[52] = {
	0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
	6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24,
	26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56,
	58, 60, 62, 64
};

/* cpu.h:1028-1033 */
static const unsigned char k_tc[52] = {
	0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
	0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3,
	3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 13,
	14, 16, 18, 20
};

/* cpu.h:1064-1067 */
int hdbo_beta(unsigned qp) { return k_beta[qp > 51 ? 51 : qp]; }
/* cpu.h:1069-1072 */
int hdbo_tc(unsigned qp) { return k_tc[qp > 51 ? 51 : qp]; }

/* cpu.h:86 (and 104 with the chroma sizes): left-to-right unsigned arithmetic. */
unsigned hdbo_num_vert_bs(unsigned width, unsigned height) { return (width / S + 1) * height / S; }
/* cpu.h:87 (and 105) */
unsigned hdbo_num_hor_bs(unsigned width, unsigned height) { return (height / S + 1) * width / S; }

/* cpu.h:92-95 (luma), 110-113 (chroma) */
void hdbo_default_vert_bs(unsigned width, unsigned height, uint8_t *dst)
{
	unsigned n = hdbo_num_vert_bs(width, height), i;
	for (i = 0; i < n; i++) dst[i] = (i % (width / S + 1) == 0) ? 0 : 2;
}

/* cpu.h:96-99 (luma), 114-117 (chroma) */
void hdbo_default_hor_bs(unsigned width, unsigned height, uint8_t *dst)
{
	unsigned n = hdbo_num_hor_bs(width, height), i;
	for (i = 0; i < n; i++) dst[i] = (i % (height / S + 1) == 0) ? 0 : 2;
}

/*
 * Samples are held as 16-bit values whatever their depth: an 8-bit frame is widened on the way in and narrowed on
 * the way out (the reference's own storage is unsigned char, cpu.h:1041-1048, its arithmetic int: no value changes).
 * The one constant of the reference that depends on the depth is the sample maximum, 255 (cpu.h:1202, 1228, 1474);
 * g_maxv carries it.  Not thread-safe on purpose: set per call, test infrastructure.
 */
typedef uint16_t px_t;
static int g_maxv = 255;
static int g_depth_shift = 0;      /* bit depth - 8: beta and tc scale with it (see deblock_oracle.h) */

static int iabs(int v) { return v < 0 ? -v : v; }
/* cpu.h:1117-1120 Clip1: [-c, c] */
static int clip_sym(int d, int c) { int lo = -c; if (d < lo) d = lo; if (d > c) d = c; return d; }
/* cpu.h:1123-1126 Clip2 with c = 255 (the sample maximum of the depth) */
static px_t clip_px(int v) { if (v < 0) v = 0; if (v > g_maxv) v = g_maxv; return (px_t)v; }

/*
 * One 4-line edge segment.  Line L (0..3), tap k (0..3):
 *   p_k = pb[L*ls - k*ts]      q_k = qb[L*ls + k*ts]
 * pb points at p0 of line 0, qb at q0 of line 0.
 */
typedef struct { px_t *pb, *qb; ptrdiff_t ts, ls; } seg_t;

#define P(L, k) ((int)sg->pb[(L) * sg->ls - (k) * sg->ts])
#define Q(L, k) ((int)sg->qb[(L) * sg->ls + (k) * sg->ts])
#define PW(L, k) sg->pb[(L) * sg->ls - (k) * sg->ts]
#define QW(L, k) sg->qb[(L) * sg->ls + (k) * sg->ts]

/* second derivative |x2 - 2 x1 + x0| used by cpu.h:1086-1087, 1099-1100, 1245, 1249 */
static int d2p(const seg_t *sg, int L) { return iabs(P(L, 2) - 2 * P(L, 1) + P(L, 0)); }
static int d2q(const seg_t *sg, int L) { return iabs(Q(L, 2) - 2 * Q(L, 1) + Q(L, 0)); }

/* cpu.h:1128-1213 ApplyStrongFilter: all reads before any write, c = 2*tc */
static void luma_strong(const seg_t *sg, int tc)
{
	int c = 2 * tc, L;
	for (L = 0; L < 4; L++) {
		int p0 = P(L, 0), p1 = P(L, 1), p2 = P(L, 2), p3 = P(L, 3);
		int q0 = Q(L, 0), q1 = Q(L, 1), q2 = Q(L, 2), q3 = Q(L, 3);
		int dp0 = (p2 + 2 * p1 - 6 * p0 + 2 * q0 + q1 + 4) >> 3; /* cpu.h:1154 */
		int dp1 = (p2 - 3 * p1 + p0 + q0 + 2) >> 2;              /* cpu.h:1160 */
		int dp2 = (2 * p3 - 5 * p2 + p1 + p0 + q0 + 4) >> 3;     /* cpu.h:1166 */
		int dq0 = (q2 + 2 * q1 - 6 * q0 + 2 * p0 + p1 + 4) >> 3; /* cpu.h:1173 */
		int dq1 = (q2 - 3 * q1 + q0 + p0 + 2) >> 2;              /* cpu.h:1179 */
		int dq2 = (2 * q3 - 5 * q2 + q1 + q0 + p0 + 4) >> 3;     /* cpu.h:1185 */
		PW(L, 0) = clip_px(p0 + clip_sym(dp0, c));               /* cpu.h:1192-1211 */
		PW(L, 1) = clip_px(p1 + clip_sym(dp1, c));
		PW(L, 2) = clip_px(p2 + clip_sym(dp2, c));
		QW(L, 0) = clip_px(q0 + clip_sym(dq0, c));
		QW(L, 1) = clip_px(q1 + clip_sym(dq1, c));
		QW(L, 2) = clip_px(q2 + clip_sym(dq2, c));
	}
}

/* cpu.h:1215-1357 ApplyNormalFilter */
static void luma_normal(const seg_t *sg, int beta, int tc)
{
	int c = 2 * tc, c2 = tc / 2, L;                              /* cpu.h:1235-1236 */
	int dEp = d2p(sg, 0) + d2p(sg, 3) < 3 * beta / 16;           /* cpu.h:1245 */
	int dEq = d2q(sg, 0) + d2q(sg, 3) < 3 * beta / 16;           /* cpu.h:1249 */
	for (L = 0; L < 4; L++) {
		int p0 = P(L, 0), p1 = P(L, 1), p2 = P(L, 2);
		int q0 = Q(L, 0), q1 = Q(L, 1), q2 = Q(L, 2);
		int d = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;        /* cpu.h:1253 */
		if (iabs(d) < 10 * tc) {                                 /* cpu.h:1254 */
			int D = clip_sym(d, c);                              /* cpu.h:1256 */
			int Dp = clip_sym((((p2 + p0 + 1) >> 1) - p1 + D) >> 1, c2); /* cpu.h:1259-1260 */
			int Dq = clip_sym((((q2 + q0 + 1) >> 1) - q1 - D) >> 1, c2); /* cpu.h:1264-1265 */
			PW(L, 0) = clip_px(p0 + D);                          /* cpu.h:1268 */
			QW(L, 0) = clip_px(q0 - D);                          /* cpu.h:1269 */
			if (dEp) PW(L, 1) = clip_px(p1 + Dp);                /* cpu.h:1272 */
			if (dEq) QW(L, 1) = clip_px(q1 + Dq);                /* cpu.h:1275 */
		}
	}
}

/* cpu.h:1359-1429 DeblockingFilterLuma with 1074-1089 (dE) and 1091-1114 (strong) */
static void luma_segment(const seg_t *sg, int beta, int tc)
{
	int dp0 = d2p(sg, 0), dp3 = d2p(sg, 3), dq0 = d2q(sg, 0), dq3 = d2q(sg, 3);
	int strong;
	if (!(dp0 + dp3 + dq0 + dq3 < beta)) return;                 /* cpu.h:1086-1087 */
	strong = (dp0 + dq0 < beta / 8) && (dp3 + dq3 < beta / 8)    /* cpu.h:1099-1100 */
		&& (iabs(P(0, 3) - P(0, 0)) + iabs(Q(0, 0) - Q(0, 3)) < beta / 8) /* cpu.h:1104 */
		&& (iabs(P(3, 3) - P(3, 0)) + iabs(Q(3, 0) - Q(3, 3)) < beta / 8) /* cpu.h:1105 */
		&& (iabs(P(0, 0) - Q(0, 0)) < 5 * tc / 2)                /* cpu.h:1109 */
		&& (iabs(P(3, 0) - Q(3, 0)) < 5 * tc / 2);               /* cpu.h:1110 */
	if (strong) luma_strong(sg, tc); else luma_normal(sg, beta, tc);
}

/* cpu.h:1431-1488 DeblockingFilterChroma (sign convention as written there) */
static void chroma_segment(const seg_t *sg, int tc)
{
	int L;
	for (L = 0; L < 4; L++) {
		int p0 = P(L, 0), p1 = P(L, 1), q0 = Q(L, 0), q1 = Q(L, 1);
		int dp = (((p0 - q0) * 4) + p1 - q1 + 4) >> 3;           /* cpu.h:1453 (<<2 == *4) */
		int dq = (((q0 - p0) * 4) + q1 - p1 + 4) >> 3;           /* cpu.h:1458 */
		PW(L, 0) = clip_px(p0 + clip_sym(dp, tc));               /* cpu.h:1475 */
		QW(L, 0) = clip_px(q0 - clip_sym(dq, tc));               /* cpu.h:1476 */
	}
}

static unsigned bs_at(const uint8_t *bs, unsigned n, unsigned idx)
{
	return idx < n ? bs[idx] : 0u; /* zero-extension */
}

/*
 * Block driver over one padded plane: cpu.h:145-448 (luma), 453-721 / 724-992
 * (chroma).  pitch is the addressing pitch the reference uses (blocks_x * 8),
 * bw = W/8 (or cw/8) for BS indexing, lim_x/lim_y the block counts used in the
 * "not last block" tests -- the LUMA counts for both plane kinds (cpu.h:515,
 * 645, 786, 916).
 */
static void filter_plane(px_t *pl, int nbx, int nby, ptrdiff_t pitch, unsigned bw,
	int lim_x, int lim_y, const uint8_t *vbs, unsigned nv, const uint8_t *hbs, unsigned nh,
	int beta, int tc, int chroma, int edges)
{
	int bx, by;
	for (bx = 0; bx < nbx; bx++) {
		for (by = 0; by < nby; by++) {
			px_t *b = pl + (ptrdiff_t)by * S * pitch + (ptrdiff_t)bx * S;
			unsigned bs[4];
			seg_t sg[4];
			int e;
			/* upper vertical edge segment: rows 0-3, p cols 3..0, q cols 4..7  cpu.h:159-220 */
			bs[0] = by > 0 ? bs_at(vbs, nv, (unsigned)(by - 1) * (bw + 1) + bx) : 0;
			sg[0].pb = b + 3; sg[0].qb = b + 4; sg[0].ts = 1; sg[0].ls = pitch;
			/* lower vertical edge segment: rows 4-7  cpu.h:223-284 */
			bs[1] = by < lim_y - 1 ? bs_at(vbs, nv, (unsigned)by * (bw + 1) + bx) : 0;
			sg[1].pb = b + 4 * pitch + 3; sg[1].qb = b + 4 * pitch + 4; sg[1].ts = 1; sg[1].ls = pitch;
			/* left horizontal edge segment: cols 0-3, p rows 3..0, q rows 4..7  cpu.h:287-365 */
			bs[2] = bx > 0 ? bs_at(hbs, nh, (unsigned)by * bw + (bx - 1)) : 0;
			sg[2].pb = b + 3 * pitch; sg[2].qb = b + 4 * pitch; sg[2].ts = pitch; sg[2].ls = 1;
			/* right horizontal edge segment: p cols 4-7 but q cols 0-3  cpu.h:368-446 */
			bs[3] = bx < lim_x - 1 ? bs_at(hbs, nh, (unsigned)by * bw + bx) : 0;
			sg[3].pb = b + 3 * pitch + 4; sg[3].qb = b + 4 * pitch; sg[3].ts = pitch; sg[3].ls = 1;
			for (e = 0; e < 4; e++) {
				if (chroma) { if (bs[e] == 2) chroma_segment(&sg[e], tc); } /* cpu.h:463 */
				else if (edges) {
					/* edge byte = bs | qp << 2: the segment's own beta / tc (deblock_oracle.h) */
					if ((bs[e] & 3u) > 0) luma_segment(&sg[e], hdbo_beta(bs[e] >> 2) << g_depth_shift, hdbo_tc(bs[e] >> 2) << g_depth_shift);
				}
				else { if (bs[e] > 0) luma_segment(&sg[e], beta, tc); }     /* cpu.h:164 */
			}
		}
	}
}

/* one row of samples between the caller's storage (1 or 2 bytes per sample) and the 16-bit working planes */
static void row_in(px_t *dst, const void *src, size_t first, unsigned n, int bytes)
{
	unsigned i;
	if (bytes == 1) for (i = 0; i < n; i++) dst[i] = ((const uint8_t *)src)[first + i];
	else memcpy(dst, (const uint16_t *)src + first, (size_t)n * 2);
}

static void row_out(void *dst, size_t first, const px_t *src, unsigned n, int bytes)
{
	unsigned i;
	if (bytes == 1) for (i = 0; i < n; i++) ((uint8_t *)dst)[first + i] = (uint8_t)src[i];
	else memcpy((uint16_t *)dst + first, src, (size_t)n * 2);
}

/*
 * bytes = 1: the reference's own 8-bit frames (depth 8).  bytes = 2: 16-bit containers holding `depth`-bit samples
 * in their low bits (deblock_oracle.h: hdbo_deblock_frame16); depth 8 in such containers is the reference's
 * arithmetic unchanged.
 */
static int deblock_frame(const void *in, void *out, unsigned width, unsigned height,
	unsigned qp, const uint8_t *vert_bs, unsigned num_vert_bs,
	const uint8_t *hor_bs, unsigned num_hor_bs, int edges, int bytes, unsigned depth)
{
	unsigned cw = width / 2, ch = height / 2;                    /* cpu.h:59-60 */
	unsigned nw = width + S, nh = height + S;                    /* cpu.h:55-56 */
	unsigned ncw = cw + S, nch = ch + S;                         /* cpu.h:61-62 */
	unsigned nV = hdbo_num_vert_bs(width, height), nH = hdbo_num_hor_bs(width, height);
	unsigned ncV = hdbo_num_vert_bs(cw, ch), ncH = hdbo_num_hor_bs(cw, ch);
	int nbx = (int)(nw / S), nby = (int)(nh / S);                /* cpu.h:141-142 */
	int cnbx = (int)(ncw / S), cnby = (int)(nch / S);            /* cpu.h:450-451 */
	int beta = hdbo_beta(qp) << (depth - 8), tc = hdbo_tc(qp) << (depth - 8);   /* cpu.h:136-137; scaled for depth > 8 */
	px_t *Y, *C;
	uint8_t *vbs, *hbs, *cvbs, *chbs;
	unsigned r, pl;

	if (width % S != 0 || height % S != 0) return HDBO_ERR_SIZE; /* cpu.h:46-48 */
	if (vert_bs || hor_bs) {
		if (!vert_bs || !hor_bs) return HDBO_ERR_BS_SIZE;
		if (num_vert_bs != nV || num_hor_bs != nH) return HDBO_ERR_BS_SIZE; /* cpu.h:122-123 */
	}
	g_maxv = (1 << depth) - 1;
	g_depth_shift = (int)depth - 8;

	/* zero padding: the reference leaves it uninitialised (cpu.h:57, 63-64) */
	Y = (px_t *)calloc((size_t)nw * nh, sizeof(px_t));
	C = (px_t *)calloc((size_t)ncw * nch, sizeof(px_t));
	vbs = (uint8_t *)malloc(nV + 1); hbs = (uint8_t *)malloc(nH + 1);
	cvbs = (uint8_t *)malloc(ncV + 1); chbs = (uint8_t *)malloc(ncH + 1);
	if (vert_bs) { memcpy(vbs, vert_bs, nV); memcpy(hbs, hor_bs, nH); } /* cpu.h:125-130 */
	else { hdbo_default_vert_bs(width, height, vbs); hdbo_default_hor_bs(width, height, hbs); }
	hdbo_default_vert_bs(cw, ch, cvbs);                          /* cpu.h:110-113 */
	hdbo_default_hor_bs(cw, ch, chbs);                           /* cpu.h:114-117 */

	/* luma: image at (4,4) of the padded plane  cpu.h:68-71 */
	for (r = 0; r < height; r++)
		row_in(Y + (size_t)(r + S / 2) * nw + S / 2, in, (size_t)r * width, width, bytes);
	filter_plane(Y, nbx, nby, (ptrdiff_t)nbx * S, width / S, nbx, nby,
		vbs, nV, hbs, nH, beta, tc, 0, edges);
	for (r = 0; r < height; r++)                                 /* cpu.h:1000-1004 */
		row_out(out, (size_t)r * width, Y + (size_t)(r + S / 2) * nw + S / 2, width, bytes);

	/* the two chroma planes, identical treatment  cpu.h:73-82, 453-992, 1007-1016 */
	for (pl = 0; pl < 2; pl++) {
		const size_t first = (size_t)width * height + (size_t)pl * cw * ch;
		memset(C, 0, (size_t)ncw * nch * sizeof(px_t));
		for (r = 0; r < ch; r++)
			row_in(C + (size_t)(r + S / 2) * ncw + S / 2, in, first + (size_t)r * cw, cw, bytes);
		/* addressing pitch is cnbx*8 (cpu.h:469-471), "last block" limits are the
		 * luma counts (cpu.h:515, 645) */
		filter_plane(C, cnbx, cnby, (ptrdiff_t)cnbx * S, cw / S, nbx, nby,
			cvbs, ncV, chbs, ncH, beta, tc, 1, 0);
		for (r = 0; r < ch; r++)
			row_out(out, first + (size_t)r * cw, C + (size_t)(r + S / 2) * ncw + S / 2, cw, bytes);
	}
	free(Y); free(C); free(vbs); free(hbs); free(cvbs); free(chbs);
	g_maxv = 255; g_depth_shift = 0;
	return HDBO_OK;
}

int hdbo_deblock_frame(const uint8_t *in, uint8_t *out, unsigned width, unsigned height,
	unsigned qp, const uint8_t *vert_bs, unsigned num_vert_bs,
	const uint8_t *hor_bs, unsigned num_hor_bs)
{
	return deblock_frame(in, out, width, height, qp, vert_bs, num_vert_bs, hor_bs, num_hor_bs, 0, 1, 8);
}

/* cpu.h unchanged except that a luma segment takes beta / tc from its own edge byte */
int hdbo_deblock_frame_edges(const uint8_t *in, uint8_t *out, unsigned width, unsigned height,
	unsigned qp, const uint8_t *vert_edges, unsigned num_vert, const uint8_t *hor_edges, unsigned num_hor)
{
	if (!vert_edges || !hor_edges) return HDBO_ERR_BS_SIZE;
	return deblock_frame(in, out, width, height, qp, vert_edges, num_vert, hor_edges, num_hor, 1, 1, 8);
}

/* 16-bit containers, bit_depth 8 .. 12 (deblock_oracle.h) */
int hdbo_deblock_frame16(const uint16_t *in, uint16_t *out, unsigned width, unsigned height,
	unsigned qp, const uint8_t *vert, unsigned num_vert, const uint8_t *hor, unsigned num_hor,
	unsigned bit_depth, int edges)
{
	if (bit_depth < 8 || bit_depth > 12) return HDBO_ERR_SIZE;
	if (edges && (!vert || !hor)) return HDBO_ERR_BS_SIZE;
	return deblock_frame(in, out, width, height, qp, vert, num_vert, hor, num_hor, edges ? 1 : 0, 2, bit_depth);
}

/* ------------------------------------------------------------------ */
/* boundary-strength and edge-QP derivation (H.265 8.7.2.4; no reference counterpart) */
/* ------------------------------------------------------------------ */
static int mv_far(const int16_t *a, const int16_t *b)
{
	return iabs((int)a[0] - (int)b[0]) >= 4 || iabs((int)a[1] - (int)b[1]) >= 4;
}

static unsigned derive_bs(const hdbo_block_info *p, const hdbo_block_info *q, int tu_edge, int pu_edge)
{
	int np, nq;
	if (!tu_edge && !pu_edge) return 0;
	if ((p->flags | q->flags) & HDBO_BLK_INTRA) return 2;
	if (tu_edge && ((p->flags | q->flags) & HDBO_BLK_CBF)) return 1;
	np = (p->ref[0] >= 0) + (p->ref[1] >= 0);
	nq = (q->ref[0] >= 0) + (q->ref[1] >= 0);
	if (np != nq) return 1;
	if (np == 1) {
		int a = p->ref[0] >= 0 ? 0 : 1, b = q->ref[0] >= 0 ? 0 : 1;
		if (p->ref[a] != q->ref[b]) return 1;
		return (unsigned)mv_far(p->mv[a], q->mv[b]);
	}
	if (np == 2) {
		if (!((p->ref[0] == q->ref[0] && p->ref[1] == q->ref[1]) || (p->ref[0] == q->ref[1] && p->ref[1] == q->ref[0])))
			return 1;
		if (p->ref[0] != p->ref[1]) {
			if (p->ref[0] == q->ref[0]) return (unsigned)(mv_far(p->mv[0], q->mv[0]) || mv_far(p->mv[1], q->mv[1]));
			return (unsigned)(mv_far(p->mv[0], q->mv[1]) || mv_far(p->mv[1], q->mv[0]));
		}
		return (unsigned)((mv_far(p->mv[0], q->mv[0]) || mv_far(p->mv[1], q->mv[1]))
			&& (mv_far(p->mv[0], q->mv[1]) || mv_far(p->mv[1], q->mv[0])));
	}
	return 0;
}

static uint8_t derive_edge(const hdbo_block_info *p, const hdbo_block_info *q, int tu_edge, int pu_edge)
{
	unsigned qp = ((unsigned)p->qp + q->qp + 1) >> 1;
	if (qp > 51) qp = 51;
	return HDBO_EDGE(derive_bs(p, q, tu_edge, pu_edge), qp);
}

void hdbo_derive_edges(const hdbo_block_info *info, unsigned width, unsigned height,
	uint8_t *vert_edges, uint8_t *hor_edges)
{
	unsigned w8 = width / S, h8 = height / S, r, x;
	/* vertical edges: [h8 rows][w8 + 1 edges] (cpu.h:86, 88); edge 0 and edge w8 are the picture's
	 * sides, which H.265 does not filter */
	for (r = 0; r < h8; r++)
		for (x = 0; x <= w8; x++) {
			const hdbo_block_info *q = info + (size_t)r * w8 + x;
			vert_edges[(size_t)r * (w8 + 1) + x] = (x == 0 || x == w8) ? 0
				: derive_edge(q - 1, q, q->flags & HDBO_BLK_TU_LEFT, q->flags & HDBO_BLK_PU_LEFT);
		}
	/* horizontal edges: [h8 + 1 edge rows][w8] (cpu.h:87, 89) */
	for (r = 0; r <= h8; r++)
		for (x = 0; x < w8; x++) {
			const hdbo_block_info *q = info + (size_t)r * w8 + x;
			hor_edges[(size_t)r * w8 + x] = (r == 0 || r == h8) ? 0
				: derive_edge(q - w8, q, q->flags & HDBO_BLK_TU_TOP, q->flags & HDBO_BLK_PU_TOP);
		}
}

int hdbo_deblock_batch(const uint8_t *in, uint8_t *out, unsigned n, unsigned width,
	unsigned height, const unsigned *qp, const uint8_t *vert_bs, const uint8_t *hor_bs)
{
	size_t fb = (size_t)width * height * 3 / 2;
	unsigned nV = hdbo_num_vert_bs(width, height), nH = hdbo_num_hor_bs(width, height), f;
	for (f = 0; f < n; f++) {
		int rc = hdbo_deblock_frame(in + f * fb, out + f * fb, width, height, qp[f],
			vert_bs ? vert_bs + (size_t)f * nV : NULL, nV,
			hor_bs ? hor_bs + (size_t)f * nH : NULL, nH);
		if (rc) return rc;
	}
	return HDBO_OK;
}

/* ------------------------------------------------------------------ */
/* sample adaptive offset (H.265 8.7.3; no reference counterpart)      */
/* ------------------------------------------------------------------ */
static int isign(int v) { return (v > 0) - (v < 0); }

/* one plane; the same statement for a byte a sample and for 16-bit containers (band shift = bit depth - 5, sample
 * maximum (1 << bit depth) - 1) */
#define HDBO_SAO_PLANE(NAME, T) \
static void NAME(const T *in, T *out, int w, int h, const hdbo_sao_params *params, int ctb_cols, int ctb, int depth) \
{ \
	/* 8.7.3.2: hPos / vPos of the two neighbours per SaoEoClass */ \
	static const int hpos[4][2] = { { -1, 1 }, { 0, 0 }, { -1, 1 }, { 1, -1 } }; \
	static const int vpos[4][2] = { { 0, 0 }, { -1, 1 }, { -1, 1 }, { -1, 1 } }; \
	const int maxv = (1 << depth) - 1; \
	int x, y; \
	for (y = 0; y < h; y++) \
		for (x = 0; x < w; x++) { \
			const hdbo_sao_params *p = params + (size_t)(y / ctb) * ctb_cols + x / ctb; \
			int c = in[(size_t)y * w + x], v = c; \
			if (p->type == 1) { \
				int k = ((c >> (depth - 5)) - p->position) & 31;      /* bandShift = bitDepth - 5 */ \
				if (k < 4) v = c + p->offset[k]; \
			} else if (p->type == 2) { \
				int cls = p->position & 3; \
				int xa = x + hpos[cls][0], ya = y + vpos[cls][0], xb = x + hpos[cls][1], yb = y + vpos[cls][1]; \
				if (xa >= 0 && xa < w && ya >= 0 && ya < h && xb >= 0 && xb < w && yb >= 0 && yb < h) { \
					int e = 2 + isign(c - in[(size_t)ya * w + xa]) + isign(c - in[(size_t)yb * w + xb]); \
					if (e != 2) v = c + p->offset[e < 2 ? e : e - 1];   /* edgeIdx 0, 1, 3, 4 -> offsets 0..3 */ \
				} \
			} \
			out[(size_t)y * w + x] = (T)(v < 0 ? 0 : (v > maxv ? maxv : v)); \
		} \
}
HDBO_SAO_PLANE(sao_plane8, uint8_t)
HDBO_SAO_PLANE(sao_plane16, uint16_t)

int hdbo_sao_frame(const uint8_t *in, uint8_t *out, unsigned width, unsigned height,
	const hdbo_sao_params *params, unsigned ctb_size)
{
	unsigned cw = width / 2, ch = height / 2;
	unsigned cols, rows;
	if (width % 8 != 0 || height % 8 != 0) return HDBO_ERR_SIZE;
	if (ctb_size != 16 && ctb_size != 32 && ctb_size != 64) return HDBO_ERR_SIZE;
	cols = (width + ctb_size - 1) / ctb_size; rows = (height + ctb_size - 1) / ctb_size;
	sao_plane8(in, out, (int)width, (int)height, params, (int)cols, (int)ctb_size, 8);
	sao_plane8(in + (size_t)width * height, out + (size_t)width * height, (int)cw, (int)ch,
		params + (size_t)rows * cols, (int)cols, (int)ctb_size / 2, 8);
	sao_plane8(in + (size_t)width * height + (size_t)cw * ch, out + (size_t)width * height + (size_t)cw * ch, (int)cw, (int)ch,
		params + 2 * (size_t)rows * cols, (int)cols, (int)ctb_size / 2, 8);
	return HDBO_OK;
}

int hdbo_sao_frame16(const uint16_t *in, uint16_t *out, unsigned width, unsigned height,
	const hdbo_sao_params *params, unsigned ctb_size, unsigned bit_depth)
{
	unsigned cw = width / 2, ch = height / 2;
	unsigned cols, rows;
	if (width % 8 != 0 || height % 8 != 0) return HDBO_ERR_SIZE;
	if (ctb_size != 16 && ctb_size != 32 && ctb_size != 64) return HDBO_ERR_SIZE;
	if (bit_depth < 8 || bit_depth > 12) return HDBO_ERR_SIZE;
	cols = (width + ctb_size - 1) / ctb_size; rows = (height + ctb_size - 1) / ctb_size;
	sao_plane16(in, out, (int)width, (int)height, params, (int)cols, (int)ctb_size, (int)bit_depth);
	sao_plane16(in + (size_t)width * height, out + (size_t)width * height, (int)cw, (int)ch,
		params + (size_t)rows * cols, (int)cols, (int)ctb_size / 2, (int)bit_depth);
	sao_plane16(in + (size_t)width * height + (size_t)cw * ch, out + (size_t)width * height + (size_t)cw * ch, (int)cw, (int)ch,
		params + 2 * (size_t)rows * cols, (int)cols, (int)ctb_size / 2, (int)bit_depth);
	return HDBO_OK;
}
