/*
 * deblock_core16.cuh -- the filter arithmetic for samples deeper than 8 bits (SURVEY.md section 8(f) rank 2:
 * 10-bit surfaces).  Same structure as deblock_core.cuh: one thread owns one padded 8x8 block in registers and
 * runs the reference's four edge segments in the reference's order (hevc_deblocking_filter_cpu.h ("cpu.h")
 * :158-446); the four lines of a segment are two "pair" registers per tap, pair X = lines (0, 3), pair Y = lines
 * (1, 2), 16 bits per line, evaluated with the packed-halfword instructions.  What changes with the depth:
 *   - a sample fills its 16-bit half, so a block is 32 words (row r = w[r][0..3], two samples each) and a tap of
 *     a vertical edge is ONE byte permute per line pair (no masking);
 *   - beta and tc are scaled by 1 << (depth - 8) and 255 becomes (1 << depth) - 1.  The reference has neither
 *     (it is 8-bit only: unsigned char planes, cpu.h:1041-1048, and 255 in its clips, cpu.h:1202, 1228, 1474);
 *     the scaling is the published rule (H.265 8.7.2.5.3) and oracle/deblock_oracle.h::hdbo_deblock_frame16
 *     defines the result.  DEPTHS ABOVE 8 ARE UNPINNED AGAINST THE REFERENCE; depth 8 in 16-bit containers
 *     must equal the 8-bit filter (and does: tests/test_depth.py);
 *   - every positive bias that keeps the packed intermediates non-negative grows with the depth (Depth<>
 *     below); the largest intermediate is 8 * max + 16 tc + 7 = 9471 at depth 10, so signed halves hold up to
 *     depth 10 and no further.
 * Compiles for the host with HDB_HOST_EMU like deblock_core.cuh (tests/emu).
 */
#pragma once

#include "deblock_core.cuh"

namespace hdb {

template <int kDepth>
struct Depth {
	static_assert(kDepth >= 8 && kDepth <= 10, "signed 16-bit halves hold the strong filter's sums up to depth 10");
	static constexpr u32 kMax = (1u << kDepth) - 1u;          /* cpu.h:1202: 255 */
	static constexpr u32 kMaxRep = kMax * 0x10001u;
	static constexpr u32 kB2 = 2u << kDepth;                  /* bias of a signed second difference (512 at depth 8) */
	static constexpr u32 kE0 = 1u << kDepth;                  /* bias of delta after its shift (256 at depth 8) */
	static constexpr u32 kB2Rep = kB2 * 0x10001u, kE0x2Rep = 2u * kE0 * 0x10001u;
	static constexpr u32 kEMask = (2u * kE0 - 1u) * 0x10001u;
};

struct FilterParams16 {
	int dE_lim, dEpq_neg;
	u32 k_b8, k_tc52;
	u32 s8_neg, s8_pos, s4_neg, s4_pos;
	u32 c2_neg, c2_pos, k_e_lo, k_tc10, k_half, h4_lo, h8_hi, h4_mid, h_neg, h_rep, c2_rep4;
	u32 ct_neg, ct_lo, ct_2;
};

/* beta / tc: the 8-bit table values (cpu.h:1021-1033), scaled here */
template <int kDepth>
HDB_FN void MakeFilterParams16(FilterParams16 &fp, int beta8, int tc8)
{
	typedef Depth<kDepth> D;
	const int beta = beta8 << (kDepth - 8), tc = tc8 << (kDepth - 8);
	fp.dE_lim = 4 * (int)D::kB2 + beta;                       /* cpu.h:1086-1087 on the biased sum */
	fp.dEpq_neg = -(2 * (int)D::kB2 + 3 * beta / 16);         /* cpu.h:1245, 1249 as a sign bit */
	fp.k_b8 = Rep2(0x8000 + beta / 8 - 1);                    /* cpu.h:1099-1105 */
	fp.k_tc52 = Rep2(0x8000 + 5 * tc / 2 - 1);                /* cpu.h:1109-1110 */
	fp.s8_neg = Rep2(-16 * tc); fp.s8_pos = Rep2(16 * tc + 7);
	fp.s4_neg = Rep2(-8 * tc); fp.s4_pos = Rep2(8 * tc + 3);
	fp.c2_neg = Rep2(-2 * tc); fp.c2_pos = Rep2(2 * tc);
	fp.k_e_lo = Rep2(2 * tc - (int)D::kE0);
	fp.k_tc10 = Rep2(10 * tc - (int)D::kE0);                  /* cpu.h:1254 */
	fp.k_half = Rep2((int)D::kB2 + 1 - 4 * tc);
	fp.h4_lo = Rep2(4 * (tc / 2) - 2 * (int)D::kB2); fp.h8_hi = Rep2(8 * (tc / 2) + 3);
	fp.h4_mid = Rep2(4 * (tc / 2) + 3);
	fp.h_neg = Rep2(-(tc / 2)); fp.h_rep = Rep2(tc / 2);
	fp.c2_rep4 = Rep2(4 * tc);
	fp.ct_neg = Rep2(-tc); fp.ct_lo = Rep2(tc - (int)D::kE0); fp.ct_2 = Rep2(2 * tc);
}

/* |a - b| per half, halves holding samples of up to 10 bits */
HDB_FN u32 AbsDiff16(u32 a, u32 b) { return Max2(a, b) - Min2(a, b); }

template <int kDepth>
HDB_FN u32 SecondDiffSigned16(u32 a2, u32 a1, u32 a0) { return a2 + a0 + Depth<kDepth>::kB2Rep - 2u * a1; }

template <int kDepth>
HDB_FN u32 SecondDiffAbs16(u32 sd) { return Max2(sd, 2u * Depth<kDepth>::kB2Rep - sd); }

/* cpu.h:1128-1213; see StrongPair in deblock_core.cuh.  A sample fills its half here, so what the final shift
 * drags across from the upper half has to be masked off. */
template <int kDepth>
HDB_FN void StrongPair16(LumaPair &t, const FilterParams16 &fp)
{
	const u32 m = Depth<kDepth>::kMaxRep;
	u32 s2 = t.p0 + t.q0 + 0x00020002u;
	u32 t1p = t.p2 + t.p1 + s2;
	u32 t1q = t.q2 + t.q1 + s2;
	u32 mm = t.p1 + t.q1 + s2;
	u32 t0p = t1p + mm;
	u32 t0q = t1q + mm;
	u32 t2p = 2u * (t.p3 + t.p2 + 0x00010001u) + t1p;
	u32 t2q = 2u * (t.q3 + t.q2 + 0x00010001u) + t1q;
	u32 a;
	a = 8u * t.p0; t.p0 = (AddMin2(a, fp.s8_pos, AddMax2(a, fp.s8_neg, t0p)) >> 3) & m;
	a = 4u * t.p1; t.p1 = (AddMin2(a, fp.s4_pos, AddMax2(a, fp.s4_neg, t1p)) >> 2) & m;
	a = 8u * t.p2; t.p2 = (AddMin2(a, fp.s8_pos, AddMax2(a, fp.s8_neg, t2p)) >> 3) & m;
	a = 8u * t.q0; t.q0 = (AddMin2(a, fp.s8_pos, AddMax2(a, fp.s8_neg, t0q)) >> 3) & m;
	a = 4u * t.q1; t.q1 = (AddMin2(a, fp.s4_pos, AddMax2(a, fp.s4_neg, t1q)) >> 2) & m;
	a = 8u * t.q2; t.q2 = (AddMin2(a, fp.s8_pos, AddMax2(a, fp.s8_neg, t2q)) >> 3) & m;
}

/* cpu.h:1215-1357; see NormalPair in deblock_core.cuh */
template <int kDepth>
HDB_FN void NormalPair16(LumaPair &t, const FilterParams16 &fp, u32 lo_p, u32 hi_p, u32 lo_q, u32 hi_q, u32 sd_p, u32 sd_q)
{
	typedef Depth<kDepth> D;
	/* 16 E0 + 8 + 9(q0-p0) - 3(q1-p1): positive in each half */
	u32 num = 3u * (3u * (t.q0 - t.p0) + t.p1 - t.q1) + (16u * D::kE0 + 8u) * 0x10001u;
	u32 e = (num >> 4) & D::kEMask;                          /* E0 + delta                 cpu.h:1253 */
	u32 na = Min2(e, D::kE0x2Rep - e);                        /* E0 - |delta| */
	u32 on = AddMinRelu2(na, fp.k_tc10, 0x00010001u);         /* 1 where |delta| < 10tc     cpu.h:1254 */
	u32 mask = on * 0x01ffu;                                  /* 9 bits per half: every value selected below is < 512 */
	u32 d = AddMinRelu2(e, fp.k_e_lo, fp.c2_rep4);            /* Clip(delta) + 2tc in [0, 4tc]  cpu.h:1256 */
	d = (d & mask) | (fp.c2_pos & ~mask);
	u32 dn = fp.c2_rep4 - d;
	u32 hp = AddMinRelu2(sd_p + fp.k_half + 2u * d, lo_p, hi_p) >> 2;      /* cpu.h:1259-1265 */
	u32 hq = AddMinRelu2(sd_q + fp.k_half + 2u * dn, lo_q, hi_q) >> 2;
	u32 dp = (hp & mask) | (fp.h_rep & ~mask);
	u32 dq = (hq & mask) | (fp.h_rep & ~mask);
	t.p1 = AddMinRelu2(t.p1 + dp, fp.h_neg, D::kMaxRep);      /* cpu.h:1272 */
	t.q1 = AddMinRelu2(t.q1 + dq, fp.h_neg, D::kMaxRep);      /* cpu.h:1275 */
	t.p0 = AddMinRelu2(t.p0 + d, fp.c2_neg, D::kMaxRep);      /* cpu.h:1268 */
	t.q0 = AddMinRelu2(t.q0 + dn, fp.c2_neg, D::kMaxRep);     /* cpu.h:1269 */
}

/* cpu.h:1359-1429 with 1074-1114; see LumaDecide / LumaNormal in deblock_core.cuh */
template <int kDepth>
HDB_FN int LumaSegment16(LumaPair &x, LumaPair &y, const FilterParams16 &fp)
{
	typedef Depth<kDepth> D;
	const u32 sdp = SecondDiffSigned16<kDepth>(x.p2, x.p1, x.p0), sdq = SecondDiffSigned16<kDepth>(x.q2, x.q1, x.q0);
	const u32 dP = SecondDiffAbs16<kDepth>(sdp), dQ = SecondDiffAbs16<kDepth>(sdq);     /* B2 + dp0 | B2 + dp3 */
	const u32 s = dP + dQ;
	if (!((int)SumHalves(s) < fp.dE_lim)) return kSegOff;                               /* cpu.h:1086-1087 */
	const u32 c3 = AbsDiff16(x.p3, x.p0) + AbsDiff16(x.q0, x.q3);
	const u32 c4 = AbsDiff16(x.p0, x.q0);
	const u32 m1 = Max2(s - 2u * D::kB2Rep, c3);
	const u32 ok = (fp.k_b8 - m1) & (fp.k_tc52 - c4) & 0x80008000u;                     /* cpu.h:1099-1110 */
	if (ok == 0x80008000u) {
		StrongPair16<kDepth>(x, fp);
		StrongPair16<kDepth>(y, fp);
		return kSegStrong;
	}
	const u32 m_p = (u32)((int)SumHalvesPlus(dP, (u32)fp.dEpq_neg) >> 31);              /* cpu.h:1245 */
	const u32 m_q = (u32)((int)SumHalvesPlus(dQ, (u32)fp.dEpq_neg) >> 31);              /* cpu.h:1249 */
	const u32 lo_p = (fp.h4_lo & m_p) | (0x40004000u & ~m_p), hi_p = (fp.h8_hi & m_p) | (fp.h4_mid & ~m_p);
	const u32 lo_q = (fp.h4_lo & m_q) | (0x40004000u & ~m_q), hi_q = (fp.h8_hi & m_q) | (fp.h4_mid & ~m_q);
	NormalPair16<kDepth>(x, fp, lo_p, hi_p, lo_q, hi_q, sdp, sdq);
	NormalPair16<kDepth>(y, fp, lo_p, hi_p, lo_q, hi_q, SecondDiffSigned16<kDepth>(y.p2, y.p1, y.p0), SecondDiffSigned16<kDepth>(y.q2, y.q1, y.q0));
	return kSegNormal;
}

/* cpu.h:1431-1488; see ChromaPairFilter in deblock_core.cuh */
template <int kDepth>
HDB_FN void ChromaPairFilter16(ChromaPair &t, const FilterParams16 &fp)
{
	typedef Depth<kDepth> D;
	u32 np = 4u * (t.p0 - t.q0) + t.p1 - t.q1 + (8u * D::kE0 + 4u) * 0x10001u;   /* cpu.h:1453 */
	u32 nq = 2u * (8u * D::kE0 + 4u) * 0x10001u - np;                             /* cpu.h:1458 */
	u32 ep = (np >> 3) & D::kEMask;
	u32 eq = (nq >> 3) & D::kEMask;
	u32 dp = AddMinRelu2(ep, fp.ct_lo, fp.ct_2);                                  /* cpu.h:1464 */
	u32 dq = AddMinRelu2(eq, fp.ct_lo, fp.ct_2);                                  /* cpu.h:1469 */
	t.p0 = AddMinRelu2(t.p0 + dp, fp.ct_neg, D::kMaxRep);                         /* cpu.h:1475 */
	t.q0 = AddMinRelu2(t.q0 + fp.ct_2 - dq, fp.ct_neg, D::kMaxRep);               /* cpu.h:1476 */
}

/* ------------------------------------------------------------------ */
/* register block (8 rows x 4 words, two samples a word) <-> line pairs */
/* ------------------------------------------------------------------ */
/* low halves of a and b / high halves of a and b */
HDB_FN u32 PairLo(u32 a, u32 b) { return Prmt(a, b, 0x5410); }
HDB_FN u32 PairHi(u32 a, u32 b) { return Prmt(a, b, 0x7632); }

/* vertical edge: taps run along the row -- p3 p2 | p1 p0 | q0 q1 | q2 q3 are the row's four words; pair = rows (ra, rb) */
HDB_FN void LoadVertPair16(LumaPair &t, const u32 (&a)[4], const u32 (&b)[4])
{
	t.p3 = PairLo(a[0], b[0]); t.p2 = PairHi(a[0], b[0]);
	t.p1 = PairLo(a[1], b[1]); t.p0 = PairHi(a[1], b[1]);
	t.q0 = PairLo(a[2], b[2]); t.q1 = PairHi(a[2], b[2]);
	t.q2 = PairLo(a[3], b[3]); t.q3 = PairHi(a[3], b[3]);
}

HDB_FN void StoreVertPair16(const LumaPair &t, int kind, u32 (&a)[4], u32 (&b)[4])
{
	a[1] = PairLo(t.p1, t.p0); b[1] = PairHi(t.p1, t.p0);
	a[2] = PairLo(t.q0, t.q1); b[2] = PairHi(t.q0, t.q1);
	if (kind == kSegStrong) {
		a[0] = PairLo(t.p3, t.p2); b[0] = PairHi(t.p3, t.p2);
		a[3] = PairLo(t.q2, t.q3); b[3] = PairHi(t.q2, t.q3);
	}
}

/* horizontal edge: the four samples of a row's word pair (w0, w1) are the four lines of one tap;
 * x = lines (0, 3), y = lines (1, 2) */
HDB_FN void SplitRow16(u32 w0, u32 w1, u32 &x, u32 &y) { x = Prmt(w0, w1, 0x7610); y = Prmt(w0, w1, 0x5432); }
HDB_FN void JoinRow16(u32 x, u32 y, u32 &w0, u32 &w1) { w0 = Prmt(x, y, 0x5410); w1 = Prmt(x, y, 0x3276); }

template <int kDepth, bool kEdgeQp>
HDB_FN void LumaBlock16(u32 (&w)[8][4], u32 bs_ver1, u32 bs_ver2, u32 bs_hor1, u32 bs_hor2, const FilterParams16 &fp)
{
	FilterParams16 fpe;
	u32 bt_ver1 = 0, bt_ver2 = 0, bt_hor1 = 0, bt_hor2 = 0;
	if (kEdgeQp) {
		bt_ver1 = BetaTcOfQp6(bs_ver1 >> 2); bt_ver2 = BetaTcOfQp6(bs_ver2 >> 2);
		bt_hor1 = BetaTcOfQp6(bs_hor1 >> 2); bt_hor2 = BetaTcOfQp6(bs_hor2 >> 2);
		bs_ver1 &= 3u; bs_ver2 &= 3u; bs_hor1 &= 3u; bs_hor2 &= 3u;
	}
#pragma unroll
	for (int s = 0; s < 2; s++) {
		const int r = 4 * s;
		if ((s == 0 ? bs_ver1 : bs_ver2) > 0) {                     /* cpu.h:164, 228 */
			if (kEdgeQp) { const u32 bt = s == 0 ? bt_ver1 : bt_ver2; MakeFilterParams16<kDepth>(fpe, (int)(bt & 0xffu), (int)(bt >> 8)); }
			LumaPair x, y;
			LoadVertPair16(x, w[r], w[r + 3]);
			LoadVertPair16(y, w[r + 1], w[r + 2]);
			const int kind = LumaSegment16<kDepth>(x, y, kEdgeQp ? fpe : fp);
			if (kind != kSegOff) {
				StoreVertPair16(x, kind, w[r], w[r + 3]);
				StoreVertPair16(y, kind, w[r + 1], w[r + 2]);
			}
		}
	}
	/* left (c = 0) and right (c = 1) horizontal edge segments; the right one pairs p = columns 4-7 of rows 3..0
	 * with q = columns 0-3 of rows 4..7, the reference's pairing (cpu.h:383-445) */
#pragma unroll
	for (int c = 0; c < 2; c++) {
		if ((c == 0 ? bs_hor1 : bs_hor2) > 0) {
			if (kEdgeQp) { const u32 bt = c == 0 ? bt_hor1 : bt_hor2; MakeFilterParams16<kDepth>(fpe, (int)(bt & 0xffu), (int)(bt >> 8)); }
			const int pc = 2 * c;
			LumaPair x, y;
			SplitRow16(w[0][pc], w[0][pc + 1], x.p3, y.p3); SplitRow16(w[1][pc], w[1][pc + 1], x.p2, y.p2);
			SplitRow16(w[2][pc], w[2][pc + 1], x.p1, y.p1); SplitRow16(w[3][pc], w[3][pc + 1], x.p0, y.p0);
			SplitRow16(w[4][0], w[4][1], x.q0, y.q0); SplitRow16(w[5][0], w[5][1], x.q1, y.q1);
			SplitRow16(w[6][0], w[6][1], x.q2, y.q2); SplitRow16(w[7][0], w[7][1], x.q3, y.q3);
			const int kind = LumaSegment16<kDepth>(x, y, kEdgeQp ? fpe : fp);
			if (kind != kSegOff) {
				JoinRow16(x.p1, y.p1, w[2][pc], w[2][pc + 1]); JoinRow16(x.p0, y.p0, w[3][pc], w[3][pc + 1]);
				JoinRow16(x.q0, y.q0, w[4][0], w[4][1]); JoinRow16(x.q1, y.q1, w[5][0], w[5][1]);
				if (kind == kSegStrong) { JoinRow16(x.p2, y.p2, w[1][pc], w[1][pc + 1]); JoinRow16(x.q2, y.q2, w[6][0], w[6][1]); }
			}
		}
	}
}

/* cpu.h:453-721: taps k = 0, 1, BS == 2 only */
template <int kDepth>
HDB_FN void ChromaBlock16(u32 (&w)[8][4], u32 bs_ver1, u32 bs_ver2, u32 bs_hor1, u32 bs_hor2, const FilterParams16 &fp)
{
#pragma unroll
	for (int s = 0; s < 2; s++) {
		const int r = 4 * s;
		if ((s == 0 ? bs_ver1 : bs_ver2) == 2) {                    /* cpu.h:463, 519 */
#pragma unroll
			for (int k = 0; k < 2; k++) {
				const int ra = k == 0 ? r : r + 1, rb = k == 0 ? r + 3 : r + 2;
				ChromaPair t;
				t.p1 = PairLo(w[ra][1], w[rb][1]); t.p0 = PairHi(w[ra][1], w[rb][1]);
				t.q0 = PairLo(w[ra][2], w[rb][2]); t.q1 = PairHi(w[ra][2], w[rb][2]);
				ChromaPairFilter16<kDepth>(t, fp);
				w[ra][1] = PairLo(t.p1, t.p0); w[rb][1] = PairHi(t.p1, t.p0);
				w[ra][2] = PairLo(t.q0, t.q1); w[rb][2] = PairHi(t.q0, t.q1);
			}
		}
	}
#pragma unroll
	for (int c = 0; c < 2; c++) {
		if ((c == 0 ? bs_hor1 : bs_hor2) == 2) {                    /* cpu.h:572, 649 */
			const int pc = 2 * c;
			ChromaPair x, y;
			SplitRow16(w[2][pc], w[2][pc + 1], x.p1, y.p1); SplitRow16(w[3][pc], w[3][pc + 1], x.p0, y.p0);
			SplitRow16(w[4][0], w[4][1], x.q0, y.q0); SplitRow16(w[5][0], w[5][1], x.q1, y.q1);
			ChromaPairFilter16<kDepth>(x, fp);
			ChromaPairFilter16<kDepth>(y, fp);
			JoinRow16(x.p0, y.p0, w[3][pc], w[3][pc + 1]); JoinRow16(x.q0, y.q0, w[4][0], w[4][1]);
		}
	}
}

} // namespace hdb
