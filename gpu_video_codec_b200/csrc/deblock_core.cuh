/*
 * deblock_core.cuh -- the filter arithmetic of the B200 deblocking path.
 *
 * One CUDA thread owns one padded 8x8 block (the reference's unit of work,
 * hevc_deblocking_filter_cpu.h ("cpu.h") :146-156) entirely in registers:
 * sixteen 32-bit words, row r = { lo[r] : columns 0..3, hi[r] : columns 4..7 }.
 * The four edge segments of the block are filtered in the reference's order
 * (cpu.h:158-446): upper vertical, lower vertical, left horizontal, right
 * horizontal.
 *
 * Arithmetic is SIMD-within-a-register: the four lines of a segment are held
 * as two "pair" registers per tap, each carrying two lines as 16-bit halves
 *     pair X = lines (0,3)   -- the two lines the decisions look at (cpu.h:1074-1114)
 *     pair Y = lines (1,2)
 * and evaluated with the sm_100a packed-halfword instructions VIADD.16x2,
 * VIMNMX.S16x2, VIADDMNMX.S16x2(.RELU), IDP.2A and PRMT.  Signed intermediate
 * values are carried with a positive bias so that plain 32-bit adds and shifts
 * act on both halves at once.  No tensor cores: this is an integer stencil.
 *
 * The same header compiles for the host when HDB_HOST_EMU is defined: the five
 * packed primitives are then plain C.  tests/emu/ uses that build to prove the
 * arithmetic bit-exact against the oracle on a machine without a GPU.  The
 * product library never compiles the host flavour.
 */
#pragma once

#include <stdint.h>


#if defined(HDB_HOST_EMU)
#define HDB_FN static inline
#define HDB_HD static inline
#else
#define HDB_FN __device__ __forceinline__
#define HDB_HD __host__ __device__ __forceinline__
#endif

namespace hdb {

typedef uint32_t u32;

/* ------------------------------------------------------------------ */
/* packed-halfword primitives                                          */
/* ------------------------------------------------------------------ */
#if defined(HDB_HOST_EMU)
static inline int16_t Lo(u32 a) { return (int16_t)(a & 0xffffu); }
static inline int16_t Hi(u32 a) { return (int16_t)(a >> 16); }
static inline u32 Pk(int lo, int hi) { return ((u32)lo & 0xffffu) | ((u32)hi << 16); }
static inline int Mn(int a, int b) { return a < b ? a : b; }
static inline int Mx(int a, int b) { return a > b ? a : b; }
HDB_FN u32 Add2(u32 a, u32 b) { return Pk(Lo(a) + Lo(b), Hi(a) + Hi(b)); }
HDB_FN u32 Max2(u32 a, u32 b) { return Pk(Mx(Lo(a), Lo(b)), Mx(Hi(a), Hi(b))); }
HDB_FN u32 Min2(u32 a, u32 b) { return Pk(Mn(Lo(a), Lo(b)), Mn(Hi(a), Hi(b))); }
HDB_FN u32 AddMax2(u32 a, u32 b, u32 c) {
	return Pk(Mx((int16_t)(Lo(a) + Lo(b)), Lo(c)), Mx((int16_t)(Hi(a) + Hi(b)), Hi(c)));
}
HDB_FN u32 AddMin2(u32 a, u32 b, u32 c) {
	return Pk(Mn((int16_t)(Lo(a) + Lo(b)), Lo(c)), Mn((int16_t)(Hi(a) + Hi(b)), Hi(c)));
}
HDB_FN u32 AddMinRelu2(u32 a, u32 b, u32 c) {
	return Pk(Mx(Mn((int16_t)(Lo(a) + Lo(b)), Lo(c)), 0), Mx(Mn((int16_t)(Hi(a) + Hi(b)), Hi(c)), 0));
}
HDB_FN u32 Prmt(u32 x, u32 y, u32 s) {
	uint64_t v = ((uint64_t)y << 32) | x;
	u32 r = 0;
	for (int i = 0; i < 4; i++) r |= (u32)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
	return r;
}
HDB_FN u32 SumHalves(u32 a) { return (a & 0xffffu) + (a >> 16); }
HDB_FN u32 SumHalvesPlus(u32 a, u32 c) { return (a & 0xffffu) + (a >> 16) + c; }
HDB_FN u32 MaxU2(u32 a, u32 b) { return (Mx(a & 0xffffu, b & 0xffffu)) | ((u32)Mx(a >> 16, b >> 16) << 16); }
HDB_FN u32 MinU2(u32 a, u32 b) { return (Mn(a & 0xffffu, b & 0xffffu)) | ((u32)Mn(a >> 16, b >> 16) << 16); }
#else
HDB_FN u32 Add2(u32 a, u32 b) { return __vadd2(a, b); }                          /* VIADD.16x2 */
HDB_FN u32 Max2(u32 a, u32 b) { return __vmaxs2(a, b); }                         /* VIMNMX.S16x2 */
HDB_FN u32 Min2(u32 a, u32 b) { return __vmins2(a, b); }
HDB_FN u32 AddMax2(u32 a, u32 b, u32 c) { return __viaddmax_s16x2(a, b, c); }    /* VIADDMNMX.S16x2 */
HDB_FN u32 AddMin2(u32 a, u32 b, u32 c) { return __viaddmin_s16x2(a, b, c); }
HDB_FN u32 AddMinRelu2(u32 a, u32 b, u32 c) { return __viaddmin_s16x2_relu(a, b, c); }
HDB_FN u32 Prmt(u32 x, u32 y, u32 s) { return __byte_perm(x, y, s); }            /* PRMT */
HDB_FN u32 SumHalves(u32 a) { return __dp2a_lo(a, 0x0101u, 0u); }                /* IDP.2A: lo*1 + hi*1 */
HDB_FN u32 SumHalvesPlus(u32 a, u32 c) { return __dp2a_lo(a, 0x0101u, c); }      /* ... + c, same instruction */
HDB_FN u32 MaxU2(u32 a, u32 b) { return __vmaxu2(a, b); }                         /* VIMNMX.U16x2 */
HDB_FN u32 MinU2(u32 a, u32 b) { return __vminu2(a, b); }
#endif

/* both halves = v (v may be negative: two's complement per half) */
HDB_FN u32 Rep2(int v) { return ((u32)v & 0xffffu) * 0x10001u; }
/* x >> k of the packed halves.  HDB_SHR_IMAD: as the high word of x * 2^(32 - k), which issues on the FMA pipe (19 %
 * busy) instead of the ALU pipe (85 % busy) */
#if defined(HDB_SHR_IMAD) && HDB_SHR_IMAD && !defined(HDB_HOST_EMU)
template <int k> HDB_FN u32 Shr(u32 x) { return __umulhi(x, 1u << (32 - k)); }
#else
template <int k> HDB_FN u32 Shr(u32 x) { return x >> k; }
#endif

/* ------------------------------------------------------------------ */
/* per-frame filter parameters (cpu.h:1021-1033, 1064-1072)            */
/* ------------------------------------------------------------------ */
struct FilterParams {
	int beta, tc;
	/* derived once per frame, all as packed or biased constants */
	int dE_lim;      /* 2048 + beta            : dE test on the biased sum      cpu.h:1086-1087 */
	int dEpq_neg;    /* -(1024 + 3*beta/16)    : dEp / dEq as a sign bit        cpu.h:1245, 1249 */
	u32 k_b8;        /* (0x8000 + beta/8 - 1) x2 : "< beta/8" on both halves    cpu.h:1099-1105 */
	u32 k_tc52;      /* (0x8000 + 5*tc/2 - 1) x2 : "< 5*tc/2" on both halves    cpu.h:1109-1110 */
	/* strong filter: the +-2tc clip applied before the >> 3 / >> 2 (scaled by 8 / 4)   cpu.h:1191-1211 */
	u32 s8_neg, s8_pos;   /* -16tc ; 16tc + 7 */
	u32 s4_neg, s4_pos;   /* -8tc  ; 8tc + 3 */
	/* normal filter */
	u32 c2_neg, c2_pos;   /* -2tc, +2tc replicated                              cpu.h:1235 */
	u32 k_e_lo;      /* (2tc - 256) x2 */
	u32 k_tc10;      /* (10tc - 256) x2                                         cpu.h:1254 */
	u32 k_half;      /* (513 - 4tc) x2 : bias of the merged p1/q1 numerator */
	u32 h4_lo, h8_hi;     /* 4*(tc/2 - 256) ; 8*(tc/2) + 3 : the +-tc/2 clip before the >> 2   cpu.h:1236 */
	u32 h4_mid;           /* 4*(tc/2) + 3 : upper bound of the window of a side that is off */
	u32 h_neg, h_rep;     /* -(tc/2) ; (tc/2) */
	u32 c2_rep4;     /* (4tc) x2 */
	u32 ct_neg, ct_lo, ct_2; /* chroma, c = tc: -tc, (tc - 256), 2tc replicated cpu.h:1442 */
};

HDB_FN void MakeFilterParams(FilterParams &fp, int beta, int tc)
{
	fp.beta = beta; fp.tc = tc;
	fp.dE_lim = 2048 + beta;
	fp.dEpq_neg = -(1024 + 3 * beta / 16);
	fp.k_b8 = Rep2(0x8000 + beta / 8 - 1);
	fp.k_tc52 = Rep2(0x8000 + 5 * tc / 2 - 1);
	fp.s8_neg = Rep2(-16 * tc); fp.s8_pos = Rep2(16 * tc + 7);
	fp.s4_neg = Rep2(-8 * tc); fp.s4_pos = Rep2(8 * tc + 3);
	fp.c2_neg = Rep2(-2 * tc); fp.c2_pos = Rep2(2 * tc);
	fp.k_e_lo = Rep2(2 * tc - 256);
	fp.k_tc10 = Rep2(10 * tc - 256);
	fp.k_half = Rep2(513 - 4 * tc);
	fp.h4_lo = Rep2(4 * (tc / 2 - 256)); fp.h8_hi = Rep2(8 * (tc / 2) + 3);
	fp.h4_mid = Rep2(4 * (tc / 2) + 3);
	fp.h_neg = Rep2(-(tc / 2)); fp.h_rep = Rep2(tc / 2);
	fp.c2_rep4 = Rep2(4 * tc);
	fp.ct_neg = Rep2(-tc); fp.ct_lo = Rep2(tc - 256); fp.ct_2 = Rep2(2 * tc);
}

/* Taps of one line pair, unsigned 0..255 in each half. */
struct LumaPair { u32 p3, p2, p1, p0, q0, q1, q2, q3; };
struct ChromaPair { u32 p1, p0, q0, q1; };

/* 512 + (a2 - 2*a1 + a0) per half, in [2, 1022] */
HDB_FN u32 SecondDiffSigned(u32 a2, u32 a1, u32 a0) { return a2 + a0 + 0x02000200u - 2u * a1; }

/* 512 + |a2 - 2*a1 + a0| per half from the signed form (cpu.h:1086: the second-derivative terms) */
HDB_FN u32 SecondDiffAbs(u32 sd) { return Max2(sd, 0x04000400u - sd); }

/*
 * |a - b| per half for halves that hold a sample (0..255, upper byte zero): the byte-wise
 * absolute difference does it in one instruction (VABSDIFF4.U8; the upper bytes give |0 - 0|).
 */
#if defined(HDB_HOST_EMU)
HDB_FN u32 AbsDiff2(u32 a, u32 b) { return Max2(a, b) - Min2(a, b); }
#else
HDB_FN u32 AbsDiff2(u32 a, u32 b) { return __vabsdiffu4(a, b); }
#endif

/*
 * Strong filter on one line pair (cpu.h:1128-1213).  With x + delta_x equal to
 * the all-positive sums below (adding a multiple of 8 commutes with >> 3),
 *   x' = Clip255(x + Clip(-2tc, 2tc, delta_x)) = min(max(F_x, x - 2tc), x + 2tc),   F_x = T_x >> 3
 * and the 0..255 clip is the identity because 0 <= F_x <= 255.  The clip is applied to
 * T_x itself, against 8(x - 2tc) and 8(x + 2tc) + 7 (floor is monotone), so that the
 * shift comes last: what it drags from the upper half into bits 13..15 of the lower half
 * needs no mask, because every consumer of a filtered tap takes bytes 0 and 2 only
 * (StoreVertPair, JoinRow).  FILTERED TAPS THEREFORE CARRY GARBAGE ABOVE BIT 7 OF EACH HALF.
 */
HDB_FN void StrongPair(LumaPair &t, const FilterParams &fp)
{
	u32 s2 = t.p0 + t.q0 + 0x00020002u;       /* p0 + q0 + 2 */
	u32 t1p = t.p2 + t.p1 + s2;               /* p2 + p1 + p0 + q0 + 2          -> >>2  cpu.h:1160 */
	u32 t1q = t.q2 + t.q1 + s2;               /*                                         cpu.h:1179 */
	u32 m = t.p1 + t.q1 + s2;
	u32 t0p = t1p + m;                      /* p2 + 2p1 + 2p0 + 2q0 + q1 + 4  -> >>3  cpu.h:1154 */
	u32 t0q = t1q + m;                      /*                                         cpu.h:1173 */
	u32 t2p = 2u * (t.p3 + t.p2 + 0x00010001u) + t1p; /* 2p3 + 3p2 + p1 + p0 + q0 + 4 -> >>3  cpu.h:1166 */
	u32 t2q = 2u * (t.q3 + t.q2 + 0x00010001u) + t1q; /*                                       cpu.h:1185 */
	u32 a;
	a = 8u * t.p0; t.p0 = Shr<3>(AddMin2(a, fp.s8_pos, AddMax2(a, fp.s8_neg, t0p)));   /* cpu.h:1192-1211 */
	a = 4u * t.p1; t.p1 = Shr<2>(AddMin2(a, fp.s4_pos, AddMax2(a, fp.s4_neg, t1p)));
	a = 8u * t.p2; t.p2 = Shr<3>(AddMin2(a, fp.s8_pos, AddMax2(a, fp.s8_neg, t2p)));
	a = 8u * t.q0; t.q0 = Shr<3>(AddMin2(a, fp.s8_pos, AddMax2(a, fp.s8_neg, t0q)));
	a = 4u * t.q1; t.q1 = Shr<2>(AddMin2(a, fp.s4_pos, AddMax2(a, fp.s4_neg, t1q)));
	a = 8u * t.q2; t.q2 = Shr<3>(AddMin2(a, fp.s8_pos, AddMax2(a, fp.s8_neg, t2q)));
}

/*
 * Normal filter on one line pair (cpu.h:1215-1357).  [lo_p, hi_p] / [lo_q, hi_q] are the
 * clip windows of the p1 / q1 corrections, which carry the segment-wide side decisions
 * dEp / dEq (cpu.h:1243-1249, see LumaNormal); the |delta| < 10*tc test is per line
 * (cpu.h:1254), hence a per-half mask.  sd_p / sd_q
 * are 512 + (x2 - 2 x1 + x0) per half.  The p1 update uses
 *   (((p2 + p0 + 1) >> 1) - p1 + D) >> 1  ==  (p2 + p0 - 2 p1 + 2 D + 1) >> 2
 * (dropping the parity bit of an even numerator does not change floor(. / 4)), and its
 * +-tc/2 clip is applied before the >> 2 like the strong filter's; the bits the shift
 * drags across the halves are removed by the select that follows (its mask keeps 9 bits).
 */
HDB_FN void NormalPair(LumaPair &t, const FilterParams &fp, u32 lo_p, u32 hi_p, u32 lo_q, u32 hi_q, u32 sd_p, u32 sd_q)
{
	/* 4096 + 9(q0-p0) - 3(q1-p1) + 8, each half in [1044, 7164] */
	u32 num = 3u * (3u * (t.q0 - t.p0) + t.p1 - t.q1) + 0x10081008u;
	u32 e = Shr<4>(num) & 0x01ff01ffu;                  /* 256 + delta              cpu.h:1253 */
	u32 na = Min2(e, 0x02000200u - e);                 /* 256 - |delta| */
	u32 on = AddMinRelu2(na, fp.k_tc10, 0x00010001u);  /* 1 where |delta| < 10tc   cpu.h:1254 */
	u32 mask = on * 0x01ffu;                           /* 9 bits per half: every value selected below is < 512 */
	u32 d = AddMinRelu2(e, fp.k_e_lo, fp.c2_rep4);     /* Clip(delta) + 2tc in [0, 4tc]  cpu.h:1256 */
	d = (d & mask) | (fp.c2_pos & ~mask);              /* lines that are off: delta = 0 */
	u32 dn = fp.c2_rep4 - d;                         /* 2tc - Clip(delta) */
	/* Clip(((x2 + x0 - 2 x1 +- 2D + 1) >> 2), tc/2) + tc/2 in [0, 2(tc/2)]   cpu.h:1259-1265 */
	u32 hp = Shr<2>(AddMinRelu2(sd_p + fp.k_half + 2u * d, lo_p, hi_p));
	u32 hq = Shr<2>(AddMinRelu2(sd_q + fp.k_half + 2u * dn, lo_q, hi_q));
	u32 dp = (hp & mask) | (fp.h_rep & ~mask);         /* line off: no change */
	u32 dq = (hq & mask) | (fp.h_rep & ~mask);
	t.p1 = AddMinRelu2(t.p1 + dp, fp.h_neg, 0x00ff00ffu);      /* cpu.h:1272 */
	t.q1 = AddMinRelu2(t.q1 + dq, fp.h_neg, 0x00ff00ffu);      /* cpu.h:1275 */
	t.p0 = AddMinRelu2(t.p0 + d, fp.c2_neg, 0x00ff00ffu);      /* cpu.h:1268 */
	t.q0 = AddMinRelu2(t.q0 + dn, fp.c2_neg, 0x00ff00ffu);     /* cpu.h:1269 */
}

enum { kSegOff = 0, kSegNormal = 1, kSegStrong = 2 };

/*
 * The decisions of one luma edge segment (cpu.h:1359-1429 with 1074-1114), taken once per
 * segment from pair x = lines (0,3): off (dE false), strong, or normal together with the side
 * masks dEp / dEq and the signed second differences of x the normal filter reuses.
 */
HDB_FN int LumaDecide(const LumaPair &x, const FilterParams &fp, u32 &m_p, u32 &m_q, u32 &sdp, u32 &sdq)
{
	sdp = SecondDiffSigned(x.p2, x.p1, x.p0); sdq = SecondDiffSigned(x.q2, x.q1, x.q0);
	const u32 dP = SecondDiffAbs(sdp);                 /* 512 + dp0 | 512 + dp3 */
	const u32 dQ = SecondDiffAbs(sdq);
	const u32 s = dP + dQ;                             /* 1024 + dp+dq per line */
	m_p = m_q = 0u;
	if (!((int)SumHalves(s) < fp.dE_lim)) return kSegOff;           /* cpu.h:1086-1087 */
	/* strong decision, lines 0 and 3 at once (cpu.h:1099-1110) */
	const u32 c3 = AbsDiff2(x.p3, x.p0) + AbsDiff2(x.q0, x.q3);
	const u32 c4 = AbsDiff2(x.p0, x.q0);
	const u32 m1 = Max2(s - 0x04000400u, c3);
	const u32 ok = (fp.k_b8 - m1) & (fp.k_tc52 - c4) & 0x80008000u;
	if (ok == 0x80008000u) return kSegStrong;
	m_p = (u32)((int)SumHalvesPlus(dP, (u32)fp.dEpq_neg) >> 31);   /* all ones iff dp0 + dp3 < 3*beta/16  cpu.h:1245 */
	m_q = (u32)((int)SumHalvesPlus(dQ, (u32)fp.dEpq_neg) >> 31);   /* cpu.h:1249 */
	return kSegNormal;
}

HDB_FN void LumaNormal(LumaPair &x, LumaPair &y, const FilterParams &fp, u32 m_p, u32 m_q, u32 sdp, u32 sdq)
{
	/* The side decisions dEp / dEq hold for the whole segment, so they go into the window the p1 / q1
	 * correction is clipped to, once per segment, instead of into a select per line pair: a side that
	 * is off gets the window [tc/2, tc/2] -- whatever the numerator, the correction comes out as "no
	 * change" -- by an offset no numerator can stay below and an upper bound of 4*(tc/2) + 3. */
	const u32 lo_p = (fp.h4_lo & m_p) | (0x40004000u & ~m_p), hi_p = (fp.h8_hi & m_p) | (fp.h4_mid & ~m_p);
	const u32 lo_q = (fp.h4_lo & m_q) | (0x40004000u & ~m_q), hi_q = (fp.h8_hi & m_q) | (fp.h4_mid & ~m_q);
	NormalPair(x, fp, lo_p, hi_p, lo_q, hi_q, sdp, sdq);
	NormalPair(y, fp, lo_p, hi_p, lo_q, hi_q, SecondDiffSigned(y.p2, y.p1, y.p0), SecondDiffSigned(y.q2, y.q1, y.q0));
}

HDB_FN void LumaStrong(LumaPair &x, LumaPair &y, const FilterParams &fp)
{
	StrongPair(x, fp);
	StrongPair(y, fp);
}

/*
 * One luma edge segment (cpu.h:1359-1429).  x = lines (0,3), y = lines (1,2).
 * Returns which taps may have changed: kSegOff none, kSegNormal p1,p0,q0,q1,
 * kSegStrong p2..q2.
 */
HDB_FN int LumaSegment(LumaPair &x, LumaPair &y, const FilterParams &fp)
{
	u32 m_p, m_q, sdp, sdq;
	const int kind = LumaDecide(x, fp, m_p, m_q, sdp, sdq);
	if (kind == kSegStrong) LumaStrong(x, y, fp);
	else if (kind == kSegNormal) LumaNormal(x, y, fp, m_p, m_q, sdp, sdq);
	return kind;
}

/* One chroma line pair (cpu.h:1431-1488); only p0 and q0 change. */
HDB_FN void ChromaPairFilter(ChromaPair &t, const FilterParams &fp)
{
	/* 2048 + 4(p0-q0) + p1 - q1 + 4 in [777, 3327]; the q numerator is 8 - (p numerator - 4) + ... */
	u32 np = 4u * (t.p0 - t.q0) + t.p1 - t.q1 + 0x08040804u;        /* cpu.h:1453 */
	u32 nq = 0x10081008u - np;                                      /* cpu.h:1458 */
	u32 ep = Shr<3>(np) & 0x01ff01ffu;                               /* 256 + delta_p */
	u32 eq = Shr<3>(nq) & 0x01ff01ffu;                               /* 256 + delta_q */
	u32 dp = AddMinRelu2(ep, fp.ct_lo, fp.ct_2);                    /* Clip(delta_p) + tc   cpu.h:1464 */
	u32 dq = AddMinRelu2(eq, fp.ct_lo, fp.ct_2);                    /* Clip(delta_q) + tc   cpu.h:1469 */
	t.p0 = AddMinRelu2(t.p0 + dp, fp.ct_neg, 0x00ff00ffu);          /* cpu.h:1475 */
	t.q0 = AddMinRelu2(t.q0 + fp.ct_2 - dq, fp.ct_neg, 0x00ff00ffu);/* cpu.h:1476 */
}

/* ------------------------------------------------------------------ */
/* register block <-> line pairs                                       */
/* ------------------------------------------------------------------ */

/*
 * Vertical edge (cpu.h:159-284): lines are rows r0..r0+3, taps run along the
 * row: p_k = column 3-k (word lo, byte 3-k), q_k = column 4+k (word hi, byte k).
 * Pair (ra, rb): half lo = row ra, half hi = row rb.
 */
HDB_FN void LoadVertPair(LumaPair &t, u32 la, u32 lb, u32 ha, u32 hb)
{
	const u32 m8 = 0x00ff00ffu;
	u32 g1 = Prmt(la, lb, 0x5410);  /* a.b0 a.b1 b.b0 b.b1 = p3a p2a p3b p2b */
	u32 g2 = Prmt(la, lb, 0x7632);  /* p1a p0a p1b p0b */
	u32 h1 = Prmt(ha, hb, 0x5410);  /* q0a q1a q0b q1b */
	u32 h2 = Prmt(ha, hb, 0x7632);  /* q2a q3a q2b q3b */
	t.p3 = g1 & m8; t.p2 = Prmt(g1, 0u, 0x4341);
	t.p1 = g2 & m8; t.p0 = Prmt(g2, 0u, 0x4341);
	t.q0 = h1 & m8; t.q1 = Prmt(h1, 0u, 0x4341);
	t.q2 = h2 & m8; t.q3 = Prmt(h2, 0u, 0x4341);
}

HDB_FN void StoreVertPair(const LumaPair &t, int kind, u32 &la, u32 &lb, u32 &ha, u32 &hb)
{
	u32 g2 = Prmt(t.p1, t.p0, 0x6240);  /* p1a p0a p1b p0b */
	u32 h1 = Prmt(t.q0, t.q1, 0x6240);  /* q0a q1a q0b q1b */
	if (kind == kSegStrong) {
		u32 wa = Prmt(la, t.p2, 0x0040);    /* a.b0 p2a . . */
		u32 wb = Prmt(lb, t.p2, 0x0060);    /* b.b0 p2b . . */
		la = Prmt(wa, g2, 0x5410);
		lb = Prmt(wb, g2, 0x7610);
		u32 va = Prmt(ha, t.q2, 0x3400);    /* . . q2a a.b3 */
		u32 vb = Prmt(hb, t.q2, 0x3600);    /* . . q2b b.b3 */
		ha = Prmt(h1, va, 0x7610);
		hb = Prmt(h1, vb, 0x7632);
	} else {
		la = Prmt(la, g2, 0x5410);
		lb = Prmt(lb, g2, 0x7610);
		ha = Prmt(h1, ha, 0x7610);
		hb = Prmt(h1, hb, 0x7632);
	}
}

/*
 * Horizontal edge (cpu.h:287-446): lines are columns, the four bytes of a row
 * word are the four lines of one tap.  x = bytes (0,3), y = bytes (1,2).
 */
HDB_FN void SplitRow(u32 w, u32 &x, u32 &y) { x = Prmt(w, 0u, 0x4340); y = Prmt(w, 0u, 0x4241); }
HDB_FN u32 JoinRow(u32 x, u32 y) { return Prmt(x, y, 0x2640); }

/* chroma, vertical edge: p1 p0 = columns 2,3 ; q0 q1 = columns 4,5 */
HDB_FN void LoadVertPairChroma(ChromaPair &t, u32 la, u32 lb, u32 ha, u32 hb)
{
	const u32 m8 = 0x00ff00ffu;
	u32 g = Prmt(la, lb, 0x7632);  /* p1a p0a p1b p0b */
	u32 h = Prmt(ha, hb, 0x5410);  /* q0a q1a q0b q1b */
	t.p1 = g & m8; t.p0 = Prmt(g, 0u, 0x4341);
	t.q0 = h & m8; t.q1 = Prmt(h, 0u, 0x4341);
}

HDB_FN void StoreVertPairChroma(const ChromaPair &t, u32 &la, u32 &lb, u32 &ha, u32 &hb)
{
	la = Prmt(la, t.p0, 0x4210);
	lb = Prmt(lb, t.p0, 0x6210);
	ha = Prmt(ha, t.q0, 0x3214);
	hb = Prmt(hb, t.q0, 0x3216);
}

/*
 * NV12: the two chroma planes arrive interleaved, U0 V0 U1 V1 ...; the
 * reference only knows planar YV12 (cpu.h:73-82) and treats both chroma planes alike (cpu.h:453-992).
 * A row of an interleaved 8-sample block is four words; SplitUv picks one plane's eight samples out of
 * them as the (lo, hi) pair ChromaBlock works on, JoinUv interleaves four samples of each plane again.
 */
HDB_FN void SplitUv(u32 w0, u32 w1, u32 w2, u32 w3, bool v_plane, u32 &lo, u32 &hi)
{
	const u32 sel = v_plane ? 0x7531u : 0x6420u;
	lo = Prmt(w0, w1, sel); hi = Prmt(w2, w3, sel);
}
HDB_FN void JoinUv(u32 u, u32 v, u32 &w_first, u32 &w_second)
{
	w_first = Prmt(u, v, 0x5140u); w_second = Prmt(u, v, 0x7362u);
}

/* ------------------------------------------------------------------ */
/* tables and boundary-strength resolution                             */
/* ------------------------------------------------------------------ */

/*
 * beta / tc lookup, QP clamped to 51 (cpu.h:1021-1033, 1064-1072).  The tables
 * depend on QP only -- in the reference the boundary strength gates the edge
 * (cpu.h:164, 463) but does not offset tc.  ptxas places them in the constant bank.
 */
HDB_HD int BetaOf(unsigned qp)
{
	static const unsigned char beta_of_qp[52] = {
		0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
		16, 17, 18, 20, 22, 24, 26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52,
		54, 56, 58, 60, 62, 64 };
	return beta_of_qp[qp > 51 ? 51 : qp];
}

HDB_HD int TcOf(unsigned qp)
{
	static const unsigned char tc_of_qp[52] = {
		0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1,
		1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16,
		18, 20 };
	return tc_of_qp[qp > 51 ? 51 : qp];
}

/* Sizes of the boundary-strength arrays (cpu.h:86-87, 104-105), unsigned left-to-right. */
HDB_HD unsigned NumVertBs(unsigned w, unsigned h) { return (w / 8 + 1) * h / 8; }
HDB_HD unsigned NumHorBs(unsigned w, unsigned h) { return (h / 8 + 1) * w / 8; }

/*
 * The four boundary strengths of luma block (bx, by), packed
 * BS_ver1 | BS_ver2 << 8 | BS_hor1 << 16 | BS_hor2 << 24  (cpu.h:159-163, 223-227, 287-291, 368-372).
 * w8 = W/8; nbx = W/8 + 1, nby = H/8 + 1 are the padded block counts.
 */
HDB_FN u32 LumaBsOfBlock(int bx, int by, int w8, int nbx, int nby,
	const unsigned char *vert_bs, const unsigned char *hor_bs)
{
	u32 r = 0;
	if (by > 0) r |= (u32)vert_bs[(by - 1) * (w8 + 1) + bx];
	if (by < nby - 1) r |= (u32)vert_bs[by * (w8 + 1) + bx] << 8;
	if (bx > 0) r |= (u32)hor_bs[by * w8 + bx - 1] << 16;
	if (bx < nbx - 1) r |= (u32)hor_bs[by * w8 + bx] << 24;
	return r;
}

/*
 * Same for a chroma block.  Chroma strengths are never caller-set: they are the
 * constructor defaults (cpu.h:110-117: 2, except index % (cw/8+1) == 0 for the
 * vertical and index % (ch/8+1) == 0 for the horizontal array), read with the
 * LUMA block counts as limits (cpu.h:515, 645) and therefore past the end of
 * the arrays on the last chroma block row / column; reads past the end are 0
 * (zero-extension).
 */
HDB_FN u32 ChromaBsOfBlock(int bx, int by, unsigned cw, unsigned ch, int luma_nbx, int luma_nby)
{
	const unsigned cw8 = cw / 8, ch8 = ch / 8;
	const unsigned ncv = NumVertBs(cw, ch), nch = NumHorBs(cw, ch);
	u32 r = 0;
	if (by > 0) {
		unsigned i = (unsigned)(by - 1) * (cw8 + 1) + bx;
		if (i < ncv && i % (cw8 + 1) != 0) r |= 2u;
	}
	if (by < luma_nby - 1) {
		unsigned i = (unsigned)by * (cw8 + 1) + bx;
		if (i < ncv && i % (cw8 + 1) != 0) r |= 2u << 8;
	}
	if (bx > 0) {
		unsigned i = (unsigned)by * cw8 + bx - 1;
		if (i < nch && i % (ch8 + 1) != 0) r |= 2u << 16;
	}
	if (bx < luma_nbx - 1) {
		unsigned i = (unsigned)by * cw8 + bx;
		if (i < nch && i % (ch8 + 1) != 0) r |= 2u << 24;
	}
	return r;
}

/* ------------------------------------------------------------------ */
/* one padded 8x8 block                                                 */
/* ------------------------------------------------------------------ */

/*
 * Edge maps (no reference counterpart -- the reference has one QP per
 * frame, cpu.h:35-37).  An edge byte is  bs | qp << 2  in the reference's own array layouts; a luma
 * segment then takes beta / tc from its own entry.  The 6-bit qp indexes a 64-entry table of
 * beta | tc << 8 whose entries 52..63 repeat entry 51 (the clamp of cpu.h:1064-1072).  On the device
 * the table sits in one 128-byte line of global memory: the index differs from lane to lane, which a
 * constant-bank access would serialise.
 */
#define HDB_BT(b, t) (uint16_t)((b) | ((t) << 8))
#define HDB_BETA_TC_TABLE { \
	HDB_BT(0, 0), HDB_BT(0, 0), HDB_BT(0, 0), HDB_BT(0, 0), HDB_BT(0, 0), HDB_BT(0, 0), HDB_BT(0, 0), HDB_BT(0, 0), \
	HDB_BT(0, 0), HDB_BT(0, 0), HDB_BT(0, 0), HDB_BT(0, 0), HDB_BT(0, 0), HDB_BT(0, 0), HDB_BT(0, 0), HDB_BT(0, 0), \
	HDB_BT(6, 0), HDB_BT(7, 0), HDB_BT(8, 1), HDB_BT(9, 1), HDB_BT(10, 1), HDB_BT(11, 1), HDB_BT(12, 1), HDB_BT(13, 1), \
	HDB_BT(14, 1), HDB_BT(15, 1), HDB_BT(16, 1), HDB_BT(17, 2), HDB_BT(18, 2), HDB_BT(20, 2), HDB_BT(22, 2), HDB_BT(24, 3), \
	HDB_BT(26, 3), HDB_BT(28, 3), HDB_BT(30, 3), HDB_BT(32, 4), HDB_BT(34, 4), HDB_BT(36, 4), HDB_BT(38, 5), HDB_BT(40, 5), \
	HDB_BT(42, 6), HDB_BT(44, 6), HDB_BT(46, 7), HDB_BT(48, 8), HDB_BT(50, 9), HDB_BT(52, 10), HDB_BT(54, 11), HDB_BT(56, 13), \
	HDB_BT(58, 14), HDB_BT(60, 16), HDB_BT(62, 18), HDB_BT(64, 20), HDB_BT(64, 20), HDB_BT(64, 20), HDB_BT(64, 20), HDB_BT(64, 20), \
	HDB_BT(64, 20), HDB_BT(64, 20), HDB_BT(64, 20), HDB_BT(64, 20), HDB_BT(64, 20), HDB_BT(64, 20), HDB_BT(64, 20), HDB_BT(64, 20) }
#if defined(HDB_HOST_EMU)
static const uint16_t g_beta_tc_of_qp6[64] = HDB_BETA_TC_TABLE;
HDB_FN u32 BetaTcOfQp6(u32 qp6) { return g_beta_tc_of_qp6[qp6 & 63u]; }
#else
static __device__ __align__(128) const uint16_t g_beta_tc_of_qp6[64] = HDB_BETA_TC_TABLE;
HDB_FN u32 BetaTcOfQp6(u32 qp6) { return __ldg(g_beta_tc_of_qp6 + (qp6 & 63u)); }
#endif

HDB_HD u32 EdgeByte(u32 bs, u32 qp) { return (bs & 3u) | ((qp > 51u ? 51u : qp) << 2); }

/*
 * Luma block, cpu.h:158-446.  The four arguments are BS_ver1, BS_ver2, BS_hor1, BS_hor2 as
 * resolved by the caller (zero where the reference's guards skip the lookup): boundary
 * strengths when kEdgeQp is false -- every segment then uses the frame's parameters fp --
 * and edge bytes when it is true.
 */
template <bool kEdgeQp>
HDB_FN void LumaBlockT(u32 (&lo)[8], u32 (&hi)[8], u32 bs_ver1, u32 bs_ver2, u32 bs_hor1, u32 bs_hor2,
	const FilterParams &fp)
{
	FilterParams fpe;
	u32 bt_ver1 = 0, bt_ver2 = 0, bt_hor1 = 0, bt_hor2 = 0;
	if (kEdgeQp) {
		/* the four table reads first, so that they are in flight together */
		bt_ver1 = BetaTcOfQp6(bs_ver1 >> 2); bt_ver2 = BetaTcOfQp6(bs_ver2 >> 2);
		bt_hor1 = BetaTcOfQp6(bs_hor1 >> 2); bt_hor2 = BetaTcOfQp6(bs_hor2 >> 2);
		bs_ver1 &= 3u; bs_ver2 &= 3u; bs_hor1 &= 3u; bs_hor2 &= 3u;
	}
	/* upper (rows 0-3) and lower (rows 4-7) vertical edge segments */
#pragma unroll
	for (int s = 0; s < 2; s++) {
		const int r = 4 * s;
		if ((s == 0 ? bs_ver1 : bs_ver2) > 0) {                     /* cpu.h:164, 228 */
			if (kEdgeQp) { const u32 bt = s == 0 ? bt_ver1 : bt_ver2; MakeFilterParams(fpe, (int)(bt & 0xffu), (int)(bt >> 8)); }
			LumaPair x, y;
			LoadVertPair(x, lo[r], lo[r + 3], hi[r], hi[r + 3]);
			LoadVertPair(y, lo[r + 1], lo[r + 2], hi[r + 1], hi[r + 2]);
			int kind = LumaSegment(x, y, kEdgeQp ? fpe : fp);
			if (kind != kSegOff) {
				StoreVertPair(x, kind, lo[r], lo[r + 3], hi[r], hi[r + 3]);
				StoreVertPair(y, kind, lo[r + 1], lo[r + 2], hi[r + 1], hi[r + 2]);
			}
		}
	}
	/* left horizontal edge segment: columns 0-3, p rows 3..0, q rows 4..7 (cpu.h:287-365) */
	if (bs_hor1 > 0) {
		if (kEdgeQp) MakeFilterParams(fpe, (int)(bt_hor1 & 0xffu), (int)(bt_hor1 >> 8));
		LumaPair x, y;
		SplitRow(lo[0], x.p3, y.p3); SplitRow(lo[1], x.p2, y.p2);
		SplitRow(lo[2], x.p1, y.p1); SplitRow(lo[3], x.p0, y.p0);
		SplitRow(lo[4], x.q0, y.q0); SplitRow(lo[5], x.q1, y.q1);
		SplitRow(lo[6], x.q2, y.q2); SplitRow(lo[7], x.q3, y.q3);
		int kind = LumaSegment(x, y, kEdgeQp ? fpe : fp);
		if (kind != kSegOff) {
			lo[2] = JoinRow(x.p1, y.p1); lo[3] = JoinRow(x.p0, y.p0);
			lo[4] = JoinRow(x.q0, y.q0); lo[5] = JoinRow(x.q1, y.q1);
			if (kind == kSegStrong) { lo[1] = JoinRow(x.p2, y.p2); lo[6] = JoinRow(x.q2, y.q2); }
		}
	}
	/* right horizontal edge segment: p = columns 4-7 of rows 3..0, q = columns 0-3 of
	 * rows 4..7 -- the reference's pairing (cpu.h:383-445), reproduced as is */
	if (bs_hor2 > 0) {
		if (kEdgeQp) MakeFilterParams(fpe, (int)(bt_hor2 & 0xffu), (int)(bt_hor2 >> 8));
		LumaPair x, y;
		SplitRow(hi[0], x.p3, y.p3); SplitRow(hi[1], x.p2, y.p2);
		SplitRow(hi[2], x.p1, y.p1); SplitRow(hi[3], x.p0, y.p0);
		SplitRow(lo[4], x.q0, y.q0); SplitRow(lo[5], x.q1, y.q1);
		SplitRow(lo[6], x.q2, y.q2); SplitRow(lo[7], x.q3, y.q3);
		int kind = LumaSegment(x, y, kEdgeQp ? fpe : fp);
		if (kind != kSegOff) {
			hi[2] = JoinRow(x.p1, y.p1); hi[3] = JoinRow(x.p0, y.p0);
			lo[4] = JoinRow(x.q0, y.q0); lo[5] = JoinRow(x.q1, y.q1);
			if (kind == kSegStrong) { hi[1] = JoinRow(x.p2, y.p2); lo[6] = JoinRow(x.q2, y.q2); }
		}
	}
}

HDB_FN void LumaBlock(u32 (&lo)[8], u32 (&hi)[8], u32 bs_ver1, u32 bs_ver2, u32 bs_hor1, u32 bs_hor2,
	const FilterParams &fp)
{
	LumaBlockT<false>(lo, hi, bs_ver1, bs_ver2, bs_hor1, bs_hor2, fp);
}

/* Chroma block, cpu.h:453-721 (and 724-992): same geometry, taps k = 0,1, BS == 2 only. */
HDB_FN void ChromaBlock(u32 (&lo)[8], u32 (&hi)[8], u32 bs_ver1, u32 bs_ver2, u32 bs_hor1, u32 bs_hor2,
	const FilterParams &fp)
{
#pragma unroll
	for (int s = 0; s < 2; s++) {
		const int r = 4 * s;
		if ((s == 0 ? bs_ver1 : bs_ver2) == 2) {                    /* cpu.h:463, 519 */
			ChromaPair x, y;
			LoadVertPairChroma(x, lo[r], lo[r + 3], hi[r], hi[r + 3]);
			LoadVertPairChroma(y, lo[r + 1], lo[r + 2], hi[r + 1], hi[r + 2]);
			ChromaPairFilter(x, fp);
			ChromaPairFilter(y, fp);
			StoreVertPairChroma(x, lo[r], lo[r + 3], hi[r], hi[r + 3]);
			StoreVertPairChroma(y, lo[r + 1], lo[r + 2], hi[r + 1], hi[r + 2]);
		}
	}
	if (bs_hor1 == 2) {                                             /* cpu.h:572 */
		ChromaPair x, y;
		SplitRow(lo[2], x.p1, y.p1); SplitRow(lo[3], x.p0, y.p0);
		SplitRow(lo[4], x.q0, y.q0); SplitRow(lo[5], x.q1, y.q1);
		ChromaPairFilter(x, fp);
		ChromaPairFilter(y, fp);
		lo[3] = JoinRow(x.p0, y.p0); lo[4] = JoinRow(x.q0, y.q0);
	}
	if (bs_hor2 == 2) {                                             /* cpu.h:649 */
		ChromaPair x, y;
		SplitRow(hi[2], x.p1, y.p1); SplitRow(hi[3], x.p0, y.p0);
		SplitRow(lo[4], x.q0, y.q0); SplitRow(lo[5], x.q1, y.q1);
		ChromaPairFilter(x, fp);
		ChromaPairFilter(y, fp);
		hi[3] = JoinRow(x.p0, y.p0); lo[4] = JoinRow(x.q0, y.q0);
	}
}

/* ------------------------------------------------------------------ */
/* boundary-strength and edge-QP derivation                            */
/* ------------------------------------------------------------------ */

/*
 * What a decoder knows about one 8x8 luma block; binary layout of hdb_block_info
 * (include/hevc_deblock_b200.h), read as three 32-bit words:
 *   w0 = mv[0] (x | y << 16)   w1 = mv[1]   w2 = ref[0] | ref[1] << 8 | flags << 16 | qp << 24
 * The reference has nothing upstream of its filter -- strengths are constants (cpu.h:91-99) -- so
 * the rule is the published one, H.265 8.7.2.4: an edge that is neither a transform nor a
 * prediction edge is off; intra on either side -> 2; a transform edge with coded coefficients on
 * either side -> 1; different reference pictures or a different number of motion vectors -> 1; a
 * motion-vector difference of four quarter samples or more in either component -> 1; else 0.
 */
struct BlockWords { u32 mv0, mv1, meta; };
enum { kBlkIntra = 1, kBlkCbf = 2, kBlkTuLeft = 4, kBlkTuTop = 8, kBlkPuLeft = 16, kBlkPuTop = 32 };

/*
 * A record is normalised once, by the thread that fetches it, and then used by the three edges it borders:
 *   - the sign bits of the vectors' halves are flipped (signed order -> unsigned order), so that |a - b| per
 *     half is max - min with no borrow between the halves, exact over the whole 16-bit range;
 *   - the vector of an unused list (reference < 0) becomes one fixed value: wherever the two sides' lists pair up
 *     (see DeriveBs) unused lists then compare as "no difference" by themselves;
 *   - both reference bytes with every negative value turned into 0xff ("list unused").
 */
HDB_FN BlockWords NormalizeBlock(const BlockWords &raw)
{
	BlockWords n;
	const u32 r = raw.meta & 0xffffu;
	const u32 refs = r | (((r >> 7) & 0x0101u) * 0xffu);
	n.mv0 = (refs & 0x80u) ? 0u : (raw.mv0 ^ 0x80008000u);
	n.mv1 = (refs & 0x8000u) ? 0u : (raw.mv1 ^ 0x80008000u);
	n.meta = (raw.meta & 0xffff0000u) | refs;
	return n;
}

/* non-zero iff the two (normalised) vectors differ by 4 or more in a component: any bit above bit 1 of |a - b| */
HDB_FN u32 MvFarBits(u32 a, u32 b) { return (MaxU2(a, b) - MinU2(a, b)) & 0xfffcfffcu; }

/*
 * Straight-line form of the rule on normalised records (the kernel evaluates two edges per thread with
 * lane-varying data: branches would diverge on every one of them).  With pr / qr the two reference bytes of each
 * side, "same pictures, same vector count" is exactly  pr == qr  (lists pair up straight) or  pr == swap(qr)
 * (lists pair up crossed); both hold when all used references are one picture, and then -- as in the standard --
 * the motion differs only if it differs under both pairings.  Hence
 *     motion differs  =  (far under the straight pairing  or  lists do not pair up straight)
 *                    and (far under the crossed pairing   or  lists do not pair up crossed).
 */
HDB_FN u32 DeriveBs(const BlockWords &p, const BlockWords &q, bool tu_edge, bool pu_edge)
{
	const u32 flags = (p.meta | q.meta) >> 16;
	const u32 pr = p.meta & 0xffffu, qr = q.meta & 0xffffu;
	const u32 qr_swapped = Prmt(qr, 0u, 0x4401);
	const bool far_straight = (MvFarBits(p.mv0, q.mv0) | MvFarBits(p.mv1, q.mv1)) != 0u || pr != qr;
	const bool far_crossed = (MvFarBits(p.mv0, q.mv1) | MvFarBits(p.mv1, q.mv0)) != 0u || pr != qr_swapped;
	const u32 motion = (far_straight && far_crossed) ? 1u : 0u;
	const u32 inter = (tu_edge && (flags & kBlkCbf)) ? 1u : motion;
	const u32 bs = (flags & kBlkIntra) ? 2u : inter;
	return (tu_edge || pu_edge) ? bs : 0u;
}

/* edge byte between block p (left / above) and block q, both normalised: strength by the rule above,
 * QP = (QpP + QpQ + 1) >> 1 */
HDB_FN u32 DeriveEdge(const BlockWords &p, const BlockWords &q, bool tu_edge, bool pu_edge)
{
	return EdgeByte(DeriveBs(p, q, tu_edge, pu_edge), ((p.meta >> 24) + (q.meta >> 24) + 1u) >> 1);
}

} // namespace hdb
