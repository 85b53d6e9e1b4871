/*
 * sao_core.cuh -- the arithmetic of sample adaptive offset (H.265 8.7.3) on packed samples, shared by the kernels
 * (sao_kernels.cu) and, compiled for the host with the packed primitives emulated (HDB_HOST_EMU), by
 * tests/emu/emu_sao.cpp, which runs it over whole planes against the oracle without a GPU.
 *
 * With S(y, i) = 1 + sign(c(y, i) - first neighbour) -- the first neighbour being the left, upper, upper-left or
 * upper-right sample for classes 0 .. 3 -- the second comparison of a sample is the negated first comparison of the
 * sample it is the first neighbour of:  category(y, i) = 2 + S(y, i) - S(y + vy, i - hx),  (vy, hx) = (0, -1),
 * (1, 0), (1, -1), (1, 1).  One S per sample.
 */
#pragma once

#include "deblock_core.cuh"

namespace hdb {

/* prmt.b32 with the sign-replicating selector nibbles (nibble | 8: the selected byte's sign in all eight bits);
 * __byte_perm documents the low three bits of a nibble only */
#if defined(HDB_HOST_EMU)
HDB_FN u32 PrmtSx(u32 x, u32 y, u32 s)
{
	const uint64_t v = ((uint64_t)y << 32) | x;
	u32 r = 0;
	for (int i = 0; i < 4; i++) {
		const u32 nib = (s >> (4 * i)) & 0xfu;
		u32 byte = (u32)((v >> (8 * (nib & 7u))) & 0xffu);
		if (nib & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
		r |= byte << (8 * i);
	}
	return r;
}
#else
HDB_FN u32 PrmtSx(u32 x, u32 y, u32 s)
{
	u32 d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(y), "r"(s));
	return d;
}
#endif

/* ---- a byte a sample: four samples a word ---- */

/* bytes (b0, b1) / (b2, b3) of a word as a 16x2 pair */
HDB_FN u32 PairA(u32 w) { return Prmt(w, 0u, 0x4140); }
HDB_FN u32 PairB(u32 w) { return Prmt(w, 0u, 0x4342); }

/* four samples of word c with their indices (0..4, a byte each) in idx4: offset picked out of the 8-byte pool,
 * added, clipped to 0..255 */
HDB_FN u32 ApplyOffsets4(u32 c, u32 idx4, u32 pool_lo, u32 pool_hi)
{
	/* index | index << 4 per byte, bit 7 set: bytes 0, 1 are the permute selector of samples 0, 1 (low byte of the
	 * half picked by the index, high byte its sign), bytes 2, 3 that of samples 2, 3 (prmt reads 16 bits) */
	const u32 sel = idx4 * 0x11u + 0x80808080u;
	const u32 oA = PrmtSx(pool_lo, pool_hi, sel), oB = PrmtSx(pool_lo, pool_hi, sel >> 16);
	return Prmt(AddMinRelu2(PairA(c), oA, 0x00ff00ffu), AddMinRelu2(PairB(c), oB, 0x00ff00ffu), 0x6420);
}

/* 1 + sign(c - a) of four samples, a byte each: 256 + c - a - 255, at most 2, at least 0 */
HDB_FN u32 Signs4(u32 c, u32 a)
{
	const u32 sA = AddMinRelu2(PairA(c) + 0x01000100u - PairA(a), 0xff01ff01u, 0x00020002u);
	const u32 sB = AddMinRelu2(PairB(c) + 0x01000100u - PairB(a), 0xff01ff01u, 0x00020002u);
	return Prmt(sA, sB, 0x6420);
}

/* four samples of word c: band offset.  Band index minus band position, modulo 32, a byte each (band_add4 =
 * 32 - position in every byte); indices 4..31 (no offset) are folded onto 4..7, which the pool holds zeros for. */
HDB_FN u32 BandOffset4(u32 c, u32 band_add4, u32 pool_lo, u32 pool_hi)
{
	const u32 k = (((c >> 3) & 0x1f1f1f1fu) + band_add4) & 0x1f1f1f1fu;
	const u32 hi = k & 0x1c1c1c1cu;
	const u32 idx = (k & 0x03030303u) | ((hi | (hi >> 1) | (hi >> 2)) & 0x04040404u);
	return ApplyOffsets4(c, idx, pool_lo, pool_hi);
}

/* the pool of a block: edge offset -- categories 0, 1, (2: none), 3, 4; band offset -- bands 0 .. 3, (4 ..: none) */
HDB_FN void OffsetPool(u32 type, u32 o0, u32 o1, u32 o2, u32 o3, u32 &pool_lo, u32 &pool_hi)
{
	if (type == 2u) { pool_lo = o0 | (o1 << 8) | (o2 << 24); pool_hi = o3; }
	else { pool_lo = o0 | (o1 << 8) | (o2 << 16) | (o3 << 24); pool_hi = 0u; }
}

/* ---- 16-bit containers: two samples a word, nothing to unpack ---- */

/* offsets picked by the index halves (0..4) out of the pool, as signed halves */
HDB_FN u32 PickOffsets2(u32 idx2, u32 pool_lo, u32 pool_hi)
{
	const u32 t = idx2 * 0x11u;                                  /* nibble | nibble << 4 in bytes 0 and 2 */
	const u32 sel = Prmt(t, 0u, 0x4420) | 0x8080u;               /* nibbles: i0, i0 | 8, i1, i1 | 8 */
	return PrmtSx(pool_lo, pool_hi, sel);
}

HDB_FN u32 ApplyOffsets2(u32 c, u32 idx2, u32 pool_lo, u32 pool_hi, u32 max2)
{
	return AddMinRelu2(c, PickOffsets2(idx2, pool_lo, pool_hi), max2);
}

/* 1 + sign(c - a) per half; bias2 = 1 << depth, unbias2 = -((1 << depth) - 1), both in both halves */
HDB_FN u32 Signs2(u32 c, u32 a, u32 bias2, u32 unbias2) { return AddMinRelu2(c + bias2 - a, unbias2, 0x00020002u); }

/* band offset: band = sample >> (depth - 5); band_add2 = 32 - position in both halves */
HDB_FN u32 BandOffset2(u32 c, u32 band_shift, u32 band_add2, u32 pool_lo, u32 pool_hi, u32 max2)
{
	const u32 k = (((c >> band_shift) & 0x001f001fu) + band_add2) & 0x001f001fu;
	return ApplyOffsets2(c, MinU2(k, 0x00040004u), pool_lo, pool_hi, max2);
}

} // namespace hdb
