/*
 * emu_deblock.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the product's arithmetic header (gpu_video_codec_b200/csrc/deblock_core.cuh)
 * for the HOST with the packed-halfword primitives emulated in C, and drives it
 * over whole frames exactly like the generic CUDA kernel does (padded planes,
 * one 8x8 block at a time).  tests/test_core_emu.py compares it with the oracle,
 * which proves the register-level formulation bit-exact without a GPU.  It is
 * not part of, nor reachable from, the product library.
 */
#define HDB_HOST_EMU 1
#include "../../gpu_video_codec_b200/csrc/deblock_core.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

using namespace hdb;

static void FilterPlane(uint8_t *pl, int nbx, int nby, size_t pitch, bool chroma,
	unsigned w, unsigned h, int luma_nbx, int luma_nby,
	const uint8_t *vbs, const uint8_t *hbs, const FilterParams &fp, bool edges = false)
{
	for (int by = 0; by < nby; by++)
		for (int bx = 0; bx < nbx; bx++) {
			uint8_t *b = pl + (size_t)by * 8 * pitch + (size_t)bx * 8;
			u32 lo[8], hi[8];
			for (int r = 0; r < 8; r++) {
				memcpy(&lo[r], b + r * pitch, 4);
				memcpy(&hi[r], b + r * pitch + 4, 4);
			}
			u32 bs = chroma ? ChromaBsOfBlock(bx, by, w, h, luma_nbx, luma_nby)
				: LumaBsOfBlock(bx, by, (int)(w / 8), nbx, nby, vbs, hbs);
			if (chroma) ChromaBlock(lo, hi, bs & 0xff, (bs >> 8) & 0xff, (bs >> 16) & 0xff, bs >> 24, fp);
			else if (edges) LumaBlockT<true>(lo, hi, bs & 0xff, (bs >> 8) & 0xff, (bs >> 16) & 0xff, bs >> 24, fp);
			else LumaBlock(lo, hi, bs & 0xff, (bs >> 8) & 0xff, (bs >> 16) & 0xff, bs >> 24, fp);
			for (int r = 0; r < 8; r++) {
				memcpy(b + r * pitch, &lo[r], 4);
				memcpy(b + r * pitch + 4, &hi[r], 4);
			}
		}
}

extern "C" int emu_beta(unsigned qp) { return BetaOf(qp); }
extern "C" int emu_tc(unsigned qp) { return TcOf(qp); }

static int DeblockFrame(const uint8_t *in, uint8_t *out, unsigned w, unsigned h, unsigned qp,
	const uint8_t *vbs, const uint8_t *hbs, bool edges)
{
	if (w % 8 || h % 8) return 1;
	unsigned cw = w / 2, ch = h / 2;
	unsigned nw = w + 8, nh = h + 8, ncw = cw + 8, nch = ch + 8;
	int nbx = nw / 8, nby = nh / 8, cnbx = ncw / 8, cnby = nch / 8;
	FilterParams fp;
	MakeFilterParams(fp, BetaOf(qp), TcOf(qp));
	std::vector<uint8_t> dv, dh;
	if (!vbs) {
		dv.resize(NumVertBs(w, h)); dh.resize(NumHorBs(w, h));
		for (size_t i = 0; i < dv.size(); i++) dv[i] = (i % (w / 8 + 1) == 0) ? 0 : 2;
		for (size_t i = 0; i < dh.size(); i++) dh[i] = (i % (h / 8 + 1) == 0) ? 0 : 2;
		vbs = dv.data(); hbs = dh.data();
	}
	std::vector<uint8_t> Y((size_t)nw * nh, 0), C((size_t)ncw * nch);
	for (unsigned r = 0; r < h; r++) memcpy(&Y[(size_t)(r + 4) * nw + 4], in + (size_t)r * w, w);
	FilterPlane(Y.data(), nbx, nby, (size_t)nbx * 8, false, w, h, nbx, nby, vbs, hbs, fp, edges);
	for (unsigned r = 0; r < h; r++) memcpy(out + (size_t)r * w, &Y[(size_t)(r + 4) * nw + 4], w);
	for (int p = 0; p < 2; p++) {
		const uint8_t *src = in + (size_t)w * h + (size_t)p * cw * ch;
		uint8_t *dst = out + (size_t)w * h + (size_t)p * cw * ch;
		std::fill(C.begin(), C.end(), 0);
		for (unsigned r = 0; r < ch; r++) memcpy(&C[(size_t)(r + 4) * ncw + 4], src + (size_t)r * cw, cw);
		FilterPlane(C.data(), cnbx, cnby, (size_t)cnbx * 8, true, cw, ch, nbx, nby, nullptr, nullptr, fp);
		for (unsigned r = 0; r < ch; r++) memcpy(dst + (size_t)r * cw, &C[(size_t)(r + 4) * ncw + 4], cw);
	}
	return 0;
}

extern "C" int emu_deblock_frame(const uint8_t *in, uint8_t *out, unsigned w, unsigned h, unsigned qp,
	const uint8_t *vbs, const uint8_t *hbs)
{
	return DeblockFrame(in, out, w, h, qp, vbs, hbs, false);
}

/* edge bytes (bs | qp << 2) instead of strengths: LumaBlockT<true> */
extern "C" int emu_deblock_frame_edges(const uint8_t *in, uint8_t *out, unsigned w, unsigned h, unsigned qp,
	const uint8_t *vert_edges, const uint8_t *hor_edges)
{
	if (!vert_edges || !hor_edges) return 2;
	return DeblockFrame(in, out, w, h, qp, vert_edges, hor_edges, true);
}

/* DeriveEdge over one frame, indexed exactly as DeriveEdgesKernel does */
extern "C" void emu_derive_edges(const uint32_t *info, unsigned w, unsigned h, uint8_t *vert, uint8_t *hor)
{
	const unsigned w8 = w / 8, h8 = h / 8;
	auto blk = [&](unsigned r, unsigned x) { const uint32_t *p = info + 3 * ((size_t)r * w8 + x); return NormalizeBlock(BlockWords{ p[0], p[1], p[2] }); };
	for (unsigned r = 0; r <= h8; r++)
		for (unsigned x = 0; x <= w8; x++) {
			if (r < h8) {
				uint8_t e = 0;
				if (x > 0 && x < w8) { const BlockWords q = blk(r, x); e = (uint8_t)DeriveEdge(blk(r, x - 1), q, (q.meta >> 16) & kBlkTuLeft, (q.meta >> 16) & kBlkPuLeft); }
				vert[(size_t)r * (w8 + 1) + x] = e;
			}
			if (x < w8) {
				uint8_t e = 0;
				if (r > 0 && r < h8) { const BlockWords q = blk(r, x); e = (uint8_t)DeriveEdge(blk(r - 1, x), q, (q.meta >> 16) & kBlkTuTop, (q.meta >> 16) & kBlkPuTop); }
				hor[(size_t)r * w8 + x] = e;
			}
		}
}

/*
 * The interleaved-chroma block code of DeblockTiledKernel<., true>, lane pair by lane pair: the "U lane"
 * and the "V lane" each pick their samples out of the four words of a row (SplitUv), filter, swap halves
 * and write two words each (JoinUv).  uv: tight interleaved plane, w bytes x h/2 rows (w, h = luma size).
 */
extern "C" int emu_deblock_uv_plane(const uint8_t *uv_in, uint8_t *uv_out, unsigned w, unsigned h, unsigned qp)
{
	if (w % 8 || h % 8) return 1;
	const unsigned cw = w / 2, ch = h / 2;
	const int nbx = (w + 8) / 8, nby = (h + 8) / 8, cnbx = (cw + 8) / 8, cnby = (ch + 8) / 8;
	const size_t pitch = (size_t)cnbx * 16;     /* bytes per padded interleaved row */
	FilterParams fp;
	MakeFilterParams(fp, BetaOf(qp), TcOf(qp));
	std::vector<uint8_t> P(pitch * cnby * 8, 0);
	for (unsigned r = 0; r < ch; r++) memcpy(&P[(size_t)(r + 4) * pitch + 8], uv_in + (size_t)r * w, w);
	for (int by = 0; by < cnby; by++)
		for (int bx = 0; bx < cnbx; bx++) {
			uint8_t *b = &P[(size_t)by * 8 * pitch + (size_t)bx * 16];
			const u32 bs = ChromaBsOfBlock(bx, by, cw, ch, nbx, nby);
			u32 lo[2][8], hi[2][8];
			for (int v = 0; v < 2; v++) {
				for (int r = 0; r < 8; r++) {
					u32 wd[4];
					memcpy(wd, b + r * pitch, 16);
					SplitUv(wd[0], wd[1], wd[2], wd[3], v != 0, lo[v][r], hi[v][r]);
				}
				ChromaBlock(lo[v], hi[v], bs & 0xff, (bs >> 8) & 0xff, (bs >> 16) & 0xff, bs >> 24, fp);
			}
			for (int r = 0; r < 8; r++) {
				u32 wd[4];
				JoinUv(lo[0][r], lo[1][r], wd[0], wd[1]);   /* written by the U lane */
				JoinUv(hi[0][r], hi[1][r], wd[2], wd[3]);   /* written by the V lane */
				memcpy(b + r * pitch, wd, 16);
			}
		}
	for (unsigned r = 0; r < ch; r++) memcpy(uv_out + (size_t)r * w, &P[(size_t)(r + 4) * pitch + 8], w);
	return 0;
}

/* ---- samples deeper than 8 bits: deblock_core16.cuh driven the same way over 16-bit planes ---- */
#include "../../gpu_video_codec_b200/csrc/deblock_core16.cuh"

template <int kDepth>
static void FilterPlane16(uint16_t *pl, int nbx, int nby, size_t pitch, bool chroma, unsigned w, unsigned h,
	int luma_nbx, int luma_nby, const uint8_t *vbs, const uint8_t *hbs, const FilterParams16 &fp, bool edges)
{
	for (int by = 0; by < nby; by++)
		for (int bx = 0; bx < nbx; bx++) {
			uint16_t *b = pl + (size_t)by * 8 * pitch + (size_t)bx * 8;
			u32 wd[8][4];
			for (int r = 0; r < 8; r++) memcpy(wd[r], b + r * pitch, 16);
			u32 bs = chroma ? ChromaBsOfBlock(bx, by, w, h, luma_nbx, luma_nby)
				: LumaBsOfBlock(bx, by, (int)(w / 8), nbx, nby, vbs, hbs);
			if (chroma) ChromaBlock16<kDepth>(wd, bs & 0xff, (bs >> 8) & 0xff, (bs >> 16) & 0xff, bs >> 24, fp);
			else if (edges) LumaBlock16<kDepth, true>(wd, bs & 0xff, (bs >> 8) & 0xff, (bs >> 16) & 0xff, bs >> 24, fp);
			else LumaBlock16<kDepth, false>(wd, bs & 0xff, (bs >> 8) & 0xff, (bs >> 16) & 0xff, bs >> 24, fp);
			for (int r = 0; r < 8; r++) memcpy(b + r * pitch, wd[r], 16);
		}
}

template <int kDepth>
static int DeblockFrame16(const uint16_t *in, uint16_t *out, unsigned w, unsigned h, unsigned qp,
	const uint8_t *vbs, const uint8_t *hbs, bool edges)
{
	if (w % 8 || h % 8) return 1;
	unsigned cw = w / 2, ch = h / 2;
	unsigned nw = w + 8, nh = h + 8, ncw = cw + 8, nch = ch + 8;
	int nbx = nw / 8, nby = nh / 8, cnbx = ncw / 8, cnby = nch / 8;
	FilterParams16 fp;
	MakeFilterParams16<kDepth>(fp, BetaOf(qp), TcOf(qp));
	std::vector<uint8_t> dv, dh;
	if (!vbs) {
		dv.resize(NumVertBs(w, h)); dh.resize(NumHorBs(w, h));
		for (size_t i = 0; i < dv.size(); i++) dv[i] = (i % (w / 8 + 1) == 0) ? 0 : 2;
		for (size_t i = 0; i < dh.size(); i++) dh[i] = (i % (h / 8 + 1) == 0) ? 0 : 2;
		vbs = dv.data(); hbs = dh.data();
	}
	std::vector<uint16_t> Y((size_t)nw * nh, 0), C((size_t)ncw * nch);
	for (unsigned r = 0; r < h; r++) memcpy(&Y[(size_t)(r + 4) * nw + 4], in + (size_t)r * w, 2 * w);
	FilterPlane16<kDepth>(Y.data(), nbx, nby, (size_t)nbx * 8, false, w, h, nbx, nby, vbs, hbs, fp, edges);
	for (unsigned r = 0; r < h; r++) memcpy(out + (size_t)r * w, &Y[(size_t)(r + 4) * nw + 4], 2 * w);
	for (int p = 0; p < 2; p++) {
		const uint16_t *src = in + (size_t)w * h + (size_t)p * cw * ch;
		uint16_t *dst = out + (size_t)w * h + (size_t)p * cw * ch;
		std::fill(C.begin(), C.end(), 0);
		for (unsigned r = 0; r < ch; r++) memcpy(&C[(size_t)(r + 4) * ncw + 4], src + (size_t)r * cw, 2 * cw);
		FilterPlane16<kDepth>(C.data(), cnbx, cnby, (size_t)cnbx * 8, true, cw, ch, nbx, nby, nullptr, nullptr, fp, false);
		for (unsigned r = 0; r < ch; r++) memcpy(dst + (size_t)r * cw, &C[(size_t)(r + 4) * ncw + 4], 2 * cw);
	}
	return 0;
}

extern "C" int emu_deblock_frame16(const uint16_t *in, uint16_t *out, unsigned w, unsigned h, unsigned qp,
	const uint8_t *vbs, const uint8_t *hbs, unsigned depth, int edges)
{
	if (edges && (!vbs || !hbs)) return 2;
	switch (depth) {
	case 8: return DeblockFrame16<8>(in, out, w, h, qp, vbs, hbs, edges != 0);
	case 9: return DeblockFrame16<9>(in, out, w, h, qp, vbs, hbs, edges != 0);
	case 10: return DeblockFrame16<10>(in, out, w, h, qp, vbs, hbs, edges != 0);
	}
	return 3;
}
