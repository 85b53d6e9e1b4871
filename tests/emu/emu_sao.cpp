/*
 * emu_sao.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the product's sample-adaptive-offset arithmetic (gpu_video_codec_b200/csrc/sao_core.cuh) for the HOST
 * with the packed primitives emulated in C and runs it over whole planes the way the kernels do: a word of samples
 * at a time, one comparison S per sample, category = 2 + S(y, i) - S(y + vy, i - hx), rows and words outside the
 * picture read from the nearest ones inside, samples with a neighbour outside kept through a mask.
 * tests/test_sao_emu.py compares it with the oracle, which proves that formulation bit-exact without a GPU (what
 * the GPU tests then add is the lanes, the shuffles and the data movement).  Not part of, nor reachable from,
 * the product library.
 */
#define HDB_HOST_EMU 1
#include "../../gpu_video_codec_b200/csrc/sao_core.cuh"

#include <algorithm>
#include <cstring>

using namespace hdb;

struct Params { uint8_t type, position; int8_t offset[4]; };

static inline u32 Funnel(u32 lo, u32 hi, int shift) { return shift ? (lo >> shift) | (hi << (32 - shift)) : lo; }

/* words of `per` samples: 4 a word for bytes, 2 a word for 16-bit containers */
template <typename T>
struct Plane {
	const T *in; int w, h;
	static constexpr int per = 4 / (int)sizeof(T);
	int words() const { return w / per; }
	u32 Word(int y, int x) const      /* clamped like the kernels' loads */
	{
		y = std::min(std::max(y, 0), h - 1); x = std::min(std::max(x, 0), words() - 1);
		u32 v;
		memcpy(&v, in + (size_t)y * w + (size_t)x * per, 4);
		return v;
	}
};

/* S of word x of row y under (vy, hx): against the first neighbour, one row up (vy) and one sample to the left /
 * right (hx = -1 / 1) */
static u32 SignsAt(const Plane<uint8_t> &p, int y, int x, int vy, int hx)
{
	const u32 c = p.Word(y, x), m = p.Word(y - vy, x);
	u32 a = m;
	if (hx < 0) a = Funnel(p.Word(y - vy, x - 1), m, 24);
	if (hx > 0) a = Funnel(m, p.Word(y - vy, x + 1), 8);
	return Signs4(c, a);
}

static u32 SignsAt(const Plane<uint16_t> &p, int y, int x, int vy, int hx, u32 bias2, u32 unbias2)
{
	const u32 c = p.Word(y, x), m = p.Word(y - vy, x);
	u32 a = m;
	if (hx < 0) a = Prmt(p.Word(y - vy, x - 1), m, 0x5432);
	if (hx > 0) a = Prmt(m, p.Word(y - vy, x + 1), 0x5432);
	return Signs2(c, a, bias2, unbias2);
}

static const int kVy[4] = { 0, 1, 1, 1 }, kHx[4] = { -1, 0, -1, 1 };

extern "C" void emu_sao_plane8(const uint8_t *in, uint8_t *out, int w, int h, const Params *params, int ctb_cols, int ctb)
{
	const Plane<uint8_t> p{ in, w, h };
	const int words = p.words();
	for (int y = 0; y < h; y++)
		for (int x = 0; x < words; x++) {
			const Params &q = params[(size_t)(y / ctb) * ctb_cols + 4 * x / ctb];
			const u32 c = p.Word(y, x);
			u32 v = c, pool_lo, pool_hi;
			OffsetPool(q.type, (uint8_t)q.offset[0], (uint8_t)q.offset[1], (uint8_t)q.offset[2], (uint8_t)q.offset[3], pool_lo, pool_hi);
			if (q.type == 1) v = BandOffset4(c, (32u - q.position) * 0x01010101u, pool_lo, pool_hi);
			else if (q.type == 2) {
				const int cls = q.position & 3, vy = kVy[cls], hx = kHx[cls];
				const u32 s = SignsAt(p, y, x, vy, hx);
				/* S of the samples these are the first neighbours of: row y + vy, one sample to the right (hx = -1) / left */
				const u32 n = SignsAt(p, y + vy, x, vy, hx);
				u32 t = n;
				if (hx < 0) t = Funnel(n, SignsAt(p, y + vy, x + 1, vy, hx), 8);
				if (hx > 0) t = Funnel(SignsAt(p, y + vy, x - 1, vy, hx), n, 24);
				const u32 r = ApplyOffsets4(c, s + 0x02020202u - t, pool_lo, pool_hi);
				u32 keep = 0;
				if (hx != 0) keep = (x == 0 ? 0x000000ffu : 0u) | (x == words - 1 ? 0xff000000u : 0u);
				if (vy != 0 && (y == 0 || y == h - 1)) keep = 0xffffffffu;
				v = (c & keep) | (r & ~keep);
			}
			memcpy(out + (size_t)y * w + 4 * (size_t)x, &v, 4);
		}
}

extern "C" void emu_sao_plane16(const uint16_t *in, uint16_t *out, int w, int h, const Params *params, int ctb_cols, int ctb, int depth)
{
	const Plane<uint16_t> p{ in, w, h };
	const int words = p.words();
	const u32 one = 1u << depth;
	const u32 bias2 = one * 0x10001u, max2 = (one - 1u) * 0x10001u, unbias2 = ((0x10000u - (one - 1u)) & 0xffffu) * 0x10001u;
	for (int y = 0; y < h; y++)
		for (int x = 0; x < words; x++) {
			const Params &q = params[(size_t)(y / ctb) * ctb_cols + 2 * x / ctb];
			const u32 c = p.Word(y, x);
			u32 v = c, pool_lo, pool_hi;
			OffsetPool(q.type, (uint8_t)q.offset[0], (uint8_t)q.offset[1], (uint8_t)q.offset[2], (uint8_t)q.offset[3], pool_lo, pool_hi);
			if (q.type == 1) v = BandOffset2(c, (u32)depth - 5u, (32u - q.position) * 0x10001u, pool_lo, pool_hi, max2);
			else if (q.type == 2) {
				const int cls = q.position & 3, vy = kVy[cls], hx = kHx[cls];
				const u32 s = SignsAt(p, y, x, vy, hx, bias2, unbias2);
				const u32 n = SignsAt(p, y + vy, x, vy, hx, bias2, unbias2);
				u32 t = n;
				if (hx < 0) t = Prmt(n, SignsAt(p, y + vy, x + 1, vy, hx, bias2, unbias2), 0x5432);
				if (hx > 0) t = Prmt(SignsAt(p, y + vy, x - 1, vy, hx, bias2, unbias2), n, 0x5432);
				const u32 r = ApplyOffsets2(c, s + 0x00020002u - t, pool_lo, pool_hi, max2);
				u32 keep = 0;
				if (hx != 0) keep = (x == 0 ? 0x0000ffffu : 0u) | (x == words - 1 ? 0xffff0000u : 0u);
				if (vy != 0 && (y == 0 || y == h - 1)) keep = 0xffffffffu;
				v = (c & keep) | (r & ~keep);
			}
			memcpy(out + (size_t)y * w + 2 * (size_t)x, &v, 4);
		}
}
