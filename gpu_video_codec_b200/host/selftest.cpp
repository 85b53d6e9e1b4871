/*
 * selftest.cpp -- the C++ mirror (hevc_deblocking_filter_b200.hpp) against the C ABI it sits on:
 * the reference's call sequence  ctor -> SetBoundaryStrenght -> DeblockingFilter -> Save  (cpu.h:35, 120,
 * 134, 995) must give the bytes hdb_deblock_frame gives for the same inputs, SetEdgeMaps those of
 * hdb_deblock_frame_edges, and the reference's three thrown messages must come out as const char *.
 *     hevc_deblock_b200_selftest frame.yuv width height Qp out_plain.yuv out_edges.yuv
 * Exit code 0 = all checks passed.  (tests/test_gpu_nv12.py runs it and compares the files with the oracle.)
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "hevc_deblocking_filter_b200.hpp"

static int Fail(const char *what) { std::fprintf(stderr, "selftest: %s\n", what); return 1; }

int main(int argc, char **argv)
{
	if (argc < 7) return Fail("usage: selftest frame.yuv width height Qp out_plain.yuv out_edges.yuv");
	const char *in = argv[1];
	const unsigned w = (unsigned)std::strtoul(argv[2], nullptr, 10), h = (unsigned)std::strtoul(argv[3], nullptr, 10);
	const unsigned qp = (unsigned)std::strtoul(argv[4], nullptr, 10);
	const unsigned nv = hdb_num_vert_bs(w, h), nh = hdb_num_hor_bs(w, h);
	const size_t bytes = (size_t)3 * w * h / 2;
	/* deterministic strengths and edge maps */
	std::vector<unsigned char> v(nv), hh(nh), ve(nv), he(nh);
	unsigned s = 12345u;
	auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
	for (auto &x : v) x = (unsigned char)(rnd() % 3);
	for (auto &x : hh) x = (unsigned char)(rnd() % 3);
	for (auto &x : ve) x = HDB_EDGE(rnd() % 3, 18 + rnd() % 34);
	for (auto &x : he) x = HDB_EDGE(rnd() % 3, 18 + rnd() % 34);
	try {
		b200::ReadYuvFrame a(in, w, h, qp);
		std::vector<unsigned char> src(a.Frame(), a.Frame() + bytes), want(bytes);
		a.SetBoundaryStrenght(v.data(), nv, hh.data(), nh);
		a.DeblockingFilter(4);
		a.Save(argv[5]);
		if (hdb_deblock_frame(a.Context(), src.data(), want.data(), w, h, qp, v.data(), nv, hh.data(), nh, nullptr) != HDB_OK)
			return Fail(hdb_last_error(a.Context()));
		if (std::memcmp(a.Frame(), want.data(), bytes) != 0) return Fail("ReadYuvFrame path differs from hdb_deblock_frame");

		b200::ReadYuvFrame e(in, w, h, qp);
		e.SetEdgeMaps(ve.data(), nv, he.data(), nh);
		e.DeblockingFilter();
		e.Save(argv[6]);
		if (hdb_deblock_frame_edges(e.Context(), src.data(), want.data(), w, h, qp, ve.data(), nv, he.data(), nh, nullptr) != HDB_OK)
			return Fail(hdb_last_error(e.Context()));
		if (std::memcmp(e.Frame(), want.data(), bytes) != 0) return Fail("SetEdgeMaps path differs from hdb_deblock_frame_edges");

		bool thrown = false;
		try { a.SetBoundaryStrenght(v.data(), nv - 1, hh.data(), nh); } catch (const char *m) { thrown = std::strstr(m, "boundary strenght") != nullptr; }   /* cpu.h:123 */
		if (!thrown) return Fail("size mismatch did not throw the reference's message");
		thrown = false;
		try { b200::ReadYuvFrame bad(in, w, h + 8, qp); } catch (const char *m) { thrown = std::strcmp(m, "Incorrect file size") == 0; }                     /* cpu.h:44 */
		if (!thrown) return Fail("wrong size did not throw \"Incorrect file size\"");
	} catch (const char *m) {
		return Fail(m);
	}
	/* dump the inputs the checks used, so that the caller can run its own checker on them */
	std::string side = std::string(argv[5]) + ".maps";
	if (FILE *f = std::fopen(side.c_str(), "wb")) {
		std::fwrite(v.data(), 1, nv, f); std::fwrite(hh.data(), 1, nh, f); std::fwrite(ve.data(), 1, nv, f); std::fwrite(he.data(), 1, nh, f);
		std::fclose(f);
	}
	std::printf("selftest ok\n");
	return 0;
}
