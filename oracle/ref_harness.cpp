/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Wraps the UNMODIFIED reference header
 *   /root/reference/hevc_deblocking_filter/hevc_deblocking_filter_cpu.h
 * (included by path through -I, never copied into this repo) so that the real
 * reference CPU filter can be called on in-memory frames.  Built only by
 * oracle/Makefile into oracle/_ref/libref_deblock.so.  Used to pin
 * oracle/deblock_oracle.c, to generate tests/golden/, and as the
 * "reference" CPU arm of bench.py.
 *
 * Determinism without editing the reference: the
 * header leaves plane padding uninitialised (cpu.h:57, 63-64) and its chroma
 * loops read past the end of the chroma boundary-strength arrays (cpu.h:517,
 * 647, 788, 918).  This TU replaces the global array allocator with a
 * zero-filling, over-allocating one, which defines both as zero.
 *
 * The class only reads and writes files (cpu.h:35-40, 995-1018), so frames go
 * through temporary files under /dev/shm (or $TMPDIR).  File I/O is outside
 * every timed region, as in the reference's own driver (main.cu:41-43).
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include <unistd.h>

static const size_t kSlack = 64 * 1024; /* > one BS row + 1 for any supported width */

void *operator new[](size_t n) {
	void *p = calloc(n + kSlack, 1);
	if (!p) throw std::bad_alloc();
	return p;
}
void operator delete[](void *p) noexcept { free(p); }
void operator delete[](void *p, size_t) noexcept { free(p); }
/* the header frees new[] memory through unique_ptr<unsigned char> => scalar delete */
void *operator new(size_t n) {
	void *p = malloc(n ? n : 1);
	if (!p) throw std::bad_alloc();
	return p;
}
void operator delete(void *p) noexcept { free(p); }
void operator delete(void *p, size_t) noexcept { free(p); }

#include "hevc_deblocking_filter_cpu.h"

namespace {

std::string TempPath(const char *tag, unsigned idx) {
	const char *dir = getenv("HDB_REF_TMPDIR");
	if (!dir) dir = access("/dev/shm", W_OK) == 0 ? "/dev/shm" : "/tmp";
	char buf[512];
	snprintf(buf, sizeof buf, "%s/hdb_ref_%d_%s_%u.yuv", dir, (int)getpid(), tag, idx);
	return buf;
}

bool WriteFile(const std::string &path, const uint8_t *data, size_t n) {
	FILE *f = fopen(path.c_str(), "wb");
	if (!f) return false;
	bool ok = fwrite(data, 1, n, f) == n;
	fclose(f);
	return ok;
}

bool ReadFile(const std::string &path, uint8_t *data, size_t n) {
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) return false;
	bool ok = fread(data, 1, n, f) == n;
	fclose(f);
	return ok;
}

} // namespace

extern "C" {

/*
 * One frame through ReadYuvFrame: ctor -> [SetBoundaryStrenght] -> DeblockingFilter(threads) -> Save.
 * Returns 0, or 1 when the reference throws (message copied to err if non-NULL),
 * or 2 on harness I/O failure.  *seconds (optional) receives the time of the
 * DeblockingFilter call alone.
 */
int ref_deblock_frame(const uint8_t *in, size_t in_bytes, uint8_t *out, unsigned width, unsigned height,
	unsigned qp, const uint8_t *vert_bs, unsigned num_vert_bs,
	const uint8_t *hor_bs, unsigned num_hor_bs, unsigned threads,
	double *seconds, char *err, unsigned err_len)
{
	static unsigned counter = 0;
	unsigned id = __sync_fetch_and_add(&counter, 1);
	size_t nbytes = (size_t)3 * width * height / 2;
	std::string fin = TempPath("in", id), fout = TempPath("out", id);
	/* the file is written with the caller's length: the reference checks it (cpu.h:43) */
	if (!WriteFile(fin, in, in_bytes)) return 2;
	int rc = 0;
	try {
		ReadYuvFrame frame(fin.c_str(), width, height, qp);
		if (vert_bs || hor_bs)
			frame.SetBoundaryStrenght(const_cast<uint8_t *>(vert_bs), num_vert_bs,
				const_cast<uint8_t *>(hor_bs), num_hor_bs);
		auto t0 = std::chrono::steady_clock::now();
		frame.DeblockingFilter(threads ? threads : 1);
		auto t1 = std::chrono::steady_clock::now();
		if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
		frame.Save(fout.c_str());
		if (!ReadFile(fout, out, nbytes)) rc = 2;
	} catch (const char *msg) {
		if (err && err_len) { strncpy(err, msg, err_len - 1); err[err_len - 1] = 0; }
		rc = 1;
	}
	unlink(fin.c_str());
	unlink(fout.c_str());
	return rc;
}

/*
 * n frames (frame f at in + f*3*w*h/2; per-frame qp; per-frame BS or NULL).
 * All ReadYuvFrame objects are constructed first; the timed region is then the
 * DeblockingFilter calls only:
 *   outer_threads == 1: frames one after another, each DeblockingFilter(inner_threads)
 *                       (the reference's own OpenMP path when inner_threads > 1);
 *   outer_threads  > 1: that many std::threads each taking every outer_threads-th
 *                       frame with DeblockingFilter(inner_threads).
 * *seconds receives the wall time of the timed region.
 */
int ref_deblock_batch(const uint8_t *in, uint8_t *out, unsigned n, unsigned width, unsigned height,
	const unsigned *qp, const uint8_t *vert_bs, const uint8_t *hor_bs,
	unsigned outer_threads, unsigned inner_threads, double *seconds)
{
	size_t fb = (size_t)3 * width * height / 2;
	unsigned nV = (width / 8 + 1) * height / 8, nH = (height / 8 + 1) * width / 8;
	if (outer_threads < 1) outer_threads = 1;
	if (inner_threads < 1) inner_threads = 1;
	std::vector<ReadYuvFrame *> frames(n, nullptr);
	int rc = 0;
	try {
		for (unsigned f = 0; f < n; f++) {
			std::string fin = TempPath("bin", f);
			if (!WriteFile(fin, in + f * fb, fb)) { rc = 2; break; }
			frames[f] = new ReadYuvFrame(fin.c_str(), width, height, qp[f]);
			unlink(fin.c_str());
			if (vert_bs && hor_bs)
				frames[f]->SetBoundaryStrenght(const_cast<uint8_t *>(vert_bs + (size_t)f * nV), nV,
					const_cast<uint8_t *>(hor_bs + (size_t)f * nH), nH);
		}
		if (rc == 0) {
			auto t0 = std::chrono::steady_clock::now();
			if (outer_threads == 1) {
				for (unsigned f = 0; f < n; f++) frames[f]->DeblockingFilter(inner_threads);
			} else {
				std::vector<std::thread> pool;
				for (unsigned t = 0; t < outer_threads; t++)
					pool.emplace_back([&, t]() {
						for (unsigned f = t; f < n; f += outer_threads)
							frames[f]->DeblockingFilter(inner_threads);
					});
				for (auto &th : pool) th.join();
			}
			auto t1 = std::chrono::steady_clock::now();
			if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
			for (unsigned f = 0; f < n && rc == 0; f++) {
				std::string fout = TempPath("bout", f);
				frames[f]->Save(fout.c_str());
				if (!ReadFile(fout, out + f * fb, fb)) rc = 2;
				unlink(fout.c_str());
			}
		}
	} catch (const char *) {
		rc = 1;
	}
	for (auto *p : frames) delete p;
	return rc;
}

} // extern "C"
