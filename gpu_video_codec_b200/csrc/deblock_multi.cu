/*
 * deblock_multi.cu -- the multi-GPU partitioner (several B200s of one box, one
 * process).  The reference has no analogue: it is single-GPU on device 0
 * (main.cu:93-95).  Both partitions rest on facts about the reference filter
 * (SURVEY.md section 8(e)): frames are independent (all state is per object,
 * cpu.h:35-137) and, inside a frame, padded 8x8 blocks are independent
 * (cpu.h:146-447), so
 *   - a batch is sharded by contiguous frame ranges, one pipeline per GPU, no
 *     collective and no exchange at all;
 *   - a single frame is cut at CTU-row boundaries (multiples of 64 luma lines);
 *     the block row that straddles a cut needs 4 input rows from the other side,
 *     which come from the neighbouring GPU's memory by peer-to-peer copy, and
 *     both GPUs compute that block row but each stores only its own rows.
 * No NCCL: there is nothing to reduce or gather on the device side.
 */
#include "deblock_api.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

constexpr int kStripSlots = 3;        /* frames of a strip stream in flight per device */

struct StripSlot {                    /* one frame's strip on one device */
	unsigned char *in = nullptr, *out = nullptr;      /* the strip's own rows: Y, U, V planes back to back, tight */
	size_t cap = 0;
	unsigned char *vbs = nullptr, *hbs = nullptr;     /* frame-sized arrays: the kernel's indices are frame-global */
	size_t vbs_cap = 0, hbs_cap = 0;
	cudaEvent_t up_done = nullptr, k_done = nullptr, d2h_done = nullptr, k0 = nullptr, k1 = nullptr;
	bool timed = false;                                /* k0 / k1 hold a kernel interval not yet read */
};

struct StripDev {
	cudaStream_t s_h2d = nullptr, s_k = nullptr, s_d2h = nullptr;
	cudaEvent_t ready = nullptr;                       /* device-resident entry point: "the inputs on this device are ready" */
	StripSlot slot[kStripSlots];
	/* host-side progress: sequence number of the last frame whose up_done / k_done event has been recorded.
	 * A wait on a neighbour's event may only be enqueued once that neighbour has recorded it for this frame. */
	std::atomic<long long> up_rec{-1}, k_rec{-1};
	double kernel_ms = 0;
};

struct hdb_multi {
	std::vector<hdb_context *> ctx;
	std::vector<StripDev *> dev;                       /* one per device, for the strip path */
	long long strip_seq = 0;                           /* frames sent through the strip stream so far */
	std::atomic<int> strip_abort{0};
	std::mutex mu;
	std::string err;
};

namespace hdb {

static int MFail(hdb_multi *m, int status, const std::string &msg)
{
	if (m) m->err = msg;
	return status;
}

struct Strip {
	int y0, y1;            /* luma rows owned: [y0, y1) */
};

/* CTU rows (64 luma lines) dealt out as evenly as possible, earlier strips first */
static std::vector<Strip> MakeStrips(int height, int parts)
{
	const int nctu = (height + 63) / 64;
	parts = std::max(1, std::min(parts, nctu));
	std::vector<Strip> out;
	int c = 0;
	for (int d = 0; d < parts; d++) {
		const int take = nctu / parts + (d < nctu % parts ? 1 : 0);
		Strip s;
		s.y0 = c * 64;
		s.y1 = std::min((c + take) * 64, height);
		out.push_back(s);
		c += take;
	}
	return out;
}

} // namespace hdb

using namespace hdb;

extern "C" {

int hdb_multi_create(hdb_multi **out, const int *devices, int n_devices)
{
	if (!out) return HDB_ERR_INVALID_ARG;
	*out = nullptr;
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return Fail(nullptr, HDB_ERR_CUDA, "no CUDA device");
	if (n_devices <= 0) n_devices = count;
	hdb_multi *m = new (std::nothrow) hdb_multi();
	if (!m) return HDB_ERR_NO_MEMORY;
	for (int i = 0; i < n_devices; i++) {
		const int dev = devices ? devices[i] : i;
		hdb_context *c = nullptr;
		int rc = hdb_create(&c, dev);
		if (rc) { hdb_multi_destroy(m); return rc; }
		m->ctx.push_back(c);
	}
	const int n = (int)m->ctx.size();
	for (int i = 0; i < n; i++) {
		DeviceGuard g(m->ctx[i]->device);
		StripDev *sd = new (std::nothrow) StripDev();
		if (!sd) { hdb_multi_destroy(m); return HDB_ERR_NO_MEMORY; }
		m->dev.push_back(sd);
		bool ok = cudaStreamCreateWithFlags(&sd->s_h2d, cudaStreamNonBlocking) == cudaSuccess
			&& cudaStreamCreateWithFlags(&sd->s_k, cudaStreamNonBlocking) == cudaSuccess
			&& cudaStreamCreateWithFlags(&sd->s_d2h, cudaStreamNonBlocking) == cudaSuccess
			&& cudaEventCreateWithFlags(&sd->ready, cudaEventDisableTiming) == cudaSuccess;
		for (auto &sl : sd->slot)
			ok = ok && cudaEventCreateWithFlags(&sl.up_done, cudaEventDisableTiming) == cudaSuccess
				&& cudaEventCreateWithFlags(&sl.k_done, cudaEventDisableTiming) == cudaSuccess
				&& cudaEventCreateWithFlags(&sl.d2h_done, cudaEventDisableTiming) == cudaSuccess
				&& cudaEventCreate(&sl.k0) == cudaSuccess && cudaEventCreate(&sl.k1) == cudaSuccess;
		if (!ok) {
			hdb_multi_destroy(m);
			return Fail(nullptr, HDB_ERR_CUDA, "multi-GPU stream/event creation failed");
		}
		/* direct peer access to the other devices (NVLink / NVSwitch): the strip kernel reads its halo rows
		 * straight out of the neighbours' buffers */
		for (int j = 0; j < n; j++) {
			if (m->ctx[j]->device == m->ctx[i]->device) continue;
			int can = 0;
			if (cudaDeviceCanAccessPeer(&can, m->ctx[i]->device, m->ctx[j]->device) == cudaSuccess && can) {
				cudaError_t e = cudaDeviceEnablePeerAccess(m->ctx[j]->device, 0);
				if (e != cudaSuccess) cudaGetLastError();   /* already enabled is fine */
			} else {
				hdb_multi_destroy(m);
				return Fail(nullptr, HDB_ERR_CUDA, "the devices of a multi-GPU context must be peers of each other");
			}
		}
	}
	*out = m;
	return HDB_OK;
}

void hdb_multi_destroy(hdb_multi *m)
{
	if (!m) return;
	for (size_t i = 0; i < m->ctx.size(); i++) {
		DeviceGuard g(m->ctx[i]->device);
		cudaDeviceSynchronize();
		if (i >= m->dev.size()) continue;
		StripDev *sd = m->dev[i];
		for (auto &sl : sd->slot) {
			cudaFree(sl.in); cudaFree(sl.out); cudaFree(sl.vbs); cudaFree(sl.hbs);
			for (cudaEvent_t ev : { sl.up_done, sl.k_done, sl.d2h_done, sl.k0, sl.k1 }) if (ev) cudaEventDestroy(ev);
		}
		if (sd->ready) cudaEventDestroy(sd->ready);
		for (cudaStream_t st : { sd->s_h2d, sd->s_k, sd->s_d2h }) if (st) cudaStreamDestroy(st);
		delete sd;
	}
	for (auto *c : m->ctx) hdb_destroy(c);
	delete m;
}

int hdb_multi_device_count(const hdb_multi *m) { return m ? (int)m->ctx.size() : 0; }
const char *hdb_multi_last_error(const hdb_multi *m) { return m ? m->err.c_str() : ""; }

int hdb_multi_deblock_batch(hdb_multi *m, const unsigned char *frames_in, unsigned char *frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *qp,
	const unsigned char *vert_bs, const unsigned char *hor_bs, hdb_timing *timing)
{
	return hdb_multi_deblock_batch_ex(m, frames_in, frames_out, n_frames, width, height, qp, vert_bs, hor_bs, 0, timing);
}

int hdb_multi_deblock_batch_ex(hdb_multi *m, const unsigned char *frames_in, unsigned char *frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *qp,
	const unsigned char *vert_bs, const unsigned char *hor_bs, unsigned flags, hdb_timing *timing)
{
	if (!m || m->ctx.empty()) return HDB_ERR_INVALID_ARG;
	if (flags & ~(HDB_FLAG_EDGE_MAPS | HDB_FLAG_NV12 | HDB_FLAG_SAMPLES16 | 0xff00u)) return MFail(m, HDB_ERR_INVALID_ARG, "unknown flag");
	if (!frames_in || !frames_out || !qp || n_frames == 0) return MFail(m, HDB_ERR_INVALID_ARG, "NULL buffer or empty batch");
	const int g = (int)std::min<size_t>(m->ctx.size(), n_frames);
	const size_t fb = (size_t)width * height * 3 / 2 * ((flags & HDB_FLAG_SAMPLES16) ? 2 : 1);
	const unsigned nv = hdb_num_vert_bs(width, height), nh = hdb_num_hor_bs(width, height);
	std::vector<int> rcs(g, HDB_OK);
	std::vector<hdb_timing> tms(g);
	std::vector<std::thread> pool;
	const auto t0 = std::chrono::steady_clock::now();
	/* contiguous frame ranges, remainder to the first devices */
	unsigned f0 = 0;
	for (int d = 0; d < g; d++) {
		const unsigned nf = n_frames / g + ((unsigned)d < n_frames % g ? 1 : 0);
		const unsigned first = f0;
		pool.emplace_back([=, &rcs, &tms]() {
			rcs[d] = BatchHost(m->ctx[d], frames_in + first * fb, frames_out + first * fb, nf, width, height,
				qp + first, vert_bs ? vert_bs + (size_t)first * nv : nullptr,
				hor_bs ? hor_bs + (size_t)first * nh : nullptr, &tms[d], flags);
		});
		f0 += nf;
	}
	for (auto &t : pool) t.join();
	const auto t1 = std::chrono::steady_clock::now();
	for (int d = 0; d < g; d++)
		if (rcs[d]) return MFail(m, rcs[d], std::string("device ") + std::to_string(m->ctx[d]->device) + ": " + m->ctx[d]->err);
	if (timing) {
		memset(timing, 0, sizeof *timing);
		for (int d = 0; d < g; d++) {
			timing->kernel_ms = std::max(timing->kernel_ms, tms[d].kernel_ms);   /* devices run concurrently */
			timing->copy_ms = std::max(timing->copy_ms, tms[d].copy_ms);
			timing->h2d_bytes += tms[d].h2d_bytes; timing->d2h_bytes += tms[d].d2h_bytes;
			timing->kernel_launches += tms[d].kernel_launches;
			timing->path = tms[d].path;
		}
		timing->total_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
	}
	return HDB_OK;
}


} // extern "C"

#define M_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
	for (auto *c__ : m->ctx) { DeviceGuard g__(c__->device); cudaDeviceSynchronize(); } \
	return MFail(m, HDB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } } while (0)

/* ------------------------------------------------------------------ */
/* one frame split into CTU-row strips: a pipelined stream of frames   */
/* ------------------------------------------------------------------ */
/*
 * Per device and frame (slot = frame number mod kStripSlots):
 *   A  upload of the strip's own rows (Y, U, V) and of its boundary-strength rows       stream s_h2d
 *   B  the filter kernel, out of place; the 4 input rows on the far side of each cut are read by the kernel
 *      itself from the neighbouring devices' input buffers (peer loads over NVLink)     stream s_k
 *   C  download of the strip's own rows                                                 stream s_d2h
 * so that frame f + 1 is uploaded while frame f is filtered and frame f - 1 is downloaded.  Dependencies
 * (all CUDA events, no host synchronisation inside the stream):
 *   A(f) after  B(f - K) of this device and of both neighbours (they read this input buffer)
 *   B(f) after  A(f) of this device and of both neighbours, and C(f - K) of this device (the output buffer)
 *   C(f) after  B(f)
 * One host thread per device enqueues its device's work; a wait on a neighbour's event is enqueued only
 * after that neighbour has recorded it for the same frame (up_rec / k_rec, host-side counters).
 */
namespace hdb {

struct StripGeom {
	int y0, y1;          /* luma rows owned */
	int lrows, crows;    /* rows held: luma, each chroma plane */
	size_t u_off, v_off, bytes;
};

struct StripJob {
	hdb_multi *m;
	int g;                                   /* strips = devices in use */
	int W, H;
	std::vector<Strip> strips;
	std::vector<StripGeom> geom;
	const unsigned char *in; unsigned char *out;     /* host frames, frame_bytes apart */
	size_t frame_bytes;
	const unsigned *qp;
	const unsigned char *vbs, *hbs;          /* host, per frame, or NULL */
	unsigned nv, nh;
	bool edge_qp, timed;
	long long base_seq;
	std::vector<int> rc;
};

static int StripFail(StripJob &job, int d, cudaError_t e, const char *what)
{
	job.m->strip_abort.store(1);
	return FailCuda(job.m->ctx[d], e, what);
}

#define S_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return StripFail(job, d, e__, #call); } while (0)

/* spin until `counter` has reached `want` (another device's host thread is a few API calls away), or abort */
static bool WaitRecorded(const std::atomic<long long> &counter, long long want, const std::atomic<int> &abort_flag)
{
	while (counter.load(std::memory_order_acquire) < want) {
		if (abort_flag.load(std::memory_order_relaxed)) return false;
		std::this_thread::yield();
	}
	return true;
}

/* the view of strip d with its buffer holding the strip's own rows only; the caller adds the halo pointers */
static FrameView StripView(const StripJob &job, int d)
{
	const Strip &s = job.strips[d];
	const StripGeom &G = job.geom[d];
	const int W = job.W, H = job.H, cw = W / 2, ch = H / 2;
	FrameView v{};
	v.width = W; v.height = H; v.whole_frame = false;
	v.luma_row0 = s.y0 / 8;
	v.luma_rows = (d == job.g - 1 ? (H + 8) / 8 : s.y1 / 8 + 1) - v.luma_row0;
	v.chroma_row0 = s.y0 / 16;
	v.chroma_rows = (d == job.g - 1 ? (ch + 8) / 8 : s.y1 / 16 + 1) - v.chroma_row0;
	v.luma_buf_rows = G.lrows; v.chroma_buf_rows = G.crows;
	/* the buffer starts at the strip's first row: its first block row begins 4 rows above it */
	v.luma_y_org = -4; v.chroma_y_org = -4;
	v.u_offset = v.u_offset_out = (long long)G.u_off; v.v_offset = v.v_offset_out = (long long)G.v_off;
	v.luma_pitch = W; v.chroma_pitch = cw;
	v.frame_stride = v.chroma_frame_stride = (long long)((G.bytes + 15) & ~(size_t)15);
	v.strip_halo = true;
	return v;
}

/* halo pointers: the last 4 rows of the strip above / the first 4 rows of the strip below, in their input planes */
static void SetHalo(FrameView &v, const StripJob &job, int d, const unsigned char *const up[3], const unsigned char *const down[3])
{
	const int W = job.W, cw = W / 2;
	if (up) {
		const StripGeom &U = job.geom[d - 1];
		v.halo_top[0] = up[0] + (size_t)(U.lrows - 4) * W;
		v.halo_top[1] = up[1] + (size_t)(U.crows - 4) * cw;
		v.halo_top[2] = up[2] + (size_t)(U.crows - 4) * cw;
	}
	if (down) { v.halo_bot[0] = down[0]; v.halo_bot[1] = down[1]; v.halo_bot[2] = down[2]; }
}

/* A: uploads of frame f on device d */
static int StripUpload(StripJob &job, int d, long long f)
{
	hdb_multi *m = job.m;
	StripDev &D = *m->dev[d];
	const long long seq = job.base_seq + f;
	StripSlot &S = D.slot[seq % kStripSlots];
	const Strip &s = job.strips[d];
	const StripGeom &G = job.geom[d];
	const int W = job.W, cw = W / 2, own = s.y1 - s.y0;
	/* the input buffer of this slot was last read by the kernels of frame seq - K: this device's and both neighbours' */
	if (seq >= kStripSlots && seq - kStripSlots >= job.base_seq) {
		for (int n = d - 1; n <= d + 1; n++) {
			if (n < 0 || n >= job.g) continue;
			if (!WaitRecorded(m->dev[n]->k_rec, seq - kStripSlots, m->strip_abort)) return HDB_ERR_CUDA;
			S_CUDA(cudaStreamWaitEvent(D.s_h2d, m->dev[n]->slot[seq % kStripSlots].k_done, 0));
		}
	}
	const unsigned char *fin = job.in + (size_t)f * job.frame_bytes;
	const unsigned char *in_u = fin + (size_t)W * job.H, *in_v = in_u + (size_t)cw * (job.H / 2);
	S_CUDA(cudaMemcpyAsync(S.in, fin + (size_t)s.y0 * W, (size_t)own * W, cudaMemcpyHostToDevice, D.s_h2d));
	S_CUDA(cudaMemcpyAsync(S.in + G.u_off, in_u + (size_t)(s.y0 / 2) * cw, (size_t)(own / 2) * cw, cudaMemcpyHostToDevice, D.s_h2d));
	S_CUDA(cudaMemcpyAsync(S.in + G.v_off, in_v + (size_t)(s.y0 / 2) * cw, (size_t)(own / 2) * cw, cudaMemcpyHostToDevice, D.s_h2d));
	if (job.vbs) {
		/* only the rows this strip's block rows read are transferred (cpu.h:161, 225, 289, 370) */
		const int w8 = W / 8, h8 = job.H / 8;
		const int by0 = s.y0 / 8, by1 = d == job.g - 1 ? h8 : s.y1 / 8;
		const int vr0 = std::max(by0 - 1, 0), vr1 = std::min(by1, h8 - 1);
		const int hr0 = by0, hr1 = std::min(by1, h8);
		const unsigned char *fv = job.vbs + (size_t)f * job.nv, *fh = job.hbs + (size_t)f * job.nh;
		if (vr1 >= vr0)
			S_CUDA(cudaMemcpyAsync(S.vbs + (size_t)vr0 * (w8 + 1), fv + (size_t)vr0 * (w8 + 1),
				(size_t)(vr1 - vr0 + 1) * (w8 + 1), cudaMemcpyHostToDevice, D.s_h2d));
		S_CUDA(cudaMemcpyAsync(S.hbs + (size_t)hr0 * w8, fh + (size_t)hr0 * w8,
			(size_t)(hr1 - hr0 + 1) * w8, cudaMemcpyHostToDevice, D.s_h2d));
	}
	S_CUDA(cudaEventRecord(S.up_done, D.s_h2d));
	D.up_rec.store(seq, std::memory_order_release);
	return HDB_OK;
}

/* B: the kernel of frame f on device d */
static int StripFilter(StripJob &job, int d, long long f)
{
	hdb_multi *m = job.m;
	hdb_context *c = m->ctx[d];
	StripDev &D = *m->dev[d];
	const long long seq = job.base_seq + f;
	const int si = (int)(seq % kStripSlots);
	StripSlot &S = D.slot[si];
	for (int n = d - 1; n <= d + 1; n++) {
		if (n < 0 || n >= job.g) continue;
		if (!WaitRecorded(m->dev[n]->up_rec, seq, m->strip_abort)) return HDB_ERR_CUDA;
		S_CUDA(cudaStreamWaitEvent(D.s_k, m->dev[n]->slot[si].up_done, 0));
	}
	/* the output buffer of this slot is free once the download of frame seq - K has finished */
	if (seq - kStripSlots >= job.base_seq) S_CUDA(cudaStreamWaitEvent(D.s_k, S.d2h_done, 0));
	if (job.timed) {
		if (S.timed) {          /* read the interval this slot still holds before its events are recorded again */
			S_CUDA(cudaEventSynchronize(S.k1));
			float ms = 0;
			cudaEventElapsedTime(&ms, S.k0, S.k1);
			D.kernel_ms += ms; S.timed = false;
		}
		S_CUDA(cudaEventRecord(S.k0, D.s_k));
	}
	FrameView v = StripView(job, d);
	{
		const unsigned char *up[3] = { nullptr, nullptr, nullptr }, *down[3] = { nullptr, nullptr, nullptr };
		if (d > 0) {
			const unsigned char *base = m->dev[d - 1]->slot[si].in;
			up[0] = base; up[1] = base + job.geom[d - 1].u_off; up[2] = base + job.geom[d - 1].v_off;
		}
		if (d < job.g - 1) {
			const unsigned char *base = m->dev[d + 1]->slot[si].in;
			down[0] = base; down[1] = base + job.geom[d + 1].u_off; down[2] = base + job.geom[d + 1].v_off;
		}
		SetHalo(v, job, d, d > 0 ? up : nullptr, d < job.g - 1 ? down : nullptr);
	}
	int rc;
	{
		std::lock_guard<std::mutex> lock(c->mu);
		rc = EnqueueFilter(c, v, S.in, S.out, 1, nullptr, job.qp[f], job.vbs ? S.vbs : nullptr, 0,
			job.vbs ? S.hbs : nullptr, 0, D.s_k, nullptr, job.edge_qp);
	}
	if (rc) { m->strip_abort.store(1); return rc; }
	if (job.timed) { S_CUDA(cudaEventRecord(S.k1, D.s_k)); S.timed = true; }
	S_CUDA(cudaEventRecord(S.k_done, D.s_k));
	D.k_rec.store(seq, std::memory_order_release);
	return HDB_OK;
}

/* C: downloads of frame f on device d */
static int StripDownload(StripJob &job, int d, long long f)
{
	hdb_multi *m = job.m;
	StripDev &D = *m->dev[d];
	const long long seq = job.base_seq + f;
	StripSlot &S = D.slot[seq % kStripSlots];
	const Strip &s = job.strips[d];
	const StripGeom &G = job.geom[d];
	const int W = job.W, cw = W / 2, own = s.y1 - s.y0;
	unsigned char *fout = job.out + (size_t)f * job.frame_bytes;
	unsigned char *out_u = fout + (size_t)W * job.H, *out_v = out_u + (size_t)cw * (job.H / 2);
	S_CUDA(cudaStreamWaitEvent(D.s_d2h, S.k_done, 0));
	S_CUDA(cudaMemcpyAsync(fout + (size_t)s.y0 * W, S.out, (size_t)own * W, cudaMemcpyDeviceToHost, D.s_d2h));
	S_CUDA(cudaMemcpyAsync(out_u + (size_t)(s.y0 / 2) * cw, S.out + G.u_off, (size_t)(own / 2) * cw, cudaMemcpyDeviceToHost, D.s_d2h));
	S_CUDA(cudaMemcpyAsync(out_v + (size_t)(s.y0 / 2) * cw, S.out + G.v_off, (size_t)(own / 2) * cw, cudaMemcpyDeviceToHost, D.s_d2h));
	S_CUDA(cudaEventRecord(S.d2h_done, D.s_d2h));
	return HDB_OK;
}

static int StripFinish(StripJob &job, int d)
{
	StripDev &D = *job.m->dev[d];
	S_CUDA(cudaStreamSynchronize(D.s_d2h));
	S_CUDA(cudaStreamSynchronize(D.s_k));
	S_CUDA(cudaStreamSynchronize(D.s_h2d));
	for (auto &S : D.slot) {
		if (!S.timed) continue;
		float ms = 0;
		cudaEventElapsedTime(&ms, S.k0, S.k1);
		D.kernel_ms += ms; S.timed = false;
	}
	return HDB_OK;
}

static std::vector<StripGeom> MakeGeom(const std::vector<Strip> &strips, int W)
{
	std::vector<StripGeom> out;
	for (const Strip &s : strips) {
		StripGeom G;
		G.y0 = s.y0; G.y1 = s.y1;
		G.lrows = s.y1 - s.y0; G.crows = (s.y1 - s.y0) / 2;
		G.u_off = (size_t)W * G.lrows;
		G.v_off = G.u_off + (size_t)(W / 2) * G.crows;
		G.bytes = G.v_off + (size_t)(W / 2) * G.crows;
		out.push_back(G);
	}
	return out;
}

/* device buffers of every slot, grown on demand (all work of earlier calls has completed: calls are blocking) */
static int EnsureStripBuffers(StripJob &job)
{
	for (int d = 0; d < job.g; d++) {
		hdb_context *c = job.m->ctx[d];
		DeviceGuard guard(c->device);
		for (auto &S : job.m->dev[d]->slot) {
			const size_t need = (job.geom[d].bytes + 15) & ~(size_t)15;
			if (S.cap < need) {
				cudaFree(S.in); cudaFree(S.out); S.in = S.out = nullptr; S.cap = 0;
				S_CUDA(cudaMalloc(&S.in, need));
				S_CUDA(cudaMalloc(&S.out, need));
				S.cap = need;
			}
			if (job.vbs && S.vbs_cap < job.nv) { cudaFree(S.vbs); S.vbs = nullptr; S_CUDA(cudaMalloc(&S.vbs, job.nv)); S.vbs_cap = job.nv; }
			if (job.vbs && S.hbs_cap < job.nh) { cudaFree(S.hbs); S.hbs = nullptr; S_CUDA(cudaMalloc(&S.hbs, job.nh)); S.hbs_cap = job.nh; }
		}
	}
	return HDB_OK;
}

} // namespace hdb

extern "C" {

int hdb_multi_strip_layout(const hdb_multi *m, unsigned height, int *first_row, int *n_rows)
{
	if (!m || m->ctx.empty() || height == 0 || height % 8 != 0) return 0;
	const std::vector<Strip> strips = MakeStrips((int)height, (int)m->ctx.size());
	for (size_t d = 0; d < strips.size(); d++) {
		if (first_row) first_row[d] = strips[d].y0;
		if (n_rows) n_rows[d] = strips[d].y1 - strips[d].y0;
	}
	return (int)strips.size();
}

int hdb_multi_deblock_stream_strips(hdb_multi *m, const unsigned char *frames_in, unsigned char *frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *qp,
	const unsigned char *vert, const unsigned char *hor, unsigned flags, hdb_timing *timing)
{
	if (!m || m->ctx.empty()) return HDB_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> call_lock(m->mu);
	if (flags & ~HDB_FLAG_EDGE_MAPS) return MFail(m, HDB_ERR_INVALID_ARG, "the strip split takes HDB_FLAG_EDGE_MAPS or 0");
	const bool edge_qp = (flags & HDB_FLAG_EDGE_MAPS) != 0;
	if (edge_qp && (!vert || !hor)) return MFail(m, HDB_ERR_INVALID_ARG, "edge maps are required: there is no default for them");
	if (!frames_in || !frames_out || !qp || n_frames == 0) return MFail(m, HDB_ERR_INVALID_ARG, "NULL buffer or empty stream");
	int rc = ValidateFrameArgs(m->ctx[0], width, height, vert, 0, hor, 0, false);
	if (rc) return MFail(m, rc, m->ctx[0]->err);
	const int W = (int)width, H = (int)height;
	const size_t fb = (size_t)W * H * 3 / 2;
	StripJob job;
	job.m = m; job.W = W; job.H = H;
	job.strips = MakeStrips(H, (int)m->ctx.size());
	job.g = (int)job.strips.size();
	/* strips need the tensor-map path; a frame it cannot describe, or a single strip, goes to one GPU whole */
	if (job.g == 1 || W % 32 != 0) {
		rc = BatchHost(m->ctx[0], frames_in, frames_out, n_frames, width, height, qp, vert, hor, timing, flags);
		return rc ? MFail(m, rc, m->ctx[0]->err) : HDB_OK;
	}
	job.geom = MakeGeom(job.strips, W);
	job.in = frames_in; job.out = frames_out; job.frame_bytes = fb;
	job.qp = qp; job.vbs = vert; job.hbs = hor;
	job.nv = hdb_num_vert_bs(width, height); job.nh = hdb_num_hor_bs(width, height);
	job.edge_qp = edge_qp; job.timed = timing != nullptr;
	job.base_seq = m->strip_seq;
	job.rc.assign(job.g, HDB_OK);
	m->strip_abort.store(0);
	for (int d = 0; d < job.g; d++) m->dev[d]->kernel_ms = 0;

	const auto t0 = std::chrono::steady_clock::now();
	PinHostRange pin_in(frames_in, fb * n_frames), pin_out(frames_out == frames_in ? nullptr : frames_out, fb * n_frames);
	PinHostRange pin_v(vert, (size_t)job.nv * n_frames), pin_h(hor, (size_t)job.nh * n_frames);
	if ((rc = EnsureStripBuffers(job))) return MFail(m, rc, "strip buffers: out of device memory");
	auto run_device = [&job, n_frames](int d) {
		DeviceGuard guard(job.m->ctx[d]->device);
		int r = HDB_OK;
		for (long long f = 0; f < (long long)n_frames && !r; f++) {
			/* one frame ahead with the upload: A(f + 1) is enqueued before B(f) so that the copy engine never
			 * waits for this thread */
			if (f == 0) r = StripUpload(job, d, 0);
			if (!r && f + 1 < (long long)n_frames) r = StripUpload(job, d, f + 1);
			if (!r) r = StripFilter(job, d, f);
			if (!r) r = StripDownload(job, d, f);
		}
		if (r) job.m->strip_abort.store(1);
		const int r2 = StripFinish(job, d);
		job.rc[d] = r ? r : r2;
	};
	if (n_frames == 1) {
		/* a single frame: three short phases from the calling thread, no thread start-up in its latency */
		for (int d = 0; d < job.g && !rc; d++) { DeviceGuard guard(m->ctx[d]->device); rc = StripUpload(job, d, 0); }
		for (int d = 0; d < job.g && !rc; d++) { DeviceGuard guard(m->ctx[d]->device); rc = StripFilter(job, d, 0); if (!rc) rc = StripDownload(job, d, 0); }
		for (int d = 0; d < job.g; d++) { DeviceGuard guard(m->ctx[d]->device); const int r2 = StripFinish(job, d); if (!rc) rc = r2; }
		job.rc[0] = rc;
	} else {
		std::vector<std::thread> pool;
		for (int d = 0; d < job.g; d++) pool.emplace_back(run_device, d);
		for (auto &t : pool) t.join();
	}
	m->strip_seq += n_frames;
	const auto t1 = std::chrono::steady_clock::now();
	for (int d = 0; d < job.g; d++)
		if (job.rc[d]) {
			for (auto *c : m->ctx) { DeviceGuard gd(c->device); cudaDeviceSynchronize(); cudaGetLastError(); }
			for (auto *sd : m->dev) for (auto &S : sd->slot) S.timed = false;
			return MFail(m, job.rc[d], std::string("device ") + std::to_string(m->ctx[d]->device) + ": " + m->ctx[d]->err);
		}
	if (timing) {
		memset(timing, 0, sizeof *timing);
		for (int d = 0; d < job.g; d++) timing->kernel_ms = std::max(timing->kernel_ms, m->dev[d]->kernel_ms);   /* devices run concurrently */
		timing->total_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
		timing->copy_ms = std::max(0.0, timing->total_ms - timing->kernel_ms);
		timing->h2d_bytes = (unsigned long long)n_frames * (fb + (vert ? job.nv + job.nh : 0));
		timing->d2h_bytes = (unsigned long long)n_frames * fb;
		timing->kernel_launches = job.g * (int)n_frames;
		timing->path = 1;
	}
	return HDB_OK;
}

int hdb_multi_deblock_frame_strips(hdb_multi *m, const unsigned char *yuv_in, unsigned char *yuv_out,
	unsigned width, unsigned height, unsigned qp, const unsigned char *vert_bs, unsigned num_vert_bs,
	const unsigned char *hor_bs, unsigned num_hor_bs, hdb_timing *timing)
{
	return hdb_multi_deblock_frame_strips_ex(m, yuv_in, yuv_out, width, height, qp, vert_bs, num_vert_bs, hor_bs, num_hor_bs, 0, timing);
}

int hdb_multi_deblock_frame_strips_ex(hdb_multi *m, const unsigned char *yuv_in, unsigned char *yuv_out,
	unsigned width, unsigned height, unsigned qp, const unsigned char *vert_bs, unsigned num_vert_bs,
	const unsigned char *hor_bs, unsigned num_hor_bs, unsigned flags, hdb_timing *timing)
{
	if (!m || m->ctx.empty()) return HDB_ERR_INVALID_ARG;
	if (!yuv_in || !yuv_out) return MFail(m, HDB_ERR_INVALID_ARG, "NULL buffer");
	int rc = ValidateFrameArgs(m->ctx[0], width, height, vert_bs, num_vert_bs, hor_bs, num_hor_bs, true);
	if (rc) return MFail(m, rc, m->ctx[0]->err);
	return hdb_multi_deblock_stream_strips(m, yuv_in, yuv_out, 1, width, height, &qp, vert_bs, hor_bs, flags, timing);
}

/*
 * The same split for strips that already sit in each GPU's memory (a decoder per GPU, or a previous stage):
 * nothing crosses the host link; per frame each GPU runs one kernel, and 4 rows per plane and cut cross NVLink
 * inside it.
 */
int hdb_multi_deblock_strips_device(hdb_multi *m, const hdb_strip_planes *in, const hdb_strip_planes *out,
	unsigned width, unsigned height, unsigned qp, const unsigned char *const *d_vert, const unsigned char *const *d_hor,
	unsigned flags, void *const *streams, hdb_timing *timing)
{
	if (!m || m->ctx.empty()) return HDB_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> call_lock(m->mu);
	if (flags & ~HDB_FLAG_EDGE_MAPS) return MFail(m, HDB_ERR_INVALID_ARG, "the strip split takes HDB_FLAG_EDGE_MAPS or 0");
	const bool edge_qp = (flags & HDB_FLAG_EDGE_MAPS) != 0;
	if (!in || !out) return MFail(m, HDB_ERR_INVALID_ARG, "NULL strip array");
	if ((d_vert != nullptr) != (d_hor != nullptr)) return MFail(m, HDB_ERR_INVALID_ARG, "vert and hor must both be given or both be NULL");
	if (edge_qp && !d_vert) return MFail(m, HDB_ERR_INVALID_ARG, "edge maps are required: there is no default for them");
	int rc = ValidateFrameArgs(m->ctx[0], width, height, nullptr, 0, nullptr, 0, false);
	if (rc) return MFail(m, rc, m->ctx[0]->err);
	if (width % 32 != 0) return MFail(m, HDB_ERR_DIMENSIONS, "device-resident strips need a width that is a multiple of 32");
	const int W = (int)width, H = (int)height, cw = W / 2;
	StripJob job;
	job.m = m; job.W = W; job.H = H;
	job.strips = MakeStrips(H, (int)m->ctx.size());
	job.g = (int)job.strips.size();
	job.geom = MakeGeom(job.strips, W);
	job.edge_qp = edge_qp;
	for (int d = 0; d < job.g; d++) {
		const uintptr_t a = (uintptr_t)in[d].y | (uintptr_t)in[d].u | (uintptr_t)in[d].v | (uintptr_t)out[d].y | (uintptr_t)out[d].u | (uintptr_t)out[d].v;
		if (!in[d].y || !in[d].u || !in[d].v || !out[d].y || !out[d].u || !out[d].v || a % 16 != 0)
			return MFail(m, HDB_ERR_INVALID_ARG, "strip planes must be non-NULL and 16-byte aligned");
		if (in[d].y == out[d].y || in[d].u == out[d].u || in[d].v == out[d].v)
			return MFail(m, HDB_ERR_INVALID_ARG, "strips are filtered out of place: the neighbours read the input rows");
		if (d_vert && (!d_vert[d] || !d_hor[d])) return MFail(m, HDB_ERR_INVALID_ARG, "NULL boundary-strength array");
	}
	auto stream_of = [&](int d) { return streams ? static_cast<cudaStream_t>(streams[d]) : m->dev[d]->s_k; };
	const auto t0 = std::chrono::steady_clock::now();
	/* a device's input rows are ready when the work already enqueued on its stream has completed; the
	 * neighbours' kernels, which read 4 of those rows, wait for that point too */
	for (int d = 0; d < job.g; d++) {
		DeviceGuard guard(m->ctx[d]->device);
		M_CUDA(cudaEventRecord(m->dev[d]->ready, stream_of(d)));
	}
	for (int d = 0; d < job.g; d++) {
		hdb_context *c = m->ctx[d];
		DeviceGuard guard(c->device);
		StripSlot &S = m->dev[d]->slot[0];
		cudaStream_t st = stream_of(d);
		if (d > 0) M_CUDA(cudaStreamWaitEvent(st, m->dev[d - 1]->ready, 0));
		if (d < job.g - 1) M_CUDA(cudaStreamWaitEvent(st, m->dev[d + 1]->ready, 0));
		FrameView v = StripView(job, d);
		const unsigned char *iy = static_cast<const unsigned char *>(in[d].y);
		unsigned char *oy = static_cast<unsigned char *>(out[d].y);
		v.u_offset = static_cast<const unsigned char *>(in[d].u) - iy; v.v_offset = static_cast<const unsigned char *>(in[d].v) - iy;
		v.u_offset_out = static_cast<unsigned char *>(out[d].u) - oy; v.v_offset_out = static_cast<unsigned char *>(out[d].v) - oy;
		v.frame_stride = (long long)W * job.geom[d].lrows; v.chroma_frame_stride = (long long)cw * job.geom[d].crows;
		const unsigned char *up[3], *down[3];
		if (d > 0) { up[0] = static_cast<const unsigned char *>(in[d - 1].y); up[1] = static_cast<const unsigned char *>(in[d - 1].u); up[2] = static_cast<const unsigned char *>(in[d - 1].v); }
		if (d < job.g - 1) { down[0] = static_cast<const unsigned char *>(in[d + 1].y); down[1] = static_cast<const unsigned char *>(in[d + 1].u); down[2] = static_cast<const unsigned char *>(in[d + 1].v); }
		SetHalo(v, job, d, d > 0 ? up : nullptr, d < job.g - 1 ? down : nullptr);
		if (timing) M_CUDA(cudaEventRecord(S.k0, st));
		{
			std::lock_guard<std::mutex> lock(c->mu);
			rc = EnqueueFilter(c, v, iy, oy, 1, nullptr, qp, d_vert ? d_vert[d] : nullptr, 0, d_hor ? d_hor[d] : nullptr, 0,
				st, nullptr, edge_qp);
		}
		if (rc) {
			for (auto *cc : m->ctx) { DeviceGuard gd(cc->device); cudaDeviceSynchronize(); }
			return MFail(m, rc, c->err);
		}
		if (timing) M_CUDA(cudaEventRecord(S.k1, st));
	}
	if (!timing) return HDB_OK;       /* asynchronous: the caller synchronises its streams */
	for (int d = 0; d < job.g; d++) {
		DeviceGuard guard(m->ctx[d]->device);
		M_CUDA(cudaStreamSynchronize(stream_of(d)));
	}
	const auto t1 = std::chrono::steady_clock::now();
	memset(timing, 0, sizeof *timing);
	for (int d = 0; d < job.g; d++) {
		float ms = 0;
		cudaEventElapsedTime(&ms, m->dev[d]->slot[0].k0, m->dev[d]->slot[0].k1);
		timing->kernel_ms = std::max(timing->kernel_ms, (double)ms);
	}
	timing->total_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
	timing->kernel_launches = job.g;
	timing->path = 1;
	return HDB_OK;
}

unsigned long long hdb_multi_strip_halo_bytes(const hdb_multi *m, unsigned width, unsigned height)
{
	if (!m || m->ctx.empty() || height % 8 != 0) return 0;
	const int cuts = (int)MakeStrips((int)height, (int)m->ctx.size()).size() - 1;
	/* per cut and direction: 4 luma rows and 4 rows of each chroma plane */
	return (unsigned long long)std::max(cuts, 0) * 2ull * (4ull * width + 2ull * 4ull * (width / 2));
}

} // extern "C"
