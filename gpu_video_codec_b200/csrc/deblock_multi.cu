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
#include <chrono>
#include <cstring>
#include <thread>

struct hdb_multi {
	std::vector<hdb_context *> ctx;
	std::vector<cudaStream_t> stream;            /* one per device, for the strip path */
	std::vector<cudaEvent_t> up_done, fetched;
	std::vector<unsigned char *> strip_buf, vbs, hbs;
	std::vector<size_t> strip_cap, vbs_cap, hbs_cap;
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
	int top, bot;          /* halo rows above / below held in the buffer (0 or 4) */
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
		s.top = d > 0 ? 4 : 0;
		s.bot = d < parts - 1 ? 4 : 0;
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
	m->stream.assign(n, nullptr); m->up_done.assign(n, nullptr); m->fetched.assign(n, nullptr);
	m->strip_buf.assign(n, nullptr); m->vbs.assign(n, nullptr); m->hbs.assign(n, nullptr);
	m->strip_cap.assign(n, 0); m->vbs_cap.assign(n, 0); m->hbs_cap.assign(n, 0);
	for (int i = 0; i < n; i++) {
		DeviceGuard g(m->ctx[i]->device);
		if (cudaStreamCreateWithFlags(&m->stream[i], cudaStreamNonBlocking) != cudaSuccess
			|| cudaEventCreateWithFlags(&m->up_done[i], cudaEventDisableTiming) != cudaSuccess
			|| cudaEventCreateWithFlags(&m->fetched[i], cudaEventDisableTiming) != cudaSuccess) {
			hdb_multi_destroy(m);
			return Fail(nullptr, HDB_ERR_CUDA, "multi-GPU stream/event creation failed");
		}
		/* direct peer access to the neighbours (NVLink / NVSwitch when present); the peer
		 * copies below also work without it, staged by the driver */
		for (int j = 0; j < n; j++) {
			if (j == i) continue;
			int can = 0;
			if (cudaDeviceCanAccessPeer(&can, m->ctx[i]->device, m->ctx[j]->device) == cudaSuccess && can) {
				cudaError_t e = cudaDeviceEnablePeerAccess(m->ctx[j]->device, 0);
				if (e != cudaSuccess) cudaGetLastError();   /* already enabled is fine */
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
		if (i < m->strip_buf.size()) { cudaFree(m->strip_buf[i]); cudaFree(m->vbs[i]); cudaFree(m->hbs[i]); }
		if (i < m->up_done.size() && m->up_done[i]) cudaEventDestroy(m->up_done[i]);
		if (i < m->fetched.size() && m->fetched[i]) cudaEventDestroy(m->fetched[i]);
		if (i < m->stream.size() && m->stream[i]) cudaStreamDestroy(m->stream[i]);
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
	if (flags & ~(HDB_FLAG_EDGE_MAPS | HDB_FLAG_NV12)) return MFail(m, HDB_ERR_INVALID_ARG, "unknown flag");
	if (!frames_in || !frames_out || !qp || n_frames == 0) return MFail(m, HDB_ERR_INVALID_ARG, "NULL buffer or empty batch");
	const int g = (int)std::min<size_t>(m->ctx.size(), n_frames);
	const size_t fb = (size_t)width * height * 3 / 2;
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

#define M_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
	for (auto *c__ : m->ctx) { DeviceGuard g__(c__->device); cudaDeviceSynchronize(); } \
	return MFail(m, HDB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } } while (0)

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
	if (flags & ~HDB_FLAG_EDGE_MAPS) return MFail(m, HDB_ERR_INVALID_ARG, "the strip split takes HDB_FLAG_EDGE_MAPS or 0");
	const bool edge_qp = (flags & HDB_FLAG_EDGE_MAPS) != 0;
	if (edge_qp && (!vert_bs || !hor_bs)) return MFail(m, HDB_ERR_INVALID_ARG, "edge maps are required: there is no default for them");
	if (!yuv_in || !yuv_out) return MFail(m, HDB_ERR_INVALID_ARG, "NULL buffer");
	int rc = ValidateFrameArgs(m->ctx[0], width, height, vert_bs, num_vert_bs, hor_bs, num_hor_bs, true);
	if (rc) return MFail(m, rc, m->ctx[0]->err);
	const int W = (int)width, H = (int)height, cw = W / 2, ch = H / 2;
	const std::vector<Strip> strips = MakeStrips(H, (int)m->ctx.size());
	const int g = (int)strips.size();
	/* strips need the tensor-map path; a frame it cannot describe, or a single strip, goes to one GPU whole */
	if (g == 1 || W % 32 != 0) {
		rc = edge_qp ? hdb_deblock_frame_edges(m->ctx[0], yuv_in, yuv_out, width, height, qp, vert_bs, num_vert_bs, hor_bs, num_hor_bs, timing)
			: hdb_deblock_frame(m->ctx[0], yuv_in, yuv_out, width, height, qp, vert_bs, num_vert_bs, hor_bs, num_hor_bs, timing);
		return rc ? MFail(m, rc, m->ctx[0]->err) : HDB_OK;
	}

	const auto t0 = std::chrono::steady_clock::now();
	const unsigned nv = hdb_num_vert_bs(width, height), nh = hdb_num_hor_bs(width, height);
	const int w8 = W / 8, h8 = H / 8;
	const unsigned char *in_u = yuv_in + (size_t)W * H, *in_v = in_u + (size_t)cw * ch;
	unsigned char *out_u = yuv_out + (size_t)W * H, *out_v = out_u + (size_t)cw * ch;
	PinHostRange pin_in(yuv_in, (size_t)W * H * 3 / 2), pin_out(yuv_out == yuv_in ? nullptr : yuv_out, (size_t)W * H * 3 / 2);

	struct Layout { int lrows, crows; size_t u_off, v_off, bytes; unsigned char *y, *u, *v; };
	std::vector<Layout> lay(g);
	/* 1. per device: buffers, upload of the strip's own rows and of its boundary-strength rows */
	for (int d = 0; d < g; d++) {
		hdb_context *c = m->ctx[d];
		DeviceGuard guard(c->device);
		const Strip &s = strips[d];
		Layout &L = lay[d];
		L.lrows = s.top + (s.y1 - s.y0) + s.bot;
		L.crows = s.top + (s.y1 - s.y0) / 2 + s.bot;
		L.u_off = (size_t)W * L.lrows;
		L.v_off = L.u_off + (size_t)cw * L.crows;
		L.bytes = L.v_off + (size_t)cw * L.crows;
		if (m->strip_cap[d] < L.bytes) {
			cudaFree(m->strip_buf[d]); m->strip_buf[d] = nullptr; m->strip_cap[d] = 0;
			M_CUDA(cudaMalloc(&m->strip_buf[d], L.bytes));
			m->strip_cap[d] = L.bytes;
		}
		L.y = m->strip_buf[d]; L.u = L.y + L.u_off; L.v = L.y + L.v_off;
		cudaStream_t st = m->stream[d];
		M_CUDA(cudaMemcpyAsync(L.y + (size_t)s.top * W, yuv_in + (size_t)s.y0 * W, (size_t)(s.y1 - s.y0) * W, cudaMemcpyHostToDevice, st));
		M_CUDA(cudaMemcpyAsync(L.u + (size_t)s.top * cw, in_u + (size_t)(s.y0 / 2) * cw, (size_t)((s.y1 - s.y0) / 2) * cw, cudaMemcpyHostToDevice, st));
		M_CUDA(cudaMemcpyAsync(L.v + (size_t)s.top * cw, in_v + (size_t)(s.y0 / 2) * cw, (size_t)((s.y1 - s.y0) / 2) * cw, cudaMemcpyHostToDevice, st));
		if (vert_bs) {
			/* full-size arrays on the device so that the kernel's frame-global indices hold;
			 * only the rows this strip's block rows read are transferred */
			if (m->vbs_cap[d] < nv) { cudaFree(m->vbs[d]); m->vbs[d] = nullptr; M_CUDA(cudaMalloc(&m->vbs[d], nv)); m->vbs_cap[d] = nv; }
			if (m->hbs_cap[d] < nh) { cudaFree(m->hbs[d]); m->hbs[d] = nullptr; M_CUDA(cudaMalloc(&m->hbs[d], nh)); m->hbs_cap[d] = nh; }
			const int by0 = s.y0 / 8, by1 = d == g - 1 ? h8 : s.y1 / 8;   /* block rows by0..by1 */
			const int vr0 = std::max(by0 - 1, 0), vr1 = std::min(by1, h8 - 1);   /* vert_bs rows (cpu.h:161, 225) */
			const int hr0 = by0, hr1 = std::min(by1, h8);                          /* hor_bs rows  (cpu.h:289, 370) */
			if (vr1 >= vr0)
				M_CUDA(cudaMemcpyAsync(m->vbs[d] + (size_t)vr0 * (w8 + 1), vert_bs + (size_t)vr0 * (w8 + 1),
					(size_t)(vr1 - vr0 + 1) * (w8 + 1), cudaMemcpyHostToDevice, st));
			M_CUDA(cudaMemcpyAsync(m->hbs[d] + (size_t)hr0 * w8, hor_bs + (size_t)hr0 * w8,
				(size_t)(hr1 - hr0 + 1) * w8, cudaMemcpyHostToDevice, st));
		}
		M_CUDA(cudaEventRecord(m->up_done[d], st));
	}
	/* 2. halo exchange: the 4 input rows on the far side of each cut, by peer copy from the
	 *    neighbour's buffer (its own first / last rows), after the neighbour's upload */
	for (int d = 0; d < g; d++) {
		hdb_context *c = m->ctx[d];
		DeviceGuard guard(c->device);
		cudaStream_t st = m->stream[d];
		const Layout &L = lay[d];
		if (d > 0) {
			const Layout &N = lay[d - 1];
			const Strip &ns = strips[d - 1];
			const int nown = ns.y1 - ns.y0;
			M_CUDA(cudaStreamWaitEvent(st, m->up_done[d - 1], 0));
			M_CUDA(cudaMemcpyPeerAsync(L.y, c->device, N.y + (size_t)(ns.top + nown - 4) * W, m->ctx[d - 1]->device, (size_t)4 * W, st));
			M_CUDA(cudaMemcpyPeerAsync(L.u, c->device, N.u + (size_t)(ns.top + nown / 2 - 4) * cw, m->ctx[d - 1]->device, (size_t)4 * cw, st));
			M_CUDA(cudaMemcpyPeerAsync(L.v, c->device, N.v + (size_t)(ns.top + nown / 2 - 4) * cw, m->ctx[d - 1]->device, (size_t)4 * cw, st));
		}
		if (d < g - 1) {
			const Layout &N = lay[d + 1];
			const Strip &ns = strips[d + 1];
			const Strip &s = strips[d];
			const int own = s.y1 - s.y0;
			M_CUDA(cudaStreamWaitEvent(st, m->up_done[d + 1], 0));
			M_CUDA(cudaMemcpyPeerAsync(L.y + (size_t)(s.top + own) * W, c->device, N.y + (size_t)ns.top * W, m->ctx[d + 1]->device, (size_t)4 * W, st));
			M_CUDA(cudaMemcpyPeerAsync(L.u + (size_t)(s.top + own / 2) * cw, c->device, N.u + (size_t)ns.top * cw, m->ctx[d + 1]->device, (size_t)4 * cw, st));
			M_CUDA(cudaMemcpyPeerAsync(L.v + (size_t)(s.top + own / 2) * cw, c->device, N.v + (size_t)ns.top * cw, m->ctx[d + 1]->device, (size_t)4 * cw, st));
		}
		M_CUDA(cudaEventRecord(m->fetched[d], st));
	}
	/* 3. filter in place once the neighbours have taken their halos from this buffer, then
	 *    download the strip's own rows only */
	std::vector<cudaEvent_t> k0(g, nullptr), k1(g, nullptr);
	for (int d = 0; d < g; d++) {
		hdb_context *c = m->ctx[d];
		DeviceGuard guard(c->device);
		cudaStream_t st = m->stream[d];
		const Strip &s = strips[d];
		const Layout &L = lay[d];
		if (d > 0) M_CUDA(cudaStreamWaitEvent(st, m->fetched[d - 1], 0));
		if (d < g - 1) M_CUDA(cudaStreamWaitEvent(st, m->fetched[d + 1], 0));
		FrameView v{};
		v.width = W; v.height = H; v.whole_frame = false;
		v.luma_row0 = s.y0 / 8;
		v.luma_rows = (d == g - 1 ? (H + 8) / 8 : s.y1 / 8 + 1) - v.luma_row0;
		v.chroma_row0 = s.y0 / 16;
		v.chroma_rows = (d == g - 1 ? (ch + 8) / 8 : s.y1 / 16 + 1) - v.chroma_row0;
		v.luma_buf_rows = L.lrows; v.chroma_buf_rows = L.crows;
		/* the buffer starts 4 rows above the strip's first block row, except at the top of the frame */
		v.luma_y_org = v.luma_row0 == 0 ? -4 : 0; v.chroma_y_org = v.chroma_row0 == 0 ? -4 : 0;
		v.u_offset = v.u_offset_out = (long long)L.u_off; v.v_offset = v.v_offset_out = (long long)L.v_off;
		v.luma_pitch = W; v.chroma_pitch = cw;
		v.frame_stride = v.chroma_frame_stride = (long long)((L.bytes + 15) & ~(size_t)15);
		if (timing) {
			M_CUDA(cudaEventCreate(&k0[d])); M_CUDA(cudaEventCreate(&k1[d]));
			M_CUDA(cudaEventRecord(k0[d], st));
		}
		{
			std::lock_guard<std::mutex> lock(c->mu);
			rc = EnqueueFilter(c, v, L.y, L.y, 1, nullptr, qp, vert_bs ? m->vbs[d] : nullptr, 0,
				vert_bs ? m->hbs[d] : nullptr, 0, st, nullptr, edge_qp);
		}
		if (rc) return MFail(m, rc, c->err);
		if (timing) M_CUDA(cudaEventRecord(k1[d], st));
		const int own = s.y1 - s.y0;
		M_CUDA(cudaMemcpyAsync(yuv_out + (size_t)s.y0 * W, L.y + (size_t)s.top * W, (size_t)own * W, cudaMemcpyDeviceToHost, st));
		M_CUDA(cudaMemcpyAsync(out_u + (size_t)(s.y0 / 2) * cw, L.u + (size_t)s.top * cw, (size_t)(own / 2) * cw, cudaMemcpyDeviceToHost, st));
		M_CUDA(cudaMemcpyAsync(out_v + (size_t)(s.y0 / 2) * cw, L.v + (size_t)s.top * cw, (size_t)(own / 2) * cw, cudaMemcpyDeviceToHost, st));
	}
	for (int d = 0; d < g; d++) {
		DeviceGuard guard(m->ctx[d]->device);
		M_CUDA(cudaStreamSynchronize(m->stream[d]));
	}
	const auto t1 = std::chrono::steady_clock::now();
	if (timing) {
		memset(timing, 0, sizeof *timing);
		for (int d = 0; d < g; d++) {
			float ms = 0;
			cudaEventElapsedTime(&ms, k0[d], k1[d]);
			timing->kernel_ms = std::max(timing->kernel_ms, (double)ms);
			cudaEventDestroy(k0[d]); cudaEventDestroy(k1[d]);
			timing->h2d_bytes += (unsigned long long)(strips[d].y1 - strips[d].y0) * W * 3 / 2;
			timing->d2h_bytes += (unsigned long long)(strips[d].y1 - strips[d].y0) * W * 3 / 2;
		}
		timing->total_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
		timing->copy_ms = timing->total_ms - timing->kernel_ms;
		timing->kernel_launches = g;
		timing->path = 1;
	}
	return HDB_OK;
}

} // extern "C"
