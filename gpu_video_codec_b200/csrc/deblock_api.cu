/*
 * deblock_api.cu -- host side of the library: context, tensor maps, the pinned
 * multi-stream pipeline and the single-GPU C-ABI entry points of
 * include/hevc_deblock_b200.h.  Replaces the reference's GPU host path
 * (gpu.cu:1074-1306: Initialize / ExecuteGpu / Release), which allocates per
 * call, copies synchronously on the default stream and keeps its state in
 * file-scope globals (gpu.cu:35-77).
 */
#include "deblock_api.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/stat.h>

namespace hdb {

static const char *kMsgFileSize = "Incorrect file size";                                             /* cpu.h:44 */
static const char *kMsgDims = "Width and height of image must be multiplier of sample block size";   /* cpu.h:47 */
static const char *kMsgBsSize = "Incorrect size of input boundary strenght array";                    /* cpu.h:123 */

static std::string g_create_error;

int Fail(hdb_context *ctx, int status, const std::string &msg)
{
	if (ctx) ctx->err = msg; else g_create_error = msg;
	return status;
}

int FailCuda(hdb_context *ctx, cudaError_t e, const char *what)
{
	return Fail(ctx, HDB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define HDB_CUDA(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return FailCuda(ctx, e__, #call); } while (0)

/* ------------------------------------------------------------------ */
/* context                                                             */
/* ------------------------------------------------------------------ */
/* cpu.h:86-87 evaluated as the reference does (unsigned, left to right); the operands are range-checked first */
static unsigned NumBs32(unsigned w, unsigned h, bool vert) { return vert ? (w / 8 + 1) * h / 8 : (h / 8 + 1) * w / 8; }

static int CreateContext(hdb_context **out, int device)
{
	if (!out) return Fail(nullptr, HDB_ERR_INVALID_ARG, "ctx is NULL");
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return Fail(nullptr, HDB_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
	if (device < 0 || device >= count) return Fail(nullptr, HDB_ERR_INVALID_ARG, "device ordinal out of range");
	hdb_context *ctx = new (std::nothrow) hdb_context();
	if (!ctx) return Fail(nullptr, HDB_ERR_NO_MEMORY, "out of host memory");
	ctx->device = device;
	DeviceGuard guard(device);
	if ((e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess) {
		delete ctx;
		return FailCuda(nullptr, e, "cudaGetDeviceProperties");
	}
	if (ctx->prop.major != 10) {
		char buf[256];
		snprintf(buf, sizeof buf, "device %d (%s) is compute capability %d.%d; this library is built for sm_100a only",
			device, ctx->prop.name, ctx->prop.major, ctx->prop.minor);
		delete ctx;
		return Fail(nullptr, HDB_ERR_CUDA, buf);
	}
	void *fn = nullptr;
	cudaDriverEntryPointQueryResult qres;
	e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
	if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
		delete ctx;
		return Fail(nullptr, HDB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
	}
	ctx->encode = reinterpret_cast<EncodeTiledFn>(fn);
	/* measurement knob: HDB_ZERO_COPY=0 sends single-slice calls through the copy engines too */
	if (const char *z = getenv("HDB_ZERO_COPY")) ctx->zero_copy = atoi(z) != 0;
	/* test / measurement knob: HDB_NO_LDGSTS sends widths off the 32 grid through the padded-scratch kernels */
	if (getenv("HDB_NO_LDGSTS")) ctx->use_ldgsts = false;
	/* tuning knob for measurements: slice size of the host pipeline in MB */
	if (const char *mb = getenv("HDB_CHUNK_MB")) {
		const long v = strtol(mb, nullptr, 10);
		if (v >= 1 && v <= 4096) ctx->chunk_bytes = (size_t)v << 20;
	}
	bool ok = cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking) == cudaSuccess
		&& cudaStreamCreateWithFlags(&ctx->s_k, cudaStreamNonBlocking) == cudaSuccess
		&& cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking) == cudaSuccess
		&& PrepareTiled(ctx->tiled_ctas_per_sm) == cudaSuccess
		&& PrepareTiled16(ctx->tiled16_ctas_per_sm) == cudaSuccess
		&& cudaStreamCreateWithFlags(&ctx->s_setup, cudaStreamNonBlocking) == cudaSuccess
		&& cudaMalloc(&ctx->sched, TiledSchedBytes() * kSchedSlots) == cudaSuccess
		&& cudaMemset(ctx->sched, 0, TiledSchedBytes() * kSchedSlots) == cudaSuccess;
	if (ok) {
		cudaMemPoolProps props{};
		props.allocType = cudaMemAllocationTypePinned;
		props.handleTypes = cudaMemHandleTypeNone;
		props.location.type = cudaMemLocationTypeDevice;
		props.location.id = device;
		unsigned long long keep = ~0ull;
		ok = cudaMemPoolCreate(&ctx->pool, &props) == cudaSuccess
			&& cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess;
	}
	for (int i = 0; ok && i < kSchedSlots; i++)
		ok = cudaEventCreateWithFlags(&ctx->sched_done[i], cudaEventDisableTiming) == cudaSuccess;
	if (!ok) {
		e = cudaGetLastError();
		DestroyContext(ctx);
		return FailCuda(nullptr, e, "context resources");
	}
	*out = ctx;
	return HDB_OK;
}

void DestroyContext(hdb_context *ctx)
{
	if (!ctx) return;
	DeviceGuard guard(ctx->device);
	cudaDeviceSynchronize();
	for (auto &s : ctx->sizes) { cudaFree(s.chroma_bs); cudaFree(s.def_vbs); cudaFree(s.def_hbs); }
	for (auto &sl : ctx->slots) {
		cudaFree(sl.frames); cudaFree(sl.vbs); cudaFree(sl.hbs); cudaFree(sl.info);
		if (sl.ev_ready) { cudaEventDestroy(sl.h2d_done); cudaEventDestroy(sl.k_done); cudaEventDestroy(sl.d2h_done); }
	}
	for (auto ev : ctx->sched_done) if (ev) cudaEventDestroy(ev);
	if (ctx->s_setup) cudaStreamDestroy(ctx->s_setup);
	if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
	cudaFree(ctx->qp_dev);
	cudaFree(ctx->derive_dev);
	for (auto ev : ctx->event_pool) cudaEventDestroy(ev);
	cudaFree(ctx->sched);
	if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
	if (ctx->s_k) cudaStreamDestroy(ctx->s_k);
	if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
	delete ctx;
}

/* per-size device tables: resolved chroma strengths, default luma strengths */
int GetSizeState(hdb_context *ctx, int w, int h, SizeState **out)
{
	for (auto &s : ctx->sizes) if (s.w == w && s.h == h) { *out = &s; return HDB_OK; }
	SizeState s{};
	s.w = w; s.h = h;
	const size_t ncb = (size_t)((w / 2 + 8) / 8) * ((h / 2 + 8) / 8);
	const unsigned nv = NumBs32((unsigned)w, (unsigned)h, true), nh = NumBs32((unsigned)w, (unsigned)h, false);
	HDB_CUDA(ctx, cudaMalloc(&s.chroma_bs, std::max<size_t>(ncb, 1) * sizeof(uint32_t)));
	HDB_CUDA(ctx, cudaMalloc(&s.def_vbs, std::max(nv, 1u)));
	HDB_CUDA(ctx, cudaMalloc(&s.def_hbs, std::max(nh, 1u)));
	/* built on a stream of the context's own and waited for here, once per frame size: visible to every stream
	 * afterwards, and no other stream of the device is touched (the public header says so) */
	HDB_CUDA(ctx, LaunchChromaBsTable(s.chroma_bs, w, h, ctx->s_setup));
	HDB_CUDA(ctx, LaunchDefaultLumaBs(s.def_vbs, s.def_hbs, w, h, ctx->s_setup));
	HDB_CUDA(ctx, cudaStreamSynchronize(ctx->s_setup));
	ctx->sizes.push_back(s);
	*out = &ctx->sizes.back();
	return HDB_OK;
}

static int EncodePlaneMap(hdb_context *ctx, CUtensorMap *map, const void *base, int pw, int ph, int n,
	long long pitch, long long frame_stride, int box_rows, unsigned box_elems = 68u)
{
	/* pw: bytes per row.  32-bit elements: a box may then be 272 (8-bit samples) or 528 bytes (16-bit) wide (a box
	 * dimension is limited to 256 elements) */
	cuuint64_t dims[3] = { (cuuint64_t)pw / 4, (cuuint64_t)ph, (cuuint64_t)n };
	cuuint64_t strides[2] = { (cuuint64_t)pitch, (cuuint64_t)frame_stride };
	cuuint32_t box[3] = { box_elems, (cuuint32_t)box_rows, 1u };
	cuuint32_t estr[3] = { 1u, 1u, 1u };
	CUresult r = ctx->encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void *>(base), dims, strides, box, estr,
		CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		char buf[160];
		snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) for a %dx%d plane, %d frames", (int)r, pw, ph, n);
		return Fail(ctx, HDB_ERR_CUDA, buf);
	}
	return HDB_OK;
}

bool TiledPathUsable(int w, int h, const void *d_in, const void *d_out)
{
	(void)h;
	return w % 32 == 0 && ((uintptr_t)d_in % 16 == 0) && ((uintptr_t)d_out % 16 == 0);
}

/* the tiled kernel with its cp.async loader: any width the reference accepts (cpu.h:46-48) from 32 up;
 * tight planes of such frames start on 4-byte boundaries when the frame does */
bool LdgstsPathUsable(int w, int h, const void *d_in, const void *d_out)
{
	(void)h;
	return w >= 32 && w <= 262136 && ((uintptr_t)d_in % 8 == 0) && ((uintptr_t)d_out % 8 == 0);
}

/* the scheduler words of the next launch on `stream` (EnqueueFilter explains the rotation); *slot for SchedLaunched */
static int SchedAcquire(hdb_context *ctx, cudaStream_t stream, void **sched, int *slot)
{
	*slot = (int)(ctx->sched_turn++ % kSchedSlots);
	*sched = static_cast<char *>(ctx->sched) + TiledSchedBytes() * *slot;
	if (ctx->sched_used[*slot] && ctx->sched_stream[*slot] != stream)
		HDB_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->sched_done[*slot], 0));
	return HDB_OK;
}

static int SchedLaunched(hdb_context *ctx, cudaStream_t stream, int slot)
{
	/* (measured: the record costs nothing on 0.2 .. 0.75 ms launches issued back to back, profiles/r02t_*) */
	HDB_CUDA(ctx, cudaEventRecord(ctx->sched_done[slot], stream));
	ctx->sched_used[slot] = true; ctx->sched_stream[slot] = stream;
	ctx->launches += 1;
	return HDB_OK;
}

/* 16-bit samples (view.bit_depth 8 .. 10): planar frames through DeblockTiled16Kernel; TMA only */
static int EnqueueFilter16(hdb_context *ctx, const FrameView &view, SizeState *ss, const unsigned char *d_in, unsigned char *d_out,
	int n, const unsigned *d_qp, unsigned qp_value, const unsigned char *d_vbs, long long vbs_stride,
	const unsigned char *d_hbs, long long hbs_stride, cudaStream_t stream, int *path, bool edge_qp)
{
	const int w = view.width, h = view.height, cw = w / 2, ch = h / 2;
	if (view.bit_depth < 8 || view.bit_depth > 10)
		return Fail(ctx, HDB_ERR_INVALID_ARG, "bit depth must be 8, 9 or 10 (16-bit containers)");
	if (view.strip_halo) return Fail(ctx, HDB_ERR_INVALID_ARG, "16-bit frames are not split into strips");
	if (view.nv12 && view.bit_depth != 10)
		return Fail(ctx, HDB_ERR_INVALID_ARG, "HDB_FLAG_NV12 with HDB_FLAG_DEPTH is P010: depth 10 only");
	if (w % 16 != 0 || w < 32 || ((uintptr_t)d_in | (uintptr_t)d_out) % 16 != 0 || view.luma_pitch % 16 != 0 || view.chroma_pitch % 16 != 0
		|| view.frame_stride % 16 != 0 || view.chroma_frame_stride % 16 != 0 || (view.u_offset | view.v_offset | view.u_offset_out | view.v_offset_out) % 16 != 0)
		return Fail(ctx, HDB_ERR_DIMENSIONS, "16-bit frames need a width that is a multiple of 16, at least 32, and 16-byte aligned buffers");
	if (path) *path = 1;
	const TiledConfig cfg = GetTiledConfig();
	TiledBatch b{};
	b.width = w; b.height = h;
	b.luma_nbx = (w + 8) / 8; b.luma_nby = (h + 8) / 8;
	b.chroma_nbx = (cw + 8) / 8; b.chroma_nby = (ch + 8) / 8;
	b.luma_row0 = view.luma_row0; b.luma_rows = view.luma_rows;
	b.chroma_row0 = view.chroma_row0; b.chroma_rows = view.chroma_rows;
	b.luma_y_org = view.luma_y_org; b.chroma_y_org = view.chroma_y_org;
	b.vert_bs_stride = vbs_stride; b.hor_bs_stride = hbs_stride;
	b.chroma_bs = ss->chroma_bs;
	b.qp_value = qp_value;
	b.nv12 = view.nv12 ? 1 : 0;
	b.n_frames = 1;
	FinalizeTiledBatch(b);
	const long long max_frames = std::max<long long>(1, ((1ll << 31) - (1 << 20)) / std::max(b.items_per_frame, 1));
	for (long long f0 = 0; f0 < n; f0 += max_frames) {
		const int nf = (int)std::min<long long>(max_frames, n - f0);
		b.n_frames = nf;
		FinalizeTiledBatch(b);
		b.qp = d_qp ? d_qp + f0 : nullptr;
		b.vert_bs = d_vbs + f0 * vbs_stride;
		b.hor_bs = d_hbs + f0 * hbs_stride;
		b.frame_stride = view.frame_stride; b.chroma_frame_stride = view.chroma_frame_stride;
		b.luma_pitch = view.luma_pitch; b.chroma_pitch = view.chroma_pitch;
		b.luma_buf_rows = view.luma_buf_rows; b.chroma_buf_rows = view.chroma_buf_rows;
		const long long off_in[3] = { 0, view.u_offset, view.v_offset }, off_out[3] = { 0, view.u_offset_out, view.v_offset_out };
		TensorMaps3 maps;
		memset(&maps, 0, sizeof maps);
		for (int p = 0; p < (view.nv12 ? 2 : 3); p++) {
			const long long pfs = p ? view.chroma_frame_stride : view.frame_stride;
			/* bytes per row: two per sample; the interleaved UV plane of a P010 frame is as wide as its luma plane */
			int rc = EncodePlaneMap(ctx, &maps.in[p], d_in + off_in[p] + f0 * pfs, 2 * ((p && !view.nv12) ? cw : w), p ? view.chroma_buf_rows : view.luma_buf_rows, nf,
				p ? view.chroma_pitch : view.luma_pitch, pfs, 8 * cfg.block_rows_per_item, 132u);
			if (rc) return rc;
			b.in_plane[p] = d_in + off_in[p] + f0 * pfs;
			b.out_plane[p] = d_out + off_out[p] + f0 * pfs;
		}
		void *sched; int slot;
		int rc = SchedAcquire(ctx, stream, &sched, &slot);
		if (rc) return rc;
		HDB_CUDA(ctx, LaunchTiled16(maps, b, sched, ctx->prop.multiProcessorCount, ctx->tiled16_ctas_per_sm, stream, view.bit_depth, edge_qp));
		if ((rc = SchedLaunched(ctx, stream, slot))) return rc;
	}
	return HDB_OK;
}

/*
 * Enqueue the filter for frames resident in device memory.  view describes which
 * rows the buffers hold (whole frames unless a strip).  No synchronisation.
 */
int EnqueueFilter(hdb_context *ctx, const FrameView &view, const unsigned char *d_in, unsigned char *d_out,
	int n, const unsigned *d_qp, unsigned qp_value, const unsigned char *d_vbs, long long vbs_stride,
	const unsigned char *d_hbs, long long hbs_stride, cudaStream_t stream, int *path, bool edge_qp)
{
	const int w = view.width, h = view.height;
	SizeState *ss = nullptr;
	int rc = GetSizeState(ctx, w, h, &ss);
	if (rc) return rc;
	if (edge_qp && (!d_vbs || !d_hbs)) return Fail(ctx, HDB_ERR_INVALID_ARG, "edge maps are required: there is no default for them");
	if (!d_vbs || !d_hbs) { d_vbs = ss->def_vbs; d_hbs = ss->def_hbs; vbs_stride = hbs_stride = 0; }
	const TiledConfig cfg = GetTiledConfig();

	if (view.bit_depth) return EnqueueFilter16(ctx, view, ss, d_in, d_out, n, d_qp, qp_value, d_vbs, vbs_stride, d_hbs, hbs_stride, stream, path, edge_qp);
	const bool tma_ok = TiledPathUsable(w, h, d_in, d_out) && view.luma_pitch % 16 == 0 && view.chroma_pitch % 16 == 0
		&& view.frame_stride % 16 == 0 && view.chroma_frame_stride % 16 == 0 && view.u_offset % 16 == 0 && view.v_offset % 16 == 0
		&& view.u_offset_out % 16 == 0 && view.v_offset_out % 16 == 0;
	const bool ldgsts_ok = !tma_ok && !view.strip_halo && ctx->use_ldgsts && LdgstsPathUsable(w, h, d_in, d_out);
	/* NV12 with a chroma width of 4 mod 8: the reference's skewed chroma blocks (deblock_kernels.cu: SkewToImage)
	 * straddle the interleaving; no kernel takes that */
	if (view.nv12 && (w % 16 != 0 || (!tma_ok && !ldgsts_ok)))
		return Fail(ctx, HDB_ERR_DIMENSIONS, "NV12 frames need a width that is a multiple of 16, at least 32, and 8-byte aligned buffers");
	if (!tma_ok && !ldgsts_ok && !view.whole_frame)
		return Fail(ctx, HDB_ERR_DIMENSIONS, "pitched or partial planes need a width of at least 32 and 4-byte aligned buffers");
	if (!tma_ok && !ldgsts_ok) {
		/* generic path: explicit padded planes in scratch memory */
		if (path) *path = 2;
		const size_t per = GenericScratchBytes(w, h);
		const int max_n = (int)std::max<size_t>(1, ((size_t)256 << 20) / per);
		const int chunk = std::min(n, max_n);
		/* the padded planes live in scratch memory that belongs to this call alone, allocated and freed in
		 * stream order: calls in flight on different streams never share it, and nothing synchronises */
		unsigned char *scratch = nullptr;
		HDB_CUDA(ctx, cudaMallocFromPoolAsync(reinterpret_cast<void **>(&scratch), per * chunk, ctx->pool, stream));
		const long long fs = (long long)w * h * 3 / 2;
		for (int f0 = 0; f0 < n; f0 += chunk) {
			GenericBatch g{};
			g.n_frames = std::min(chunk, n - f0); g.width = w; g.height = h;
			g.in = d_in + f0 * fs; g.out = d_out + f0 * fs; g.frame_stride = fs;
			g.scratch = scratch;
			g.qp = d_qp ? d_qp + f0 : nullptr; g.qp_value = qp_value;
			g.vert_bs = d_vbs + f0 * vbs_stride; g.hor_bs = d_hbs + f0 * hbs_stride;
			g.vert_bs_stride = vbs_stride; g.hor_bs_stride = hbs_stride;
			g.edge_qp = edge_qp;
			const cudaError_t e = LaunchGeneric(g, stream);
			if (e != cudaSuccess) { cudaFreeAsync(scratch, stream); return FailCuda(ctx, e, "LaunchGeneric"); }
			ctx->launches += 3;
		}
		HDB_CUDA(ctx, cudaFreeAsync(scratch, stream));
		return HDB_OK;
	}

	if (path) *path = tma_ok ? 1 : 3;
	const int cw = w / 2, ch = h / 2;
	TiledBatch b{};
	b.load_mode = tma_ok ? (view.strip_halo ? 1 : 0) : 2;
	for (int p = 0; p < 3; p++) { b.halo_top[p] = view.halo_top[p]; b.halo_bot[p] = view.halo_bot[p]; }
	b.width = w; b.height = h;
	b.luma_nbx = (w + 8) / 8; b.luma_nby = (h + 8) / 8;
	b.chroma_nbx = (cw + 8) / 8; b.chroma_nby = (ch + 8) / 8;
	b.luma_row0 = view.luma_row0; b.luma_rows = view.luma_rows;
	b.chroma_row0 = view.chroma_row0; b.chroma_rows = view.chroma_rows;
	b.luma_y_org = view.luma_y_org;
	b.chroma_y_org = view.chroma_y_org;
	b.vert_bs_stride = vbs_stride; b.hor_bs_stride = hbs_stride;
	b.chroma_bs = ss->chroma_bs;
	b.qp_value = qp_value;
	b.nv12 = view.nv12 ? 1 : 0;
	/* frames per launch bounded by the 31-bit item counter */
	b.n_frames = 1;
	FinalizeTiledBatch(b);
	const long long max_frames = std::max<long long>(1, ((1ll << 31) - (1 << 20)) / std::max(b.items_per_frame, 1));
	const long long fs = view.frame_stride, cfs = view.chroma_frame_stride;
	for (long long f0 = 0; f0 < n; f0 += max_frames) {
		const int nf = (int)std::min<long long>(max_frames, n - f0);
		b.n_frames = nf;
		FinalizeTiledBatch(b);
		b.qp = d_qp ? d_qp + f0 : nullptr;
		b.vert_bs = d_vbs + f0 * vbs_stride;
		b.hor_bs = d_hbs + f0 * hbs_stride;
		TensorMaps3 maps;
		b.frame_stride = fs; b.chroma_frame_stride = cfs;
		b.luma_pitch = view.luma_pitch; b.chroma_pitch = view.chroma_pitch;
		b.luma_buf_rows = view.luma_buf_rows; b.chroma_buf_rows = view.chroma_buf_rows;
		const long long off_in[3] = { 0, view.u_offset, view.v_offset };
		const long long off_out[3] = { 0, view.u_offset_out, view.v_offset_out };
		const int pw[3] = { w, view.nv12 ? w : cw, cw };     /* bytes per row: the interleaved UV plane is as wide as luma */
		const int ph[3] = { view.luma_buf_rows, view.chroma_buf_rows, view.chroma_buf_rows };
		memset(&maps, 0, sizeof maps);
		for (int p = 0; p < (view.nv12 ? 2 : 3); p++) {
			const long long pfs = p ? cfs : fs;
			if (tma_ok && (rc = EncodePlaneMap(ctx, &maps.in[p], d_in + off_in[p] + f0 * pfs, pw[p], ph[p], nf,
					p ? view.chroma_pitch : view.luma_pitch, pfs, 8 * cfg.block_rows_per_item))) return rc;
			b.in_plane[p] = d_in + off_in[p] + f0 * pfs;
			b.out_plane[p] = d_out + off_out[p] + f0 * pfs;
		}
		/* launches take the scheduler words in turn: two launches that overlap on the device (different
		 * streams) must not share a work counter, and each launch leaves its own words zeroed */
		/* the launch that had these words kSchedSlots launches ago may still be running on another stream: the new
		 * one waits for it there (the same stream orders itself) */
		void *sched; int slot;
		if ((rc = SchedAcquire(ctx, stream, &sched, &slot))) return rc;
		HDB_CUDA(ctx, LaunchTiled(maps, b, sched, ctx->prop.multiProcessorCount, ctx->tiled_ctas_per_sm, stream, edge_qp));
		if ((rc = SchedLaunched(ctx, stream, slot))) return rc;
	}
	return HDB_OK;
}

FrameView WholeFrameView(int w, int h, bool nv12, int bit_depth)
{
	FrameView v{};
	v.nv12 = nv12;
	v.bit_depth = bit_depth;
	v.width = w; v.height = h; v.whole_frame = true;
	v.luma_row0 = 0; v.luma_rows = (h + 8) / 8;
	v.chroma_row0 = 0; v.chroma_rows = (h / 2 + 8) / 8;
	v.luma_buf_rows = h; v.chroma_buf_rows = h / 2;
	v.luma_y_org = -4; v.chroma_y_org = -4;
	v.u_offset = (long long)w * h; v.v_offset = (long long)w * h + (long long)(w / 2) * (h / 2);
	v.u_offset_out = v.u_offset; v.v_offset_out = v.v_offset;
	v.luma_pitch = w; v.chroma_pitch = nv12 ? w : w / 2;
	v.frame_stride = v.chroma_frame_stride = (long long)w * h * 3 / 2;
	if (bit_depth) {      /* two bytes a sample: every byte quantity doubles */
		v.u_offset *= 2; v.v_offset *= 2; v.u_offset_out *= 2; v.v_offset_out *= 2;
		v.luma_pitch *= 2; v.chroma_pitch *= 2; v.frame_stride *= 2; v.chroma_frame_stride *= 2;
	}
	return v;
}

int ValidateFrameArgs(hdb_context *ctx, unsigned w, unsigned h, const unsigned char *vbs, unsigned nv,
	const unsigned char *hbs, unsigned nh, bool sizes_given)
{
	if (w == 0 || h == 0 || w % 8 != 0 || h % 8 != 0) return Fail(ctx, HDB_ERR_DIMENSIONS, kMsgDims);
	/* the item descriptor of the tiled kernel numbers 32-block chunks in 10 bits and row groups in 20; the
	 * reference's own unsigned array sizes (cpu.h:86-87) must not wrap either */
	if (w > 262136u || h > (1u << 20) || (unsigned long long)(w / 8 + 1) * h >= (1ull << 32)
		|| (unsigned long long)(h / 8 + 1) * w >= (1ull << 32) || (unsigned long long)w * h * 3 / 2 >= (1ull << 40))
		return Fail(ctx, HDB_ERR_INVALID_ARG, "frame too large (width up to 262136, height up to 1048576, (width/8+1)*height below 2^32)");
	if ((vbs != nullptr) != (hbs != nullptr))
		return Fail(ctx, HDB_ERR_INVALID_ARG, "vert_bs and hor_bs must both be given or both be NULL");
	if (vbs && sizes_given && (nv != (w / 8 + 1) * h / 8 || nh != (h / 8 + 1) * w / 8))
		return Fail(ctx, HDB_ERR_BS_SIZE, kMsgBsSize);
	return HDB_OK;
}

/* ------------------------------------------------------------------ */
/* pinned multi-stream pipeline                                        */
/* ------------------------------------------------------------------ */
static int EnsureSlots(hdb_context *ctx, size_t frame_bytes, int chunk, unsigned nv, unsigned nh, bool with_bs,
	size_t info_bytes_per_frame)
{
	for (auto &sl : ctx->slots) {
		if (info_bytes_per_frame && sl.info_cap < info_bytes_per_frame * chunk) {
			cudaFree(sl.info); sl.info = nullptr; sl.info_cap = 0;
			HDB_CUDA(ctx, cudaMalloc(&sl.info, info_bytes_per_frame * chunk));
			sl.info_cap = info_bytes_per_frame * chunk;
		}
		if (sl.frames_cap < frame_bytes * chunk) {
			cudaFree(sl.frames); sl.frames = nullptr; sl.frames_cap = 0;
			HDB_CUDA(ctx, cudaMalloc(&sl.frames, frame_bytes * chunk));
			sl.frames_cap = frame_bytes * chunk;
		}
		if (with_bs && sl.vbs_cap < (size_t)nv * chunk) {
			cudaFree(sl.vbs); sl.vbs = nullptr; sl.vbs_cap = 0;
			HDB_CUDA(ctx, cudaMalloc(&sl.vbs, std::max<size_t>((size_t)nv * chunk, 1)));
			sl.vbs_cap = (size_t)nv * chunk;
		}
		if (with_bs && sl.hbs_cap < (size_t)nh * chunk) {
			cudaFree(sl.hbs); sl.hbs = nullptr; sl.hbs_cap = 0;
			HDB_CUDA(ctx, cudaMalloc(&sl.hbs, std::max<size_t>((size_t)nh * chunk, 1)));
			sl.hbs_cap = (size_t)nh * chunk;
		}
	}
	return HDB_OK;
}

constexpr size_t kZeroCopyMaxBytes = (size_t)4 << 20;

/* device-side address of a host buffer that is page-locked and mapped, or false */
static bool MappedDevicePointer(const void *host, unsigned char **dev)
{
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return false; }
	if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
	*dev = static_cast<unsigned char *>(a.devicePointer);
	return true;
}

/*
 * Slice plan of a batch of n frames, max_f frames per full slice.  Slices grow geometrically
 * from max_f/8 at the head and shrink again at the tail: the first H2D copy and the last D2H
 * copy are the only ones nothing overlaps, so they are kept short, while the body uses large
 * slices (per-copy and per-launch overheads cost ~10 % of the link rate at 32 MB, ~1 % at 128 MB).
 */
static std::vector<int> PlanSlices(unsigned n, int max_f)
{
	std::vector<int> head, tail;
	long long rem = n;
	int k = std::max(1, max_f / 8);
	while (rem > 0) {
		const int size = std::min(k, max_f);
		if (rem <= 2ll * size) {
			const int a = (int)((rem + 1) / 2), b2 = (int)(rem - a);
			head.push_back(a);
			if (b2) tail.push_back(b2);
			break;
		}
		head.push_back(size); tail.push_back(size);
		rem -= 2ll * size;
		if (k < max_f) k *= 2;
	}
	head.insert(head.end(), tail.rbegin(), tail.rend());
	return head;
}

static int GetEvents(hdb_context *ctx, size_t n)
{
	while (ctx->event_pool.size() < n) {
		cudaEvent_t ev;
		HDB_CUDA(ctx, cudaEventCreate(&ev));
		ctx->event_pool.push_back(ev);
	}
	return HDB_OK;
}

/*
 * Host buffers in, host buffers out.  The batch is cut into slices (PlanSlices) that
 * cycle through kSlots device slots: slice i's H2D copy, slice i-1's kernel and
 * slice i-2's D2H copy run concurrently on three streams, chained by events.  A
 * single slice runs on one stream.
 */
/* On any failure after work has been enqueued, nothing may still be in flight when the caller gets its
 * buffers back (and before they are un-pinned): declared after the PinHostRange objects, destroyed before them. */
struct DrainOnFailure {
	hdb_context *ctx;
	bool armed = true;
	explicit DrainOnFailure(hdb_context *c) : ctx(c) {}
	~DrainOnFailure()
	{
		if (!armed) return;
		cudaStreamSynchronize(ctx->s_h2d); cudaStreamSynchronize(ctx->s_k); cudaStreamSynchronize(ctx->s_d2h);
		cudaGetLastError();
	}
};

int BatchHost(hdb_context *ctx, const unsigned char *in, unsigned char *out, unsigned n, unsigned w, unsigned h,
	const unsigned *qp, const unsigned char *vbs, const unsigned char *hbs, hdb_timing *timing, unsigned flags,
	const hdb_block_info *info)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lock(ctx->mu);
	return BatchHostLocked(ctx, in, out, n, w, h, qp, vbs, hbs, timing, flags, info);
}

int BatchHostLocked(hdb_context *ctx, const unsigned char *in, unsigned char *out, unsigned n, unsigned w, unsigned h,
	const unsigned *qp, const unsigned char *vbs, const unsigned char *hbs, hdb_timing *timing, unsigned flags,
	const hdb_block_info *info)
{
	/* block data given: the edge maps of every slice are derived on the device, into the slot's own arrays */
	const bool edge_qp = (flags & HDB_FLAG_EDGE_MAPS) != 0 || info != nullptr, nv12 = (flags & HDB_FLAG_NV12) != 0;
	const int depth = DepthOfFlags(flags);
	if ((flags & HDB_FLAG_SAMPLES16) && (depth < 8 || depth > 10)) return Fail(ctx, HDB_ERR_INVALID_ARG, "bit depth must be 8, 9 or 10 (HDB_FLAG_DEPTH)");
	if (depth && nv12 && depth != 10) return Fail(ctx, HDB_ERR_INVALID_ARG, "HDB_FLAG_NV12 with HDB_FLAG_DEPTH is P010: depth 10 only");
	if (!in || !out || !qp || n == 0) return Fail(ctx, HDB_ERR_INVALID_ARG, "NULL buffer or empty batch");
	if (info && (vbs || hbs)) return Fail(ctx, HDB_ERR_INVALID_ARG, "give block data or edge maps, not both");
	if (edge_qp && !info && (!vbs || !hbs)) return Fail(ctx, HDB_ERR_INVALID_ARG, "edge maps are required: there is no default for them");
	int rc = ValidateFrameArgs(ctx, w, h, vbs, 0, hbs, 0, false);
	if (rc) return rc;
	DeviceGuard guard(ctx->device);
	const auto t_begin = std::chrono::steady_clock::now();
	const size_t fb = (size_t)w * h * 3 / 2 * (depth ? 2 : 1);
	const unsigned nv = (w / 8 + 1) * h / 8, nh = (h / 8 + 1) * w / 8;
	const int chunk = (int)std::min<size_t>(n, std::max<size_t>(1, ((size_t)ctx->chunk_bytes) / fb));
	const std::vector<int> slices = PlanSlices(n, chunk);
	const int nchunks = (int)slices.size();
	const size_t ib = info ? sizeof(hdb_block_info) * (size_t)(w / 8) * (h / 8) : 0;     /* block data per frame */
	const bool with_bs = vbs != nullptr || info != nullptr;
	if ((rc = EnsureSlots(ctx, fb, chunk, nv, nh, with_bs, ib))) return rc;
	const bool uniform_qp = (n == 1);
	if (!uniform_qp && ctx->qp_cap < n) {
		cudaFree(ctx->qp_dev); ctx->qp_dev = nullptr; ctx->qp_cap = 0;
		HDB_CUDA(ctx, cudaMalloc(&ctx->qp_dev, sizeof(unsigned) * n));
		ctx->qp_cap = n;
	}
	const FrameView view = WholeFrameView((int)w, (int)h, nv12, depth);
	if (depth && (w % 16 != 0 || w < 32)) return Fail(ctx, HDB_ERR_DIMENSIONS, "16-bit frames need a width that is a multiple of 16 and at least 32");
	if (nv12 && (w < 32 || w % 16 != 0)) return Fail(ctx, HDB_ERR_DIMENSIONS, "NV12 frames need a width that is a multiple of 16 and at least 32");

	PinHostRange pin_in(in, fb * n), pin_out(out != in ? out : nullptr, fb * n);
	PinHostRange pin_v(vbs, (size_t)nv * n), pin_h(hbs, (size_t)nh * n), pin_i(info, ib * n);
	const unsigned char *info_bytes = reinterpret_cast<const unsigned char *>(info);
	DrainOnFailure drain(ctx);

	const bool timed = timing != nullptr;
	if (timed && (rc = GetEvents(ctx, (size_t)nchunks * 6))) return rc;
	const bool single = nchunks == 1;
	cudaStream_t s_h2d = single ? ctx->s_k : ctx->s_h2d, s_k = ctx->s_k, s_d2h = single ? ctx->s_k : ctx->s_d2h;
	int path = 0;
	unsigned long long launches0 = ctx->launches;

	/* every frame's QP in one copy, ahead of the first slice on the same stream */
	if (!uniform_qp)
		HDB_CUDA(ctx, cudaMemcpyAsync(ctx->qp_dev, qp, sizeof(unsigned) * n, cudaMemcpyHostToDevice, s_h2d));

	/*
	 * One slice: there is no second slice to overlap a copy with, so H2D, kernel and D2H would run
	 * back to back.  When both host buffers are page-locked and mapped, the kernel works on them
	 * directly instead -- TMA loads and vector stores over the link -- and the two transfers and
	 * the arithmetic overlap inside the one launch.  (Boundary strengths still go through a copy:
	 * the kernel reads them bytewise.)  Measured on B200 / PCIe Gen5 (tools/latency.py): 75 -> 47 us
	 * for 352x288, 109 -> 67 us for 768x576, 233 -> 178 us for 1080p; but SM-issued transfers reach
	 * only ~19 GB/s per direction against the copy engines' ~50, so a 4K frame loses (523 -> 668 us):
	 * hence the size limit.
	 */
	unsigned char *z_in = nullptr, *z_out = nullptr;
	if (single && ctx->zero_copy && fb * n <= kZeroCopyMaxBytes
		&& MappedDevicePointer(in, &z_in) && MappedDevicePointer(out, &z_out)) {
		Slot &sl = ctx->slots[0];
		if (vbs) {
			HDB_CUDA(ctx, cudaMemcpyAsync(sl.vbs, vbs, (size_t)nv * n, cudaMemcpyHostToDevice, s_k));
			HDB_CUDA(ctx, cudaMemcpyAsync(sl.hbs, hbs, (size_t)nh * n, cudaMemcpyHostToDevice, s_k));
		}
		if (info) {
			HDB_CUDA(ctx, cudaMemcpyAsync(sl.info, info_bytes, ib * n, cudaMemcpyHostToDevice, s_k));
			HDB_CUDA(ctx, LaunchDeriveEdges(sl.info, (int)w, (int)h, (int)n, sl.vbs, sl.hbs, s_k));
			ctx->launches += 1;
		}
		cudaEvent_t *ev = timed ? &ctx->event_pool[0] : nullptr;
		if (timed) HDB_CUDA(ctx, cudaEventRecord(ev[0], s_k));
		rc = EnqueueFilter(ctx, view, z_in, z_out, (int)n, uniform_qp ? nullptr : ctx->qp_dev, qp[0],
			with_bs ? sl.vbs : nullptr, nv, with_bs ? sl.hbs : nullptr, nh, s_k, &path, edge_qp);
		if (rc) { cudaDeviceSynchronize(); return rc; }
		if (timed) HDB_CUDA(ctx, cudaEventRecord(ev[1], s_k));
		HDB_CUDA(ctx, cudaStreamSynchronize(s_k));
		const auto t_done = std::chrono::steady_clock::now();
		if (timed) {
			memset(timing, 0, sizeof *timing);
			float k = 0;
			cudaEventElapsedTime(&k, ev[0], ev[1]);
			timing->kernel_ms = k;                       /* includes both transfers: they happen inside the kernel */
			timing->copy_ms = 0;
			timing->total_ms = std::chrono::duration<double, std::milli>(t_done - t_begin).count();
			timing->h2d_bytes = (unsigned long long)fb * n + (n == 1 ? 0 : sizeof(unsigned) * (unsigned long long)n)
				+ (vbs ? (unsigned long long)(nv + nh) * n : 0) + (unsigned long long)ib * n;
			timing->d2h_bytes = (unsigned long long)fb * n;
			timing->kernel_launches = (int)(ctx->launches - launches0);
			timing->path = path;
		}
		drain.armed = false;
		return HDB_OK;
	}
	unsigned f0 = 0;
	for (int c = 0; c < nchunks; f0 += (unsigned)slices[c], c++) {
		Slot &sl = ctx->slots[c % kSlots];
		const int nf = slices[c];
		cudaEvent_t *ev = timed ? &ctx->event_pool[(size_t)c * 6] : nullptr;
		if (!sl.ev_ready) {
			HDB_CUDA(ctx, cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
			HDB_CUDA(ctx, cudaEventCreateWithFlags(&sl.k_done, cudaEventDisableTiming));
			HDB_CUDA(ctx, cudaEventCreateWithFlags(&sl.d2h_done, cudaEventDisableTiming));
			sl.ev_ready = true;
		}
		/* the slot is free once the D2H copy of the slice that used it last has finished */
		if (!single && c >= kSlots) HDB_CUDA(ctx, cudaStreamWaitEvent(s_h2d, sl.d2h_done, 0));
		if (timed) HDB_CUDA(ctx, cudaEventRecord(ev[0], s_h2d));
		HDB_CUDA(ctx, cudaMemcpyAsync(sl.frames, in + (size_t)f0 * fb, fb * nf, cudaMemcpyHostToDevice, s_h2d));
		if (vbs) {
			HDB_CUDA(ctx, cudaMemcpyAsync(sl.vbs, vbs + (size_t)f0 * nv, (size_t)nv * nf, cudaMemcpyHostToDevice, s_h2d));
			HDB_CUDA(ctx, cudaMemcpyAsync(sl.hbs, hbs + (size_t)f0 * nh, (size_t)nh * nf, cudaMemcpyHostToDevice, s_h2d));
		}
		if (info) HDB_CUDA(ctx, cudaMemcpyAsync(sl.info, info_bytes + (size_t)f0 * ib, ib * nf, cudaMemcpyHostToDevice, s_h2d));
		if (timed) HDB_CUDA(ctx, cudaEventRecord(ev[1], s_h2d));
		if (!single) {
			HDB_CUDA(ctx, cudaEventRecord(sl.h2d_done, s_h2d));
			HDB_CUDA(ctx, cudaStreamWaitEvent(s_k, sl.h2d_done, 0));
		}
		if (timed) HDB_CUDA(ctx, cudaEventRecord(ev[2], s_k));
		if (info) {
			HDB_CUDA(ctx, LaunchDeriveEdges(sl.info, (int)w, (int)h, nf, sl.vbs, sl.hbs, s_k));
			ctx->launches += 1;
		}
		rc = EnqueueFilter(ctx, view, sl.frames, sl.frames, nf, uniform_qp ? nullptr : ctx->qp_dev + f0, qp[0],
			with_bs ? sl.vbs : nullptr, nv, with_bs ? sl.hbs : nullptr, nh, s_k, &path, edge_qp);
		if (rc) { cudaDeviceSynchronize(); return rc; }
		if (timed) HDB_CUDA(ctx, cudaEventRecord(ev[3], s_k));
		if (!single) {
			HDB_CUDA(ctx, cudaEventRecord(sl.k_done, s_k));
			HDB_CUDA(ctx, cudaStreamWaitEvent(s_d2h, sl.k_done, 0));
		}
		if (timed) HDB_CUDA(ctx, cudaEventRecord(ev[4], s_d2h));
		HDB_CUDA(ctx, cudaMemcpyAsync(out + (size_t)f0 * fb, sl.frames, fb * nf, cudaMemcpyDeviceToHost, s_d2h));
		if (timed) HDB_CUDA(ctx, cudaEventRecord(ev[5], s_d2h));
		if (!single) HDB_CUDA(ctx, cudaEventRecord(sl.d2h_done, s_d2h));
	}
	HDB_CUDA(ctx, cudaStreamSynchronize(s_d2h));
	HDB_CUDA(ctx, cudaStreamSynchronize(s_k));
	const auto t_end = std::chrono::steady_clock::now();
	if (timed) {
		memset(timing, 0, sizeof *timing);
		for (int c = 0; c < nchunks; c++) {
			cudaEvent_t *ev = &ctx->event_pool[(size_t)c * 6];
			float a = 0, k = 0, d = 0;
			cudaEventElapsedTime(&a, ev[0], ev[1]);
			cudaEventElapsedTime(&k, ev[2], ev[3]);
			cudaEventElapsedTime(&d, ev[4], ev[5]);
			timing->copy_ms += a + d;
			timing->kernel_ms += k;
		}
		timing->total_ms = std::chrono::duration<double, std::milli>(t_end - t_begin).count();
		timing->h2d_bytes = (unsigned long long)fb * n + (n == 1 ? 0 : sizeof(unsigned) * (unsigned long long)n)
			+ (vbs ? (unsigned long long)(nv + nh) * n : 0) + (unsigned long long)ib * n;
		timing->d2h_bytes = (unsigned long long)fb * n;
		timing->kernel_launches = (int)(ctx->launches - launches0);
		timing->path = path;
	}
	drain.armed = false;
	return HDB_OK;
}

/*
 * One frame given as three separate planes (the reference keeps ext_Y / ext_U / ext_V, cpu.h:1041-1048): the
 * planes go straight into one device slot with three copies and come back with three, on one stream, under
 * the context lock from the first copy to the last.
 */
int PlanesHost(hdb_context *ctx, const unsigned char *y_in, const unsigned char *u_in, const unsigned char *v_in,
	unsigned char *y_out, unsigned char *u_out, unsigned char *v_out, unsigned w, unsigned h, unsigned qp,
	const unsigned char *vbs, const unsigned char *hbs, hdb_timing *timing)
{
	std::lock_guard<std::mutex> lock(ctx->mu);
	DeviceGuard guard(ctx->device);
	const auto t_begin = std::chrono::steady_clock::now();
	const size_t ysz = (size_t)w * h, csz = (size_t)(w / 2) * (h / 2), fb = ysz + 2 * csz;
	const unsigned nv = (w / 8 + 1) * h / 8, nh = (h / 8 + 1) * w / 8;
	int rc = EnsureSlots(ctx, fb, 1, nv, nh, vbs != nullptr, 0);
	if (rc) return rc;
	if (timing && (rc = GetEvents(ctx, 2))) return rc;
	Slot &sl = ctx->slots[0];
	cudaStream_t st = ctx->s_k;
	DrainOnFailure drain(ctx);
	const unsigned long long launches0 = ctx->launches;
	HDB_CUDA(ctx, cudaMemcpyAsync(sl.frames, y_in, ysz, cudaMemcpyHostToDevice, st));
	HDB_CUDA(ctx, cudaMemcpyAsync(sl.frames + ysz, u_in, csz, cudaMemcpyHostToDevice, st));
	HDB_CUDA(ctx, cudaMemcpyAsync(sl.frames + ysz + csz, v_in, csz, cudaMemcpyHostToDevice, st));
	if (vbs) {
		HDB_CUDA(ctx, cudaMemcpyAsync(sl.vbs, vbs, nv, cudaMemcpyHostToDevice, st));
		HDB_CUDA(ctx, cudaMemcpyAsync(sl.hbs, hbs, nh, cudaMemcpyHostToDevice, st));
	}
	if (timing) HDB_CUDA(ctx, cudaEventRecord(ctx->event_pool[0], st));
	int path = 0;
	rc = EnqueueFilter(ctx, WholeFrameView((int)w, (int)h), sl.frames, sl.frames, 1, nullptr, qp,
		vbs ? sl.vbs : nullptr, nv, vbs ? sl.hbs : nullptr, nh, st, &path, false);
	if (rc) return rc;
	if (timing) HDB_CUDA(ctx, cudaEventRecord(ctx->event_pool[1], st));
	HDB_CUDA(ctx, cudaMemcpyAsync(y_out, sl.frames, ysz, cudaMemcpyDeviceToHost, st));
	HDB_CUDA(ctx, cudaMemcpyAsync(u_out, sl.frames + ysz, csz, cudaMemcpyDeviceToHost, st));
	HDB_CUDA(ctx, cudaMemcpyAsync(v_out, sl.frames + ysz + csz, csz, cudaMemcpyDeviceToHost, st));
	HDB_CUDA(ctx, cudaStreamSynchronize(st));
	drain.armed = false;
	if (timing) {
		memset(timing, 0, sizeof *timing);
		float k = 0;
		cudaEventElapsedTime(&k, ctx->event_pool[0], ctx->event_pool[1]);
		timing->kernel_ms = k;
		timing->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
		timing->copy_ms = std::max(0.0, timing->total_ms - timing->kernel_ms);
		timing->h2d_bytes = fb + (vbs ? nv + nh : 0);
		timing->d2h_bytes = fb;
		timing->kernel_launches = (int)(ctx->launches - launches0);
		timing->path = path;
	}
	return HDB_OK;
}

} // namespace hdb

/* ------------------------------------------------------------------ */
/* C ABI                                                               */
/* ------------------------------------------------------------------ */
using namespace hdb;

extern "C" {

int hdb_create(hdb_context **ctx, int device) { return CreateContext(ctx, device); }
void hdb_destroy(hdb_context *ctx) { DestroyContext(ctx); }
const char *hdb_last_error(const hdb_context *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

const char *hdb_status_string(int status)
{
	switch (status) {
	case HDB_OK: return "ok";
	case HDB_ERR_INVALID_ARG: return "invalid argument";
	case HDB_ERR_FILE_SIZE: return kMsgFileSize;
	case HDB_ERR_DIMENSIONS: return kMsgDims;
	case HDB_ERR_BS_SIZE: return kMsgBsSize;
	case HDB_ERR_CUDA: return "CUDA error";
	case HDB_ERR_IO: return "file I/O error";
	case HDB_ERR_NO_MEMORY: return "out of memory";
	}
	return "unknown status";
}

int hdb_get_device_info(const hdb_context *ctx, hdb_device_info *info)
{
	if (!ctx || !info) return HDB_ERR_INVALID_ARG;
	memset(info, 0, sizeof *info);
	snprintf(info->name, sizeof info->name, "%s", ctx->prop.name);
	info->cc_major = ctx->prop.major; info->cc_minor = ctx->prop.minor;
	info->sm_count = ctx->prop.multiProcessorCount; info->warp_size = ctx->prop.warpSize;
	info->max_threads_per_block = ctx->prop.maxThreadsPerBlock;
	info->total_global_mem = ctx->prop.totalGlobalMem; info->total_const_mem = ctx->prop.totalConstMem;
	info->shared_mem_per_block = ctx->prop.sharedMemPerBlock;
	return HDB_OK;
}

unsigned hdb_num_vert_bs(unsigned width, unsigned height) { return (width / 8 + 1) * height / 8; }
unsigned hdb_num_hor_bs(unsigned width, unsigned height) { return (height / 8 + 1) * width / 8; }

int hdb_default_boundary_strength(unsigned width, unsigned height, unsigned char *vert_bs, unsigned char *hor_bs)
{
	if (width % 8 != 0 || height % 8 != 0) return HDB_ERR_DIMENSIONS;
	const unsigned nv = hdb_num_vert_bs(width, height), nh = hdb_num_hor_bs(width, height);
	if (vert_bs) for (unsigned i = 0; i < nv; i++) vert_bs[i] = (i % (width / 8 + 1) == 0) ? 0 : 2;   /* cpu.h:92-95 */
	if (hor_bs) for (unsigned i = 0; i < nh; i++) hor_bs[i] = (i % (height / 8 + 1) == 0) ? 0 : 2;    /* cpu.h:96-99 */
	return HDB_OK;
}

int hdb_beta(unsigned qp) { return HostBeta(qp); }
int hdb_tc(unsigned qp) { return HostTc(qp); }

void *hdb_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return p;
}

void hdb_host_free(void *p) { if (p) cudaFreeHost(p); }

int hdb_deblock_frame(hdb_context *ctx, const unsigned char *yuv_in, unsigned char *yuv_out,
	unsigned width, unsigned height, unsigned qp, const unsigned char *vert_bs, unsigned num_vert_bs,
	const unsigned char *hor_bs, unsigned num_hor_bs, hdb_timing *timing)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	int rc = ValidateFrameArgs(ctx, width, height, vert_bs, num_vert_bs, hor_bs, num_hor_bs, true);
	if (rc) return rc;
	return BatchHost(ctx, yuv_in, yuv_out, 1, width, height, &qp, vert_bs, hor_bs, timing);
}

int hdb_deblock_planes(hdb_context *ctx, const unsigned char *y_in, const unsigned char *u_in, const unsigned char *v_in,
	unsigned char *y_out, unsigned char *u_out, unsigned char *v_out, unsigned width, unsigned height, unsigned qp,
	const unsigned char *vert_bs, unsigned num_vert_bs, const unsigned char *hor_bs, unsigned num_hor_bs,
	hdb_timing *timing)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	if (!y_in || !u_in || !v_in || !y_out || !u_out || !v_out) return Fail(ctx, HDB_ERR_INVALID_ARG, "NULL plane");
	int rc = ValidateFrameArgs(ctx, width, height, vert_bs, num_vert_bs, hor_bs, num_hor_bs, true);
	if (rc) return rc;
	return PlanesHost(ctx, y_in, u_in, v_in, y_out, u_out, v_out, width, height, qp, vert_bs, hor_bs, timing);
}

int hdb_execute_gpu(hdb_context *ctx, const char *input_file_name, const char *output_file_name,
	unsigned width, unsigned height, unsigned qp, hdb_timing *timing)
{
	if (!ctx || !input_file_name || !output_file_name) return HDB_ERR_INVALID_ARG;
	FILE *f = fopen(input_file_name, "rb");
	if (!f) return Fail(ctx, HDB_ERR_IO, std::string("cannot open ") + input_file_name);
	fseek(f, 0, SEEK_END);
	const long length = ftell(f);
	fseek(f, 0, SEEK_SET);
	const size_t fb = (size_t)3 * width * height / 2;
	if (length < 0 || (unsigned)length != 3 * width * height / 2) { fclose(f); return Fail(ctx, HDB_ERR_FILE_SIZE, kMsgFileSize); } /* cpu.h:43 */
	if (width % 8 != 0 || height % 8 != 0) { fclose(f); return Fail(ctx, HDB_ERR_DIMENSIONS, kMsgDims); }                            /* cpu.h:46 */
	unsigned char *buf = static_cast<unsigned char *>(hdb_host_alloc(fb));                                                              /* gpu.cu:1103 */
	if (!buf) { fclose(f); return Fail(ctx, HDB_ERR_NO_MEMORY, "cudaMallocHost failed"); }
	const bool read_ok = fread(buf, 1, fb, f) == fb;
	fclose(f);
	int rc = read_ok ? HDB_OK : Fail(ctx, HDB_ERR_IO, "short read");
	if (!rc) rc = BatchHost(ctx, buf, buf, 1, width, height, &qp, nullptr, nullptr, timing);
	if (!rc) {
		FILE *o = fopen(output_file_name, "wb");
		if (!o || fwrite(buf, 1, fb, o) != fb) rc = Fail(ctx, HDB_ERR_IO, std::string("cannot write ") + output_file_name);
		if (o) fclose(o);
	}
	hdb_host_free(buf);
	return rc;
}

int hdb_deblock_batch(hdb_context *ctx, const unsigned char *frames_in, unsigned char *frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *qp,
	const unsigned char *vert_bs, const unsigned char *hor_bs, hdb_timing *timing)
{
	return BatchHost(ctx, frames_in, frames_out, n_frames, width, height, qp, vert_bs, hor_bs, timing);
}

int hdb_deblock_batch_device(hdb_context *ctx, const void *d_frames_in, void *d_frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *d_qp,
	const unsigned char *d_vert_bs, const unsigned char *d_hor_bs, void *stream)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lock(ctx->mu);
	if (!d_frames_in || !d_frames_out || !d_qp || n_frames == 0) return Fail(ctx, HDB_ERR_INVALID_ARG, "NULL buffer or empty batch");
	int rc = ValidateFrameArgs(ctx, width, height, d_vert_bs, 0, d_hor_bs, 0, false);
	if (rc) return rc;
	DeviceGuard guard(ctx->device);
	const FrameView view = WholeFrameView((int)width, (int)height);
	return EnqueueFilter(ctx, view, static_cast<const unsigned char *>(d_frames_in),
		static_cast<unsigned char *>(d_frames_out), (int)n_frames, d_qp, 0,
		d_vert_bs, hdb_num_vert_bs(width, height), d_hor_bs, hdb_num_hor_bs(width, height),
		static_cast<cudaStream_t>(stream), nullptr);
}

int hdb_deblock_surfaces_device(hdb_context *ctx, const hdb_surfaces *in, const hdb_surfaces *out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *d_qp,
	const unsigned char *d_vert_bs, const unsigned char *d_hor_bs, void *stream)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lock(ctx->mu);
	if (!in || !out || !d_qp || n_frames == 0 || !in->y || !in->u || !in->v || !out->y || !out->u || !out->v)
		return Fail(ctx, HDB_ERR_INVALID_ARG, "NULL surface, plane or QP pointer, or empty batch");
	int rc = ValidateFrameArgs(ctx, width, height, d_vert_bs, 0, d_hor_bs, 0, false);
	if (rc) return rc;
	if (width % 32 != 0) return Fail(ctx, HDB_ERR_DIMENSIONS, "pitched surfaces need a width that is a multiple of 32");
	const hdb_surfaces *both[2] = { in, out };
	for (const hdb_surfaces *s : both) {
		const uintptr_t a = (uintptr_t)s->y | (uintptr_t)s->u | (uintptr_t)s->v | s->pitch_y | s->pitch_c
			| (n_frames > 1 ? s->frame_stride_y | s->frame_stride_c : 0);
		if (a % 16 != 0) return Fail(ctx, HDB_ERR_INVALID_ARG, "surface addresses, pitches and frame strides must be multiples of 16 bytes");
		if (s->pitch_y < width || s->pitch_c < width / 2 || s->pitch_y > (1u << 30) || s->pitch_c > (1u << 30))
			return Fail(ctx, HDB_ERR_INVALID_ARG, "pitch smaller than the plane width");
	}
	if (in->pitch_y != out->pitch_y || in->pitch_c != out->pitch_c
		|| (n_frames > 1 && (in->frame_stride_y != out->frame_stride_y || in->frame_stride_c != out->frame_stride_c)))
		return Fail(ctx, HDB_ERR_INVALID_ARG, "input and output surfaces must share pitches and frame strides");
	DeviceGuard guard(ctx->device);
	FrameView view = WholeFrameView((int)width, (int)height);
	const unsigned char *iy = static_cast<const unsigned char *>(in->y);
	unsigned char *oy = static_cast<unsigned char *>(out->y);
	view.u_offset = static_cast<const unsigned char *>(in->u) - iy;
	view.v_offset = static_cast<const unsigned char *>(in->v) - iy;
	view.u_offset_out = static_cast<unsigned char *>(out->u) - oy;
	view.v_offset_out = static_cast<unsigned char *>(out->v) - oy;
	view.luma_pitch = (int)in->pitch_y; view.chroma_pitch = (int)in->pitch_c;
	/* one frame: any 16-byte multiple serves as the (unused) stride of the frame dimension */
	view.frame_stride = n_frames > 1 ? (long long)in->frame_stride_y : (long long)in->pitch_y * height;
	view.chroma_frame_stride = n_frames > 1 ? (long long)in->frame_stride_c : (long long)in->pitch_c * (height / 2);
	view.whole_frame = false;   /* pitched planes have no generic path: the checks above admit the tiled kernel only */
	return EnqueueFilter(ctx, view, iy, oy, (int)n_frames, d_qp, 0,
		d_vert_bs, hdb_num_vert_bs(width, height), d_hor_bs, hdb_num_hor_bs(width, height),
		static_cast<cudaStream_t>(stream), nullptr);
}

/* ---- edge maps: per-edge QP and boundary-strength derivation ---- */

int hdb_deblock_frame_edges(hdb_context *ctx, const unsigned char *yuv_in, unsigned char *yuv_out,
	unsigned width, unsigned height, unsigned qp, const unsigned char *vert_edges, unsigned num_vert_edges,
	const unsigned char *hor_edges, unsigned num_hor_edges, hdb_timing *timing)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	if (!vert_edges || !hor_edges) return Fail(ctx, HDB_ERR_INVALID_ARG, "edge maps are required: there is no default for them");
	int rc = ValidateFrameArgs(ctx, width, height, vert_edges, num_vert_edges, hor_edges, num_hor_edges, true);
	if (rc) return rc;
	return BatchHost(ctx, yuv_in, yuv_out, 1, width, height, &qp, vert_edges, hor_edges, timing, HDB_FLAG_EDGE_MAPS);
}

int hdb_deblock_batch_edges(hdb_context *ctx, const unsigned char *frames_in, unsigned char *frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *qp,
	const unsigned char *vert_edges, const unsigned char *hor_edges, hdb_timing *timing)
{
	return BatchHost(ctx, frames_in, frames_out, n_frames, width, height, qp, vert_edges, hor_edges, timing, HDB_FLAG_EDGE_MAPS);
}

int hdb_deblock_batch_edges_device(hdb_context *ctx, const void *d_frames_in, void *d_frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *d_qp,
	const unsigned char *d_vert_edges, const unsigned char *d_hor_edges, void *stream)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lock(ctx->mu);
	if (!d_frames_in || !d_frames_out || !d_qp || !d_vert_edges || !d_hor_edges || n_frames == 0)
		return Fail(ctx, HDB_ERR_INVALID_ARG, "NULL buffer or empty batch");
	int rc = ValidateFrameArgs(ctx, width, height, d_vert_edges, 0, d_hor_edges, 0, false);
	if (rc) return rc;
	DeviceGuard guard(ctx->device);
	const FrameView view = WholeFrameView((int)width, (int)height);
	return EnqueueFilter(ctx, view, static_cast<const unsigned char *>(d_frames_in),
		static_cast<unsigned char *>(d_frames_out), (int)n_frames, d_qp, 0,
		d_vert_edges, hdb_num_vert_bs(width, height), d_hor_edges, hdb_num_hor_bs(width, height),
		static_cast<cudaStream_t>(stream), nullptr, true);
}

/* caller holds ctx->mu */
static int DeriveEdgesLocked(hdb_context *ctx, const hdb_block_info *d_info, unsigned n_frames,
	unsigned width, unsigned height, unsigned char *d_vert_edges, unsigned char *d_hor_edges, cudaStream_t stream)
{
	if ((uintptr_t)d_info % 4 != 0) return Fail(ctx, HDB_ERR_INVALID_ARG, "block data must be 4-byte aligned");
	DeviceGuard guard(ctx->device);
	HDB_CUDA(ctx, LaunchDeriveEdges(d_info, (int)width, (int)height, (int)n_frames, d_vert_edges, d_hor_edges, stream));
	ctx->launches += 1;
	return HDB_OK;
}

int hdb_derive_edges_device(hdb_context *ctx, const hdb_block_info *d_info, unsigned n_frames,
	unsigned width, unsigned height, unsigned char *d_vert_edges, unsigned char *d_hor_edges, void *stream)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lock(ctx->mu);
	if (!d_info || !d_vert_edges || !d_hor_edges || n_frames == 0) return Fail(ctx, HDB_ERR_INVALID_ARG, "NULL buffer or empty batch");
	int rc = ValidateFrameArgs(ctx, width, height, nullptr, 0, nullptr, 0, false);
	if (rc) return rc;
	return DeriveEdgesLocked(ctx, d_info, n_frames, width, height, d_vert_edges, d_hor_edges, static_cast<cudaStream_t>(stream));
}

int hdb_derive_edges(hdb_context *ctx, const hdb_block_info *info, unsigned n_frames,
	unsigned width, unsigned height, unsigned char *vert_edges, unsigned char *hor_edges)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lock(ctx->mu);
	if (!info || !vert_edges || !hor_edges || n_frames == 0) return Fail(ctx, HDB_ERR_INVALID_ARG, "NULL buffer or empty batch");
	int rc = ValidateFrameArgs(ctx, width, height, nullptr, 0, nullptr, 0, false);
	if (rc) return rc;
	const size_t ni = sizeof(hdb_block_info) * (size_t)(width / 8) * (height / 8) * n_frames;
	const size_t nv = (size_t)hdb_num_vert_bs(width, height) * n_frames, nh = (size_t)hdb_num_hor_bs(width, height) * n_frames;
	const size_t off_v = (ni + 255) & ~(size_t)255, off_h = (off_v + nv + 255) & ~(size_t)255, total = off_h + nh;
	DeviceGuard guard(ctx->device);
	if (ctx->derive_cap < total) {
		cudaFree(ctx->derive_dev); ctx->derive_dev = nullptr; ctx->derive_cap = 0;
		HDB_CUDA(ctx, cudaMalloc(&ctx->derive_dev, total));
		ctx->derive_cap = total;
	}
	unsigned char *d = static_cast<unsigned char *>(ctx->derive_dev);
	HDB_CUDA(ctx, cudaMemcpyAsync(d, info, ni, cudaMemcpyHostToDevice, ctx->s_k));
	rc = DeriveEdgesLocked(ctx, reinterpret_cast<const hdb_block_info *>(d), n_frames, width, height, d + off_v, d + off_h, ctx->s_k);
	if (rc) return rc;
	HDB_CUDA(ctx, cudaMemcpyAsync(vert_edges, d + off_v, nv, cudaMemcpyDeviceToHost, ctx->s_k));
	HDB_CUDA(ctx, cudaMemcpyAsync(hor_edges, d + off_h, nh, cudaMemcpyDeviceToHost, ctx->s_k));
	HDB_CUDA(ctx, cudaStreamSynchronize(ctx->s_k));
	return HDB_OK;
}

/* ---- NV12 and the flagged general forms ---- */

int hdb_deblock_batch_ex(hdb_context *ctx, const unsigned char *frames_in, unsigned char *frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *qp,
	const unsigned char *vert, const unsigned char *hor, unsigned flags, hdb_timing *timing)
{
	if (flags & ~(HDB_FLAG_EDGE_MAPS | HDB_FLAG_NV12 | HDB_FLAG_SAMPLES16 | 0xff00u)) return ctx ? Fail(ctx, HDB_ERR_INVALID_ARG, "unknown flag") : HDB_ERR_INVALID_ARG;
	return BatchHost(ctx, frames_in, frames_out, n_frames, width, height, qp, vert, hor, timing, flags);
}

int hdb_deblock_batch_blocks(hdb_context *ctx, const unsigned char *frames_in, unsigned char *frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *qp,
	const hdb_block_info *info, unsigned flags, hdb_timing *timing)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	if (flags & ~HDB_FLAG_NV12) return Fail(ctx, HDB_ERR_INVALID_ARG, "unknown flag");
	if (!info) return Fail(ctx, HDB_ERR_INVALID_ARG, "block data is required");
	return BatchHost(ctx, frames_in, frames_out, n_frames, width, height, qp, nullptr, nullptr, timing, flags, info);
}

int hdb_deblock_batch_device_ex(hdb_context *ctx, const void *d_frames_in, void *d_frames_out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *d_qp,
	const unsigned char *d_vert, const unsigned char *d_hor, unsigned flags, void *stream)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lock(ctx->mu);
	if (flags & ~(HDB_FLAG_EDGE_MAPS | HDB_FLAG_NV12 | HDB_FLAG_SAMPLES16 | 0xff00u)) return Fail(ctx, HDB_ERR_INVALID_ARG, "unknown flag");
	if (!d_frames_in || !d_frames_out || !d_qp || n_frames == 0) return Fail(ctx, HDB_ERR_INVALID_ARG, "NULL buffer or empty batch");
	int rc = ValidateFrameArgs(ctx, width, height, d_vert, 0, d_hor, 0, false);
	if (rc) return rc;
	const int depth = DepthOfFlags(flags);
	if ((flags & HDB_FLAG_SAMPLES16) && (depth < 8 || depth > 10)) return Fail(ctx, HDB_ERR_INVALID_ARG, "bit depth must be 8, 9 or 10 (HDB_FLAG_DEPTH)");
	DeviceGuard guard(ctx->device);
	const FrameView view = WholeFrameView((int)width, (int)height, (flags & HDB_FLAG_NV12) != 0, depth);
	return EnqueueFilter(ctx, view, static_cast<const unsigned char *>(d_frames_in),
		static_cast<unsigned char *>(d_frames_out), (int)n_frames, d_qp, 0,
		d_vert, hdb_num_vert_bs(width, height), d_hor, hdb_num_hor_bs(width, height),
		static_cast<cudaStream_t>(stream), nullptr, (flags & HDB_FLAG_EDGE_MAPS) != 0);
}

int hdb_deblock_nv12_surfaces_device(hdb_context *ctx, const hdb_nv12_surfaces *in, const hdb_nv12_surfaces *out,
	unsigned n_frames, unsigned width, unsigned height, const unsigned *d_qp,
	const unsigned char *d_vert, const unsigned char *d_hor, unsigned flags, void *stream)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lock(ctx->mu);
	if (flags & ~HDB_FLAG_EDGE_MAPS) return Fail(ctx, HDB_ERR_INVALID_ARG, "unknown flag");
	if (!in || !out || !d_qp || n_frames == 0 || !in->y || !in->uv || !out->y || !out->uv)
		return Fail(ctx, HDB_ERR_INVALID_ARG, "NULL surface, plane or QP pointer, or empty batch");
	int rc = ValidateFrameArgs(ctx, width, height, d_vert, 0, d_hor, 0, false);
	if (rc) return rc;
	if (width % 32 != 0) return Fail(ctx, HDB_ERR_DIMENSIONS, "NV12 surfaces need a width that is a multiple of 32");
	const hdb_nv12_surfaces *both[2] = { in, out };
	for (const hdb_nv12_surfaces *s : both) {
		const uintptr_t a = (uintptr_t)s->y | (uintptr_t)s->uv | s->pitch_y | s->pitch_uv
			| (n_frames > 1 ? s->frame_stride_y | s->frame_stride_uv : 0);
		if (a % 16 != 0) return Fail(ctx, HDB_ERR_INVALID_ARG, "surface addresses, pitches and frame strides must be multiples of 16 bytes");
		if (s->pitch_y < width || s->pitch_uv < width || s->pitch_y > (1u << 30) || s->pitch_uv > (1u << 30))
			return Fail(ctx, HDB_ERR_INVALID_ARG, "pitch smaller than the plane width");
	}
	if (in->pitch_y != out->pitch_y || in->pitch_uv != out->pitch_uv
		|| (n_frames > 1 && (in->frame_stride_y != out->frame_stride_y || in->frame_stride_uv != out->frame_stride_uv)))
		return Fail(ctx, HDB_ERR_INVALID_ARG, "input and output surfaces must share pitches and frame strides");
	DeviceGuard guard(ctx->device);
	FrameView view = WholeFrameView((int)width, (int)height, true);
	const unsigned char *iy = static_cast<const unsigned char *>(in->y);
	unsigned char *oy = static_cast<unsigned char *>(out->y);
	view.u_offset = static_cast<const unsigned char *>(in->uv) - iy;
	view.u_offset_out = static_cast<unsigned char *>(out->uv) - oy;
	view.v_offset = view.u_offset; view.v_offset_out = view.u_offset_out;
	view.luma_pitch = (int)in->pitch_y; view.chroma_pitch = (int)in->pitch_uv;
	view.frame_stride = n_frames > 1 ? (long long)in->frame_stride_y : (long long)in->pitch_y * height;
	view.chroma_frame_stride = n_frames > 1 ? (long long)in->frame_stride_uv : (long long)in->pitch_uv * (height / 2);
	view.whole_frame = false;
	return EnqueueFilter(ctx, view, iy, oy, (int)n_frames, d_qp, 0,
		d_vert, hdb_num_vert_bs(width, height), d_hor, hdb_num_hor_bs(width, height),
		static_cast<cudaStream_t>(stream), nullptr, (flags & HDB_FLAG_EDGE_MAPS) != 0);
}

/*
 * A raw YUV sequence: the file holds any number of frames back to back (rank 4 of SURVEY.md section 8(f);
 * the reference's reader takes exactly one, cpu.h:43).  Frames go through the batch pipeline a bounded
 * chunk at a time, so the file may be larger than host or device memory.
 */
int hdb_execute_gpu_sequence(hdb_context *ctx, const char *input_file_name, const char *output_file_name,
	unsigned width, unsigned height, unsigned qp, unsigned flags, unsigned *n_frames_out, hdb_timing *timing)
{
	if (!ctx || !input_file_name || !output_file_name) return HDB_ERR_INVALID_ARG;
	if (flags & ~HDB_FLAG_NV12) return Fail(ctx, HDB_ERR_INVALID_ARG, "unknown flag");
	FILE *f = fopen(input_file_name, "rb");
	if (!f) return Fail(ctx, HDB_ERR_IO, std::string("cannot open ") + input_file_name);
	fseeko(f, 0, SEEK_END);
	const long long length = ftello(f);
	fseeko(f, 0, SEEK_SET);
	const size_t fb = (size_t)3 * width * height / 2;
	/* the reference's order: length first (cpu.h:43), then the dimensions (cpu.h:46) */
	if (fb == 0 || length <= 0 || (size_t)length % fb != 0) { fclose(f); return Fail(ctx, HDB_ERR_FILE_SIZE, kMsgFileSize); }
	if (width % 8 != 0 || height % 8 != 0) { fclose(f); return Fail(ctx, HDB_ERR_DIMENSIONS, kMsgDims); }
	const unsigned long long total = (unsigned long long)length / fb;
	const unsigned chunk = (unsigned)std::min<unsigned long long>(total, std::max<size_t>(1, ((size_t)512 << 20) / fb));
	unsigned char *buf = static_cast<unsigned char *>(hdb_host_alloc(fb * chunk));
	if (!buf) { fclose(f); return Fail(ctx, HDB_ERR_NO_MEMORY, "cudaMallocHost failed"); }
	/* the output is written while the input is still being read: the two must be different files */
	struct stat st_in, st_out;
	if (stat(input_file_name, &st_in) == 0 && stat(output_file_name, &st_out) == 0
		&& st_in.st_dev == st_out.st_dev && st_in.st_ino == st_out.st_ino) {
		fclose(f); hdb_host_free(buf);
		return Fail(ctx, HDB_ERR_INVALID_ARG, "input and output name the same file");
	}
	FILE *o = fopen(output_file_name, "wb");
	int rc = o ? HDB_OK : Fail(ctx, HDB_ERR_IO, std::string("cannot write ") + output_file_name);
	std::vector<unsigned> qps(chunk, qp);
	hdb_timing sum{}, t{};
	for (unsigned long long f0 = 0; !rc && f0 < total; f0 += chunk) {
		const unsigned n = (unsigned)std::min<unsigned long long>(chunk, total - f0);
		if (fread(buf, 1, fb * n, f) != fb * n) { rc = Fail(ctx, HDB_ERR_IO, "short read"); break; }
		rc = BatchHost(ctx, buf, buf, n, width, height, qps.data(), nullptr, nullptr, &t, flags);
		if (rc) break;
		if (fwrite(buf, 1, fb * n, o) != fb * n) { rc = Fail(ctx, HDB_ERR_IO, std::string("cannot write ") + output_file_name); break; }
		sum.kernel_ms += t.kernel_ms; sum.copy_ms += t.copy_ms; sum.total_ms += t.total_ms;
		sum.h2d_bytes += t.h2d_bytes; sum.d2h_bytes += t.d2h_bytes; sum.kernel_launches += t.kernel_launches; sum.path = t.path;
	}
	fclose(f);
	if (o) fclose(o);
	hdb_host_free(buf);
	if (!rc) {
		if (timing) *timing = sum;
		if (n_frames_out) *n_frames_out = (unsigned)total;
	}
	return rc;
}

/* ---- sample adaptive offset ---- */

/* *depth16: 0 for a byte a sample, else the bit depth of samples in 16-bit containers */
static int ValidateSao(hdb_context *ctx, unsigned width, unsigned height, unsigned ctb_size, unsigned flags, int max_depth, int *depth16)
{
	if (flags & ~(HDB_FLAG_SAMPLES16 | (0xffu << 8))) return Fail(ctx, HDB_ERR_INVALID_ARG, "unknown flag");
	*depth16 = DepthOfFlags(flags);
	if ((flags & HDB_FLAG_SAMPLES16) && (*depth16 < 8 || *depth16 > max_depth)) return Fail(ctx, HDB_ERR_INVALID_ARG, "bit depth out of range (HDB_FLAG_DEPTH)");
	int rc = ValidateFrameArgs(ctx, width, height, nullptr, 0, nullptr, 0, false);
	if (rc) return rc;
	if (ctb_size != 16 && ctb_size != 32 && ctb_size != 64) return Fail(ctx, HDB_ERR_INVALID_ARG, "coding tree block size must be 16, 32 or 64");
	return HDB_OK;
}

/* the kernel's ticket counters: a stream-ordered allocation of the call's own, so that launches in flight on
 * different streams never share one */
static cudaError_t LaunchSaoWithTickets(hdb_context *ctx, const unsigned char *d_in, unsigned char *d_out, int n_frames, int width, int height,
	const void *d_params, int ctb, int depth16, cudaStream_t st)
{
	unsigned *tickets = nullptr;
	cudaError_t e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&tickets), sizeof(unsigned) * kSaoTicketWords, ctx->pool, st);
	if (e != cudaSuccess) return e;
	e = LaunchSao(d_in, d_out, n_frames, width, height, d_params, ctb, depth16 ? depth16 : 8, depth16 ? 2 : 1, tickets, ctx->prop.multiProcessorCount, st);
	cudaFreeAsync(tickets, st);
	return e;
}

int hdb_sao_batch_device_ex(hdb_context *ctx, const void *d_frames_in, void *d_frames_out, unsigned n_frames,
	unsigned width, unsigned height, const hdb_sao_params *d_params, unsigned ctb_size, unsigned flags, void *stream)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lock(ctx->mu);
	if (!d_frames_in || !d_frames_out || !d_params || n_frames == 0) return Fail(ctx, HDB_ERR_INVALID_ARG, "NULL buffer or empty batch");
	if (d_frames_in == d_frames_out) return Fail(ctx, HDB_ERR_INVALID_ARG, "sample adaptive offset runs out of place: a sample's neighbours are input samples");
	int depth16 = 0;
	int rc = ValidateSao(ctx, width, height, ctb_size, flags, 12, &depth16);
	if (rc) return rc;
	if (((uintptr_t)d_frames_in | (uintptr_t)d_frames_out) % (depth16 ? 8 : 4) != 0)
		return Fail(ctx, HDB_ERR_INVALID_ARG, depth16 ? "frames of 16-bit samples must be 8-byte aligned" : "frames must be 4-byte aligned");
	DeviceGuard guard(ctx->device);
	HDB_CUDA(ctx, LaunchSaoWithTickets(ctx, static_cast<const unsigned char *>(d_frames_in), static_cast<unsigned char *>(d_frames_out), (int)n_frames,
		(int)width, (int)height, d_params, (int)ctb_size, depth16, static_cast<cudaStream_t>(stream)));
	ctx->launches += 1;
	return HDB_OK;
}

int hdb_sao_batch_device(hdb_context *ctx, const void *d_frames_in, void *d_frames_out, unsigned n_frames,
	unsigned width, unsigned height, const hdb_sao_params *d_params, unsigned ctb_size, void *stream)
{
	return hdb_sao_batch_device_ex(ctx, d_frames_in, d_frames_out, n_frames, width, height, d_params, ctb_size, 0, stream);
}

int hdb_sao_batch_ex(hdb_context *ctx, const void *frames_in, void *frames_out, unsigned n_frames,
	unsigned width, unsigned height, const hdb_sao_params *params, unsigned ctb_size, unsigned flags)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lock(ctx->mu);
	if (!frames_in || !frames_out || !params || n_frames == 0) return Fail(ctx, HDB_ERR_INVALID_ARG, "NULL buffer or empty batch");
	int depth16 = 0;
	int rc = ValidateSao(ctx, width, height, ctb_size, flags, 12, &depth16);
	if (rc) return rc;
	DeviceGuard guard(ctx->device);
	const size_t fb = (size_t)width * height * 3 / 2 * (depth16 ? 2 : 1);
	const size_t pb = sizeof(hdb_sao_params) * 3 * ((height + ctb_size - 1) / ctb_size) * ((width + ctb_size - 1) / ctb_size);
	/* a bounded number of frames at a time, buffers of the call's own in stream order */
	const unsigned chunk = (unsigned)std::min<size_t>(n_frames, std::max<size_t>(1, ((size_t)256 << 20) / fb));
	cudaStream_t st = ctx->s_k;
	unsigned char *d_in = nullptr, *d_out = nullptr, *d_par = nullptr;
	/* an allocation that fails must not strand the ones before it: every exit goes through the frees below */
	cudaError_t e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&d_in), fb * chunk, ctx->pool, st);
	if (e == cudaSuccess) e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&d_out), fb * chunk, ctx->pool, st);
	if (e == cudaSuccess) e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&d_par), pb * chunk, ctx->pool, st);
	for (unsigned f0 = 0; f0 < n_frames && e == cudaSuccess; f0 += chunk) {
		const unsigned nf = std::min(chunk, n_frames - f0);
		e = cudaMemcpyAsync(d_in, static_cast<const unsigned char *>(frames_in) + (size_t)f0 * fb, fb * nf, cudaMemcpyHostToDevice, st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(d_par, reinterpret_cast<const unsigned char *>(params) + (size_t)f0 * pb, pb * nf, cudaMemcpyHostToDevice, st);
		if (e == cudaSuccess) e = LaunchSaoWithTickets(ctx, d_in, d_out, (int)nf, (int)width, (int)height, d_par, (int)ctb_size, depth16, st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(static_cast<unsigned char *>(frames_out) + (size_t)f0 * fb, d_out, fb * nf, cudaMemcpyDeviceToHost, st);
		ctx->launches += 1;
	}
	if (d_in) cudaFreeAsync(d_in, st);
	if (d_out) cudaFreeAsync(d_out, st);
	if (d_par) cudaFreeAsync(d_par, st);
	const cudaError_t e2 = cudaStreamSynchronize(st);
	if (e != cudaSuccess) return FailCuda(ctx, e, "sample adaptive offset");
	if (e2 != cudaSuccess) return FailCuda(ctx, e2, "cudaStreamSynchronize");
	return HDB_OK;
}

int hdb_sao_batch(hdb_context *ctx, const unsigned char *frames_in, unsigned char *frames_out, unsigned n_frames,
	unsigned width, unsigned height, const hdb_sao_params *params, unsigned ctb_size)
{
	return hdb_sao_batch_ex(ctx, frames_in, frames_out, n_frames, width, height, params, ctb_size, 0);
}

int hdb_deblock_sao_batch_device(hdb_context *ctx, const void *d_frames_in, void *d_frames_out, unsigned n_frames,
	unsigned width, unsigned height, const unsigned *d_qp, const unsigned char *d_vert, const unsigned char *d_hor,
	unsigned flags, const hdb_sao_params *d_params, unsigned ctb_size, void *stream)
{
	if (!ctx) return HDB_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lock(ctx->mu);
	if (flags & ~(HDB_FLAG_EDGE_MAPS | HDB_FLAG_SAMPLES16 | (0xffu << 8))) return Fail(ctx, HDB_ERR_INVALID_ARG, "unknown flag");
	if (!d_frames_in || !d_frames_out || !d_qp || !d_params || n_frames == 0) return Fail(ctx, HDB_ERR_INVALID_ARG, "NULL buffer or empty batch");
	int depth16 = 0;
	int rc = ValidateFrameArgs(ctx, width, height, d_vert, 0, d_hor, 0, false);
	if (!rc) rc = ValidateSao(ctx, width, height, ctb_size, flags & ~HDB_FLAG_EDGE_MAPS, 10, &depth16);
	if (rc) return rc;
	DeviceGuard guard(ctx->device);
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	const size_t fb = (size_t)width * height * 3 / 2 * (depth16 ? 2 : 1);
	unsigned char *scratch = nullptr;
	HDB_CUDA(ctx, cudaMallocFromPoolAsync(reinterpret_cast<void **>(&scratch), fb * n_frames, ctx->pool, st));
	const FrameView view = WholeFrameView((int)width, (int)height, false, depth16);
	rc = EnqueueFilter(ctx, view, static_cast<const unsigned char *>(d_frames_in), scratch, (int)n_frames, d_qp, 0,
		d_vert, hdb_num_vert_bs(width, height), d_hor, hdb_num_hor_bs(width, height), st, nullptr, (flags & HDB_FLAG_EDGE_MAPS) != 0);
	cudaError_t e = cudaSuccess;
	if (!rc) {
		e = LaunchSaoWithTickets(ctx, scratch, static_cast<unsigned char *>(d_frames_out), (int)n_frames, (int)width, (int)height, d_params,
			(int)ctb_size, depth16, st);
		ctx->launches += 1;
	}
	cudaFreeAsync(scratch, st);
	if (rc) return rc;
	if (e != cudaSuccess) return FailCuda(ctx, e, "LaunchSao");
	return HDB_OK;
}

unsigned long long hdb_kernel_launch_count(const hdb_context *ctx) { return ctx ? ctx->launches : 0; }

} // extern "C"
