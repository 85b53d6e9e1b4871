/*
 * deblock_api.h -- internal declarations shared by deblock_api.cu and
 * deblock_multi.cu.  The public boundary is include/hevc_deblock_b200.h.
 */
#pragma once

#include "../../include/hevc_deblock_b200.h"
#include "deblock_kernels.cuh"

#include <mutex>
#include <string>
#include <vector>

namespace hdb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
	const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
	CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct DeviceGuard {   /* make `device` current for the scope */
	int prev = -1;
	explicit DeviceGuard(int device) { cudaGetDevice(&prev); if (prev != device) cudaSetDevice(device); }
	~DeviceGuard() { int cur = -1; cudaGetDevice(&cur); if (prev >= 0 && cur != prev) cudaSetDevice(prev); }
};

struct SizeState {     /* device tables that depend on the frame size only */
	int w, h;
	uint32_t *chroma_bs;          /* resolved chroma strengths per chroma block */
	unsigned char *def_vbs, *def_hbs; /* reference-default luma strengths (cpu.h:92-99) */
};

constexpr int kSlots = 4;
constexpr int kSchedSlots = 64;      /* scheduler-word sets handed to successive launches in turn */

struct Slot {          /* one slice of a batch in flight */
	unsigned char *frames = nullptr, *vbs = nullptr, *hbs = nullptr;
	unsigned char *info = nullptr;             /* block data of the slice (edge maps derived on the device) */
	size_t frames_cap = 0, vbs_cap = 0, hbs_cap = 0, info_cap = 0;
	cudaEvent_t h2d_done = nullptr, k_done = nullptr, d2h_done = nullptr;
	bool ev_ready = false;
};

/* Page-lock a caller's host buffer for the duration of a call when it is large and not
 * already page-locked (small or already pinned buffers are left alone). */
struct PinHostRange {
	void *p = nullptr;
	PinHostRange(const void *ptr, size_t bytes)
	{
		if (!ptr || bytes < ((size_t)8 << 20)) return;
		cudaPointerAttributes a;
		if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return; }
		if (a.type != cudaMemoryTypeUnregistered) return;
		if (cudaHostRegister(const_cast<void *>(ptr), bytes, cudaHostRegisterDefault) == cudaSuccess) p = const_cast<void *>(ptr);
		else cudaGetLastError();
	}
	~PinHostRange() { if (p) cudaHostUnregister(p); }
	PinHostRange(const PinHostRange &) = delete;
	PinHostRange &operator=(const PinHostRange &) = delete;
};

/* Which rows of a frame a set of plane buffers holds (whole frame, or a CTU-row strip). */
struct FrameView {
	int width, height;             /* of the whole frame */
	bool whole_frame;
	int luma_row0, luma_rows;      /* padded-block rows to filter, frame-global numbering */
	int chroma_row0, chroma_rows;
	int luma_buf_rows, chroma_buf_rows;   /* sample rows present in the buffers */
	int luma_y_org, chroma_y_org;  /* buffer row of the first sample row of block row *_row0 (-4: the buffer starts at image row 0) */
	long long u_offset, v_offset;  /* byte offsets of the input chroma planes from the input luma plane */
	long long u_offset_out, v_offset_out;   /* the same for the output planes */
	int luma_pitch, chroma_pitch;  /* bytes between rows */
	long long frame_stride, chroma_frame_stride;   /* bytes between frames, luma plane / chroma planes */
	bool nv12;                     /* two planes: Y and interleaved UV at u_offset (v_offset unused) */
	/* a strip of a frame split across GPUs whose buffers hold its own rows only: per plane, the 4 rows above /
	 * below them in the neighbouring strips' buffers (peer pointers, same pitch); all NULL otherwise */
	const unsigned char *halo_top[3], *halo_bot[3];
	bool strip_halo;
	int bit_depth;                 /* 0: 8-bit samples, one byte each; 8 .. 10: that many bits in 16-bit containers (planar) */
};

} // namespace hdb

struct hdb_context {
	int device = 0;
	cudaDeviceProp prop{};
	hdb::EncodeTiledFn encode = nullptr;
	cudaStream_t s_h2d = nullptr, s_k = nullptr, s_d2h = nullptr;
	cudaStream_t s_setup = nullptr;          /* per-size tables are built here */
	cudaMemPool_t pool = nullptr;            /* stream-ordered scratch of calls (padded planes, the frames between deblocking and SAO): kept, not returned to the OS at every synchronisation */
	void *sched = nullptr;
	/* scheduler-word sets are handed to successive launches in turn; each remembers the stream that used it
	 * last and an event recorded behind that launch, which a launch on another stream waits for */
	unsigned long long sched_turn = 0;
	cudaEvent_t sched_done[hdb::kSchedSlots] = {};
	cudaStream_t sched_stream[hdb::kSchedSlots] = {};
	bool sched_used[hdb::kSchedSlots] = {};
	int tiled_ctas_per_sm[hdb::kTiledVariants] = {};   /* resident CTAs per SM of the tiled kernel's variants on this device */
	int tiled16_ctas_per_sm[hdb::kTiled16Variants] = {};
	std::vector<hdb::SizeState> sizes;
	hdb::Slot slots[hdb::kSlots];
	std::vector<cudaEvent_t> event_pool;
	void *derive_dev = nullptr;              /* block data and edge maps of a host-side hdb_derive_edges call */
	size_t derive_cap = 0;
	unsigned *qp_dev = nullptr;              /* the QPs of the batch in flight */
	size_t qp_cap = 0;
	size_t chunk_bytes = (size_t)128 << 20;  /* full slice size of the host pipeline (HDB_CHUNK_MB) */
	bool use_ldgsts = true;                  /* widths off the 32 grid take the tiled kernel's cp.async loader */
	bool zero_copy = true;                   /* single-slice calls on mapped pinned buffers skip the copy engines */
	unsigned long long launches = 0;
	std::string err;
	std::mutex mu;
};

namespace hdb {

int Fail(hdb_context *ctx, int status, const std::string &msg);
int FailCuda(hdb_context *ctx, cudaError_t e, const char *what);
void DestroyContext(hdb_context *ctx);
int GetSizeState(hdb_context *ctx, int w, int h, SizeState **out);
FrameView WholeFrameView(int w, int h, bool nv12 = false, int bit_depth = 0);
/* bit depth carried in the flags of the _ex entry points (0 = 8-bit samples) */
inline int DepthOfFlags(unsigned flags) { return (flags & HDB_FLAG_SAMPLES16) ? (int)((flags >> 8) & 0xffu) : 0; }
bool TiledPathUsable(int w, int h, const void *d_in, const void *d_out);
bool LdgstsPathUsable(int w, int h, const void *d_in, const void *d_out);
int ValidateFrameArgs(hdb_context *ctx, unsigned w, unsigned h, const unsigned char *vbs, unsigned nv,
	const unsigned char *hbs, unsigned nh, bool sizes_given);
int EnqueueFilter(hdb_context *ctx, const FrameView &view, const unsigned char *d_in, unsigned char *d_out,
	int n, const unsigned *d_qp, unsigned qp_value, const unsigned char *d_vbs, long long vbs_stride,
	const unsigned char *d_hbs, long long hbs_stride, cudaStream_t stream, int *path, bool edge_qp = false);
int BatchHost(hdb_context *ctx, const unsigned char *in, unsigned char *out, unsigned n, unsigned w, unsigned h,
	const unsigned *qp, const unsigned char *vbs, const unsigned char *hbs, hdb_timing *timing, unsigned flags = 0,
	const hdb_block_info *info = nullptr);

/* the same with ctx->mu already held by the caller */
int BatchHostLocked(hdb_context *ctx, const unsigned char *in, unsigned char *out, unsigned n, unsigned w, unsigned h,
	const unsigned *qp, const unsigned char *vbs, const unsigned char *hbs, hdb_timing *timing, unsigned flags = 0,
	const hdb_block_info *info = nullptr);

/* beta / tc on the host (cpu.h:1021-1033, 1064-1072); defined in deblock_kernels.cu next to the device tables */
int HostBeta(unsigned qp);
int HostTc(unsigned qp);

} // namespace hdb
