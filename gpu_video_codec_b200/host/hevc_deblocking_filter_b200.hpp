/*
 * hevc_deblocking_filter_b200.hpp -- C++ host-side mirror of the reference's interface for the
 * deblocking path, on top of the C ABI (include/hevc_deblock_b200.h).
 *
 * Same names, argument meaning and error behaviour as the reference
 * (hevc_deblocking_filter/hevc_deblocking_filter_cpu.h "cpu.h", hevc_deblocking_filter_gpu.cu
 * "gpu.cu", main.cu), in namespace b200 so that both can be linked into one program:
 *
 *   b200::ReadYuvFrame            cpu.h:33-1489   ctor / SetBoundaryStrenght / DeblockingFilter / Save
 *   b200::ExecuteGpu              gpu.cu:1230     file in, file out, prints the reference's three timings
 *   b200::GetGpuDeviceInfo        main.cu:92-107
 *
 * Differences that are the point of the exercise: the frame lives in pinned host memory as tight
 * YV12 planes (no padded copies -- the padding exists only as TMA out-of-bounds fill on the GPU),
 * DeblockingFilter runs on the B200 (its num_threads argument is accepted and ignored), and nothing
 * falls back to the CPU: without a compute-capability-10 device the constructor throws.
 * Errors are thrown as const char* like the reference does (cpu.h:44, 47, 123).
 */
#pragma once

#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "../../include/hevc_deblock_b200.h"

namespace b200 {

class ReadYuvFrame {
public:
	/* cpu.h:35: reads one YV12 frame; throws "Incorrect file size" / "Width and height ..." */
	ReadYuvFrame(char const *file_name, unsigned int width, unsigned int height, unsigned int Qp = 20)
		: _width(width), _height(height), _Qp(Qp)
	{
		FILE *f = std::fopen(file_name, "rb");
		long length = -1;
		if (f) { std::fseek(f, 0, SEEK_END); length = std::ftell(f); std::fseek(f, 0, SEEK_SET); }
		if (length < 0 || (unsigned int)length != 3 * width * height / 2) {
			if (f) std::fclose(f);
			throw "Incorrect file size";                                                    /* cpu.h:43-45 */
		}
		if (width % 8 != 0 || height % 8 != 0) {
			std::fclose(f);
			throw "Width and height of image must be multiplier of sample block size";      /* cpu.h:46-48 */
		}
		if (hdb_create(&_ctx, 0) != HDB_OK) { std::fclose(f); throw "No Blackwell (sm_100) CUDA device available"; }
		_bytes = (size_t)3 * width * height / 2;
		_frame = static_cast<unsigned char *>(hdb_host_alloc(_bytes));                      /* pinned, gpu.cu:1103 */
		if (!_frame || std::fread(_frame, 1, _bytes, f) != _bytes) { std::fclose(f); Release(); throw "Cannot read input file"; }
		std::fclose(f);
		_num_vert_bs = hdb_num_vert_bs(width, height);                                      /* cpu.h:86 */
		_num_hor_bs = hdb_num_hor_bs(width, height);                                        /* cpu.h:87 */
	}
	~ReadYuvFrame() { Release(); }
	ReadYuvFrame(const ReadYuvFrame &) = delete;
	ReadYuvFrame &operator=(const ReadYuvFrame &) = delete;

	/* cpu.h:120: copies the caller's luma strengths; throws on a size mismatch */
	void SetBoundaryStrenght(unsigned char *vert_bs, unsigned int num_vert_bs,
		unsigned char *hor_bs, unsigned int num_hor_bs)
	{
		if (_num_hor_bs != num_hor_bs || _num_vert_bs != num_vert_bs)
			throw "Incorrect size of input boundary strenght array";                        /* cpu.h:122-123 */
		_vert_bs.assign(vert_bs, vert_bs + num_vert_bs);
		_hor_bs.assign(hor_bs, hor_bs + num_hor_bs);
		_edge_maps = false;
	}

	/*
	 * Beyond the reference: the same two arrays with an edge byte HDB_EDGE(bs, qp) in every entry, so that
	 * every luma edge has its own QP (include/hevc_deblock_b200.h).  Same size rule and message as the
	 * strengths-only setter.
	 */
	void SetEdgeMaps(unsigned char *vert_edges, unsigned int num_vert, unsigned char *hor_edges, unsigned int num_hor)
	{
		SetBoundaryStrenght(vert_edges, num_vert, hor_edges, num_hor);
		_edge_maps = true;
	}

	/* cpu.h:134: filters luma then both chroma planes, in place */
	void DeblockingFilter(unsigned int num_threads = 1)
	{
		(void)num_threads;
		const bool custom = !_vert_bs.empty();
		const unsigned char *v = custom ? (const unsigned char *)_vert_bs.data() : nullptr;
		const unsigned char *h = custom ? (const unsigned char *)_hor_bs.data() : nullptr;
		int rc = _edge_maps
			? hdb_deblock_frame_edges(_ctx, _frame, _frame, _width, _height, _Qp, v, _num_vert_bs, h, _num_hor_bs, &_timing)
			: hdb_deblock_frame(_ctx, _frame, _frame, _width, _height, _Qp, v, _num_vert_bs, h, _num_hor_bs, &_timing);
		if (rc != HDB_OK) { _error = hdb_last_error(_ctx); throw _error.c_str(); }
	}

	const unsigned char *Frame() const { return _frame; }          /* tight YV12, 3 * width * height / 2 bytes */

	/* cpu.h:995: writes the tight YV12 frame */
	void Save(char const *output_file_name)
	{
		FILE *f = std::fopen(output_file_name, "wb");
		if (!f || std::fwrite(_frame, 1, _bytes, f) != _bytes) { if (f) std::fclose(f); throw "Cannot write output file"; }
		std::fclose(f);
	}

	const hdb_timing &Timing() const { return _timing; }
	hdb_context *Context() const { return _ctx; }

private:
	void Release()
	{
		if (_frame) hdb_host_free(_frame);
		if (_ctx) hdb_destroy(_ctx);
		_frame = nullptr; _ctx = nullptr;
	}
	unsigned int _width, _height, _Qp;
	unsigned int _num_vert_bs = 0, _num_hor_bs = 0;
	size_t _bytes = 0;
	unsigned char *_frame = nullptr;
	hdb_context *_ctx = nullptr;
	std::basic_string<unsigned char> _vert_bs, _hor_bs;
	hdb_timing _timing{};
	std::string _error;
	bool _edge_maps = false;
};

/* main.cu:92-107 */
inline void GetGpuDeviceInfo()
{
	hdb_context *ctx = nullptr;
	if (hdb_create(&ctx, 0) != HDB_OK) { std::printf("No usable device: %s\n", hdb_last_error(nullptr)); return; }
	hdb_device_info p;
	hdb_get_device_info(ctx, &p);
	std::printf("==============================================\n");
	std::printf("Device %d: %s\n", 0, p.name);
	std::printf("Compute Capability: %d.%d\n", p.cc_major, p.cc_minor);
	std::printf("Number of multiprocessors: %d\n", p.sm_count);
	std::printf("Total amount of constant memory: %4.2f KB\n", p.total_const_mem / 1024.0);
	std::printf("Total amount of global memory: %4.2f KB\n", p.total_global_mem / 1024.0);
	std::printf("Total amount of shared memory per block: %4.2f KB\n", p.shared_mem_per_block / 1024.0);
	std::printf("Warp size: %d\n", p.warp_size);
	std::printf("Maximum number of threads per block: %d\n", p.max_threads_per_block);
	std::printf("==============================================\n\n\n");
	hdb_destroy(ctx);
}

/*
 * gpu.cu:1230: file in, file out, and the reference's three timing lines (gpu.cu:1292, 1302-1303).
 * The four thread-block dimensions of the reference's signature are accepted for source
 * compatibility and ignored (the kernel shapes itself).
 */
inline void ExecuteGpu(std::string const &input_file_name, std::string const &output_file_name,
	unsigned int width, unsigned int height, unsigned int Qp,
	unsigned dimx1 = 16, unsigned int dimy1 = 16, unsigned dimx2 = 16, unsigned int dimy2 = 16)
{
	(void)dimx1; (void)dimy1; (void)dimx2; (void)dimy2;
	hdb_context *ctx = nullptr;
	if (hdb_create(&ctx, 0) != HDB_OK) { std::printf("Error: %s\n", hdb_last_error(nullptr)); return; }
	hdb_timing t{};
	/* first call: allocations and per-size tables, as the reference's Initialize + cudaMalloc block */
	int rc = hdb_execute_gpu(ctx, input_file_name.c_str(), output_file_name.c_str(), width, height, Qp, &t);
	if (rc == HDB_OK) rc = hdb_execute_gpu(ctx, input_file_name.c_str(), output_file_name.c_str(), width, height, Qp, &t);
	if (rc != HDB_OK) {
		std::string msg = hdb_last_error(ctx);
		hdb_destroy(ctx);
		if (rc == HDB_ERR_FILE_SIZE) throw "Incorrect file size";
		if (rc == HDB_ERR_DIMENSIONS) throw "Width and height of image must be multiplier of sample block size";
		std::printf("Error: %s\n", msg.c_str());
		return;
	}
	std::cout << "Execution Time without copy on GPU: " << t.kernel_ms / 1e3 << "s" << std::endl;
	std::cout << "Execution Time with copy on GPU: " << t.total_ms / 1e3 << "s" << std::endl;
	std::cout << "Copy Operation Time with GPU buffers: " << t.copy_ms / 1e3 << "s" << std::endl;
	hdb_destroy(ctx);
}

/*
 * ExecuteGpu for a raw sequence: the file may hold any number of frames (the reference's driver runs
 * one single-frame file per build, main.cu:111-133).  Prints the frame count and the same three lines.
 */
inline void ExecuteGpuSequence(std::string const &input_file_name, std::string const &output_file_name,
	unsigned int width, unsigned int height, unsigned int Qp)
{
	hdb_context *ctx = nullptr;
	if (hdb_create(&ctx, 0) != HDB_OK) { std::printf("Error: %s\n", hdb_last_error(nullptr)); return; }
	hdb_timing t{};
	unsigned frames = 0;
	int rc = hdb_execute_gpu_sequence(ctx, input_file_name.c_str(), output_file_name.c_str(), width, height, Qp, 0, &frames, &t);
	if (rc != HDB_OK) {
		std::string msg = hdb_last_error(ctx);
		hdb_destroy(ctx);
		if (rc == HDB_ERR_FILE_SIZE) throw "Incorrect file size";
		if (rc == HDB_ERR_DIMENSIONS) throw "Width and height of image must be multiplier of sample block size";
		std::printf("Error: %s\n", msg.c_str());
		return;
	}
	std::cout << "Frames: " << frames << std::endl;
	std::cout << "Execution Time without copy on GPU: " << t.kernel_ms / 1e3 << "s" << std::endl;
	std::cout << "Execution Time with copy on GPU: " << t.total_ms / 1e3 << "s" << std::endl;
	std::cout << "Copy Operation Time with GPU buffers: " << t.copy_ms / 1e3 << "s" << std::endl;
	hdb_destroy(ctx);
}

} // namespace b200
