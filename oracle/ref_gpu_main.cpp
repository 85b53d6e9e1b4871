/*
 * ref_gpu_main.cpp -- TEST / MEASUREMENT INFRASTRUCTURE ONLY: a command line around the reference's own GPU entry
 * point, ExecuteGpu (hevc_deblocking_filter_gpu.cu:1230-1306, declared in main.cu:87-90), so that the UNMODIFIED
 * reference .cu -- compiled for sm_100a where it lies under /root/reference by oracle/Makefile -- can be timed on the
 * same B200 as a courtesy datapoint.  It is NOT an oracle: the reference's GPU kernels do
 * not compute what its CPU filter computes, and its thread grid has no bounds guard
 * (gpu.cu:540-541 with the rounded-up grid of gpu.cu:1268), so the caller passes block shapes that divide the
 * block counts exactly.  ExecuteGpu prints its own three timings (gpu.cu:1292, 1302-1303); bench.py parses them.
 *
 *   ref_gpu_demo <in.yuv> <out.yuv> <width> <height> <qp> <dimx1> <dimy1> <dimx2> <dimy2> [repetitions]
 */
#include <cstdio>
#include <cstdlib>
#include <string>

void ExecuteGpu(std::string const &input_file_name, std::string const &output_file_name,
	unsigned int width, unsigned int height, unsigned int Qp,
	unsigned dimx1, unsigned int dimy1, unsigned dimx2, unsigned int dimy2);

int main(int argc, char **argv)
{
	if (argc < 10) {
		fprintf(stderr, "usage: %s in.yuv out.yuv width height qp dimx1 dimy1 dimx2 dimy2 [repetitions]\n", argv[0]);
		return 2;
	}
	const int reps = argc > 10 ? atoi(argv[10]) : 1;
	try {
		for (int r = 0; r < reps; r++)
			ExecuteGpu(argv[1], argv[2], (unsigned)atoi(argv[3]), (unsigned)atoi(argv[4]), (unsigned)atoi(argv[5]),
				(unsigned)atoi(argv[6]), (unsigned)atoi(argv[7]), (unsigned)atoi(argv[8]), (unsigned)atoi(argv[9]));
	} catch (const char *msg) {
		fprintf(stderr, "reference threw: %s\n", msg);
		return 1;
	}
	return 0;
}
