/*
 * deblock_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's single-thread HEVC-style
 * deblocking filter, hevc_deblocking_filter/hevc_deblocking_filter_cpu.h
 * ("cpu.h" below) of RomanKazantsev/gpu_video_codec.  It is the checker the
 * CUDA path is compared against.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call it; nothing under
 * gpu_video_codec_b200/ links or loads it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement
 *   (a) against the golden md5/sha256 of the unmodified reference run on its
 *       three bundled frames, stored in tests/golden/, and
 *   (b) when oracle/_ref/libref_deblock.so exists (built from the unmodified
 *       reference header by oracle/Makefile), byte for byte against it on
 *       random frames, QPs and boundary-strength maps.
 *
 * Defined behaviour where the reference reads indeterminate memory
 * (SURVEY.md section 0 facts 4-5): plane padding is zero, and the chroma
 * boundary-strength arrays are zero-extended past their end.
 */
#ifndef HDB_DEBLOCK_ORACLE_H
#define HDB_DEBLOCK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
	HDBO_OK = 0,
	HDBO_ERR_SIZE = 1,   /* cpu.h:46-48  width/height not multiples of 8 */
	HDBO_ERR_BS_SIZE = 2 /* cpu.h:122-123 boundary-strength array size mismatch */
};

/* cpu.h:1021-1033, 1064-1072: beta/tc lookup, QP clamped to 51. */
int hdbo_beta(unsigned qp);
int hdbo_tc(unsigned qp);

/* cpu.h:86-87, 104-105: boundary-strength array sizes (C integer arithmetic kept). */
unsigned hdbo_num_vert_bs(unsigned width, unsigned height);
unsigned hdbo_num_hor_bs(unsigned width, unsigned height);

/* cpu.h:92-99: constructor defaults (all 2, except i % (W/8+1) == 0 -> 0 for
 * vertical and i % (H/8+1) == 0 -> 0 for horizontal).  For the chroma arrays
 * (cpu.h:110-117) call with width/2, height/2. */
void hdbo_default_vert_bs(unsigned width, unsigned height, uint8_t *dst);
void hdbo_default_hor_bs(unsigned width, unsigned height, uint8_t *dst);

/*
 * One frame, cpu.h:35-118 (layout) + cpu.h:134-993 (filter) + cpu.h:995-1018 (crop).
 * in/out: tight YV12 frame, 3*w*h/2 bytes (Y, then the two chroma planes).
 * vert_bs/hor_bs: luma boundary strengths in the reference layout, or NULL for
 * the constructor defaults.  out may alias in.
 */
int hdbo_deblock_frame(const uint8_t *in, uint8_t *out, unsigned width, unsigned height,
	unsigned qp, const uint8_t *vert_bs, unsigned num_vert_bs,
	const uint8_t *hor_bs, unsigned num_hor_bs);

/* n frames back to back (frame f at in + f*3*w*h/2), per-frame QP and BS
 * (vert_bs + f*num_vert_bs, ...; NULL = defaults).  Serial loop over
 * hdbo_deblock_frame -- the batch has no semantics of its own. */
int hdbo_deblock_batch(const uint8_t *in, uint8_t *out, unsigned n, unsigned width,
	unsigned height, const unsigned *qp, const uint8_t *vert_bs, const uint8_t *hor_bs);

#ifdef __cplusplus
}
#endif
#endif
