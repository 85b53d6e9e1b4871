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

/* ---------------------------------------------------------------------------
 * Edge maps: SURVEY.md section 8(f) rank 1 (per-block QP and boundary-strength
 * derivation upstream of the filter).  The reference stubs both: one QP per
 * frame (cpu.h:35-37, 136-137) and constant strengths (cpu.h:91-99).
 *
 * An edge byte is  e = bs | qp << 2  (bs 0..3, qp 0..51), stored in the
 * reference's own luma array layouts (cpu.h:86-89).  The luma filter of an
 * edge segment then takes beta / tc from that entry's qp instead of the frame
 * QP; everything else (geometry, order, decisions, filters, chroma with the
 * frame QP and the constructor's chroma strengths) is cpu.h unchanged.
 *
 * Parity status of hdbo_deblock_frame_edges: PINNED through the unmodified
 * reference.  tests/test_edges.py checks it (a) with one QP in every entry
 * against the reference itself, (b) with arbitrary per-entry QPs against a
 * chain of reference runs over the luma plane, each at one QP with only that
 * QP's edges enabled (oracle_py.ref_compose_edges: vertical edges by QP, then
 * horizontal edges column by column, which reproduces each block's segment
 * order exactly; maps with the left and bottom picture edges off, since plane
 * padding written by one run is not seen by the next), and (c) against golden hashes of such chains
 * (tests/golden/edges.json, made by oracle/make_golden_edges.py).
 *
 * Parity status of hdbo_derive_edges: UNPINNED against the reference -- it has
 * no counterpart there.  It restates the published H.265 rule (8.7.2.4: intra
 * -> 2; transform edge with coded coefficients -> 1; different reference
 * pictures / motion-vector count, or a motion-vector difference of >= 4
 * quarter samples -> 1; else 0; QP of an edge = (QpP + QpQ + 1) >> 1) and is
 * anchored only on the cases the reference does define: all-intra input gives
 * the reference's "every inner edge 2" at one QP.  The risk of a transcription
 * error is bounded by a second, independently written statement of the clause
 * (oracle/derive_spec.py: set-based, sentence by sentence) that must agree with
 * this one -- and with the product's branch-free form -- on an exhaustive
 * reduced domain of 683 k cases (tests/test_derive_spec.py).
 */
#define HDBO_EDGE(bs, qp) ((uint8_t)(((bs) & 3u) | ((qp) << 2)))

int hdbo_deblock_frame_edges(const uint8_t *in, uint8_t *out, unsigned width, unsigned height,
	unsigned qp, const uint8_t *vert_edges, unsigned num_vert, const uint8_t *hor_edges, unsigned num_hor);

/* ---------------------------------------------------------------------------
 * Samples deeper than 8 bits: SURVEY.md section 8(f) rank 2.  The reference is 8-bit only -- unsigned char planes
 * (cpu.h:1041-1048) and the sample maximum 255 written into its three clips (cpu.h:1202, 1228, 1474) -- so this
 * entry point DEFINES the deeper filter, by the one rule the published standard adds for it (H.265 8.7.2.5.3:
 * beta = beta' * (1 << (BitDepth - 8)), tc = tc' * (1 << (BitDepth - 8))) and the sample maximum
 * (1 << bit_depth) - 1 in place of 255; geometry, order, decisions and filter formulas are cpu.h unchanged.
 *
 * in / out: a planar frame like the 8-bit one (Y, then the two chroma planes, tight rows) of 16-bit little-endian
 * containers holding bit_depth-bit samples in their low bits; 3*w*h/2 samples.  edges != 0: vert / hor hold
 * edge bytes (bs | qp << 2) as in hdbo_deblock_frame_edges.
 *
 * Parity status: bit_depth == 8 is PINNED -- it must equal hdbo_deblock_frame (hence the unmodified reference)
 * on the same samples widened to 16 bits (tests/test_depth.py).  bit_depth > 8 is UNPINNED against the
 * reference: it has nothing there to be compared with.  Anchors: the depth-8 identity above; scale covariance --
 * a frame whose samples are all multiples of 4 filtered at depth 10 is NOT 4x the depth-8 result in general
 * (rounding), but flat regions, QP <= 17 and all-zero strengths are identities at every depth.
 */
int hdbo_deblock_frame16(const uint16_t *in, uint16_t *out, unsigned width, unsigned height,
	unsigned qp, const uint8_t *vert, unsigned num_vert, const uint8_t *hor, unsigned num_hor,
	unsigned bit_depth, int edges);

/* what a decoder knows about one 8x8 luma block (12 bytes) */
typedef struct hdbo_block_info {
	int16_t mv[2][2];   /* [list][x, y] in quarter samples */
	int8_t ref[2];      /* reference picture id per list, < 0 = list unused */
	uint8_t flags;      /* HDBO_BLK_* */
	uint8_t qp;         /* QpY of the block */
} hdbo_block_info;
#define HDBO_BLK_INTRA   1u   /* intra-predicted */
#define HDBO_BLK_CBF     2u   /* its transform block has non-zero coefficients */
#define HDBO_BLK_TU_LEFT 4u   /* its left edge is a transform-block edge */
#define HDBO_BLK_TU_TOP  8u
#define HDBO_BLK_PU_LEFT 16u  /* its left edge is a prediction-block edge */
#define HDBO_BLK_PU_TOP  32u

/* info[(height/8)][(width/8)] -> vert_edges[hdbo_num_vert_bs], hor_edges[hdbo_num_hor_bs] */
void hdbo_derive_edges(const hdbo_block_info *info, unsigned width, unsigned height,
	uint8_t *vert_edges, uint8_t *hor_edges);

/* ---------------------------------------------------------------------------
 * Sample adaptive offset: SURVEY.md section 8(f) rank 3, the in-loop stage that follows deblocking in HEVC.
 * ABSENT FROM THE REFERENCE (its in-loop chain ends with the deblocking filter, cpu.h:134-993), so this is a
 * restatement of the published process (H.265 8.7.3) and PARITY IS UNPINNED BY NATURE.  What bounds the risk of a
 * transcription error: oracle/sao_spec.py, a second statement written independently in numpy terms, must agree
 * with this one on random and exhaustive-small inputs (tests/test_sao.py).
 *
 * Per coding tree block (ctb_size x ctb_size luma samples, ctb_size/2 square for each chroma plane of a 4:2:0
 * frame) and colour component: type 0 off; 1 band offset -- band = sample >> (bit depth - 5), the four
 * consecutive bands starting at band_position (wrapping at 32) get offset[0..3]; 2 edge offset -- the sample is
 * compared with its two neighbours along the class's direction (0 horizontal, 1 vertical, 2 the 135-degree
 * diagonal (up-left, down-right), 3 the 45-degree one (up-right, down-left)), edgeIdx = 2 + sign(c - a) +
 * sign(c - b), categories 0 / 1 / 3 / 4 get offset[0..3] (a local minimum, a concave corner, a convex corner, a
 * local maximum), category 2 nothing; a sample with a neighbour outside the picture is left alone.  Neighbours
 * are always the INPUT (deblocked) samples.  Results are clipped to the sample range.  8-bit samples.
 * params[3][ctb_rows][ctb_cols], ctb_rows = ceil(height / ctb_size), ctb_cols = ceil(width / ctb_size).
 */
typedef struct hdbo_sao_params {
	uint8_t type;        /* 0 off, 1 band offset, 2 edge offset */
	uint8_t position;    /* band offset: first band, 0..31; edge offset: class, 0..3 */
	int8_t offset[4];
} hdbo_sao_params;      /* 6 bytes */

int hdbo_sao_frame(const uint8_t *in, uint8_t *out, unsigned width, unsigned height,
	const hdbo_sao_params *params, unsigned ctb_size);

/* The same for bit_depth-bit samples (8 .. 12) in 16-bit containers, planar like hdbo_deblock_frame16's frames: the
 * band of a sample is sample >> (bit_depth - 5) (8.7.3: bandShift), results are clipped to (1 << bit_depth) - 1,
 * offsets are final values (the standard's SaoOffsetVal; |offset| <= 31 at ten bits).  bit_depth 8 must equal
 * hdbo_sao_frame on the same samples widened (tests/test_sao.py). */
int hdbo_sao_frame16(const uint16_t *in, uint16_t *out, unsigned width, unsigned height,
	const hdbo_sao_params *params, unsigned ctb_size, unsigned bit_depth);

#ifdef __cplusplus
}
#endif
#endif
