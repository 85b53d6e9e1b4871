"""TEST INFRASTRUCTURE ONLY -- a second, independently written statement of the boundary-strength rule.

The reference has no boundary-strength derivation (strengths are constants, hevc_deblocking_filter_cpu.h:91-99),
so oracle/deblock_oracle.c::hdbo_derive_edges cannot be pinned on it.  What can be done is to bound the
risk of a transcription error: this file restates the published rule (ITU-T H.265, clause 8.7.2.4, "Derivation
process of boundary filtering strength") a second time, sentence by sentence and in terms of SETS of reference
pictures the way the clause is worded, sharing no code and no case analysis with the C oracle (which works
list by list) nor with the product's branch-free DeriveBs (deblock_core.cuh).  tests/test_derive_spec.py
enumerates a reduced domain exhaustively -- every flag combination x every used / unused reference pair over
three pictures x motion vectors at the threshold and at both ends of the 16-bit range -- and demands that the
three agree on every case.  Status of the derivation stays "parity unpinned against the reference".

Pure Python on purpose: small cases only.
"""
from collections import Counter

INTRA, CBF, TU_LEFT, TU_TOP, PU_LEFT, PU_TOP = 1, 2, 4, 8, 16, 32


def _motion(block):
    """The motion vectors a block uses, each with the picture it points into: [(picture, (mvx, mvy)), ...]."""
    return [(block["ref"][l], tuple(block["mv"][l])) for l in (0, 1) if block["ref"][l] >= 0]


def _far(a, b):
    """'the absolute difference between the horizontal or vertical component of the motion vectors used is
    greater than or equal to 4 in units of quarter luma samples'"""
    return abs(a[0] - b[0]) >= 4 or abs(a[1] - b[1]) >= 4


def boundary_strength(p, q, transform_edge, prediction_edge):
    """bS of the edge between the block holding sample p0 and the block holding sample q0 (dicts with
    "mv" [[x, y], [x, y]], "ref" [r0, r1] (< 0: list unused), "flags")."""
    # only edges of the 8x8 grid that are transform block edges or prediction block edges are filtered at all
    if not (transform_edge or prediction_edge):
        return 0
    # "the sample p0 or q0 is in the coding block of a coding unit coded with intra prediction mode": 2
    if (p["flags"] & INTRA) or (q["flags"] & INTRA):
        return 2
    # "the block edge is also a transform block edge and the sample p0 or q0 is in a luma transform block which
    #  contains one or more non-zero transform coefficient levels": 1
    if transform_edge and ((p["flags"] & CBF) or (q["flags"] & CBF)):
        return 1
    mp, mq = _motion(p), _motion(q)
    # "different reference pictures or a different number of motion vectors are used" -- which pictures are
    # referenced, without regard to the list the index came through (NOTE 1 of the clause): compare multisets
    pics_p, pics_q = Counter(r for r, _ in mp), Counter(r for r, _ in mq)
    if len(mp) != len(mq) or pics_p != pics_q:
        return 1
    # "one motion vector is used to predict [each of the two blocks]" and they are far apart
    if len(mp) == 1:
        return 1 if _far(mp[0][1], mq[0][1]) else 0
    if len(mp) == 2:
        if len(pics_p) == 2:
            # "two motion vectors and two different reference pictures ... two motion vectors for the same two
            #  reference pictures": compare the vectors that point into the same picture
            for pic in pics_p:
                vp = [v for r, v in mp if r == pic][0]
                vq = [v for r, v in mq if r == pic][0]
                if _far(vp, vq):
                    return 1
            return 0
        # "two motion vectors for the same reference picture are used [in both blocks]": 1 only when BOTH the
        # list0-list0 / list1-list1 pairing and the list0-list1 / list1-list0 pairing have a far pair
        straight = _far(mp[0][1], mq[0][1]) or _far(mp[1][1], mq[1][1])
        crossed = _far(mp[0][1], mq[1][1]) or _far(mp[1][1], mq[0][1])
        return 1 if (straight and crossed) else 0
    return 0     # neither block uses a motion vector and neither is intra: nothing distinguishes them


def edge_qp(p, q):
    """QpL of clause 8.7.2.5.3: ((QpQ + QpP + 1) >> 1); the tables stop at 51."""
    return min((p["qp"] + q["qp"] + 1) >> 1, 51)


def edge_byte(p, q, transform_edge, prediction_edge):
    return boundary_strength(p, q, transform_edge, prediction_edge) | (edge_qp(p, q) << 2)
