// zl_enc_match.cuh -- match-finder definitions shared by the CUDA kernels (zl_enc_kernels.cu), the host API and the
// CPU emulation in tests/emul.
//
// The B200 match finder is NOT a transcription of ZSTD_compressBlock_fast / _doubleFast (zstd.c:30724, 29909); it
// keeps their ingredients -- a short hash of `mls` bytes and (level 3) a second, long hash of 8 bytes, both with the
// reference's multiplicative hash (zstd.c:19809-19841) -- but restructures the search for a SIMT machine:
//   stage 1 (data-parallel): EVERY position is inserted and looks up the nearest earlier position with the same
//            hash in each table; the candidate is verified and its length measured (capped at ZL_M_CAP);
//            result: one u32 per position, M[p] = offset << 8 | length (0 = no match of at least minLen bytes);
//   stage 2 (greedy walk): from p = 0, jump to the next position with M[p] != 0, take that match (extending it past
//            the cap by direct comparison), continue after it.  Repeat-offset codes are assigned while walking
//            (zstd.c:19648-19652, ZSTD_updateRep 19688).
// Tables hold the low 16 bits of positions (blocks are <= 128 KiB): a candidate is rebuilt as the latest position
// <= p with those low bits, so offsets are limited to 65,535 and stale or never-written entries are harmless
// because every candidate is verified against the input.
#pragma once
#include "zl_common.cuh"

#define ZL_M_CAP 255u
// In-block candidates are only verified up to ZL_M_VERIFY bytes per position: a result of ZL_M_VERIFY means "at least that many",
// and the walk (stage 2) extends such a match with the whole warp, 256 bytes per round.  Inside a long run or repeat EVERY position
// has a maximal candidate, and measuring each to 255 bytes made the match kernel 3x slower on such data (rle corpus: 31 -> 11 ms
// per 4,096 blocks) for lengths the walk never looks at.  Matches into a dictionary are measured up to ZL_M_CAP (they are not extended).
#define ZL_M_VERIFY 64u

struct ZlEncParams {       // derived from the compression level by zl_enc_params()
    u32 level;
    u32 mls;               // bytes hashed by the short table (also the minimum match length)
    u32 hlogS;             // log2 entries of the short table
    u32 hlogL;             // log2 entries of the long (8-byte) table, 0 = no long table
    u32 farMaxOff;         // far candidates (below): offsets stay under this; 0 = none.  ZSTD_c_windowLog lowers it.
};
// levels map onto three table layouts (cf. the reference's rows for <= 128 KB inputs, zstd.c:29579-29582):
//   1: fast-like   short 2^14            (32 KB of u16)
//   2: fast-like   short 2^15            (64 KB)
//   3: dfast-like  short 2^14 + long 2^15 (96 KB)
static inline ZlEncParams zl_enc_params(int level)
{
    ZlEncParams p;
    p.level = (u32)level; p.mls = 5; p.farMaxOff = 1u << 24;
    if (level <= 1) { p.hlogS = 14; p.hlogL = 0; }
    else if (level == 2) { p.hlogS = 15; p.hlogL = 0; }
    else { p.hlogS = 14; p.hlogL = 15; }
    return p;
}

// Blocks of at most ZL_SMALL_BLOCK bytes (the small objects of configs[3]: a few hundred bytes each) use 2^10 / 2^11-entry tables whatever
// the level: a CTA zeroing 96 KB of tables for 400 bytes of input was most of their match time, and with 6 KB of tables a WARP can take a
// block of its own (zl_k_match_small).  The table size of a block depends on the block alone, never on what else is in the batch.
#define ZL_SMALL_BLOCK 2048u
#define ZL_SMALL_HLOG_S 10u
#define ZL_SMALL_HLOG_L 11u
ZL_HD void zl_block_hlog(const ZlEncParams& P, u32 blockSize, u32& hlogS, u32& hlogL)
{
    const bool small = blockSize <= ZL_SMALL_BLOCK;
    hlogS = small ? ZL_SMALL_HLOG_S : P.hlogS;
    hlogL = P.hlogL ? (small ? ZL_SMALL_HLOG_L : P.hlogL) : 0u;
}
ZL_HD u32 zl_hash_short(u32 lo, u32 hi, u32 mls, u32 hlog)
{
    const u64 v = ((u64)hi << 32) | lo;
    if (mls == 4) return (lo * 2654435761u) >> (32 - hlog);
    if (mls == 5) return (u32)(((v << 24) * 889523592379ULL) >> (64 - hlog));
    return (u32)(((v << 16) * 227718039650203ULL) >> (64 - hlog));
}
ZL_HD u32 zl_hash_long(u32 lo, u32 hi, u32 hlog)
{
    const u64 v = ((u64)hi << 32) | lo;
    return (u32)((v * 0xCF1BBCDCB7A56463ULL) >> (64 - hlog));
}
// candidate position from a 16-bit table entry: the latest position < p whose low 16 bits are e (-1 if none)
ZL_HD i32 zl_cand_pos(u32 e, u32 p)
{
    i32 q = (i32)((p & ~0xFFFFu) | e);
    if (q >= (i32)p) q -= 65536;
    return q;
}
// number of equal leading bytes of two 8-byte little-endian words (8 when identical)
ZL_HD u32 zl_common8(u32 alo, u32 ahi, u32 blo, u32 bhi)
{
    const u32 xl = alo ^ blo, xh = ahi ^ bhi;
#if defined(__CUDA_ARCH__)
    if (xl) return (u32)(__ffs((int)xl) - 1) >> 3;
    if (xh) return 4 + ((u32)(__ffs((int)xh) - 1) >> 3);
#else
    if (xl) return (u32)__builtin_ctz(xl) >> 3;
    if (xh) return 4 + ((u32)__builtin_ctz(xh) >> 3);
#endif
    return 8;
}

// ---- far candidates: frames of more than one block -----------------------------------------------------------------------
// Blocks are searched independently with 16-bit positions, so a match never reaches further back than 64 KiB nor across a block
// edge.  For 128 KiB frames that is the reference's own window; for a large single buffer -- zstd_compress / zstd_serialize of a
// real object -- libzstd uses a 2 MB window (zstd.c:29527, 23978) and text-like data lost 12 % to it.  Such frames now get ONE more
// table: for every 8-byte hash the EARLIEST position of the frame where it occurs (zl_k_far_build: an atomicMin per position, so the
// table does not depend on any order).  Every position looks its hash up; an earlier occurrence beyond the reach of the block tables
// is verified like any candidate and taken when it is clearly longer than the near one (its offset costs more bits).  The frame
// header then declares a window that covers the frame, at most 16 MiB (decoders accept 2^27 by default, zstd.c:42431).
#define ZL_FAR_MAX_OFF (1u << 24)          // M[p] keeps 24 bits of offset
#define ZL_FAR_MIN_LOG 17u
// Frames beyond 8 MiB are covered REGION by region (8 MiB each, one table per region, the earliest occurrence inside that region): a
// position consults its own region's table and, when that gives nothing, the previous region's -- every such candidate is less than
// 16 MiB back, which is what M[p] can hold and what the window descriptor then declares.  (One frame-wide table of earliest
// occurrences stops helping once the earliest occurrence is out of reach: text of 48 MiB was back at 1.07x libzstd.)
#define ZL_FAR_REGION_LOG 23u
#ifndef ZL_FAR_MULTI_LOG
#define ZL_FAR_MULTI_LOG 23u               // entries per region table of a multi-region frame (one per position)
#endif
// the previous region's table is only consulted from the first ZL_FAR_PREV_SPAN bytes of a region: further in, the region's own table
// already covers that much history (libzstd's level-3 window is 2 MiB, zstd.c:29527)
#ifndef ZL_FAR_PREV_SPAN
#define ZL_FAR_PREV_SPAN (1u << 21)
#endif
ZL_HD bool zl_far_use_prev(u32 pos) { return pos >= (1u << ZL_FAR_REGION_LOG) && (pos & ((1u << ZL_FAR_REGION_LOG) - 1)) < ZL_FAR_PREV_SPAN; }
ZL_HD u32 zl_far_regions(u64 n) { return (u32)((n + (1ull << ZL_FAR_REGION_LOG) - 1) >> ZL_FAR_REGION_LOG); }
// table size (log2 entries per region) for a frame of n bytes: a single region gets twice as many entries as positions (earliest-wins:
// a crowded table loses the later content)
ZL_HD u32 zl_far_log(u64 n)
{
    if (n > (1ull << ZL_FAR_REGION_LOG)) return ZL_FAR_MULTI_LOG;
    u32 l = ZL_FAR_MIN_LOG;
    while (l < ZL_FAR_REGION_LOG + 1 && (1ull << l) < 2 * n) l++;
    return l;
}
// A table entry is (position inside the region) << 8 | 8 more bits of the hash: atomicMin still keeps the earliest position, and a lookup
// whose tag differs is a different string -- it is dropped without touching the candidate's bytes (a second random DRAM sector).
// zl_far_hash: table index << 8 | tag (log <= 24, so it fits 32 bits)
ZL_HD u32 zl_far_hash(u32 lo, u32 hi, u32 flog) { const u64 v = ((u64)hi << 32) | lo; return (u32)((v * 0xCF1BBCDCB7A56463ULL) >> (56 - flog)); }
#define ZL_FAR_EMPTY 0xFFFFFFFFu
ZL_HD u64 zl_far_entries(u64 n) { return (u64)zl_far_regions(n) << zl_far_log(n); }
// Is a far match of lenFar bytes better than the near one of bestLen bytes (0 = none)?  A far offset costs 2-3 bytes more than a near or
// repeated one, so it must be clearly longer: lenFar > bestLen + 2 + bestLen / 2, and at least 8 bytes.  Measured on 2 - 16 MB buffers against
// libzstd level 3 (tests/emul, ours / theirs): text 1.116 -> 1.013, columnar 0.965 -> 0.973; an eager rule (longer by 2) gives text 1.010 but
// columnar 1.000 -- far matches then displace the cheap repeat-offset matches such data lives on.
#ifndef ZL_FAR_MARGIN
#define ZL_FAR_MARGIN 2u
#endif
#ifndef ZL_FAR_MINLEN
#define ZL_FAR_MINLEN 8u
#endif
#ifndef ZL_FAR_REL
#define ZL_FAR_REL 2u
#endif
ZL_HD bool zl_far_better(u32 lenFar, u32 bestLen, u32 mls) { (void)mls; return lenFar >= ZL_FAR_MINLEN && lenFar > bestLen + ZL_FAR_MARGIN + ((bestLen * ZL_FAR_REL) >> 2); }
// window descriptor byte of a multi-block frame of n bytes (zstd.c:41115-41121: window = (1 << (10 + exponent)) * (1 + mantissa / 8))
ZL_HD u32 zl_window_descriptor(u64 n, u32 farMaxOff)
{
    u32 wlog = 17;
    while ((1ull << wlog) < farMaxOff && (1ull << wlog) < n) wlog++;      // far offsets stay below farMaxOff (0: no far candidates)
    return (wlog - 10) << 3;
}

// The greedy walk (stage 2) is a serial chain; a block is walked as independent SEGMENTS of ZL_PARSE_SEG bytes (one warp each,
// zl_k_parse).  A segment starts with an unknown repeat-offset history (only the first segment of a frame knows the decoder's)
// and clips its matches at its end; the literals it ends with are added to the litLength of the next sequence of the block.
#ifndef ZL_PARSE_SEG
#define ZL_PARSE_SEG 16384u
#endif
#define ZL_PARSE_SEG_RECS (ZL_PARSE_SEG / 5 + 3)          // records of a segment: matches are >= 5 bytes, the clipped last one >= 3

// repeat-offset bookkeeping while walking (zstd.c:19648-19652 offBase, 19688 ZSTD_updateRep): returns offBase
struct ZlReps { u32 r0, r1, r2; };
ZL_HD u32 zl_rep_encode(ZlReps& r, u32 off, u32 ll)
{
    u32 ob;
    if (ll) {
        if (off == r.r0) ob = 1; else if (off == r.r1) ob = 2; else if (off == r.r2) ob = 3; else ob = off + 3;
        if (ob == 2) { const u32 t = r.r1; r.r1 = r.r0; r.r0 = t; }
        else if (ob == 3) { const u32 t = r.r2; r.r2 = r.r1; r.r1 = r.r0; r.r0 = t; }
        else if (ob > 3) { r.r2 = r.r1; r.r1 = r.r0; r.r0 = off; }
    } else {                                        // litLength == 0 shifts the meaning of the codes (zstd.c:44311-44324)
        if (off == r.r1) ob = 1; else if (off == r.r2) ob = 2; else if (off == r.r0 - 1 && off) ob = 3; else ob = off + 3;
        if (ob == 1) { const u32 t = r.r1; r.r1 = r.r0; r.r0 = t; }
        else { r.r2 = r.r1; r.r1 = r.r0; r.r0 = off; }                   // codes 2 (old r2), 3 (r0 - 1) and new offsets all push the history
    }
    return ob;
}

// ---- frame / block headers (zstd.c:27089-27135 ZSTD_writeFrameHeader, 19580 block header) ------------------------
// Frames of <= 128 KiB are single-segment like the reference's (window >= content).  Larger inputs are written as one
// frame of 128 KiB blocks; its window descriptor says 128 KiB when nothing refers back across a block, or covers the frame
// when far candidates are in use (above).
ZL_HD u32 zl_write_frame_header(u8* dst, u64 contentSize, u32 dictID, u32 checksumFlag, u32 farMaxOff = 0)
{
    const u32 single = contentSize <= ZL_BLOCKSIZE_MAX ? 1u : 0u;
    const u32 didCode = dictID == 0 ? 0u : (dictID < 256 ? 1u : (dictID < 65536 ? 2u : 3u));
    u32 fcsCode;
    if (single) fcsCode = contentSize >= 256 ? (contentSize >= 65536 + 256 ? 2u : 1u) : 0u;
    else fcsCode = contentSize >= 0xFFFFFFFFull ? 3u : 2u;
    u32 p = 0;
    dst[p++] = 0x28; dst[p++] = 0xB5; dst[p++] = 0x2F; dst[p++] = 0xFD;
    dst[p++] = (u8)(didCode | (checksumFlag << 2) | (single << 5) | (fcsCode << 6));
    if (!single) dst[p++] = (u8)zl_window_descriptor(contentSize, farMaxOff);
    if (didCode == 1) dst[p++] = (u8)dictID;
    else if (didCode == 2) { dst[p++] = (u8)dictID; dst[p++] = (u8)(dictID >> 8); }
    else if (didCode == 3) { for (u32 i = 0; i < 4; i++) dst[p++] = (u8)(dictID >> (8 * i)); }
    if (fcsCode == 0) { if (single) dst[p++] = (u8)contentSize; }
    else if (fcsCode == 1) { const u32 v = (u32)contentSize - 256; dst[p++] = (u8)v; dst[p++] = (u8)(v >> 8); }
    else if (fcsCode == 2) { for (u32 i = 0; i < 4; i++) dst[p++] = (u8)(contentSize >> (8 * i)); }
    else { for (u32 i = 0; i < 8; i++) dst[p++] = (u8)(contentSize >> (8 * i)); }
    return p;
}
ZL_HD void zl_write_block_header(u8* dst, u32 last, u32 type, u32 size)
{
    const u32 v = last | (type << 1) | (size << 3);
    dst[0] = (u8)v; dst[1] = (u8)(v >> 8); dst[2] = (u8)(v >> 16);
}
