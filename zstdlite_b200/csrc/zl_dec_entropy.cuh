// zl_dec_entropy.cuh -- entropy stage of the B200 Zstandard decoder (kernel K1).
//
// Work unit: one FRAME per QUAD (4 lanes), 8 frames per warp.  Lane roles inside a quad:
//   lane 0      parses frame/block/literal/sequence headers (tiny, serial), runs the serial
//               3-state FSE sequence decode and writes sequence records
//   lanes 0..3  fill the Huffman table cooperatively and decode the 4 Huffman streams
//               (one stream per lane -- HUF_decompress4X, zstd.c:38653)
//   lanes 0..2  build the LL / OF / ML FSE decode tables concurrently (zstd.c:43497)
// All tables live in shared memory (ZlFrameSm, ~10 KB per frame).  Output goes to scratch
// arenas in HBM: decoded literals, u64 sequence records (+ one checkpoint per 32 records)
// and one ZlBlockHdr per block; the execute kernel (zl_dec_exec.cuh) consumes them.
//
// Reference behaviour restated here (file:line in /root/reference/src/zstd/zstd.c):
//   frame header 41050-41152, block header 43075, literals section 43146-43347,
//   HUF_readStats 3470, FSE_readNCount 3269, HUF_readDTableX1 38436, sequence headers 43707,
//   ZSTD_buildFSETable 43497, ZSTD_decodeSequence 44241, end-of-stream checks 44686 / 38647.
#pragma once
#include "zl_common.cuh"

// ---- packed FSE decode cell: newStateBase[0:10) nbBits[10:14) nbAdditionalBits[14:19) symbol[19:25)
ZL_HD u32 zl_fse_pack(u32 base, u32 nb, u32 add, u32 sym) { return base | (nb << 10) | (add << 14) | (sym << 19); }
#define ZL_FSE_BASE(e) ((e) & 1023u)
#define ZL_FSE_NB(e) (((e) >> 10) & 15u)
#define ZL_FSE_ADD(e) (((e) >> 14) & 31u)
#define ZL_FSE_SYM(e) (((e) >> 19) & 63u)

struct ZlEntCtl {
    u32 err;            // sticky ZlErr
    u32 done;
    u32 pos;            // byte offset of the next unread byte in the frame
    u32 srcSize;
    u32 outPos;         // bytes regenerated so far in this frame
    u32 blockIdx;
    u32 blockSizeMax;
    u32 hufValid;
    u32 fseValid;       // bit t: table t (0 LL, 1 OF, 2 ML) holds a usable table
    u32 rep[3];
    u32 recUsed, ckUsed, litUsed;
    u32 dictSize;       // bytes of history available before the frame start
    u64 contentSize;
    u32 checksumFlag;
    // block level
    u32 isCompressed, last, blockEnd;
    u32 litMode, litSize, litOff, litSrcOff, rleByte;
    u32 needHufFill, hufLog, nsym;
    u32 nStreams;
    u32 sBeg[4], sEnd[4], sOut[4], sLen[4], sErr[4];
    u32 nbSeq;
    u32 needBuild;      // bit t: build table t from norm[t]
    u32 tlog[3], maxSym[3];
    u32 bitBeg;
};

struct ZlFrameSm {
    u32 fseLL[512];
    u32 fseML[512];
    u32 fseOF[256];
    u16 huf[2048];
    ZlEntCtl ctl;
    i16 norm[3][64];
    u8 weights[256];
    union {
        u16 symStart[256];
        u32 wtbl[64];
    } u;
};

// ---- backward bit reader over aligned 32-bit words (restates BIT_DStream_t, zstd.c:2352-2550) -------
// `wbase` is the frame's src pointer rounded down to 4 bytes, `bias` = src - wbase (0..3); stream
// positions are byte offsets relative to src.  Bits are consumed from the MSB side of `acc`.
struct ZlBitR {
    u64 acc;
    i32 avail;
    i32 wi;      // next (lower) word to load
    i32 wlow;    // word holding the first byte of the stream
};

ZL_HD bool zl_br_init(ZlBitR& b, const u32* wbase, u32 bias, u32 beg, u32 end)
{
    if (end <= beg) return false;
    u32 a = end - 1 + bias;
    u32 W = wbase[a >> 2];
    u32 bsel = a & 3;
    u32 L = (W >> (8 * bsel)) & 0xFF;
    if (!L) return false;                      // zstd.c:2369 endMark missing
    u32 hb = zl_highbit(L);
    u32 sh = 32 + 8 * (3 - bsel) + (8 - hb);   // drop bytes above the stream, padding and end mark
    b.acc = sh >= 64 ? 0 : ((u64)W << sh);
    b.avail = (i32)(8 * bsel + hb);
    b.wi = (i32)(a >> 2) - 1;
    b.wlow = (i32)((beg + bias) >> 2);
    return true;
}
ZL_HD void zl_br_refill(ZlBitR& b, const u32* wbase)
{
    if (b.avail <= 32) {
        u32 w = (b.wi >= b.wlow) ? wbase[b.wi] : 0u;
        b.wi--;
        b.acc |= (u64)w << (32 - b.avail);
        b.avail += 32;
    }
}
ZL_HD u32 zl_br_take(ZlBitR& b, u32 n)         // n <= 32, n may be 0
{
    u32 v = (u32)((b.acc >> 1) >> (63 - n));
    b.acc <<= n;
    b.avail -= (i32)n;
    return v;
}
// bits of the stream not consumed yet; negative when the reader ran past the start
ZL_HD i32 zl_br_remaining(const ZlBitR& b, u32 bias, u32 beg)
{
    return b.avail + 8 * (4 * (b.wi + 1) - (i32)(beg + bias));
}

// ---- FSE normalized-count header (forward, LSB first): zstd.c:3269-3409 ----------------------------
// Returns bytes consumed, or 0 on error.
ZL_HD u32 zl_read_ncount(const u8* src, u32 srcSize, i16* norm, u32* maxSymIO, u32* tableLog)
{
    if (srcSize < 1) return 0;
    u32 maxSV1 = *maxSymIO + 1, sym = 0;
    u32 bitpos = 4;
    for (u32 i = 0; i < maxSV1; i++) norm[i] = 0;
    i32 nbBits = (i32)(src[0] & 15) + 5;
    if (nbBits > 15) return 0;
    *tableLog = (u32)nbBits;
    i32 remaining = (1 << nbBits) + 1, threshold = 1 << nbBits;
    nbBits++;
    bool prev0 = false;
    for (;;) {
        if (prev0) {
            for (;;) {
                u32 by = bitpos >> 3;
                u32 w = (by < srcSize ? src[by] : 0u) | ((by + 1 < srcSize ? (u32)src[by + 1] : 0u) << 8);
                w = (w >> (bitpos & 7)) & 3;
                bitpos += 2;
                sym += w;
                if (w != 3) break;
                if (sym >= maxSV1) break;
            }
            if (sym >= maxSV1) break;
        }
        {
            u32 by = bitpos >> 3;
            u32 v = 0;
            for (u32 i = 0; i < 4; i++) v |= (by + i < srcSize ? (u32)src[by + i] : 0u) << (8 * i);
            v >>= (bitpos & 7);                     // >= 25 valid bits, fields are <= 16 bits
            i32 max = (2 * threshold - 1) - remaining, count;
            if ((i32)(v & (u32)(threshold - 1)) < max) {
                count = (i32)(v & (u32)(threshold - 1));
                bitpos += (u32)nbBits - 1;
            } else {
                count = (i32)(v & (u32)(2 * threshold - 1));
                if (count >= threshold) count -= max;
                bitpos += (u32)nbBits;
            }
            count--;
            if (count >= 0) remaining -= count; else remaining += count;
            norm[sym++] = (i16)count;
            prev0 = (count == 0);
            if (remaining < threshold) {
                if (remaining <= 1) break;
                nbBits = (i32)zl_highbit((u32)remaining) + 1;
                threshold = 1 << (nbBits - 1);
            }
            if (sym >= maxSV1) break;
        }
    }
    if (remaining != 1) return 0;
    if (sym > maxSV1) return 0;
    *maxSymIO = sym - 1;
    u32 used = (bitpos + 7) >> 3;
    if (used > srcSize) return 0;
    return used;
}

// ---- FSE decode table build (one lane per table): zstd.c:43497-43613 -----------------------------------
// kind 0 LL, 1 OF, 2 ML decides the additional-bits field.  `tbl` first receives symbols, then cells.
ZL_HD bool zl_fse_build(u32* tbl, const i16* norm, u32 maxSym, u32 log, u32 kind, const ZlConstTables& ct)
{
    u32 size = 1u << log, high = size - 1, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    u16 next[64];
    for (u32 s = 0; s <= maxSym; s++) {
        if (norm[s] == -1) { tbl[high--] = s; next[s] = 1; }
        else next[s] = (u16)norm[s];
    }
    u32 pos = 0;
    for (u32 s = 0; s <= maxSym; s++) {
        i32 n = norm[s];
        for (i32 i = 0; i < n; i++) {
            tbl[pos] = s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    }
    if (pos != 0) return false;
    for (u32 u = 0; u < size; u++) {
        u32 s = tbl[u];
        u32 ns = next[s]++;
        u32 nb = log - zl_highbit(ns);
        u32 add = kind == 0 ? ct.llBits[s] : (kind == 1 ? s : ct.mlBits[s]);
        tbl[u] = zl_fse_pack((ns << nb) - size, nb, add, s);
    }
    return true;
}

// ---- frame header (lane 0): zstd.c:41050-41152 --------------------------------------------------------
ZL_HD void zl_ent_begin_frame(ZlFrameSm& f, const ZlFrameDesc& d, ZlFrameInfo& info, u32 dictSize, u32 dictID)
{
    ZlEntCtl& c = f.ctl;
    c.err = 0; c.done = 0; c.pos = 0; c.srcSize = d.srcSize; c.outPos = 0; c.blockIdx = 0;
    c.hufValid = 0; c.fseValid = 0; c.rep[0] = 1; c.rep[1] = 4; c.rep[2] = 8;     // zstd.c:15416
    c.recUsed = 0; c.ckUsed = 0; c.litUsed = 0; c.dictSize = dictSize;
    c.contentSize = ~0ull; c.checksumFlag = 0; c.isCompressed = 0; c.blockSizeMax = ZL_BLOCKSIZE_MAX;
    info.err = 0; info.nblocks = 0; info.contentSize = ~0ull; info.totalOut = 0;
    info.checksumFlag = 0; info.checksum = 0; info.dictID = 0;
    const u8* ip = d.src;
    if (d.srcSize < 5) { c.err = ZL_E_srcSize_wrong; return; }
    if (zl_rd32(ip) != ZL_MAGIC) { c.err = ZL_E_prefix_unknown; return; }
    u32 fhd = ip[4];
    u32 didCode = fhd & 3, single = (fhd >> 5) & 1, fcsID = fhd >> 6;
    u32 didSz = didCode == 3 ? 4 : didCode, fcsSz = fcsID == 0 ? (single ? 1u : 0u) : (1u << fcsID);
    u32 hs = 5 + (single ? 0 : 1) + didSz + fcsSz;
    if (d.srcSize < hs) { c.err = ZL_E_srcSize_wrong; return; }
    if (fhd & 8) { c.err = ZL_E_frameParameter_unsupported; return; }
    u32 p = 5;
    u64 window = 0;
    if (!single) {
        u32 wl = ip[p++];
        u32 wlog = (wl >> 3) + 10;
        if (wlog > 31) { c.err = ZL_E_frameParameter_windowTooLarge; return; }
        window = 1ull << wlog;
        window += (window >> 3) * (wl & 7);
    }
    u32 did = 0;
    if (didSz == 1) did = ip[p]; else if (didSz == 2) did = zl_rd16(ip + p); else if (didSz == 4) did = zl_rd32(ip + p);
    p += didSz;
    u64 fcs = ~0ull;
    if (fcsID == 0) { if (single) fcs = ip[p]; }
    else if (fcsID == 1) fcs = zl_rd16(ip + p) + 256;
    else if (fcsID == 2) fcs = zl_rd32(ip + p);
    else fcs = zl_rd64(ip + p);
    if (single) window = fcs;
    if (did != 0 && did != dictID) { c.err = ZL_E_dictionary_wrong; return; }     // zstd.c:41318
    c.blockSizeMax = window < ZL_BLOCKSIZE_MAX ? (u32)window : ZL_BLOCKSIZE_MAX;
    c.contentSize = fcs; c.checksumFlag = (fhd >> 2) & 1;
    c.pos = hs;
    info.contentSize = fcs; info.checksumFlag = c.checksumFlag; info.dictID = did;
}

ZL_HD void zl_ent_finish_frame(ZlFrameSm& f, const ZlFrameDesc& d, ZlFrameInfo& info)
{
    ZlEntCtl& c = f.ctl;
    c.done = 1;
    if (!c.err) {
        if (c.contentSize != ~0ull && c.contentSize != (u64)c.outPos) c.err = ZL_E_corruption_detected;   // zstd.c:41646
        else if (c.checksumFlag) {
            if (c.pos + 4 > c.srcSize) c.err = ZL_E_checksum_wrong;                                          // zstd.c:41651
            else { info.checksum = zl_rd32(d.src + c.pos); c.pos += 4; }
        }
        if (!c.err && c.pos != c.srcSize) c.err = ZL_E_srcSize_wrong;       // one frame per batch item
    }
    info.err = c.err; info.nblocks = c.blockIdx; info.totalOut = c.outPos;
}

// ---- Huffman tree description (lane 0): zstd.c:3470-3540, weights via FSE 3875/3790 ----------------------
// Fills f.weights / f.u.symStart, ctl.hufLog, ctl.nsym.  Returns bytes consumed or 0 on error.
ZL_HD u32 zl_huf_read_stats(ZlFrameSm& f, const u8* src, u32 srcSize, const u32* wbase, u32 bias, u32 srcOff)
{
    if (!srcSize) return 0;
    u32 hdr = src[0], nsym = 0, iSize;
    u8* w = f.weights;
    if (hdr >= 128) {
        nsym = hdr - 127;
        iSize = (nsym + 1) / 2;
        if (iSize + 1 > srcSize) return 0;
        for (u32 n = 0; n < nsym; n++) { u32 b = src[1 + n / 2]; w[n] = (u8)((n & 1) ? (b & 15) : (b >> 4)); }
    } else {
        iSize = hdr;
        if (iSize + 1 > srcSize) return 0;
        i16 norm[16];
        u32 maxSym = 12, tlog;                 // weights are <= HUF_TABLELOG_MAX (12); FSE_MAX_SYMBOL_VALUE is 255 in
        // the reference but any symbol > 12 is rejected at zstd.c:3509, so a 13-symbol alphabet is equivalent.
        u32 nc = zl_read_ncount(src + 1, iSize, norm, &maxSym, &tlog);
        if (!nc || tlog > 6) return 0;
        u32* t = f.u.wtbl;
        {   // tiny FSE table (zstd.c:3692-3790)
            u32 size = 1u << tlog, high = size - 1, mask = size - 1, step = (size >> 1) + (size >> 3) + 3, pos = 0;
            u16 next[16];
            for (u32 s = 0; s <= maxSym; s++) { if (norm[s] == -1) { t[high--] = s; next[s] = 1; } else next[s] = (u16)norm[s]; }
            for (u32 s = 0; s <= maxSym; s++) for (i32 i = 0; i < norm[s]; i++) {
                t[pos] = s; pos = (pos + step) & mask; while (pos > high) pos = (pos + step) & mask; }
            if (pos != 0) return 0;
            for (u32 u = 0; u < size; u++) { u32 s = t[u]; u32 ns = next[s]++; u32 nb = tlog - zl_highbit(ns);
                t[u] = zl_fse_pack((ns << nb) - size, nb, 0, s); }
        }
        ZlBitR b;
        u32 beg = srcOff + 1 + nc, end = srcOff + 1 + iSize;
        if (!zl_br_init(b, wbase, bias, beg, end)) return 0;
        zl_br_refill(b, wbase);
        u32 s1 = zl_br_take(b, tlog), s2 = zl_br_take(b, tlog);
        for (;;) {                                   // alternate the two states until the stream overflows
            if (nsym > 253) return 0;
            zl_br_refill(b, wbase);
            u32 e = t[s1]; w[nsym++] = (u8)ZL_FSE_SYM(e); s1 = ZL_FSE_BASE(e) + zl_br_take(b, ZL_FSE_NB(e));
            if (zl_br_remaining(b, bias, beg) < 0) { w[nsym++] = (u8)ZL_FSE_SYM(t[s2]); break; }
            if (nsym > 253) return 0;
            e = t[s2]; w[nsym++] = (u8)ZL_FSE_SYM(e); s2 = ZL_FSE_BASE(e) + zl_br_take(b, ZL_FSE_NB(e));
            if (zl_br_remaining(b, bias, beg) < 0) { w[nsym++] = (u8)ZL_FSE_SYM(t[s1]); break; }
        }
    }
    u32 rank[13], total = 0;
    for (u32 i = 0; i < 13; i++) rank[i] = 0;
    for (u32 n = 0; n < nsym; n++) { if (w[n] > 12) return 0; rank[w[n]]++; total += (1u << w[n]) >> 1; }
    if (!total) return 0;
    u32 tableLog = zl_highbit(total) + 1;
    if (tableLog > 11) return 0;       // format limit (RFC 8878 4.2.1); libzstd tolerates 12, see DESIGN.md
    u32 rest = (1u << tableLog) - total, last = zl_highbit(rest) + 1;
    if ((1u << zl_highbit(rest)) != rest) return 0;
    w[nsym++] = (u8)last; rank[last]++;
    if (rank[1] < 2 || (rank[1] & 1)) return 0;
    u32 start[13], acc = 0;
    for (u32 wv = 1; wv <= tableLog; wv++) { start[wv] = acc; acc += rank[wv] << (wv - 1); }
    for (u32 n = 0; n < nsym; n++) { u32 wv = w[n]; if (wv) { f.u.symStart[n] = (u16)start[wv]; start[wv] += 1u << (wv - 1); } }
    f.ctl.hufLog = tableLog; f.ctl.nsym = nsym;
    return iSize + 1;
}

// cooperative table fill, lane q of 4: zstd.c:38506-38568 (cell = nbBits<<8 | symbol)
ZL_HD void zl_huf_fill(ZlFrameSm& f, u32 q)
{
    u32 tlog = f.ctl.hufLog, nsym = f.ctl.nsym;
    for (u32 n = q; n < nsym; n += 4) {
        u32 wv = f.weights[n];
        if (!wv) continue;
        u32 len = 1u << (wv - 1), st = f.u.symStart[n];
        u16 cell = (u16)(((tlog + 1 - wv) << 8) | n);
        for (u32 k = 0; k < len; k++) f.huf[st + k] = cell;
    }
}

// one Huffman stream per lane: zstd.c:38626-38650.  Returns 0 when the stream ends exactly.
ZL_HD u32 zl_huf_stream(const u16* huf, u32 tlog, const u32* wbase, u32 bias, u32 beg, u32 end, u8* out, u32 n)
{
    ZlBitR b;
    if (!zl_br_init(b, wbase, bias, beg, end)) return 1;
    const u32 sh = 64 - tlog;
    u32 i = 0;
#define ZL_HUF_SYM(dst) { u32 e = huf[b.acc >> sh]; u32 nb = e >> 8; b.acc <<= nb; b.avail -= (i32)nb; dst = e & 0xFF; }
    while (i < n && (((size_t)(out + i)) & 3)) { u32 s; zl_br_refill(b, wbase); ZL_HUF_SYM(s); out[i++] = (u8)s; }
    while (i + 4 <= n) {
        u32 s0, s1, s2, s3;
        zl_br_refill(b, wbase); ZL_HUF_SYM(s0); ZL_HUF_SYM(s1);
        zl_br_refill(b, wbase); ZL_HUF_SYM(s2); ZL_HUF_SYM(s3);
        *(u32*)(out + i) = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
        i += 4;
    }
    while (i < n) { u32 s; zl_br_refill(b, wbase); ZL_HUF_SYM(s); out[i++] = (u8)s; }
#undef ZL_HUF_SYM
    return zl_br_remaining(b, bias, beg) == 0 ? 0u : 1u;
}

// ---- block header + literals section header (lane 0) --------------------------------------------------
ZL_HD void zl_ent_block_head(ZlFrameSm& f, const ZlFrameDesc& d, ZlFrameInfo& info, ZlBlockHdr* hdrs,
                             const u32* wbase, u32 bias)
{
    ZlEntCtl& c = f.ctl;
    if (c.done) return;
    c.isCompressed = 0; c.needHufFill = 0; c.nStreams = 0;
    if (c.err) { zl_ent_finish_frame(f, d, info); return; }
    const u8* src = d.src;
    if (c.pos + 3 > c.srcSize) { c.err = ZL_E_srcSize_wrong; zl_ent_finish_frame(f, d, info); return; }
    u32 bh = zl_rd24(src + c.pos);
    u32 last = bh & 1, type = (bh >> 1) & 3, csize = bh >> 3;
    c.pos += 3; c.last = last;
    if (c.blockIdx >= d.hdrCap) { c.err = ZL_E_GENERIC; zl_ent_finish_frame(f, d, info); return; }
    ZlBlockHdr h;
    h.flags = 0; h.regenSize = 0; h.srcOff = 0; h.litOff = 0; h.litSize = 0; h.nrec = 0; h.recOff = c.recUsed; h.ckOff = c.ckUsed;
    if (type == 3) c.err = ZL_E_corruption_detected;
    else if (type == 1) {                                                    // RLE block, zstd.c:41510
        if (c.pos + 1 > c.srcSize) c.err = ZL_E_srcSize_wrong;
        else if (csize > d.dstCap - c.outPos) c.err = ZL_E_dstSize_tooSmall;
        else { h.flags = 1u | ((u32)src[c.pos] << 8); h.regenSize = csize; c.pos += 1; }
    } else if (type == 0) {                                                  // raw block, zstd.c:41497
        if (csize > c.srcSize - c.pos) c.err = ZL_E_srcSize_wrong;
        else if (csize > d.dstCap - c.outPos) c.err = ZL_E_dstSize_tooSmall;
        else { h.flags = 0; h.regenSize = csize; h.srcOff = c.pos; c.pos += csize; }
    } else {
        if (csize > c.srcSize - c.pos) c.err = ZL_E_srcSize_wrong;
        else if (csize > c.blockSizeMax) c.err = ZL_E_srcSize_wrong;         // zstd.c:45099
        else if (csize < 2) c.err = ZL_E_corruption_detected;                // MIN_CBLOCK_SIZE, zstd.c:43150
        else {
            c.isCompressed = 1; c.blockEnd = c.pos + csize;
            const u8* ip = src + c.pos;
            u32 ltype = ip[0] & 3, fmt = (ip[0] >> 2) & 3, lhSize, litSize, litCSize = 0;
            u32 outCap = d.dstCap - c.outPos;
            u32 litCap = outCap < c.blockSizeMax ? outCap : c.blockSizeMax;
            if (ltype >= 2) {
                u32 single = 0;
                if (ltype == 3 && !c.hufValid) c.err = ZL_E_dictionary_corrupted;      // zstd.c:43160
                else if (csize < 5) c.err = ZL_E_corruption_detected;
                else {
                    u32 lhc = zl_rd32(ip);
                    if (fmt < 2) { single = !fmt; lhSize = 3; litSize = (lhc >> 4) & 0x3FF; litCSize = (lhc >> 14) & 0x3FF; }
                    else if (fmt == 2) { lhSize = 4; litSize = (lhc >> 4) & 0x3FFF; litCSize = lhc >> 18; }
                    else { lhSize = 5; litSize = (lhc >> 4) & 0x3FFFF; litCSize = (lhc >> 22) + ((u32)ip[4] << 10); }
                    if (litSize > c.blockSizeMax) c.err = ZL_E_corruption_detected;
                    else if (!single && litSize < 6) c.err = ZL_E_literals_headerWrong;
                    else if (litCSize + lhSize > csize) c.err = ZL_E_corruption_detected;
                    else if (litCap < litSize) c.err = ZL_E_dstSize_tooSmall;
                    else if (litSize == 0 || litCSize == 0) c.err = ZL_E_corruption_detected;
                    else if (c.litUsed + litSize > d.litCap) c.err = ZL_E_GENERIC;
                    else {
                        u32 hs = c.pos + lhSize, hsz = litCSize;
                        if (ltype == 2) {
                            u32 th = zl_huf_read_stats(f, src + hs, hsz, wbase, bias, hs);
                            if (!th || th >= hsz) c.err = ZL_E_corruption_detected;
                            else { hs += th; hsz -= th; c.needHufFill = 1; c.hufValid = 1; }
                        }
                        if (!c.err) {
                            c.litMode = 2; c.litSize = litSize; c.litOff = c.litUsed; c.litUsed += litSize;
                            if (single) {
                                c.nStreams = 1; c.sBeg[0] = hs; c.sEnd[0] = hs + hsz; c.sOut[0] = c.litOff; c.sLen[0] = litSize;
                            } else if (hsz < 10) c.err = ZL_E_corruption_detected;          // zstd.c:38659
                            else {
                                u32 l1 = zl_rd16(src + hs), l2 = zl_rd16(src + hs + 2), l3 = zl_rd16(src + hs + 4);
                                u32 seg = (litSize + 3) / 4;
                                if (6 + l1 + l2 + l3 > hsz || seg * 3 > litSize) c.err = ZL_E_corruption_detected;
                                else {
                                    c.nStreams = 4;
                                    c.sBeg[0] = hs + 6; c.sEnd[0] = c.sBeg[0] + l1;
                                    c.sBeg[1] = c.sEnd[0]; c.sEnd[1] = c.sBeg[1] + l2;
                                    c.sBeg[2] = c.sEnd[1]; c.sEnd[2] = c.sBeg[2] + l3;
                                    c.sBeg[3] = c.sEnd[2]; c.sEnd[3] = hs + hsz;
                                    for (u32 k = 0; k < 4; k++) { c.sOut[k] = c.litOff + k * seg; c.sLen[k] = seg; }
                                    c.sLen[3] = litSize - 3 * seg;
                                }
                            }
                            c.pos += lhSize + litCSize;
                        }
                    }
                }
            } else {
                bool ok = true;
                if (fmt == 0 || fmt == 2) { lhSize = 1; litSize = ip[0] >> 3; }
                else if (fmt == 1) { lhSize = 2; litSize = zl_rd16(ip) >> 4; }
                else { lhSize = 3; if (csize < 3) { ok = false; litSize = 0; } else litSize = zl_rd24(ip) >> 4; }
                if (!ok) c.err = ZL_E_corruption_detected;
                else if (litSize > c.blockSizeMax) c.err = ZL_E_corruption_detected;
                else if (litCap < litSize) c.err = ZL_E_dstSize_tooSmall;
                else if (ltype == 0) {
                    if (lhSize + litSize > csize) c.err = ZL_E_corruption_detected;
                    else { c.litMode = 0; c.litSize = litSize; c.litSrcOff = c.pos + lhSize; c.pos += lhSize + litSize; }
                } else {
                    if (lhSize + 1 > csize) c.err = ZL_E_corruption_detected;
                    else { c.litMode = 1; c.litSize = litSize; c.rleByte = ip[lhSize]; c.pos += lhSize + 1; }
                }
            }
        }
    }
    if (c.err) { c.isCompressed = 0; zl_ent_finish_frame(f, d, info); return; }
    if (!c.isCompressed) {
        hdrs[c.blockIdx] = h;
        c.outPos += h.regenSize; c.blockIdx++;
        if (last) zl_ent_finish_frame(f, d, info);
    }
}

// ---- sequences section header (lane 0): zstd.c:43707-43790 ------------------------------------------------
ZL_HD void zl_ent_seq_head(ZlFrameSm& f, const ZlFrameDesc& d, const ZlConstTables& ct)
{
    ZlEntCtl& c = f.ctl;
    c.needBuild = 0; c.nbSeq = 0;
    for (u32 k = 0; k < c.nStreams; k++) if (c.sErr[k]) c.err = ZL_E_corruption_detected;
    if (c.err) return;
    const u8* src = d.src;
    u32 ip = c.pos, iend = c.blockEnd;
    if (ip >= iend) { c.err = ZL_E_srcSize_wrong; return; }
    u32 nbSeq = src[ip++];
    if (nbSeq > 0x7F) {
        if (nbSeq == 0xFF) { if (ip + 2 > iend) { c.err = ZL_E_srcSize_wrong; return; } nbSeq = zl_rd16(src + ip) + 0x7F00; ip += 2; }
        else { if (ip >= iend) { c.err = ZL_E_srcSize_wrong; return; } nbSeq = ((nbSeq - 0x80) << 8) + src[ip++]; }
    }
    c.nbSeq = nbSeq;
    if (nbSeq == 0) { if (ip != iend) c.err = ZL_E_corruption_detected; c.pos = ip; return; }
    if (ip + 1 > iend) { c.err = ZL_E_srcSize_wrong; return; }
    u32 modes = src[ip++];
    if (modes & 3) { c.err = ZL_E_corruption_detected; return; }
    const u32 maxSymT[3] = {35, 31, 52}, maxLogT[3] = {9, 8, 9}, defLogT[3] = {6, 5, 6};
    for (u32 t = 0; t < 3; t++) {
        u32 mode = (modes >> (6 - 2 * t)) & 3;
        u32* tbl = t == 0 ? f.fseLL : (t == 1 ? f.fseOF : f.fseML);
        if (mode == 0) {                               // predefined
            const i16* def = t == 0 ? ct.llDef : (t == 1 ? ct.ofDef : ct.mlDef);
            u32 nsym = t == 0 ? 36 : (t == 1 ? 29 : 53);
            for (u32 s = 0; s < nsym; s++) f.norm[t][s] = def[s];
            c.maxSym[t] = nsym - 1; c.tlog[t] = defLogT[t]; c.needBuild |= 1u << t;
        } else if (mode == 1) {                        // RLE
            if (ip >= iend) { c.err = ZL_E_corruption_detected; return; }
            u32 s = src[ip++];
            if (s > maxSymT[t]) { c.err = ZL_E_corruption_detected; return; }
            u32 add = t == 0 ? ct.llBits[s] : (t == 1 ? s : ct.mlBits[s]);
            tbl[0] = zl_fse_pack(0, 0, add, s); c.tlog[t] = 0;
        } else if (mode == 2) {                        // FSE-compressed
            u32 ms = maxSymT[t], tl;
            u32 h = zl_read_ncount(src + ip, iend - ip, f.norm[t], &ms, &tl);
            if (!h || tl > maxLogT[t]) { c.err = ZL_E_corruption_detected; return; }
            ip += h; c.maxSym[t] = ms; c.tlog[t] = tl; c.needBuild |= 1u << t;
        } else {                                       // repeat
            if (!((c.fseValid >> t) & 1)) { c.err = ZL_E_corruption_detected; return; }
        }
        c.fseValid |= 1u << t;
    }
    c.bitBeg = ip; c.pos = ip;
}

ZL_HD void zl_ent_fse_build(ZlFrameSm& f, u32 t, const ZlConstTables& ct)
{
    ZlEntCtl& c = f.ctl;
    if (!((c.needBuild >> t) & 1)) return;
    u32* tbl = t == 0 ? f.fseLL : (t == 1 ? f.fseOF : f.fseML);
    if (!zl_fse_build(tbl, f.norm[t], c.maxSym[t], c.tlog[t], t, ct)) c.sErr[t] = 1; else c.sErr[t] = 0;
}

// ---- serial sequence decode (lane 0): zstd.c:44241-44358 + 44627-44700 -------------------------------------
ZL_HD void zl_ent_seq_decode(ZlFrameSm& f, const ZlFrameDesc& d, ZlFrameInfo& info, ZlBlockHdr* hdrs, u64* recs, u64* cks,
                             const u32* wbase, u32 bias, const ZlConstTables& ct)
{
    ZlEntCtl& c = f.ctl;
    if (!c.err && c.needBuild) { for (u32 t = 0; t < 3; t++) if (((c.needBuild >> t) & 1) && c.sErr[t]) c.err = ZL_E_corruption_detected; }
    u32 nrec = 0, outPos = 0, litPos = 0;
    const u32 litSize = c.litSize;
    u32 outCap = d.dstCap - c.outPos;
    bool capIsDst = true;
    if (outCap > ZL_BLOCKSIZE_MAX) { outCap = ZL_BLOCKSIZE_MAX; capIsDst = false; }
    u64* rp = recs + c.recUsed;
    u64* cp = cks + c.ckUsed;
    const u32 recCap = d.recCap - c.recUsed, ckCap = d.ckCap - c.ckUsed;
    if (!c.err && c.nbSeq) {
        ZlBitR b;
        if (!zl_br_init(b, wbase, bias, c.bitBeg, c.blockEnd)) c.err = ZL_E_corruption_detected;
        else {
            const u32 logLL = c.tlog[0], logOF = c.tlog[1], logML = c.tlog[2];
            zl_br_refill(b, wbase);
            u32 sLL = zl_br_take(b, logLL);
            u32 sOF = zl_br_take(b, logOF);
            zl_br_refill(b, wbase);
            u32 sML = zl_br_take(b, logML);
            u32 rep0 = c.rep[0], rep1 = c.rep[1], rep2 = c.rep[2];
            const u32 hist = c.outPos + c.dictSize;          // history before this block
            u32 err = 0;
            const u32 nbSeq = c.nbSeq;
            for (u32 i = 0; i < nbSeq; i++) {
                zl_br_refill(b, wbase);
                const u32 eLL = f.fseLL[sLL], eOF = f.fseOF[sOF], eML = f.fseML[sML];
                const u32 ofCode = ZL_FSE_SYM(eOF), llCode = ZL_FSE_SYM(eLL), mlCode = ZL_FSE_SYM(eML);
                const u32 ofx = zl_br_take(b, ZL_FSE_ADD(eOF));
                zl_br_refill(b, wbase);
                const u32 mlx = zl_br_take(b, ZL_FSE_ADD(eML));
                const u32 llx = zl_br_take(b, ZL_FSE_ADD(eLL));
                zl_br_refill(b, wbase);
                if (i + 1 < nbSeq) {                        // LL, ML, OF order: zstd.c:44347-44353
                    sLL = ZL_FSE_BASE(eLL) + zl_br_take(b, ZL_FSE_NB(eLL));
                    sML = ZL_FSE_BASE(eML) + zl_br_take(b, ZL_FSE_NB(eML));
                    sOF = ZL_FSE_BASE(eOF) + zl_br_take(b, ZL_FSE_NB(eOF));
                }
                u32 ll = ct.llBase[llCode] + llx;
                u32 ml = ct.mlBase[mlCode] + mlx;
                u32 offset;
                if (ofCode > 1) {
                    offset = (1u << ofCode) - 3 + ofx;
                    rep2 = rep1; rep1 = rep0; rep0 = offset;
                } else {
                    const u32 ll0 = (llCode == 0);
                    if (ofCode == 0) {
                        if (ll0) { offset = rep1; rep1 = rep0; rep0 = offset; } else offset = rep0;
                    } else {
                        const u32 idx = 1 + ll0 + ofx;
                        u32 t = idx == 3 ? rep0 - 1 : (idx == 1 ? rep1 : rep2);
                        if (t == 0) { err = ZL_E_corruption_detected; t = 1; }      // zstd.c:44320
                        if (idx != 1) rep2 = rep1;
                        rep1 = rep0; rep0 = t; offset = t;
                    }
                }
                // validation that ZSTD_execSequence performs (zstd.c:44024-44066)
                if (ll > litSize - litPos) err = ZL_E_corruption_detected;
                else if ((u64)outPos + ll + ml > outCap) err = capIsDst ? ZL_E_dstSize_tooSmall : ZL_E_corruption_detected;
                else if (offset > hist + outPos + ll) err = ZL_E_corruption_detected;
                if (err) break;
                // emit (split lengths >= 65536 into several records)
                while (ll > 65535) {
                    if (nrec >= recCap || (nrec >> 5) >= ckCap) { err = ZL_E_GENERIC; break; }
                    if ((nrec & 31) == 0) cp[nrec >> 5] = (u64)outPos | ((u64)litPos << 32);
                    rp[nrec++] = zl_pack_rec(65535, 0, 0);
                    outPos += 65535; litPos += 65535; ll -= 65535;
                }
                do {
                    u32 m = ml > 65535 ? 65535 : ml;
                    if (nrec >= recCap || (nrec >> 5) >= ckCap) { err = ZL_E_GENERIC; break; }
                    if ((nrec & 31) == 0) cp[nrec >> 5] = (u64)outPos | ((u64)litPos << 32);
                    rp[nrec++] = zl_pack_rec(ll, m, offset);
                    outPos += ll + m; litPos += ll; ll = 0; ml -= m;
                } while (ml);
                if (err) break;
            }
            if (!err && zl_br_remaining(b, bias, c.bitBeg) != 0) err = ZL_E_corruption_detected;      // zstd.c:44686
            if (err) c.err = err;
            c.rep[0] = rep0; c.rep[1] = rep1; c.rep[2] = rep2;
        }
    }
    if (!c.err) {
        u32 lastLL = litSize - litPos;
        if ((u64)outPos + lastLL > outCap) c.err = capIsDst ? ZL_E_dstSize_tooSmall : ZL_E_corruption_detected;
        else outPos += lastLL;
    }
    if (c.err) { zl_ent_finish_frame(f, d, info); return; }
    ZlBlockHdr h;
    h.flags = 2u | (c.litMode << 4) | (c.litMode == 1 ? (c.rleByte << 8) : 0u);
    h.regenSize = outPos;
    h.srcOff = c.litSrcOff; h.litOff = c.litOff; h.litSize = litSize;
    h.nrec = nrec; h.recOff = c.recUsed; h.ckOff = c.ckUsed;
    hdrs[c.blockIdx] = h;
    c.recUsed += nrec; c.ckUsed += (nrec + 31) >> 5;
    c.outPos += outPos; c.blockIdx++; c.pos = c.blockEnd;
    if (c.last) zl_ent_finish_frame(f, d, info);
}
