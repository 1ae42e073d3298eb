/*
 * zl_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT part of the product path.
 *
 * A plain-C, single-threaded restatement of the Zstandard *decoder* that the
 * reference (coolbutuseless/zstdlite) reaches through its vendored libzstd 1.5.6
 * (`/root/reference/src/zstd/zstd.c`, called `zstd.c` below).  It is written in
 * the simplest possible form (bit positions instead of a refilled container, one
 * table layout, no SIMD/cache tricks) so that it can serve as an independent
 * checker for the CUDA decoder.  Only `tests/`, `__graft_entry__.smoke()` and
 * `bench.py`'s cpu_baseline leg may load this file's library.
 *
 * Pinning: tests/test_oracle.py checks this restatement against
 *   - the reference's own known-answer vector man/figures/data.json.zst
 *     (expected bytes in README.md:199-205, committed under tests/golden/),
 *   - frames produced by oracle/_ref/libzstd_ref.so (the reference's libzstd
 *     compiled from its own sources) at levels 1..19, with/without checksum,
 *     with/without dictionary, byte-for-byte against ZSTD_decompressDCtx.
 *
 * Each function cites the zstd.c lines it restates.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

typedef uint8_t  u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t  i64;

/* error codes: enum ZSTD_ErrorCode, zstd.c:1463-1499 */
enum {
    ZLO_OK = 0, ZLO_GENERIC = 1, ZLO_prefix_unknown = 10, ZLO_frameParameter_unsupported = 14,
    ZLO_frameParameter_windowTooLarge = 16, ZLO_corruption_detected = 20, ZLO_checksum_wrong = 22,
    ZLO_literals_headerWrong = 24, ZLO_dictionary_corrupted = 30, ZLO_dictionary_wrong = 32,
    ZLO_tableLog_tooLarge = 44, ZLO_maxSymbolValue_tooSmall = 48, ZLO_dstSize_tooSmall = 70,
    ZLO_srcSize_wrong = 72
};
#define ERR(c) ((size_t)0 - (size_t)(c))
#define ISERR(r) ((r) > ERR(120))
#define CONTENTSIZE_UNKNOWN ((u64)0 - 1)
#define CONTENTSIZE_ERROR   ((u64)0 - 2)

static u32 rd16(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8); }
static u32 rd24(const u8* p) { return rd16(p) | ((u32)p[2] << 16); }
static u32 rd32(const u8* p) { return rd24(p) | ((u32)p[3] << 24); }
static u64 rd64(const u8* p) { return (u64)rd32(p) | ((u64)rd32(p + 4) << 32); }
static int highbit(u32 v) { int r = -1; while (v) { v >>= 1; r++; } return r; }

/* ------------------------------------------------------------------ XXH64
 * zstd.c:11509-11513 (primes), 11549-11664 (round/merge/avalanche/finalize). */
#define P1 0x9E3779B185EBCA87ULL
#define P2 0xC2B2AE3D27D4EB4FULL
#define P3 0x165667B19E3779F9ULL
#define P4 0x85EBCA77C2B2AE63ULL
#define P5 0x27D4EB2F165667C5ULL
static u64 rotl(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
static u64 xround(u64 acc, u64 in) { return rotl(acc + in * P2, 31) * P1; }
static u64 xmerge(u64 h, u64 v) { return (h ^ xround(0, v)) * P1 + P4; }
u64 zlo_xxh64(const void* data, size_t len, u64 seed)
{
    const u8* p = (const u8*)data; const u8* end = p + len; u64 h;
    if (len >= 32) {
        u64 v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        do { v1 = xround(v1, rd64(p)); v2 = xround(v2, rd64(p + 8));
             v3 = xround(v3, rd64(p + 16)); v4 = xround(v4, rd64(p + 24)); p += 32; } while (p + 32 <= end);
        h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
        h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
    } else h = seed + P5;
    h += (u64)len;
    while (p + 8 <= end) { h ^= xround(0, rd64(p)); h = rotl(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= end) { h ^= (u64)rd32(p) * P1; h = rotl(h, 23) * P2 + P3; p += 4; }
    while (p < end) { h ^= (*p++) * P5; h = rotl(h, 11) * P1; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

/* ------------------------------------------------------------------ frame header
 * zstd.c:41019-41152 (ZSTD_frameHeaderSize_internal, ZSTD_getFrameHeader_advanced). */
typedef struct {
    u64 contentSize;      /* CONTENTSIZE_UNKNOWN if absent */
    u64 windowSize;
    u32 blockSizeMax;
    u32 dictID;
    u32 checksumFlag;
    u32 headerSize;
    u32 skippable;        /* 1: skippable frame, contentSize = payload size */
} zlo_frame_header;

/* returns 0, an error, or (when srcSize is too small) the number of bytes wanted */
size_t zlo_get_frame_header(zlo_frame_header* h, const void* src, size_t srcSize)
{
    const u8* ip = (const u8*)src;
    memset(h, 0, sizeof(*h));
    if (srcSize < 5) return 5;                                   /* zstd.c:41056-41077 minimum input */
    {   u32 magic = (srcSize >= 4) ? rd32(ip) : 0;
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {              /* skippable, zstd.c:41079-41090 */
            if (srcSize < 8) return 8;
            h->skippable = 1; h->contentSize = rd32(ip + 4); h->headerSize = 8;
            return 0;
        }
        if (magic != 0xFD2FB528u) return ERR(ZLO_prefix_unknown);
    }
    {   u8 fhd = ip[4];
        u32 dictIDSizeCode = fhd & 3, checksum = (fhd >> 2) & 1, single = (fhd >> 5) & 1, fcsID = fhd >> 6;
        static const u32 did_sz[4] = {0, 1, 2, 4}, fcs_sz[4] = {0, 2, 4, 8};
        u32 hs = 5 + !single + did_sz[dictIDSizeCode] + fcs_sz[fcsID] + (single && !fcsID);   /* zstd.c:41019-41030 */
        size_t pos = 5; u64 windowSize = 0, fcs = CONTENTSIZE_UNKNOWN; u32 dictID = 0;
        if (srcSize < hs) return hs;
        h->headerSize = hs;
        if (fhd & 0x08) return ERR(ZLO_frameParameter_unsupported);      /* reserved bit, zstd.c:41106 */
        if (!single) {
            u8 wl = ip[pos++]; u32 windowLog = (wl >> 3) + 10;           /* zstd.c:41115-41121 */
            if (windowLog > 31) return ERR(ZLO_frameParameter_windowTooLarge);
            windowSize = 1ULL << windowLog; windowSize += (windowSize >> 3) * (wl & 7);
        }
        switch (dictIDSizeCode) { case 1: dictID = ip[pos]; pos += 1; break; case 2: dictID = rd16(ip + pos); pos += 2; break;
                                  case 3: dictID = rd32(ip + pos); pos += 4; break; default: break; }
        switch (fcsID) { case 0: if (single) fcs = ip[pos]; break; case 1: fcs = rd16(ip + pos) + 256; break;
                         case 2: fcs = rd32(ip + pos); break; default: fcs = rd64(ip + pos); break; }   /* zstd.c:41137-41140 */
        if (single) windowSize = fcs;
        h->contentSize = fcs; h->windowSize = windowSize;
        h->blockSizeMax = (u32)(windowSize < (1u << 17) ? windowSize : (1u << 17));             /* zstd.c:41147 */
        h->dictID = dictID; h->checksumFlag = checksum;
    }
    return 0;
}

/* zstd.c:41170 ZSTD_getFrameContentSize */
u64 zlo_get_frame_content_size(const void* src, size_t srcSize)
{
    zlo_frame_header h;
    if (zlo_get_frame_header(&h, src, srcSize) != 0) return CONTENTSIZE_ERROR;
    if (h.skippable) return 0;
    return h.contentSize;
}

/* zstd.c:41335-41410 ZSTD_findFrameSizeInfo / ZSTD_findFrameCompressedSize */
size_t zlo_find_frame_compressed_size(const void* src, size_t srcSize)
{
    const u8* ip = (const u8*)src; zlo_frame_header h; size_t r, remaining;
    if (srcSize >= 8 && (rd32(ip) & 0xFFFFFFF0u) == 0x184D2A50u) {       /* zstd.c:41188 readSkippableFrameSize */
        u64 sz = (u64)rd32(ip + 4) + 8;
        if (sz > srcSize) return ERR(ZLO_srcSize_wrong);
        return (size_t)sz;
    }
    r = zlo_get_frame_header(&h, src, srcSize);
    if (ISERR(r)) return r;
    if (r > 0) return ERR(ZLO_srcSize_wrong);
    ip += h.headerSize; remaining = srcSize - h.headerSize;
    for (;;) {                                                            /* zstd.c:41370-41384 */
        u32 bh, type, last, csize;
        if (remaining < 3) return ERR(ZLO_srcSize_wrong);
        bh = rd24(ip); last = bh & 1; type = (bh >> 1) & 3; csize = bh >> 3;
        if (type == 3) return ERR(ZLO_corruption_detected);
        if (type == 1) csize = 1;
        if (3 + (size_t)csize > remaining) return ERR(ZLO_srcSize_wrong);
        ip += 3 + csize; remaining -= 3 + csize;
        if (last) break;
    }
    if (h.checksumFlag) { if (remaining < 4) return ERR(ZLO_srcSize_wrong); ip += 4; }
    return (size_t)(ip - (const u8*)src);
}

/* ------------------------------------------------------------------ backward bit reader
 * Restates BIT_DStream_t (zstd.c:2352-2550) with an explicit bit position: `pos` is the
 * number of not-yet-consumed bits; bits that would lie below the start read as 0 and
 * make pos negative ("overflow" in the reference's vocabulary). */
typedef struct { const u8* buf; i64 pos; } bitr;
static int br_init(bitr* b, const u8* buf, size_t size)
{
    if (size == 0 || buf[size - 1] == 0) return -1;               /* zstd.c:2354, 2369 */
    b->buf = buf; b->pos = (i64)(size - 1) * 8 + highbit(buf[size - 1]);
    return 0;
}
static u32 br_bits_at(const bitr* b, i64 bitpos, int n)           /* n <= 32 */
{
    u64 v = 0; int i;
    i64 byte = bitpos >> 3;                                          /* floor, bitpos may be negative */
    int sh = (int)(bitpos & 7);
    for (i = 0; i < 6; i++) { i64 k = byte + i; if (k >= 0) v |= (u64)b->buf[k] << (8 * i); }
    v >>= sh;
    return n ? (u32)(v & ((1ULL << n) - 1)) : 0;
}
/* The reader never dereferences past the sentinel byte: callers only ask for bits below
 * `pos`, and br_bits_at() touches at most 6 bytes starting at bitpos/8 -- we therefore keep
 * a private zero-padded copy of each stream (see br_copy). */
static u32 br_read(bitr* b, int n) { b->pos -= n; return br_bits_at(b, b->pos, n); }
static u32 br_peek(const bitr* b, int n) { return br_bits_at(b, b->pos - n, n); }

typedef struct { u8* mem; bitr b; } brc;
static int brc_open(brc* c, const u8* src, size_t size)
{
    c->mem = (u8*)calloc(size + 16, 1);
    if (!c->mem) return -1;
    memcpy(c->mem, src, size);
    if (br_init(&c->b, c->mem, size)) { free(c->mem); c->mem = NULL; return -1; }
    return 0;
}
static void brc_close(brc* c) { free(c->mem); c->mem = NULL; }

/* ------------------------------------------------------------------ FSE
 * NCount reader: zstd.c:3269-3409.  Forward LSB-first bit position over the header. */
static size_t read_ncount(short* norm, u32* maxSymPtr, u32* tableLogPtr, const u8* src, size_t srcSize)
{
    u64 bitpos = 0; u32 maxSV1 = *maxSymPtr + 1, sym = 0; int nbBits, remaining, threshold, prev0 = 0;
    if (srcSize < 1) return ERR(ZLO_srcSize_wrong);
    memset(norm, 0, maxSV1 * sizeof(short));
    {   u32 v = 0; size_t i; for (i = 0; i < 4 && i < srcSize; i++) v |= (u32)src[i] << (8 * i);
        nbBits = (int)(v & 15) + 5; }
    if (nbBits > 15) return ERR(ZLO_tableLog_tooLarge);
    *tableLogPtr = (u32)nbBits; bitpos = 4;
    remaining = (1 << nbBits) + 1; threshold = 1 << nbBits; nbBits++;
    for (;;) {
        u32 w; size_t i; u64 by;
        if (prev0) {                                                /* zero-run repeat flags, zstd.c:3311-3355 */
            for (;;) {
                by = bitpos >> 3; w = 0; for (i = 0; i < 4; i++) if (by + i < srcSize) w |= (u32)src[by + i] << (8 * i);
                w = (w >> (bitpos & 7)) & 3; bitpos += 2;
                sym += w;
                if (w != 3) break;
            }
            if (sym >= maxSV1) break;
        }
        {   int max = (2 * threshold - 1) - remaining, count; u64 v = 0;
            by = bitpos >> 3; for (i = 0; i < 5; i++) if (by + i < srcSize) v |= (u64)src[by + i] << (8 * i);
            v >>= (bitpos & 7);
            if ((int)(v & (u64)(threshold - 1)) < max) { count = (int)(v & (u64)(threshold - 1)); bitpos += (u64)nbBits - 1; }
            else { count = (int)(v & (u64)(2 * threshold - 1)); if (count >= threshold) count -= max; bitpos += (u64)nbBits; }
            count--;
            if (count >= 0) remaining -= count; else remaining += count;
            norm[sym++] = (short)count; prev0 = !count;
            if (remaining < threshold) {
                if (remaining <= 1) break;
                nbBits = highbit((u32)remaining) + 1; threshold = 1 << (nbBits - 1);
            }
            if (sym >= maxSV1) break;
        }
    }
    if (remaining != 1) return ERR(ZLO_corruption_detected);
    if (sym > maxSV1) return ERR(ZLO_maxSymbolValue_tooSmall);
    *maxSymPtr = sym - 1;
    {   size_t used = (size_t)((bitpos + 7) >> 3);
        if (used > srcSize) return ERR(ZLO_corruption_detected);
        return used; }
}

typedef struct { u16 newState; u8 symbol; u8 nbBits; } fse_cell;

/* zstd.c:3692-3790 FSE_buildDTable_internal and zstd.c:43497-43613 ZSTD_buildFSETable_body
 * (same spread; the latter only differs in what is stored per cell). */
static int fse_build(fse_cell* t, const short* norm, u32 maxSym, u32 tableLog)
{
    u32 size = 1u << tableLog, high = size - 1, step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    u16 next[256]; u32 s, pos = 0, u;
    for (s = 0; s <= maxSym; s++) {
        if (norm[s] == -1) { t[high--].symbol = (u8)s; next[s] = 1; } else next[s] = (u16)norm[s];
    }
    for (s = 0; s <= maxSym; s++) { int i; for (i = 0; i < norm[s]; i++) {
        t[pos].symbol = (u8)s; pos = (pos + step) & mask; while (pos > high) pos = (pos + step) & mask; } }
    if (pos != 0) return -1;
    for (u = 0; u < size; u++) {
        u32 ns = next[t[u].symbol]++;
        t[u].nbBits = (u8)(tableLog - (u32)highbit(ns));
        t[u].newState = (u16)((ns << t[u].nbBits) - size);
    }
    return 0;
}

/* ------------------------------------------------------------------ Huffman
 * zstd.c:3470-3540 HUF_readStats_body; 3875-3920 FSE_decompress_wksp_body; 3790-3870 decode loop;
 * 38436-38570 HUF_readDTableX1_wksp (we keep one {symbol, nbBits} cell per tableLog-bit prefix). */
typedef struct { u8 symbol, nbBits; } huf_cell;
typedef struct { huf_cell cell[1 << 12]; u32 tableLog; int valid; } huf_table;

static size_t huf_read_table(huf_table* ht, const u8* src, size_t srcSize)
{
    u8 w[256]; u32 rank[16] = {0}, nsym = 0, total = 0, n; size_t iSize, hdr;
    if (!srcSize) return ERR(ZLO_srcSize_wrong);
    hdr = src[0];
    if (hdr >= 128) {                                               /* direct 4-bit weights */
        nsym = (u32)hdr - 127; iSize = (nsym + 1) / 2;
        if (iSize + 1 > srcSize) return ERR(ZLO_srcSize_wrong);
        for (n = 0; n < nsym; n++) w[n] = (n & 1) ? (src[1 + n / 2] & 15) : (src[1 + n / 2] >> 4);
    } else {                                                        /* FSE-compressed weights */
        short norm[256]; u32 maxSym = 255, tlog; fse_cell tbl[64]; size_t nc; brc c; u32 s1, s2;
        iSize = hdr;
        if (iSize + 1 > srcSize) return ERR(ZLO_srcSize_wrong);
        nc = read_ncount(norm, &maxSym, &tlog, src + 1, iSize);
        if (ISERR(nc)) return nc;
        if (tlog > 6) return ERR(ZLO_tableLog_tooLarge);
        if (fse_build(tbl, norm, maxSym, tlog)) return ERR(ZLO_GENERIC);
        if (brc_open(&c, src + 1 + nc, iSize - nc)) return ERR(ZLO_corruption_detected);
        s1 = br_read(&c.b, (int)tlog); s2 = br_read(&c.b, (int)tlog);
        for (;;) {                                                  /* zstd.c:3846-3862 tail semantics */
            if (nsym > 253) { brc_close(&c); return ERR(ZLO_dstSize_tooSmall); }
            w[nsym++] = tbl[s1].symbol; s1 = tbl[s1].newState + br_read(&c.b, tbl[s1].nbBits);
            if (c.b.pos < 0) { w[nsym++] = tbl[s2].symbol; break; }
            if (nsym > 253) { brc_close(&c); return ERR(ZLO_dstSize_tooSmall); }
            w[nsym++] = tbl[s2].symbol; s2 = tbl[s2].newState + br_read(&c.b, tbl[s2].nbBits);
            if (c.b.pos < 0) { w[nsym++] = tbl[s1].symbol; break; }
        }
        brc_close(&c);
    }
    for (n = 0; n < nsym; n++) { if (w[n] > 12) return ERR(ZLO_corruption_detected); rank[w[n]]++; total += (1u << w[n]) >> 1; }
    if (total == 0) return ERR(ZLO_corruption_detected);
    {   u32 tableLog = (u32)highbit(total) + 1, rest, last;
        if (tableLog > 12) return ERR(ZLO_corruption_detected);
        rest = (1u << tableLog) - total; last = (u32)highbit(rest) + 1;
        if ((1u << highbit(rest)) != rest) return ERR(ZLO_corruption_detected);
        w[nsym++] = (u8)last; rank[last]++;
        if (rank[1] < 2 || (rank[1] & 1)) return ERR(ZLO_corruption_detected);
        /* fill: weight 1 symbols first (longest codes), in symbol order inside one weight */
        {   u32 start[16], wv, acc = 0;
            for (wv = 1; wv <= tableLog; wv++) { start[wv] = acc; acc += rank[wv] << (wv - 1); }
            for (n = 0; n < nsym; n++) if (w[n]) {
                u32 len = 1u << (w[n] - 1), k;
                for (k = 0; k < len; k++) { ht->cell[start[w[n]] + k].symbol = (u8)n; ht->cell[start[w[n]] + k].nbBits = (u8)(tableLog + 1 - w[n]); }
                start[w[n]] += len;
            }
        }
        ht->tableLog = tableLog; ht->valid = 1;
    }
    return iSize + 1;
}

/* one stream: zstd.c:38626-38650 (HUF_decompress1X1_usingDTable_internal_body) */
static int huf_stream(u8* dst, size_t n, const u8* src, size_t srcSize, const huf_table* ht)
{
    brc c; size_t i; int ok;
    if (brc_open(&c, src, srcSize)) return -1;
    for (i = 0; i < n; i++) {
        huf_cell e = ht->cell[br_peek(&c.b, (int)ht->tableLog)];
        c.b.pos -= e.nbBits; dst[i] = e.symbol;
    }
    ok = (c.b.pos == 0);                                            /* BIT_endOfDStream, zstd.c:38647 */
    brc_close(&c);
    return ok ? 0 : -1;
}
/* four streams: zstd.c:38653-38770 */
static int huf_4streams(u8* dst, size_t n, const u8* src, size_t srcSize, const huf_table* ht)
{
    size_t l1, l2, l3, l4, seg = (n + 3) / 4;
    if (srcSize < 10) return -1;                                    /* zstd.c:38659 */
    if (n < 6) return -1;                                           /* zstd.c:38661 */
    l1 = rd16(src); l2 = rd16(src + 2); l3 = rd16(src + 4);
    if (6 + l1 + l2 + l3 > srcSize) return -1;
    l4 = srcSize - 6 - l1 - l2 - l3;
    if (seg * 3 > n) return -1;                                     /* opStart4 > oend, zstd.c:38690 */
    if (huf_stream(dst, seg, src + 6, l1, ht)) return -1;
    if (huf_stream(dst + seg, seg, src + 6 + l1, l2, ht)) return -1;
    if (huf_stream(dst + 2 * seg, seg, src + 6 + l1 + l2, l3, ht)) return -1;
    if (huf_stream(dst + 3 * seg, n - 3 * seg, src + 6 + l1 + l2 + l3, l4, ht)) return -1;
    return 0;
}

/* ------------------------------------------------------------------ sequences
 * base/bits tables zstd.c:15470-15495 and 40060-40075; default norms 15497-15515. */
static const u32 LL_base[36] = {0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,18,20,22,24,28,32,40,48,64,0x80,0x100,0x200,0x400,0x800,0x1000,0x2000,0x4000,0x8000,0x10000};
static const u8  LL_bits[36] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16};
static const u32 ML_base[53] = {3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,32,33,34,35,37,39,41,43,47,51,59,67,83,99,0x83,0x103,0x203,0x403,0x803,0x1003,0x2003,0x4003,0x8003,0x10003};
static const u8  ML_bits[53] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,4,5,7,8,9,10,11,12,13,14,15,16};
static const short LL_defaultNorm[36] = {4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1};
static const short ML_defaultNorm[53] = {1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1,-1,-1};
static const short OF_defaultNorm[29] = {1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1};

typedef struct { fse_cell cell[512]; u32 log; int valid; } seq_table;

typedef struct {
    huf_table huf;
    seq_table ll, of, ml;
    u32 rep[3];
    const u8* dict; size_t dictSize;       /* raw content part */
    u32 dictID;
} zlo_state;

/* zstd.c:43659-43705 ZSTD_buildSeqTable */
static size_t build_seq_table(seq_table* t, u32 type, u32 maxSym, u32 maxLog, const u8* src, size_t srcSize,
                              const short* defNorm, u32 defLog)
{
    switch (type) {
    case 1: /* rle */
        if (!srcSize) return ERR(ZLO_srcSize_wrong);
        if (src[0] > maxSym) return ERR(ZLO_corruption_detected);
        t->cell[0].symbol = src[0]; t->cell[0].nbBits = 0; t->cell[0].newState = 0; t->log = 0; t->valid = 1;
        return 1;
    case 0: /* predefined */
        fse_build(t->cell, defNorm, maxSym, defLog); t->log = defLog; t->valid = 1;
        return 0;
    case 3: /* repeat */
        if (!t->valid) return ERR(ZLO_corruption_detected);
        return 0;
    default: {
        short norm[64]; u32 ms = maxSym, tlog; size_t h = read_ncount(norm, &ms, &tlog, src, srcSize);
        if (ISERR(h)) return ERR(ZLO_corruption_detected);
        if (tlog > maxLog) return ERR(ZLO_corruption_detected);
        if (fse_build(t->cell, norm, ms, tlog)) return ERR(ZLO_corruption_detected);
        t->log = tlog; t->valid = 1;
        return h; }
    }
}

/* one compressed block: zstd.c:45084 ZSTD_decompressBlock_internal, 43146 literals, 43707 seq headers,
 * 44241 ZSTD_decodeSequence, 44013 ZSTD_execSequence, 44627 sequence loop.
 * dst..: `out` points at the start of the frame's output, `op` is the running position. */
static size_t decode_block(zlo_state* st, u8* out, size_t op, size_t cap, const u8* src, size_t srcSize, u32 blockSizeMax)
{
    const u8* ip = src; const u8* lit = NULL; u8* litbuf = NULL; size_t litSize = 0, lhSize, litCSize = 0;
    size_t ret = 0, maxOut = cap - op, litCap; u32 type, fmt;
    litCap = maxOut < blockSizeMax ? maxOut : blockSizeMax;          /* expectedWriteSize, zstd.c:43170 */
    if (srcSize > blockSizeMax) return ERR(ZLO_srcSize_wrong);                 /* zstd.c:45099 */
    if (srcSize < 2) return ERR(ZLO_corruption_detected);                       /* MIN_CBLOCK_SIZE */
    type = ip[0] & 3; fmt = (ip[0] >> 2) & 3;
    if (type >= 2) {                                                            /* compressed / repeat(treeless) */
        u32 single = 0; u64 lhc;
        if (type == 3 && !st->huf.valid) return ERR(ZLO_dictionary_corrupted);
        if (srcSize < 5) return ERR(ZLO_corruption_detected);
        lhc = rd32(ip);
        switch (fmt) {
        case 0: case 1: single = !fmt; lhSize = 3; litSize = (lhc >> 4) & 0x3FF; litCSize = (lhc >> 14) & 0x3FF; break;
        case 2: lhSize = 4; litSize = (lhc >> 4) & 0x3FFF; litCSize = lhc >> 18; break;
        default: lhSize = 5; litSize = (lhc >> 4) & 0x3FFFF; litCSize = (lhc >> 22) + ((size_t)ip[4] << 10); break;
        }
        if (litSize > blockSizeMax) return ERR(ZLO_corruption_detected);
        if (!single && litSize < 6) return ERR(ZLO_literals_headerWrong);
        if (litCSize + lhSize > srcSize) return ERR(ZLO_corruption_detected);
        if (litCap < litSize) return ERR(ZLO_dstSize_tooSmall);
        litbuf = (u8*)malloc(litSize + 8);
        {   const u8* hs = ip + lhSize; size_t hsz = litCSize;
            if (type == 2) {
                size_t th = huf_read_table(&st->huf, hs, hsz);
                if (ISERR(th)) { free(litbuf); return ERR(ZLO_corruption_detected); }
                hs += th; hsz -= th;
            }
            if (hsz == 0 && litSize > 0) { /* HUF_decompress*: cSrcSize==0 -> corruption */ }
            if ((single ? huf_stream(litbuf, litSize, hs, hsz, &st->huf) : huf_4streams(litbuf, litSize, hs, hsz, &st->huf))) {
                free(litbuf); return ERR(ZLO_corruption_detected); }
        }
        lit = litbuf; ip += lhSize + litCSize;
    } else {
        switch (fmt) {
        case 0: case 2: lhSize = 1; litSize = ip[0] >> 3; break;
        case 1: lhSize = 2; litSize = rd16(ip) >> 4; break;
        default: if (srcSize < 3) return ERR(ZLO_corruption_detected); lhSize = 3; litSize = rd24(ip) >> 4; break;
        }
        if (litSize > blockSizeMax) return ERR(ZLO_corruption_detected);
        if (litCap < litSize) return ERR(ZLO_dstSize_tooSmall);
        if (type == 0) {
            if (lhSize + litSize > srcSize) return ERR(ZLO_corruption_detected);
            lit = ip + lhSize; ip += lhSize + litSize;
        } else {
            if (fmt == 1 && srcSize < 3) return ERR(ZLO_corruption_detected);
            if (fmt == 3 && srcSize < 4) return ERR(ZLO_corruption_detected);
            litbuf = (u8*)malloc(litSize + 8); memset(litbuf, ip[lhSize], litSize); lit = litbuf; ip += lhSize + 1;
        }
    }
    /* sequences section */
    {   const u8* iend = src + srcSize; size_t nbSeq, litPos = 0, o = op, oend = op + maxOut;
        if (ip >= iend) { ret = ERR(ZLO_srcSize_wrong); goto done; }           /* MIN_SEQUENCES_SIZE 1 */
        nbSeq = *ip++;
        if (nbSeq > 0x7F) {
            if (nbSeq == 0xFF) { if (ip + 2 > iend) { ret = ERR(ZLO_srcSize_wrong); goto done; } nbSeq = rd16(ip) + 0x7F00; ip += 2; }
            else { if (ip >= iend) { ret = ERR(ZLO_srcSize_wrong); goto done; } nbSeq = ((nbSeq - 0x80) << 8) + *ip++; }
        }
        if (nbSeq == 0) {
            if (ip != iend) { ret = ERR(ZLO_corruption_detected); goto done; }
        } else {
            u32 modes, sLL, sOF, sML; brc c; size_t h, i;
            if (ip + 1 > iend) { ret = ERR(ZLO_srcSize_wrong); goto done; }
            modes = *ip++;
            if (modes & 3) { ret = ERR(ZLO_corruption_detected); goto done; }
            h = build_seq_table(&st->ll, modes >> 6, 35, 9, ip, (size_t)(iend - ip), LL_defaultNorm, 6);
            if (ISERR(h)) { ret = ERR(ZLO_corruption_detected); goto done; } ip += h;
            h = build_seq_table(&st->of, (modes >> 4) & 3, 31, 8, ip, (size_t)(iend - ip), OF_defaultNorm, 5);
            if (ISERR(h)) { ret = ERR(ZLO_corruption_detected); goto done; } ip += h;
            h = build_seq_table(&st->ml, (modes >> 2) & 3, 52, 9, ip, (size_t)(iend - ip), ML_defaultNorm, 6);
            if (ISERR(h)) { ret = ERR(ZLO_corruption_detected); goto done; } ip += h;
            if (brc_open(&c, ip, (size_t)(iend - ip))) { ret = ERR(ZLO_corruption_detected); goto done; }
            sLL = br_read(&c.b, (int)st->ll.log); sOF = br_read(&c.b, (int)st->of.log); sML = br_read(&c.b, (int)st->ml.log);
            for (i = 0; i < nbSeq; i++) {
                fse_cell eLL = st->ll.cell[sLL], eOF = st->of.cell[sOF], eML = st->ml.cell[sML];
                u32 ofCode = eOF.symbol, ll, ml; u64 offset;
                /* offset first, then match length, then literal length: zstd.c:44290-44338 */
                if (ofCode > 1) {
                    offset = ((u64)1 << ofCode) - 3 + ((ofCode > 32) ? 0 : (u64)br_read(&c.b, (int)(ofCode > 32 ? 32 : ofCode)));
                    st->rep[2] = st->rep[1]; st->rep[1] = st->rep[0]; st->rep[0] = (u32)offset;
                } else {
                    u32 ll0 = (LL_base[eLL.symbol] == 0);
                    if (ofCode == 0) { offset = st->rep[ll0]; st->rep[1] = st->rep[!ll0]; st->rep[0] = (u32)offset; }
                    else {
                        u32 idx = 1 + ll0 + br_read(&c.b, 1);
                        u32 t = (idx == 3) ? st->rep[0] - 1 : st->rep[idx];
                        if (t == 0) t = 0xFFFFFFFFu;                     /* forces corruption at execution */
                        if (idx != 1) st->rep[2] = st->rep[1];
                        st->rep[1] = st->rep[0]; st->rep[0] = t; offset = t;
                    }
                }
                ml = ML_base[eML.symbol] + br_read(&c.b, ML_bits[eML.symbol]);
                ll = LL_base[eLL.symbol] + br_read(&c.b, LL_bits[eLL.symbol]);
                if (i + 1 < nbSeq) {                                        /* state update order LL, ML, OF: zstd.c:44347-44353 */
                    sLL = eLL.newState + br_read(&c.b, eLL.nbBits);
                    sML = eML.newState + br_read(&c.b, eML.nbBits);
                    sOF = eOF.newState + br_read(&c.b, eOF.nbBits);
                }
                /* execute: zstd.c:44013-44100 */
                if (ll > litSize - litPos) { ret = ERR(ZLO_corruption_detected); brc_close(&c); goto done; }
                if ((u64)ll + ml > oend - o) { ret = ERR(ZLO_dstSize_tooSmall); brc_close(&c); goto done; }
                memcpy(out + o, lit + litPos, ll); o += ll; litPos += ll;
                if (offset > o + st->dictSize) { ret = ERR(ZLO_corruption_detected); brc_close(&c); goto done; }
                {   u32 k; for (k = 0; k < ml; k++) {
                        u64 srcpos = o - offset;        /* may wrap below 0 => dictionary */
                        if (offset > o) out[o] = st->dict[st->dictSize - (size_t)(offset - o)]; else out[o] = out[srcpos];
                        o++; } }
            }
            if (c.b.pos != 0) { ret = ERR(ZLO_corruption_detected); brc_close(&c); goto done; }   /* zstd.c:44686 */
            brc_close(&c);
        }
        if (litSize - litPos > oend - o) { ret = ERR(ZLO_dstSize_tooSmall); goto done; }
        memcpy(out + o, lit + litPos, litSize - litPos); o += litSize - litPos;
        ret = o - op;
    }
done:
    free(litbuf);
    return ret;
}

/* dictionary: zstd.c:42053-42137 ZSTD_loadDEntropy, 42140-42159 ZSTD_decompress_insertDictionary */
static size_t load_dict(zlo_state* st, const u8* dict, size_t dictSize)
{
    st->dict = dict; st->dictSize = dictSize; st->dictID = 0;
    if (dictSize < 8 || rd32(dict) != 0xEC30A437u) return 0;             /* raw content dictionary */
    st->dictID = rd32(dict + 4);
    {   const u8* p = dict + 8; const u8* end = dict + dictSize; size_t h; short norm[64]; u32 ms, tl;
        h = huf_read_table(&st->huf, p, (size_t)(end - p)); if (ISERR(h)) return ERR(ZLO_dictionary_corrupted); p += h;
        ms = 31; h = read_ncount(norm, &ms, &tl, p, (size_t)(end - p)); if (ISERR(h) || tl > 8) return ERR(ZLO_dictionary_corrupted);
        fse_build(st->of.cell, norm, ms, tl); st->of.log = tl; st->of.valid = 1; p += h;
        ms = 52; h = read_ncount(norm, &ms, &tl, p, (size_t)(end - p)); if (ISERR(h) || tl > 9) return ERR(ZLO_dictionary_corrupted);
        fse_build(st->ml.cell, norm, ms, tl); st->ml.log = tl; st->ml.valid = 1; p += h;
        ms = 35; h = read_ncount(norm, &ms, &tl, p, (size_t)(end - p)); if (ISERR(h) || tl > 9) return ERR(ZLO_dictionary_corrupted);
        fse_build(st->ll.cell, norm, ms, tl); st->ll.log = tl; st->ll.valid = 1; p += h;
        if (p + 12 > end) return ERR(ZLO_dictionary_corrupted);
        {   size_t content = (size_t)(end - (p + 12)); int i;
            for (i = 0; i < 3; i++) { u32 r = rd32(p + 4 * i); if (r == 0 || r > content) return ERR(ZLO_dictionary_corrupted); st->rep[i] = r; }
            st->dict = p + 12; st->dictSize = content; }
    }
    return 0;
}

/* zstd.c:41554-41665 ZSTD_decompressFrame (+41671 multi-frame loop).  flags bit0: ignore checksum. */
size_t zlo_decompress(void* dstv, size_t cap, const void* srcv, size_t srcSize, const void* dict, size_t dictSize, int flags)
{
    u8* dst = (u8*)dstv; const u8* src = (const u8*)srcv; size_t total = 0; int first = 1;
    while (srcSize > 0) {
        zlo_frame_header h; size_t r; zlo_state* st; size_t op = 0; u8* out = dst + total; size_t ocap = cap - total;
        const u8* ip; size_t remaining;
        if (srcSize >= 4 && (rd32(src) & 0xFFFFFFF0u) == 0x184D2A50u) {
            size_t sk = zlo_find_frame_compressed_size(src, srcSize);
            if (ISERR(sk)) return sk;
            src += sk; srcSize -= sk; continue;
        }
        r = zlo_get_frame_header(&h, src, srcSize);
        if (ISERR(r)) { if (!first && r == ERR(ZLO_prefix_unknown)) return ERR(ZLO_srcSize_wrong); return r; }
        if (r > 0) return ERR(ZLO_srcSize_wrong);
        st = (zlo_state*)calloc(1, sizeof(*st));
        st->rep[0] = 1; st->rep[1] = 4; st->rep[2] = 8;                    /* zstd.c:15416 */
        if (dict && dictSize) { size_t e = load_dict(st, (const u8*)dict, dictSize); if (ISERR(e)) { free(st); return e; } }
        if (h.dictID && st->dictID != h.dictID) { free(st); return ERR(ZLO_dictionary_wrong); }   /* zstd.c:41318 */
        ip = src + h.headerSize; remaining = srcSize - h.headerSize;
        for (;;) {
            u32 bh, last, type, csize;
            if (remaining < 3) { free(st); return ERR(ZLO_srcSize_wrong); }
            bh = rd24(ip); last = bh & 1; type = (bh >> 1) & 3; csize = bh >> 3; ip += 3; remaining -= 3;
            if (type == 3) { free(st); return ERR(ZLO_corruption_detected); }
            if (type == 1) {                                                  /* RLE, zstd.c:41510 */
                if (remaining < 1) { free(st); return ERR(ZLO_srcSize_wrong); }
                if (csize > ocap - op) { free(st); return ERR(ZLO_dstSize_tooSmall); }
                memset(out + op, ip[0], csize); op += csize; ip += 1; remaining -= 1;
            } else if (type == 0) {                                           /* raw, zstd.c:41497 */
                if (csize > remaining) { free(st); return ERR(ZLO_srcSize_wrong); }
                if (csize > ocap - op) { free(st); return ERR(ZLO_dstSize_tooSmall); }
                memcpy(out + op, ip, csize); op += csize; ip += csize; remaining -= csize;
            } else {
                size_t d;
                if (csize > remaining) { free(st); return ERR(ZLO_srcSize_wrong); }
                d = decode_block(st, out, op, ocap, ip, csize, h.blockSizeMax);
                if (ISERR(d)) { free(st); return d; }
                op += d; ip += csize; remaining -= csize;
            }
            if (last) break;
        }
        free(st);
        if (h.contentSize != CONTENTSIZE_UNKNOWN && h.contentSize != op) return ERR(ZLO_corruption_detected);   /* zstd.c:41646 */
        if (h.checksumFlag) {
            if (remaining < 4) return ERR(ZLO_checksum_wrong);
            if (!(flags & 1) && (u32)zlo_xxh64(out, op, 0) != rd32(ip)) return ERR(ZLO_checksum_wrong);        /* zstd.c:41650-41657 */
            ip += 4; remaining -= 4;
        }
        total += op; src = ip; srcSize = remaining; first = 0;
    }
    return total;
}

/* zstd.c:22583 / macro 4548 ZSTD_compressBound */
size_t zlo_compress_bound(size_t n)
{
    return n + (n >> 8) + (n < (128u << 10) ? (((128u << 10) - n) >> 11) : 0);
}
