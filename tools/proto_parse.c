/*
 * proto_parse.c -- DEVELOPMENT TOOL (not product, not oracle): CPU model of the GPU match finder + parse used to
 * evaluate compression ratio before writing CUDA.  Sequences are entropy-coded by the reference's own
 * ZSTD_compressSequences (oracle/_ref/libzstd_ref.so) so that only the parse quality differs.
 *
 *   gcc -O2 tools/proto_parse.c -o /tmp/proto_parse -Loracle/_ref -l:libzstd_ref.so -Wl,-rpath,$PWD/oracle/_ref
 *   /tmp/proto_parse file frameSize
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { unsigned offset, litLength, matchLength, rep; } ZSTD_Sequence;
typedef struct ZSTD_CCtx_s ZSTD_CCtx;
ZSTD_CCtx* ZSTD_createCCtx(void);
size_t ZSTD_CCtx_setParameter(ZSTD_CCtx*, int, int);
size_t ZSTD_compress2(ZSTD_CCtx*, void*, size_t, const void*, size_t);
size_t ZSTD_compressSequences(ZSTD_CCtx*, void*, size_t, const ZSTD_Sequence*, size_t, const void*, size_t);
unsigned ZSTD_isError(size_t);
const char* ZSTD_getErrorName(size_t);

typedef uint8_t u8; typedef uint32_t u32; typedef uint64_t u64;
static u64 rd64(const u8* p) { u64 v; memcpy(&v, p, 8); return v; }
static u32 rd32(const u8* p) { u32 v; memcpy(&v, p, 4); return v; }
static const u64 prime4 = 2654435761U, prime5 = 889523592379ULL, prime6 = 227718039650203ULL, prime7 = 58295818150454627ULL, prime8 = 0xCF1BBCDCB7A56463ULL;
static u32 hashN(const u8* p, int mls, int hlog)
{
    u64 v = rd64(p);
    switch (mls) {
    case 4: return (rd32(p) * (u32)prime4) >> (32 - hlog);
    case 5: return (u32)(((v << 24) * prime5) >> (64 - hlog));
    case 6: return (u32)(((v << 16) * prime6) >> (64 - hlog));
    case 7: return (u32)(((v << 8) * prime7) >> (64 - hlog));
    default: return (u32)((v * prime8) >> (64 - hlog));
    }
}
static u32 mlen(const u8* a, const u8* b, const u8* end)
{
    u32 n = 0;
    while (a + n < end && a[n] == b[n]) n++;
    return n;
}

typedef struct { int mls, hlog, mlsLong, hlogLong, repCheck, lazy1, minLen, tileReset; } opts_t;

/* candidates for every position: nearest previous position with the same hash (insert-all, exact) */
static size_t parse(const u8* src, u32 n, const opts_t* o, ZSTD_Sequence* seqs)
{
    u32* cand = malloc(4 * (size_t)n + 64); u32* candL = malloc(4 * (size_t)n + 64);
    u32* tab = calloc((size_t)1 << o->hlog, 4); u32* tabL = o->mlsLong ? calloc((size_t)1 << o->hlogLong, 4) : NULL;
    const u8* end = src + n;
    for (u32 p = 0; p < n; p++) { cand[p] = 0xFFFFFFFFu; candL[p] = 0xFFFFFFFFu; }
    for (u32 p = 0; p + 8 <= n; p++) {
        u32 h = hashN(src + p, o->mls, o->hlog);
        cand[p] = tab[h] ? tab[h] - 1 : 0xFFFFFFFFu; tab[h] = p + 1;
        if (tabL) { u32 hl = hashN(src + p, o->mlsLong, o->hlogLong); candL[p] = tabL[hl] ? tabL[hl] - 1 : 0xFFFFFFFFu; tabL[hl] = p + 1; }
    }
    size_t ns = 0; u32 p = 0, anchor = 0, rep0 = 0, rep1 = 0;
    while (p + 8 <= n) {
        u32 bestLen = 0, bestOff = 0, start = p;
        /* explicit repcode check at p+1 (dfast style, zstd.c:29989) */
        if (o->repCheck && rep0 && p + 1 >= rep0 && p + 1 + 4 <= n && rd32(src + p + 1) == rd32(src + p + 1 - rep0)) {
            bestLen = mlen(src + p + 1, src + p + 1 - rep0, end); bestOff = rep0; start = p + 1;
        } else {
            if (candL[p] != 0xFFFFFFFFu) { u32 l = mlen(src + p, src + candL[p], end); if (l >= 8 || (o->mlsLong && l >= (u32)o->mlsLong)) { bestLen = l; bestOff = p - candL[p]; } }
            if (cand[p] != 0xFFFFFFFFu) { u32 l = mlen(src + p, src + cand[p], end); if (l >= (u32)o->minLen && l > bestLen) { bestLen = l; bestOff = p - cand[p]; } }
            if (bestLen && o->lazy1 && p + 1 + 8 <= n && candL[p + 1] != 0xFFFFFFFFu && bestLen < 8) {       /* zstd.c:30047: long match at p+1 beats a short one at p */
                u32 l = mlen(src + p + 1, src + candL[p + 1], end);
                if (l >= 8 && l > bestLen) { bestLen = l; bestOff = p + 1 - candL[p + 1]; start = p + 1; }
            }
        }
        if (!bestLen) { p++; continue; }
        /* backward extension (zstd.c:30947) */
        while (start > anchor && start - bestOff > 0 && src[start - 1] == src[start - 1 - bestOff]) { start--; bestLen++; }
        seqs[ns].offset = bestOff; seqs[ns].litLength = start - anchor; seqs[ns].matchLength = bestLen; seqs[ns].rep = 0; ns++;
        rep1 = rep0; rep0 = bestOff;
        p = start + bestLen; anchor = p;
        /* immediate rep1 matches with zero literals (zstd.c:30964-30980) */
        if (o->repCheck) while (p + 4 <= n && rep1 && p >= rep1 && rd32(src + p) == rd32(src + p - rep1)) {
            u32 l = mlen(src + p, src + p - rep1, end);
            seqs[ns].offset = rep1; seqs[ns].litLength = 0; seqs[ns].matchLength = l; seqs[ns].rep = 0; ns++;
            u32 t = rep1; rep1 = rep0; rep0 = t; p += l; anchor = p;
        }
    }
    (void)rep1;
    free(cand); free(candL); free(tab); free(tabL);
    return ns;
}

int main(int argc, char** argv)
{
    if (argc < 3) return 1;
    FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); size_t size = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
    u8* data = malloc(size + 64); if (fread(data, 1, size, f) != size) return 2; fclose(f);
    size_t fs = (size_t)atol(argv[2]);
    ZSTD_Sequence* seqs = malloc(sizeof(ZSTD_Sequence) * (fs / 3 + 16));
    u8* dst = malloc(fs * 2 + 1024);
    opts_t variants[] = {
        /* mls hlog mlsL hlogL rep lazy minLen */
        {5, 13, 0, 0, 0, 0, 5, 0},      /* L1-like, no rep */
        {5, 13, 0, 0, 1, 0, 5, 0},      /* L1-like, rep */
        {6, 13, 0, 0, 1, 0, 6, 0},      /* reference L1 minMatch 6 @128K */
        {5, 14, 0, 0, 1, 0, 5, 0},
        {5, 15, 8, 16, 0, 0, 5, 0},     /* L3-like dfast, no rep */
        {5, 15, 8, 16, 1, 0, 5, 0},
        {5, 15, 8, 16, 1, 1, 5, 0},
        {4, 14, 8, 15, 1, 1, 4, 0},
        {5, 13, 8, 14, 1, 1, 5, 0},     /* smaller tables */
    };
    int nv = sizeof(variants) / sizeof(variants[0]);
    double tot[16] = {0}, ref1 = 0, ref3 = 0, raw = 0; double nseq[16] = {0};
    ZSTD_CCtx* c1 = ZSTD_createCCtx(); ZSTD_CCtx_setParameter(c1, 100, 1);
    ZSTD_CCtx* c3 = ZSTD_createCCtx(); ZSTD_CCtx_setParameter(c3, 100, 3);
    ZSTD_CCtx* cs = ZSTD_createCCtx(); ZSTD_CCtx_setParameter(cs, 100, 3);
    for (size_t off = 0; off + fs <= size; off += fs) {
        raw += (double)fs;
        ref1 += (double)ZSTD_compress2(c1, dst, fs * 2, data + off, fs);
        ref3 += (double)ZSTD_compress2(c3, dst, fs * 2, data + off, fs);
        for (int v = 0; v < nv; v++) {
            size_t ns = parse(data + off, (u32)fs, &variants[v], seqs);
            size_t r = ZSTD_compressSequences(cs, dst, fs * 2, seqs, ns, data + off, fs);
            if (ZSTD_isError(r)) { printf("variant %d: %s\n", v, ZSTD_getErrorName(r)); r = fs; }
            tot[v] += (double)r; nseq[v] += (double)ns;
        }
    }
    printf("%s frame %zu: ref L1 ratio %.3f  L3 ratio %.3f\n", argv[1], fs, raw / ref1, raw / ref3);
    for (int v = 0; v < nv; v++)
        printf("  v%d mls%d hl%d long%d/%d rep%d lazy%d: ratio %.3f (vs L1 %+.1f%%, vs L3 %+.1f%%) seq/frame %.0f\n", v, variants[v].mls, variants[v].hlog,
               variants[v].mlsLong, variants[v].hlogLong, variants[v].repCheck, variants[v].lazy1, raw / tot[v],
               100.0 * (ref1 / tot[v] - 1.0), 100.0 * (ref3 / tot[v] - 1.0), nseq[v] / (raw / (double)fs));
    return 0;
}
