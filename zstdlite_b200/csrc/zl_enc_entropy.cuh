// zl_enc_entropy.cuh -- entropy stage of the B200 Zstandard compressor: everything that is per-lane serial logic
// (table construction, headers, the bit writers, the stream encoders).  Like zl_dec_entropy.cuh this is ZL_HD code:
// the kernels in zl_enc_kernels.cu run it one block per QUAD (lane 0 drives, lanes 0..3 take one Huffman stream /
// one FSE state chain each) and tests/emul compiles the same source with g++ to check it against libzstd on the CPU.
//
// Reference behaviour restated (file:line in /root/reference/src/zstd/zstd.c):
//   FSE_optimalTableLog 16205-16222, FSE_normalizeCount 16314-16359 (+ our own fallback where the reference uses
//   FSE_normalizeM2 16228), FSE_writeNCount 16083-16177, FSE_buildCTable_wksp 15917-16035, FSE_initCState2 2772,
//   FSE_encodeSymbol 2783, FSE_flushCState 2792, HUF_buildCTable_wksp 17415 / HUF_setMaxHeight 17035 (we use a
//   two-queue Huffman + Kraft repair instead), HUF_writeCTable_wksp 16907 / HUF_compressWeights 16806,
//   HUF_compress1X/4X 17715/17827, ZSTD_compressLiterals 20692-20798, ZSTD_seqToCodes 25195, ZSTD_LLcode 19533,
//   ZSTD_MLcode 19550, ZSTD_selectEncodingType 21012, ZSTD_buildCTable 21098, ZSTD_encodeSequences_body 21146-21235,
//   ZSTD_entropyCompressSeqStore_internal 25390-25505.
#pragma once
#include "zl_common.cuh"

// ---- constant tables for the encoder (code lookups + the decoder's base/bits tables) ------------------------------
struct ZlEncConst {
    u8 llCode[64];        // litLength < 64 -> code            (zstd.c:19533)
    u8 mlCode[128];       // matchLength-3 < 128 -> code       (zstd.c:19550)
    u8 llBits[36];
    u8 mlBits[53];
    u32 llBase[36];
    u32 mlBase[53];       // in matchLength units (3..)
    i16 llDef[36];
    i16 mlDef[53];
    i16 ofDef[29];
};
// builds the tables from the format's base/bits definitions (host side, once)
static inline void zl_enc_const_init(ZlEncConst* c)
{
    const u32 llBase[36] = ZL_LL_BASE_INIT; const u8 llBits[36] = ZL_LL_BITS_INIT;
    const u32 mlBase[53] = ZL_ML_BASE_INIT; const u8 mlBits[53] = ZL_ML_BITS_INIT;
    const i16 llDef[36] = ZL_LL_DEFNORM_INIT; const i16 mlDef[53] = ZL_ML_DEFNORM_INIT; const i16 ofDef[29] = ZL_OF_DEFNORM_INIT;
    for (int i = 0; i < 36; i++) { c->llBase[i] = llBase[i]; c->llBits[i] = llBits[i]; c->llDef[i] = llDef[i]; }
    for (int i = 0; i < 53; i++) { c->mlBase[i] = mlBase[i]; c->mlBits[i] = mlBits[i]; c->mlDef[i] = mlDef[i]; }
    for (int i = 0; i < 29; i++) c->ofDef[i] = ofDef[i];
    for (u32 v = 0; v < 64; v++) { u32 k = 0; while (k + 1 < 36 && llBase[k + 1] <= v) k++; c->llCode[v] = (u8)k; }
    for (u32 v = 0; v < 128; v++) { u32 k = 0; while (k + 1 < 53 && mlBase[k + 1] <= v + 3) k++; c->mlCode[v] = (u8)k; }
}
ZL_HD u32 zl_ll_code(const ZlEncConst& c, u32 ll) { return ll < 64 ? c.llCode[ll] : zl_highbit(ll) + 19; }
ZL_HD u32 zl_ml_code(const ZlEncConst& c, u32 ml) { const u32 b = ml - 3; return b < 128 ? c.mlCode[b] : zl_highbit(b) + 36; }

// A sequence record produced by the parse kernel: litLength[0:18) matchLength[18:36) offBase[36:56)
// (offBase = offset + 3, or 1..3 for a repeat offset; zstd.c:19648-19652)
ZL_HD u64 zl_enc_rec(u32 ll, u32 ml, u32 offBase) { return (u64)ll | ((u64)ml << 18) | ((u64)offBase << 36); }
#define ZL_REC_LL(r) ((u32)(r) & 0x3FFFFu)
#define ZL_REC_ML(r) ((u32)((r) >> 18) & 0x3FFFFu)
#define ZL_REC_OB(r) ((u32)((r) >> 36))

// ---- forward bit writer (restates BIT_CStream_t, zstd.c:2253-2345) onto a 4-byte aligned word buffer --------------
struct ZlBitW {
    u64 acc;
    u32 n;          // valid bits in acc, < 32 between calls
    u32* out;
    u32 pos, cap;   // words written / capacity in words
    u32 ovf;
};
ZL_HD void zl_bw_init(ZlBitW& w, u32* out, u32 capWords) { w.acc = 0; w.n = 0; w.out = out; w.pos = 0; w.cap = capWords; w.ovf = 0; }
ZL_HD void zl_bw_add(ZlBitW& w, u32 v, u32 nb)            // nb <= 32, v < 2^nb; then n <= 63
{
    w.acc |= (u64)v << w.n;
    w.n += nb;
}
ZL_HD void zl_bw_flush(ZlBitW& w)
{
    if (w.n >= 32) {
        if (w.pos < w.cap) w.out[w.pos] = (u32)w.acc; else w.ovf = 1;
        w.pos++;
        w.acc >>= 32; w.n -= 32;
    }
}
// end mark + padding (zstd.c:2334-2342); returns the stream size in bytes (0 on overflow)
ZL_HD u32 zl_bw_close(ZlBitW& w)
{
    zl_bw_add(w, 1, 1);
    zl_bw_flush(w);
    const u32 bytes = w.pos * 4 + ((w.n + 7) >> 3);
    if (w.n) { if (w.pos < w.cap) w.out[w.pos] = (u32)w.acc; else w.ovf = 1; }
    return w.ovf ? 0u : bytes;
}

// ---- FSE table construction ------------------------------------------------------------------------------------------
ZL_HD u32 zl_fse_min_log(u32 total, u32 maxSym)
{
    const u32 a = zl_highbit(total) + 1, b = zl_highbit(maxSym ? maxSym : 1) + 2;       // zstd.c:16196
    return a < b ? a : b;
}
ZL_HD u32 zl_fse_optimal_log(u32 maxLog, u32 total, u32 maxSym)                         // zstd.c:16205 (minus = 2)
{
    u32 srcBits = zl_highbit(total - 1) - 2, log = maxLog, minBits = zl_fse_min_log(total, maxSym);
    if (srcBits < log) log = srcBits;
    if (minBits > log) log = minBits;
    if (log < 5) log = 5;
    if (log > 12) log = 12;
    return log;
}
// Normalises count[0..maxSym] (sum `total`) to a distribution summing to 1<<log.  Returns false when one symbol holds
// every count (RLE) -- callers handle that before.  Primary method as FSE_normalizeCount; when its correction would
// take more than half of the largest bucket we fall back to our own proportional scheme (any distribution with the
// right sum and no zero for a present symbol is valid for the format).
ZL_HD bool zl_fse_normalize(i16* norm, u32 log, const u32* count, u32 total, u32 maxSym, bool lowProb)
{
    const u32 rtb[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};
    const i16 low = lowProb ? -1 : 1;
    const u32 scale = 62 - log;
    const u64 step = ((u64)1 << 62) / total, vStep = (u64)1 << (scale - 20);
    i32 still = 1 << log;
    u32 largest = 0; i16 largestP = 0;
    const u32 lowThreshold = total >> log;
    for (u32 s = 0; s <= maxSym; s++) {
        if (count[s] == total) return false;
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= lowThreshold) { norm[s] = low; still--; continue; }
        i16 p = (i16)((count[s] * step) >> scale);
        if (p < 8) { const u64 beat = vStep * rtb[p]; p += (count[s] * step) - ((u64)p << scale) > beat ? 1 : 0; }
        if (p > largestP) { largestP = p; largest = s; }
        norm[s] = p; still -= p;
    }
    if (-still < (norm[largest] >> 1)) { norm[largest] = (i16)(norm[largest] + still); return true; }
    // fallback: floor share with a minimum of one slot, then move single slots until the sum fits
    i32 sum = 0;
    for (u32 s = 0; s <= maxSym; s++) {
        if (!count[s]) { norm[s] = 0; continue; }
        u32 share = (u32)(((u64)count[s] << log) / total);
        norm[s] = (i16)(share ? share : 1); sum += norm[s];
    }
    const i32 size = 1 << log;
    while (sum > size) { u32 b = 0; for (u32 s = 1; s <= maxSym; s++) if (norm[s] > norm[b]) b = s; norm[b]--; sum--; }
    if (sum < size) { u32 b = 0; for (u32 s = 1; s <= maxSym; s++) if (count[s] > count[b]) b = s; norm[b] = (i16)(norm[b] + (size - sum)); }
    return true;
}
// NCount header (inverse of zl_read_ncount).  Returns bytes written (<= cap) or 0.
ZL_HD u32 zl_fse_write_ncount(u8* out, u32 cap, const i16* norm, u32 maxSym, u32 log)
{
    u64 acc = (u64)(log - 5); u32 nacc = 4, pos = 0;
    i32 remaining = (1 << log) + 1, threshold = 1 << log, nbBits = (i32)log + 1;
    u32 sym = 0; const u32 alphabet = maxSym + 1; bool prev0 = false;
#define ZL_NC_DRAIN() while (nacc >= 8) { if (pos >= cap) return 0; out[pos++] = (u8)acc; acc >>= 8; nacc -= 8; }
    while (sym < alphabet && remaining > 1) {
        if (prev0) {
            u32 start = sym;
            while (sym < alphabet && norm[sym] == 0) sym++;
            if (sym == alphabet) break;
            u32 run = sym - start;
            while (run >= 3) { acc |= (u64)3 << nacc; nacc += 2; run -= 3; ZL_NC_DRAIN(); }
            acc |= (u64)run << nacc; nacc += 2; ZL_NC_DRAIN();
        }
        i32 count = norm[sym++];
        const i32 max = (2 * threshold - 1) - remaining;
        remaining -= count < 0 ? -count : count;
        count++;
        if (count >= threshold) count += max;
        acc |= (u64)(u32)count << nacc; nacc += (u32)nbBits - (count < max ? 1u : 0u);
        prev0 = (count == 1);
        if (remaining < 1) return 0;
        while (remaining < threshold) { nbBits--; threshold >>= 1; }
        ZL_NC_DRAIN();
    }
    if (remaining != 1) return 0;
    while (nacc > 0) { if (pos >= cap) return 0; out[pos++] = (u8)acc; acc >>= 8; nacc = nacc >= 8 ? nacc - 8 : 0; }
#undef ZL_NC_DRAIN
    return pos;
}
// Encoding table: stateTbl[1<<log] (next state, already offset by the table size), dNb / dFS per symbol.
// `scratch` needs (1<<log) bytes + 2*(maxSym+2) bytes.
ZL_HD void zl_fse_build_ctable(u16* stateTbl, u32* dNb, i32* dFS, const i16* norm, u32 maxSym, u32 log, u8* symOf, u16* cumul)
{
    const u32 size = 1u << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    u32 high = size - 1;
    cumul[0] = 0;
    for (u32 u = 1; u <= maxSym + 1; u++) {
        if (norm[u - 1] == -1) { cumul[u] = (u16)(cumul[u - 1] + 1); symOf[high--] = (u8)(u - 1); }
        else cumul[u] = (u16)(cumul[u - 1] + norm[u - 1]);
    }
    u32 pos = 0;
    for (u32 s = 0; s <= maxSym; s++) {
        for (i32 i = 0; i < norm[s]; i++) {
            symOf[pos] = (u8)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    }
    for (u32 u = 0; u < size; u++) { const u32 s = symOf[u]; stateTbl[cumul[s]++] = (u16)(size + u); }
    u32 total = 0;
    for (u32 s = 0; s <= maxSym; s++) {
        const i32 n = norm[s];
        if (n == 0) { dNb[s] = ((log + 1) << 16) - size; dFS[s] = 0; }
        else if (n == -1 || n == 1) { dNb[s] = (log << 16) - size; dFS[s] = (i32)total - 1; total++; }
        else {
            const u32 maxBitsOut = log - zl_highbit((u32)n - 1), minStatePlus = (u32)n << maxBitsOut;
            dNb[s] = (maxBitsOut << 16) - minStatePlus; dFS[s] = (i32)total - n; total += (u32)n;
        }
    }
}
ZL_HD u32 zl_fse_init_state(const u16* stateTbl, u32 dNb, i32 dFS)                       // zstd.c:2772
{
    const u32 nb = (dNb + (1u << 15)) >> 16, value = (nb << 16) - dNb;
    return stateTbl[(i32)(value >> nb) + dFS];
}
// one transition: returns nbBits<<16 | bits, updates state (zstd.c:2783)
ZL_HD u32 zl_fse_step(const u16* stateTbl, u32 dNb, i32 dFS, u32& state)
{
    const u32 nb = (state + dNb) >> 16, bits = state & ((1u << nb) - 1);
    state = stateTbl[(i32)(state >> nb) + dFS];
    return (nb << 16) | bits;
}
// cost in 1/256 bit of coding a symbol whose normalised count is n in a table of 1<<log (n = -1 counts as 1)
ZL_HD u32 zl_fse_cost256(i32 n, u32 log)
{
    if (n == 0) return (log + 4) << 8;                 // not representable: callers never pick such a table
    const u32 v = n < 0 ? 1u : (u32)n, hb = zl_highbit(v);
    const u32 frac = hb ? (((v - (1u << hb)) << 8) >> hb) : 0;      // linear interpolation of log2 between powers of two
    return (log << 8) - ((hb << 8) + frac);
}

// ---- digested dictionary for the compressor (zstd.c:27450 ZSTD_loadCEntropy, 27550 ZSTD_loadZstdDictionary) -----------
// Built once on the host by ZSTD_CCtx_loadDictionary and kept in device memory.  The first block of every frame may
// (a) find matches in the dictionary content through the two prebuilt hash tables and (b) reuse the dictionary's
// entropy tables through the format's "repeat" modes (literals type 3, sequence table mode 3), which is where most of
// the gain on small inputs comes from.
struct ZlEncDictDev {
    const u8* content;       // device pointer, 16-byte aligned, followed by >= 16 zero bytes
    const u32* tabS;         // short-hash table: position + 1 of the last occurrence (0 = empty), 1 << hlogS entries
    const u32* tabL;         // long (8-byte) hash table, 1 << hlogL entries (null when the level has no long table)
    u32 contentSize;
    u32 hlogS, hlogL;
    u32 dictID;
    u32 hasEntropy;
    u32 rep[3];
    u32 hufLog;
    u16 hufCode[256];        // nbBits << 12 | code value, 0 when the dictionary's tree lacks the symbol
    u8 hufNbBits[256];
    i16 norm[3][64];         // LL, OF, ML normalized counts
    u32 log[3], maxSym[3];
    u16 state[3][512];
    u32 dNb[3][64];
    i32 dFS[3][64];
};

// =================================================================================================== Huffman (literals)
struct ZlHufCtl {
    u32 nLit, maxSym, nPresent, maxBits;
    u32 mode;            // 0 raw, 1 rle, 2 compressed
    u32 repeat;          // mode 2 only: 1 = coded with the dictionary's tree, no description (literals type 3)
    u32 nStreams;
    u32 descSize;        // tree description bytes in ZlHufSm.desc
    u32 sBeg[4], sEnd[4];// literal index ranges per stream
    u32 sBytes[4];
    u32 ovf;
};
struct ZlHufSm {
    u32 count[256];
    u32 key[256];        // sort keys count<<8|sym; the first nPresent entries after sorting
    u32 nodeCnt[256];    // internal nodes of the tree
    u16 parLeaf[256];
    u16 parInt[256];
    u8 depthInt[256];
    u8 nbBits[256];
    u16 code[256];       // nbBits<<12 | code value; 0 when the symbol is absent
    u8 weights[256];
    u8 desc[160];
    // FSE over the weights (zstd.c:16806 HUF_compressWeights): 13 symbols, tableLog <= 6
    u32 wCount[16]; i16 wNorm[16]; u16 wState[64]; u32 wdNb[16]; i32 wdFS[16]; u8 wSymOf[64]; u16 wCumul[18];
    u32 wBits[72];       // weight bitstream staging (<= 255 weights * 6 bits)
    ZlHufCtl ctl;
};

// Code lengths (<= 11 bits) for the symbols with count > 0, canonical code values in the decoder's table order
// (weight ascending, symbol ascending: zstd.c:38506-38568), and the weight array.  Requires >= 2 present symbols.
ZL_HD void zl_huf_build(ZlHufSm& f)
{
    u32 n = 0, maxSym = 0;
    for (u32 s = 0; s < 256; s++) { f.nbBits[s] = 0; f.code[s] = 0; f.weights[s] = 0; if (f.count[s]) { f.key[n++] = (f.count[s] << 8) | s; maxSym = s; } }
    f.ctl.nPresent = n; f.ctl.maxSym = maxSym;
    if (n < 2) return;                                         // callers use rle / raw literals for these
    // shell sort, ascending (count, symbol)
    const u32 gaps[6] = {109, 41, 19, 5, 3, 1};
    for (u32 g = 0; g < 6; g++) {
        const u32 gap = gaps[g];
        for (u32 i = gap; i < n; i++) { const u32 k = f.key[i]; u32 j = i; while (j >= gap && f.key[j - gap] > k) { f.key[j] = f.key[j - gap]; j -= gap; } f.key[j] = k; }
    }
    // two-queue Huffman: leaves in key[] (ascending), internal nodes appended to nodeCnt[] (non-decreasing)
    u32 li = 0, ii = 0, ni = 0;
    while (ni < n - 1) {
        u32 c = 0;
        for (u32 k = 0; k < 2; k++) {
            const bool takeLeaf = li < n && (ii >= ni || (f.key[li] >> 8) <= f.nodeCnt[ii]);
            if (takeLeaf) { c += f.key[li] >> 8; f.parLeaf[li++] = (u16)ni; }
            else { c += f.nodeCnt[ii]; f.parInt[ii++] = (u16)ni; }
        }
        f.nodeCnt[ni++] = c;
    }
    f.depthInt[n - 2] = 0;
    for (i32 k = (i32)n - 3; k >= 0; k--) f.depthInt[k] = (u8)(f.depthInt[f.parInt[k]] + 1);
    u32 numLen[40];
    for (u32 i = 0; i < 40; i++) numLen[i] = 0;
    u32 maxLen = 0;
    for (u32 k = 0; k < n; k++) { u32 d = (u32)f.depthInt[f.parLeaf[k]] + 1; if (d > 39) d = 39; f.parLeaf[k] = (u16)d; numLen[d]++; if (d > maxLen) maxLen = d; }
    if (maxLen > 11) {                                         // Kraft repair on the length histogram, then re-deal lengths by rank
        for (u32 i = 12; i < 40; i++) { numLen[11] += numLen[i]; numLen[i] = 0; }
        u32 total = 0;
        for (u32 i = 11; i >= 1; i--) total += numLen[i] << (11 - i);
        while (total != (1u << 11)) {
            numLen[11]--;
            for (u32 i = 10; i >= 1; i--) if (numLen[i]) { numLen[i]--; numLen[i + 1] += 2; break; }
            total--;
        }
        u32 k = 0;
        for (u32 len = 11; len >= 1; len--) for (u32 c = 0; c < numLen[len]; c++) f.parLeaf[k++] = (u16)len;   // least frequent first
        maxLen = 11;
    }
    while (maxLen > 1 && numLen[maxLen] == 0) maxLen--;
    f.ctl.maxBits = maxLen;
    u32 rank[13];
    for (u32 i = 0; i < 13; i++) rank[i] = 0;
    for (u32 k = 0; k < n; k++) { const u32 s = f.key[k] & 0xFF, len = f.parLeaf[k]; f.nbBits[s] = (u8)len; const u32 w = maxLen + 1 - len; f.weights[s] = (u8)w; rank[w]++; }
    u32 start[13], acc = 0;
    for (u32 w = 1; w <= maxLen; w++) { start[w] = acc; acc += rank[w] << (w - 1); }
    for (u32 s = 0; s <= maxSym; s++) {
        const u32 w = f.weights[s];
        if (!w) continue;
        f.code[s] = (u16)((f.nbBits[s] << 12) | (start[w] >> (w - 1)));
        start[w] += 1u << (w - 1);
    }
}

// Tree description (zstd.c:16907-16950).  Writes f.desc, sets ctl.descSize; returns false when the tree cannot be described.
ZL_HD bool zl_huf_write_desc(ZlHufSm& f)
{
    const u32 nw = f.ctl.maxSym;                     // weights of symbols 0..maxSym-1; the last one is implied
    f.ctl.descSize = 0;
    u32 hSize = 0;
    if (nw > 1) {                                    // try FSE-compressed weights (HUF_compressWeights)
        u32 maxW = 0, maxCount = 0;
        for (u32 i = 0; i < 16; i++) f.wCount[i] = 0;
        for (u32 s = 0; s < nw; s++) { f.wCount[f.weights[s]]++; if (f.weights[s] > maxW) maxW = f.weights[s]; }
        for (u32 i = 0; i <= maxW; i++) if (f.wCount[i] > maxCount) maxCount = f.wCount[i];
        if (maxCount != nw && maxCount > 1) {
            const u32 log = zl_fse_optimal_log(6, nw, maxW);
            if (zl_fse_normalize(f.wNorm, log, f.wCount, nw, maxW, false)) {
                const u32 nc = zl_fse_write_ncount(f.desc + 1, 60, f.wNorm, maxW, log);
                if (nc) {
                    zl_fse_build_ctable(f.wState, f.wdNb, f.wdFS, f.wNorm, maxW, log, f.wSymOf, f.wCumul);
                    // two interleaved states, symbols taken from the end (zstd.c:2802-2860): even index -> state 1, odd -> state 2
                    ZlBitW w; zl_bw_init(w, f.wBits, 72);
                    u32 st[2] = {0, 0}; bool have[2] = {false, false};
                    for (i32 i = (i32)nw - 1; i >= 0; i--) {
                        const u32 sym = f.weights[i], k = (u32)i & 1;
                        if (!have[k]) { st[k] = zl_fse_init_state(f.wState, f.wdNb[sym], f.wdFS[sym]); have[k] = true; }
                        else { const u32 r = zl_fse_step(f.wState, f.wdNb[sym], f.wdFS[sym], st[k]); zl_bw_add(w, r & 0xFFFF, r >> 16); zl_bw_flush(w); }
                    }
                    const u32 m = (1u << log) - 1;
                    zl_bw_add(w, st[1] & m, log); zl_bw_flush(w);
                    zl_bw_add(w, st[0] & m, log); zl_bw_flush(w);
                    const u32 bs = zl_bw_close(w);
                    if (bs && 1 + nc + bs <= 128) {
                        for (u32 i = 0; i < bs; i++) f.desc[1 + nc + i] = (u8)(f.wBits[i >> 2] >> (8 * (i & 3)));
                        hSize = nc + bs;
                    }
                }
            }
        }
    }
    if (hSize > 1 && hSize < nw / 2) { f.desc[0] = (u8)hSize; f.ctl.descSize = 1 + hSize; return true; }
    if (nw > 128) return false;                      // direct 4-bit weights only describe up to 128 symbols
    f.desc[0] = (u8)(128 + (nw - 1));
    for (u32 s = 0; s < nw; s += 2) f.desc[1 + s / 2] = (u8)((f.weights[s] << 4) | (s + 1 < nw ? f.weights[s + 1] : 0));
    f.ctl.descSize = 1 + (nw + 1) / 2;
    return true;
}

// One Huffman stream (lane q): literals lit[beg..end) encoded from the last to the first so that the decoder, reading
// backwards, sees them in order (zstd.c:17715-17780).  Returns bytes written.
ZL_HD u32 zl_huf_encode_stream(const u16* code, const u8* lit, u32 beg, u32 end, u32* out, u32 capWords, u32* ovf)
{
    ZlBitW w; zl_bw_init(w, out, capWords);
    u32 i = end;
    while (i > beg && (i & 3)) { const u32 e = code[lit[--i]]; zl_bw_add(w, e & 0xFFF, e >> 12); zl_bw_flush(w); }
    while (i >= beg + 4) {                             // i is a multiple of 4 here (relative to the literal buffer base, which is 4-aligned)
        i -= 4;
        const u32 v = *(const u32*)(lit + i);
        const u32 e3 = code[v >> 24], e2 = code[(v >> 16) & 0xFF], e1 = code[(v >> 8) & 0xFF], e0 = code[v & 0xFF];
        zl_bw_add(w, e3 & 0xFFF, e3 >> 12); zl_bw_add(w, e2 & 0xFFF, e2 >> 12); zl_bw_flush(w);
        zl_bw_add(w, e1 & 0xFFF, e1 >> 12); zl_bw_add(w, e0 & 0xFFF, e0 >> 12); zl_bw_flush(w);
    }
    while (i > beg) { const u32 e = code[lit[--i]]; zl_bw_add(w, e & 0xFFF, e >> 12); zl_bw_flush(w); }
    const u32 bytes = zl_bw_close(w);
    if (w.ovf) *ovf = 1;
    return bytes;
}

// Chunked form of the same stream, for many lanes per stream: a stream's literal range is cut into consecutive chunks;
// every lane first sums the code lengths of its chunk (zl_huf_chunk_bits), an exclusive scan over the chunks that are
// written EARLIER (the ones further towards the end of the range) gives its starting bit, and it then writes its codes
// there.  The first and last word a lane touches may be shared with its neighbours: those are OR-ed atomically into
// the zero-initialised buffer, the words in between are plain stores.
#if defined(__CUDA_ARCH__)
#define ZL_ATOMIC_OR(p, v) atomicOr((p), (v))
#else
#define ZL_ATOMIC_OR(p, v) (*(p) |= (v))
#endif
ZL_HD u32 zl_huf_chunk_bits(const u8* nbBits, const u8* lit, u32 beg, u32 end)
{
    u32 bits = 0, i = beg;
    while (i < end && (i & 3)) bits += nbBits[lit[i++]];
    // (16 bytes a step, the four loads issued together: every lane streams through its own chunk, so each load is a cache miss of its
    //  own and one after the other they were most of this kernel's time)
    for (; i + 16 <= end; i += 16) {
        u32 v[4];
#pragma unroll
        for (u32 q = 0; q < 4; q++) v[q] = *(const u32*)(lit + i + 4 * q);
#pragma unroll
        for (u32 q = 0; q < 4; q++) bits += nbBits[v[q] & 0xFF] + nbBits[(v[q] >> 8) & 0xFF] + nbBits[(v[q] >> 16) & 0xFF] + nbBits[v[q] >> 24];
    }
    for (; i + 4 <= end; i += 4) { const u32 v = *(const u32*)(lit + i); bits += nbBits[v & 0xFF] + nbBits[(v >> 8) & 0xFF] + nbBits[(v >> 16) & 0xFF] + nbBits[v >> 24]; }
    while (i < end) bits += nbBits[lit[i++]];
    return bits;
}
// writes the codes of lit[beg..end) (last literal first) starting at bit `bitOff` of `out`; `endMark` appends the
// closing 1 bit (the chunk at the start of the range is written last).  capWords bounds the buffer.
ZL_HD void zl_huf_encode_chunk(const u16* code, const u8* lit, u32 beg, u32 end, u32* out, u32 capWords, u32 bitOff, bool endMark, u32* ovf)
{
    u64 acc = 0; u32 n = bitOff & 31, pos = bitOff >> 5; bool first = true;
#define ZL_CH_FLUSH() if (n >= 32) { if (pos < capWords) { if (first) ZL_ATOMIC_OR(out + pos, (u32)acc); else out[pos] = (u32)acc; } else *ovf = 1; first = false; pos++; acc >>= 32; n -= 32; }
#define ZL_CH_SYM(s) { const u32 e = code[s]; acc |= (u64)(e & 0xFFF) << n; n += e >> 12; }
    u32 i = end;
    while (i > beg && (i & 3)) { ZL_CH_SYM(lit[--i]); ZL_CH_FLUSH(); }
    while (i >= beg + 16) {                                  // 16 bytes a step, the four loads issued together (see zl_huf_chunk_bits)
        i -= 16;
        u32 vv[4];
#pragma unroll
        for (u32 q = 0; q < 4; q++) vv[q] = *(const u32*)(lit + i + 4 * q);
#pragma unroll
        for (int q = 3; q >= 0; q--) {
            const u32 v = vv[q];
            ZL_CH_SYM(v >> 24); ZL_CH_SYM((v >> 16) & 0xFF); ZL_CH_FLUSH();
            ZL_CH_SYM((v >> 8) & 0xFF); ZL_CH_SYM(v & 0xFF); ZL_CH_FLUSH();
        }
    }
    while (i >= beg + 4) {
        i -= 4;
        const u32 v = *(const u32*)(lit + i);
        ZL_CH_SYM(v >> 24); ZL_CH_SYM((v >> 16) & 0xFF); ZL_CH_FLUSH();
        ZL_CH_SYM((v >> 8) & 0xFF); ZL_CH_SYM(v & 0xFF); ZL_CH_FLUSH();
    }
    while (i > beg) { ZL_CH_SYM(lit[--i]); ZL_CH_FLUSH(); }
    if (endMark) { acc |= (u64)1 << n; n += 1; ZL_CH_FLUSH(); }
    if (n) { if (pos < capWords) ZL_ATOMIC_OR(out + pos, (u32)acc); else *ovf = 1; }
#undef ZL_CH_FLUSH
#undef ZL_CH_SYM
}
// chunk k of `nchunks` of the range [beg, end): boundaries are multiples of 4 literals so the word loads stay aligned
ZL_HD void zl_huf_chunk_range(u32 beg, u32 end, u32 k, u32 nchunks, u32* cbeg, u32* cend)
{
    const u32 len = end - beg, cs = ((len + nchunks - 1) / nchunks + 3) & ~3u;
    u32 a = beg + k * cs, b = a + cs;
    if (a > end) a = end;
    if (b > end) b = end;
    *cbeg = a; *cend = b;
}

// Per-block outputs of the entropy kernels, consumed by the plan / assemble kernels.
struct ZlEncBlockOut {
    u32 litHeadSize;       // bytes in litHead: literals section header (+ tree description + jump table)
    u32 litBodyMode;       // 0 none (rle byte is in litHead), 1 raw literals (copy from the literal buffer), 2 Huffman streams
    u32 nLit;
    u32 nStreams;
    u32 sBytes[4];
    u32 seqHeadSize;       // nbSeq + modes byte + table descriptions
    u32 seqBitsSize;
    u32 flags;             // literals kernel: nonzero = store the block raw
    u32 seqOvf;            // sequences kernel: nonzero = the sequence bitstream overflowed (-> raw block).  Separate words: the two
                           // kernels may run side by side
    u8 litHead[176];
    u8 seqHead[304];
};

// Literals: decide raw / rle / Huffman from the histogram, build the code and the section header (lane 0).
// `lit` is only needed for the rle byte.  Stream ranges land in ctl.sBeg / sEnd.  `dict` (null unless this is the
// first block of a frame compressed with an entropy-carrying dictionary) offers its tree as a "repeat" table
// (zstd.c:20714-20760: HUF_repeat_valid lowers the minimum to 6 literals and, with a 3-byte header, forces one stream).
ZL_HD void zl_lit_plan(ZlHufSm& f, ZlEncBlockOut& o, const u8* lit, u32 nLit, const ZlEncDictDev* dict)
{
    ZlHufCtl& c = f.ctl;
    c.nLit = nLit; c.mode = 0; c.repeat = 0; c.nStreams = 0; c.ovf = 0; c.descSize = 0;
    o.nLit = nLit; o.nStreams = 0; o.flags = 0;
    for (u32 k = 0; k < 4; k++) { o.sBytes[k] = 0; c.sBytes[k] = 0; }
    u32 maxCount = 0, present = 0;
    for (u32 s = 0; s < 256; s++) { if (f.count[s]) present++; if (f.count[s] > maxCount) maxCount = f.count[s]; }
    const u32 minGain = (nLit >> 6) + 2;                                                 // zstd.c:19607
    u32 mode = 0, estBest = 0xFFFFFFFFu;
    bool dictTaken = false;
    if (nLit >= 1 && present == 1 && nLit > 2) mode = 1;                                 // zstd.c:20644 rle literals
    else if (dict && nLit >= 6 && nLit <= 1024) {
        // small literal sections prefer a valid dictionary tree without building their own (HUF_compress_internal's
        // preferRepeat heuristic, zstd.c:18030-18036 with 20714: strategy < lazy and srcSize <= 1024)
        u64 bits = 0; bool ok = true;
        for (u32 s = 0; s < 256; s++) if (f.count[s]) { if (!dict->hufNbBits[s]) { ok = false; break; } bits += (u64)f.count[s] * dict->hufNbBits[s]; }
        const u32 est = (u32)((bits + 7) >> 3) + 1u;
        if (ok && est + minGain < nLit) {
            mode = 2; c.repeat = 1; c.nStreams = 1; c.descSize = 0; dictTaken = true;
            for (u32 s = 0; s < 256; s++) { f.code[s] = dict->hufCode[s]; f.nbBits[s] = dict->hufNbBits[s]; }
        }
    }
    if (mode == 0 && !dictTaken) {
        if (nLit >= 64 && present >= 2 && maxCount > (nLit >> 7) + 4) {                  // zstd.c:20685, 18043 ("probably not compressible")
            zl_huf_build(f);
            if (zl_huf_write_desc(f)) {
                u64 bits = 0;
                for (u32 s = 0; s < 256; s++) bits += (u64)f.count[s] * f.nbBits[s];
                const u32 nStreams = nLit < 256 ? 1u : 4u;                               // zstd.c:20705
                const u32 est = (u32)((bits + 7) >> 3) + c.descSize + (nStreams == 4 ? 6u + 4u : 1u);
                if (est + minGain < nLit) { mode = 2; c.nStreams = nStreams; estBest = est; }
            }
        }
        if (dict && nLit >= 6) {                                                         // the dictionary's tree, if it covers every literal
            u64 bits = 0; bool ok = true;
            for (u32 s = 0; s < 256; s++) if (f.count[s]) { if (!dict->hufNbBits[s]) { ok = false; break; } bits += (u64)f.count[s] * dict->hufNbBits[s]; }
            const u32 nStreams = nLit < 1024 ? 1u : 4u;
            const u32 est = (u32)((bits + 7) >> 3) + (nStreams == 4 ? 6u + 4u : 1u);
            if (ok && est + minGain < nLit && est <= estBest) {
                mode = 2; c.repeat = 1; c.nStreams = nStreams; c.descSize = 0;
                for (u32 s = 0; s < 256; s++) { f.code[s] = dict->hufCode[s]; f.nbBits[s] = dict->hufNbBits[s]; }
            }
        }
    }
    c.mode = mode;
    u8* h = o.litHead;
    if (mode != 2) {                                   // raw / rle headers: zstd.c:20602-20675
        u32 hs;
        if (nLit < 32) { h[0] = (u8)(mode | (nLit << 3)); hs = 1; }
        else if (nLit < 4096) { const u32 v = mode | (1u << 2) | (nLit << 4); h[0] = (u8)v; h[1] = (u8)(v >> 8); hs = 2; }
        else { const u32 v = mode | (3u << 2) | (nLit << 4); h[0] = (u8)v; h[1] = (u8)(v >> 8); h[2] = (u8)(v >> 16); hs = 3; }
        if (mode == 1) h[hs++] = lit[0];
        o.litHeadSize = hs; o.litBodyMode = mode == 0 ? 1u : 0u;
        return;
    }
    const u32 seg = (nLit + 3) / 4;
    if (c.nStreams == 1) { c.sBeg[0] = 0; c.sEnd[0] = nLit; }
    else for (u32 k = 0; k < 4; k++) { c.sBeg[k] = k * seg; c.sEnd[k] = k == 3 ? nLit : (k + 1) * seg; }
    o.litBodyMode = 2; o.nStreams = c.nStreams;
}
// After the streams are encoded (sizes in ctl.sBytes): write the compressed-literals header, the tree description and
// the jump table (zstd.c:20772-20795, 17838-17873).  Falls back to raw literals when the result is not smaller.
ZL_HD void zl_lit_finish(ZlHufSm& f, ZlEncBlockOut& o)
{
    ZlHufCtl& c = f.ctl;
    if (c.mode != 2) return;
    const u32 nLit = c.nLit;
    u32 body = 0;
    for (u32 k = 0; k < c.nStreams; k++) { body += c.sBytes[k]; if (!c.sBytes[k] || c.sBytes[k] > 65535) c.ovf = 1; }
    const u32 cSize = c.descSize + (c.nStreams == 4 ? 6u : 0u) + body;
    const u32 lhSize = 3 + (nLit >= 1024 ? 1u : 0u) + (nLit >= 16384 ? 1u : 0u);
    const u32 minGain = (nLit >> 6) + 2;
    if (c.ovf || cSize + minGain >= nLit || (lhSize == 3 && cSize > 1023) || (lhSize == 4 && cSize > 16383) || cSize > 262143) {
        // not worth it: raw literals
        u8* h = o.litHead; u32 hs;
        if (nLit < 32) { h[0] = (u8)(nLit << 3); hs = 1; }
        else if (nLit < 4096) { const u32 v = (1u << 2) | (nLit << 4); h[0] = (u8)v; h[1] = (u8)(v >> 8); hs = 2; }
        else { const u32 v = (3u << 2) | (nLit << 4); h[0] = (u8)v; h[1] = (u8)(v >> 8); h[2] = (u8)(v >> 16); hs = 3; }
        o.litHeadSize = hs; o.litBodyMode = 1; o.nStreams = 0;
        return;
    }
    u8* h = o.litHead;
    const u32 ty = c.repeat ? 3u : 2u;                 // compressed literals with / without their own tree description
    if (lhSize == 3) { const u32 v = ty | ((c.nStreams == 1 ? 0u : 1u) << 2) | (nLit << 4) | (cSize << 14); h[0] = (u8)v; h[1] = (u8)(v >> 8); h[2] = (u8)(v >> 16); }
    else if (lhSize == 4) { const u32 v = ty | (2u << 2) | (nLit << 4) | (cSize << 18); h[0] = (u8)v; h[1] = (u8)(v >> 8); h[2] = (u8)(v >> 16); h[3] = (u8)(v >> 24); }
    else { const u64 v = ty | (3u << 2) | ((u64)nLit << 4) | ((u64)cSize << 22); for (u32 i = 0; i < 5; i++) h[i] = (u8)(v >> (8 * i)); }
    u32 p = lhSize;
    for (u32 i = 0; i < c.descSize; i++) h[p++] = f.desc[i];
    if (c.nStreams == 4) for (u32 k = 0; k < 3; k++) { h[p++] = (u8)c.sBytes[k]; h[p++] = (u8)(c.sBytes[k] >> 8); }
    o.litHeadSize = p;
    for (u32 k = 0; k < 4; k++) o.sBytes[k] = k < c.nStreams ? c.sBytes[k] : 0;
}

// =================================================================================================== sequences
struct ZlSeqEncCtl {
    u32 nbSeq;
    u32 mode[3];          // 0 predefined, 1 rle, 2 FSE-compressed, 3 repeat (the dictionary's table)
    u32 log[3];
    u32 maxSym[3];
    u32 hdrSize[3];
};
struct ZlSeqEncSm {
    u32 count[3][64];
    i16 norm[3][64];
    u16 state[3][512];
    u32 dNb[3][64];
    i32 dFS[3][64];
    u8 hdr[3][96];
    u8 symOf[3][512];
    u16 cumul[3][66];
    ZlSeqEncCtl ctl;
};
// Encoding tables that exist before any block is looked at: the three predefined ones (built once per device, zl_enc_init_const) -- the
// dictionary's live in ZlEncDictDev in the same form.  The device coder copies them with the whole warp instead of building them per block.
struct ZlEncDefTables { u16 state[3][64]; u32 dNb[3][64]; i32 dFS[3][64]; };
// In the first block of a frame compressed with a dictionary, blocks with few sequences do not get a table of their own (its description
// would cost more than it saves next to the dictionary's table): below these counts the choice is between the dictionary's table and the
// predefined one, and nothing is normalised or built.  The reference decides the same way at these levels (zstd.c:21040-21059: a valid
// repeat table is taken below 1,000 sequences; dynamicFse_nbSeq_min = (1 << defaultNormLog) * mult >> 3, mult = 10 - strategy).  Without
// a dictionary the cost comparison stays: on 400-byte objects a table of their own is worth 10 % of the size.
#define ZL_SEQ_FSE_MIN_LLML 64u
#define ZL_SEQ_FSE_MIN_OF 32u
// code of table t (0 LL, 1 OF, 2 ML) for a record
ZL_HD u32 zl_seq_code(const ZlEncConst& k, u32 t, u64 rec)
{
    return t == 0 ? zl_ll_code(k, ZL_REC_LL(rec)) : (t == 1 ? zl_highbit(ZL_REC_OB(rec)) : zl_ml_code(k, ZL_REC_ML(rec)));
}
// Table t (0 LL, 1 OF, 2 ML) from its code histogram f.count[t] (maxSym = largest code present, lastCode = code of
// the last sequence): pick the table type, build the encoding table and the table description.  The choice compares
// the exact description cost plus the estimated symbol cost of an FSE-compressed table with the predefined one (the
// reference uses sequence-count heuristics at these levels, zstd.c:21034-21059).
ZL_HD void zl_seq_build_from_hist(ZlSeqEncSm& f, u32 t, u32 nbSeq, u32 maxSym, u32 lastCode, const ZlEncConst& k, const ZlEncDictDev* dict)
{
    const u32 maxLogT = t == 1 ? 8u : 9u, defLog = t == 1 ? 5u : 6u;
    const i16* def = t == 0 ? k.llDef : (t == 1 ? k.ofDef : k.mlDef);
    const u32 defMax = t == 0 ? 35u : (t == 1 ? 28u : 52u);
    u32* cnt = f.count[t];
    u32 maxCount = 0;
    for (u32 s = 0; s <= maxSym; s++) if (cnt[s] > maxCount) maxCount = cnt[s];
    ZlSeqEncCtl& c = f.ctl;
    c.hdrSize[t] = 0;
    const bool defOk = maxSym <= defMax;
    if (maxCount == nbSeq && !(defOk && nbSeq <= 2)) {                       // one symbol only: rle table (zstd.c:21021-21031)
        c.mode[t] = 1; c.log[t] = 0; c.maxSym[t] = maxSym;
        f.hdr[t][0] = (u8)maxSym; c.hdrSize[t] = 1;
        f.state[t][0] = 0; f.state[t][1] = 0;
        for (u32 s = 0; s < 64; s++) { f.dNb[t][s] = 0; f.dFS[t][s] = 0; }
        return;
    }
    u32 costDef = 0xFFFFFFFFu;
    if (defOk) { u64 cst = 0; for (u32 s = 0; s <= maxSym; s++) cst += (u64)cnt[s] * zl_fse_cost256(def[s], defLog); costDef = (u32)(cst >> 8); }
    u32 costRep = 0xFFFFFFFFu;                                                // the dictionary's table as a "repeat" table (zstd.c:21034-21076)
    if (dict && maxSym <= dict->maxSym[t]) {
        u64 cst = 0; bool ok = true;
        for (u32 s = 0; s <= maxSym; s++) if (cnt[s]) { if (!dict->norm[t][s]) { ok = false; break; } cst += (u64)cnt[s] * zl_fse_cost256(dict->norm[t][s], dict->log[t]); }
        if (ok) costRep = (u32)(cst >> 8);
    }
    u32 costFse = 0xFFFFFFFFu, log = 0, nc = 0;
    const bool fewSeqs = defOk && costRep != 0xFFFFFFFFu && nbSeq < (t == 1 ? ZL_SEQ_FSE_MIN_OF : ZL_SEQ_FSE_MIN_LLML);
    if (nbSeq >= 8 && maxCount != nbSeq && !fewSeqs) {
        // the last sequence's symbol only initialises the state, it is not coded as a transition (zstd.c:21126-21129)
        u32 total = nbSeq;
        if (cnt[lastCode] > 1) { cnt[lastCode]--; total--; }
        log = zl_fse_optimal_log(maxLogT, nbSeq, maxSym);
        if (zl_fse_normalize(f.norm[t], log, cnt, total, maxSym, nbSeq >= 2048)) {
            nc = zl_fse_write_ncount(f.hdr[t], 96, f.norm[t], maxSym, log);
            if (nc) { u64 cst = 0; for (u32 s = 0; s <= maxSym; s++) cst += (u64)cnt[s] * zl_fse_cost256(f.norm[t][s], log); costFse = (u32)(cst >> 8) + 8 * nc; }
        }
    }
    {
        if (costRep != 0xFFFFFFFFu && costRep <= costDef && costRep <= costFse) {
            const u32 dl = dict->log[t];
            c.mode[t] = 3; c.log[t] = dl; c.maxSym[t] = dict->maxSym[t]; c.hdrSize[t] = 0;
#if !defined(__CUDA_ARCH__)                                                      // (the device coder copies the dictionary's tables with the whole warp)
            for (u32 i = 0; i < (1u << dl); i++) f.state[t][i] = dict->state[t][i];
            for (u32 s = 0; s < 64; s++) { f.dNb[t][s] = dict->dNb[t][s]; f.dFS[t][s] = dict->dFS[t][s]; }
#endif
            return;
        }
    }
    if (costFse < costDef) {
        c.mode[t] = 2; c.log[t] = log; c.maxSym[t] = maxSym; c.hdrSize[t] = nc;
        zl_fse_build_ctable(f.state[t], f.dNb[t], f.dFS[t], f.norm[t], maxSym, log, f.symOf[t], f.cumul[t]);
        return;
    }
    // predefined table (always available for our offsets: the largest offset code is 17 < 28)
    const u32 defSyms = t == 0 ? 36u : (t == 1 ? 29u : 53u);
    c.mode[t] = 0; c.log[t] = defLog; c.maxSym[t] = defSyms - 1;
#if !defined(__CUDA_ARCH__)                                                      // (the device coder has the predefined tables prebuilt: ZlEncDefTables)
    for (u32 s = 0; s < defSyms; s++) f.norm[t][s] = def[s];
    zl_fse_build_ctable(f.state[t], f.dNb[t], f.dFS[t], f.norm[t], defSyms - 1, defLog, f.symOf[t], f.cumul[t]);
#endif
}
// serial form (CPU emulation): histogram, then the above
ZL_HD void zl_seq_build_table(ZlSeqEncSm& f, u32 t, const u64* recs, u32 nbSeq, const ZlEncConst& k, const ZlEncDictDev* dict)
{
    u32* cnt = f.count[t];
    for (u32 s = 0; s < 64; s++) cnt[s] = 0;
    u32 maxSym = 0, lastCode = 0;
    for (u32 i = 0; i < nbSeq; i++) { const u32 c = zl_seq_code(k, t, recs[i]); cnt[c]++; if (c > maxSym) maxSym = c; lastCode = c; }
    zl_seq_build_from_hist(f, t, nbSeq, maxSym, lastCode, k, dict);
}
// sequences section header (lane 0): nbSeq, modes byte, table descriptions in LL, OF, ML order (zstd.c:25446-25490)
ZL_HD void zl_seq_write_head(const ZlSeqEncSm& f, ZlEncBlockOut& o)
{
    const ZlSeqEncCtl& c = f.ctl;
    u8* h = o.seqHead; u32 p = 0;
    const u32 n = c.nbSeq;
    if (n < 128) h[p++] = (u8)n;
    else if (n < 0x7F00) { h[p++] = (u8)((n >> 8) + 0x80); h[p++] = (u8)n; }
    else { h[p++] = 0xFF; h[p++] = (u8)(n - 0x7F00); h[p++] = (u8)((n - 0x7F00) >> 8); }
    if (n) {
        h[p++] = (u8)((c.mode[0] << 6) | (c.mode[1] << 4) | (c.mode[2] << 2));
        for (u32 t = 0; t < 3; t++) for (u32 i = 0; i < c.hdrSize[t]; i++) h[p++] = f.hdr[t][i];
    }
    o.seqHeadSize = p;
}
// additional bits of a record for table t: value and count
ZL_HD u32 zl_seq_extra(const ZlEncConst& k, u32 t, u64 rec, u32 code, u32* nb)
{
    if (t == 0) { *nb = k.llBits[code]; return ZL_REC_LL(rec) - k.llBase[code]; }
    if (t == 1) { *nb = code; return ZL_REC_OB(rec) - (1u << code); }
    *nb = k.mlBits[code]; return ZL_REC_ML(rec) - k.mlBase[code];
}

// The sequences bitstream (zstd.c:21146-21235): sequences are written last to first; for each one the three state
// transitions (OF, ML, LL) then the additional bits (LL, ML, OF); the first one written only initialises the states.
// One caller runs all three chains: they are independent, so their shared-memory latencies overlap.
ZL_HD u32 zl_seq_encode(const ZlSeqEncSm& f, const ZlEncConst& k, const u64* recs, u32 nbSeq, u32* out, u32 capWords, u32* ovf)
{
    ZlBitW w; zl_bw_init(w, out, capWords);
    const u16* tLL = f.state[0]; const u16* tOF = f.state[1]; const u16* tML = f.state[2];
    u64 rec = recs[nbSeq - 1];
    u32 cLL = zl_seq_code(k, 0, rec), cOF = zl_seq_code(k, 1, rec), cML = zl_seq_code(k, 2, rec);
    u32 sML = zl_fse_init_state(tML, f.dNb[2][cML], f.dFS[2][cML]);
    u32 sOF = zl_fse_init_state(tOF, f.dNb[1][cOF], f.dFS[1][cOF]);
    u32 sLL = zl_fse_init_state(tLL, f.dNb[0][cLL], f.dFS[0][cLL]);
    u32 nb, v;
    v = zl_seq_extra(k, 0, rec, cLL, &nb); zl_bw_add(w, v, nb);
    v = zl_seq_extra(k, 2, rec, cML, &nb); zl_bw_add(w, v, nb); zl_bw_flush(w);
    v = zl_seq_extra(k, 1, rec, cOF, &nb); zl_bw_add(w, v, nb); zl_bw_flush(w);
    u64 next = nbSeq >= 2 ? recs[nbSeq - 2] : 0;
    for (u32 n = nbSeq - 1; n-- > 0;) {
        rec = next;
        if (n) next = recs[n - 1];
        cLL = zl_seq_code(k, 0, rec); cOF = zl_seq_code(k, 1, rec); cML = zl_seq_code(k, 2, rec);
        const u32 rOF = zl_fse_step(tOF, f.dNb[1][cOF], f.dFS[1][cOF], sOF);
        const u32 rML = zl_fse_step(tML, f.dNb[2][cML], f.dFS[2][cML], sML);
        const u32 rLL = zl_fse_step(tLL, f.dNb[0][cLL], f.dFS[0][cLL], sLL);
        zl_bw_add(w, rOF & 0xFFFF, rOF >> 16); zl_bw_add(w, rML & 0xFFFF, rML >> 16); zl_bw_add(w, rLL & 0xFFFF, rLL >> 16);
        zl_bw_flush(w);
        v = zl_seq_extra(k, 0, rec, cLL, &nb); zl_bw_add(w, v, nb);
        v = zl_seq_extra(k, 2, rec, cML, &nb); zl_bw_add(w, v, nb); zl_bw_flush(w);
        v = zl_seq_extra(k, 1, rec, cOF, &nb); zl_bw_add(w, v, nb); zl_bw_flush(w);
    }
    const ZlSeqEncCtl& c = f.ctl;
    zl_bw_add(w, sML & ((1u << c.log[2]) - 1), c.log[2]); zl_bw_flush(w);
    zl_bw_add(w, sOF & ((1u << c.log[1]) - 1), c.log[1]); zl_bw_flush(w);
    zl_bw_add(w, sLL & ((1u << c.log[0]) - 1), c.log[0]); zl_bw_flush(w);
    const u32 bytes = zl_bw_close(w);
    if (w.ovf) *ovf = 1;
    return bytes;
}

// Block-level decision (zstd.c:25535-25536, 19607-19613, 25496-25502): returns the compressed payload size, or 0 when
// the block must be stored raw.
ZL_HD u32 zl_enc_block_payload(const ZlEncBlockOut& o, u32 srcSize, u32 nbSeq)
{
    if (o.flags || o.seqOvf) return 0;
    u32 body = 0;
    if (o.litBodyMode == 1) body = o.nLit;
    else if (o.litBodyMode == 2) for (u32 k = 0; k < o.nStreams; k++) body += o.sBytes[k];
    const u32 total = o.litHeadSize + body + o.seqHeadSize + o.seqBitsSize;
    if (nbSeq && o.seqHeadSize + o.seqBitsSize < 4) return 0;
    const u32 minGain = (srcSize >> 6) + 2;
    if (total + minGain >= srcSize) return 0;
    return total;
}
