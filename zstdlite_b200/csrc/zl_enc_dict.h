// zl_enc_dict.h -- host-side digest of a dictionary for the compressor (plain C++, no CUDA calls): shared by
// ZSTD_CCtx_loadDictionary's device upload (zl_api_compress.cu) and the CPU emulation in tests/emul.
//   entropy section -> encoding tables                  (zstd.c:27450 ZSTD_loadCEntropy)
//   content         -> padded copy + two hash tables    (the role of ZSTD_fillHashTable 30663 / ZSTD_fillDoubleHashTable
//                      29894: every position is inserted, the last occurrence of a hash wins)
#pragma once
#include <string.h>
#include <vector>
#include "zl_dec_entropy.cuh"     // ZL_HD parsers of the entropy section (zl_huf_read_stats, zl_read_ncount)
#include "zl_enc_entropy.cuh"
#include "zl_enc_match.cuh"

#define ZL_DICT_MAX_CONTENT (8u << 20)

// Fills *hd (device pointers left null), `content` (contentSize + 64 bytes, zero padded) and the tables.
// Returns 0 or a ZlErr code.
static inline u32 zl_dict_digest_host(const u8* dict, size_t dictSize, const ZlEncParams& P, ZlEncDictDev* hd, std::vector<u8>& content,
                                      std::vector<u32>& tabS, std::vector<u32>& tabL)
{
    std::vector<u8> pad(dictSize + 32, 0);
    u8* d = pad.data() + ((4 - (((size_t)pad.data()) & 3)) & 3);       // 4-aligned copy for the word reader
    if (dictSize) memcpy(d, dict, dictSize);
    memset(hd, 0, sizeof(*hd));
    size_t contentOff = 0;
    if (dictSize >= 8 && zl_rd32(d) == ZL_MAGIC_DICT) {
        hd->dictID = zl_rd32(d + 4);
        ZlLitSm* f = new ZlLitSm();
        memset(f, 0, sizeof(*f));
        size_t p = 8; bool ok = true;
        const u32 th = zl_huf_read_stats(*f, d + p, (u32)(dictSize - p), (const u32*)d, 0, (u32)p);
        if (!th) ok = false;
        if (ok) {   // canonical codes in the decoder's table order (weight ascending, symbol ascending), as zl_huf_build assigns them
            const u32 tl = f->ctl.hufLog, nsym = f->ctl.nsym;
            u32 rank[13] = {0}, start[13] = {0}, acc = 0;
            for (u32 s = 0; s < nsym; s++) rank[f->weights[s]]++;
            for (u32 w = 1; w <= tl; w++) { start[w] = acc; acc += rank[w] << (w - 1); }
            for (u32 s = 0; s < nsym; s++) {
                const u32 w = f->weights[s];
                if (!w) continue;
                const u32 nb = tl + 1 - w;
                hd->hufNbBits[s] = (u8)nb; hd->hufCode[s] = (u16)((nb << 12) | (start[w] >> (w - 1)));
                start[w] += 1u << (w - 1);
            }
            hd->hufLog = tl; p += th;
        }
        delete f;
        const u32 order[3] = {1, 2, 0}, maxSymT[3] = {35, 31, 52}, maxLog[3] = {9, 8, 9};      // OF, ML, LL in the file
        std::vector<u8> symOf(512); std::vector<u16> cumul(66);
        for (int k = 0; ok && k < 3; k++) {
            const u32 t = order[k]; u32 ms = maxSymT[t], tl = 0;
            const u32 h = p < dictSize ? zl_read_ncount(d + p, (u32)(dictSize - p), hd->norm[t], &ms, &tl) : 0;
            if (!h || tl > maxLog[t]) { ok = false; break; }
            hd->log[t] = tl; hd->maxSym[t] = ms;
            zl_fse_build_ctable(hd->state[t], hd->dNb[t], hd->dFS[t], hd->norm[t], ms, tl, symOf.data(), cumul.data());
            p += h;
        }
        if (ok && p + 12 > dictSize) ok = false;
        if (ok) {
            const size_t csz = dictSize - (p + 12);
            for (int i = 0; i < 3; i++) { const u32 r = zl_rd32(d + p + 4 * i); if (r == 0 || r > csz) ok = false; hd->rep[i] = r; }
            contentOff = p + 12;
        }
        if (!ok) return ZL_E_dictionary_corrupted;
        hd->hasEntropy = 1;
    }
    size_t contentSize = dictSize - contentOff;
    if (contentSize > ZL_DICT_MAX_CONTENT) { contentOff += contentSize - ZL_DICT_MAX_CONTENT; contentSize = ZL_DICT_MAX_CONTENT; }   // the nearest 8 MiB
    content.assign(contentSize + 64, 0);
    if (contentSize) memcpy(content.data(), d + contentOff, contentSize);
    hd->contentSize = (u32)contentSize;
    u32 hl = 10; while (hl < 18 && ((size_t)1 << hl) < 2 * contentSize) hl++;
    hd->hlogS = hl; hd->hlogL = P.hlogL ? hl : 0;
    tabS.assign((size_t)1 << hd->hlogS, 0); tabL.assign(P.hlogL ? (size_t)1 << hd->hlogL : 1, 0);
    for (size_t i = 0; i + 8 <= contentSize; i++) {
        const u32 lo = zl_rd32(content.data() + i), hi = zl_rd32(content.data() + i + 4);
        tabS[zl_hash_short(lo, hi, P.mls, hd->hlogS)] = (u32)i + 1;
        if (P.hlogL) tabL[zl_hash_long(lo, hi, hd->hlogL)] = (u32)i + 1;
    }
    return 0;
}
