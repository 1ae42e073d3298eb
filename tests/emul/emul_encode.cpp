// TEST INFRASTRUCTURE ONLY.  CPU emulation of the CUDA compressor: the product's ZL_HD entropy logic
// (zl_enc_entropy.cuh, zl_enc_match.cuh) compiled with g++, driven serially; the two warp-parallel kernels
// (match candidates, greedy walk) are restated as the plain serial loops whose results they must reproduce
// (zl_enc_match.cuh header comment).  Output must decode with the reference's libzstd; on a GPU box the tests also
// require the CUDA path to produce byte-identical frames.  Never linked into the product.
#include <vector>
#include <cstring>
#include <cstdlib>
#include "../../zstdlite_b200/csrc/zl_enc_entropy.cuh"
#include "../../zstdlite_b200/csrc/zl_enc_match.cuh"
#include "../../zstdlite_b200/csrc/zl_enc_dict.h"

// digested dictionary of the emulation (host pointers)
struct EmulDict { ZlEncDictDev d; std::vector<u8> content; std::vector<u32> tabS, tabL; };

static u32 rd32(const u8* p) { u32 v; memcpy(&v, p, 4); return v; }

static u32 match_len_capped(const u8* src, u32 n, u32 p, i32 q)
{
    u32 lim = n - p; if (lim > ZL_M_VERIFY) lim = ZL_M_VERIFY;      // in-block candidates: measured up to ZL_M_VERIFY, extended by the walk
    u32 l = 0;
    while (l < lim && src[p + l] == src[(u32)q + l]) l++;
    return l;
}

static u32 dict_match(const EmulDict& D, const std::vector<u32>& tab, u32 h, const u8* src, u32 n, u32 p, u32 mls, u32* off)
{
    const u32 e = tab[h];
    if (!e) return 0;
    const u32 q = e - 1, room = D.d.contentSize - q;
    u32 lim = n - p; if (lim > ZL_M_CAP) lim = ZL_M_CAP; if (lim > room) lim = room;
    u32 l = 0;
    while (l < lim && src[p + l] == D.content[q + l]) l++;
    if (l < mls) return 0;
    *off = p + room;
    return l;
}
// far candidates (frames of more than one block): the frame-wide table of EARLIEST occurrences of every 8-byte hash (zl_k_far_build)
struct EmulFar { const u8* frame = nullptr; u32 frameSize = 0, blockOff = 0, log = 0; const u32* tab = nullptr; };
static EmulFar g_far;
// stage 1: M[p] for every position
static void emul_match(const u8* src, u32 n, const ZlEncParams& P, std::vector<u32>& M, const EmulDict* D)
{
    M.assign(n, 0);
    u32 hlogS, hlogL;
    zl_block_hlog(P, n, hlogS, hlogL);                       // small blocks: small tables (zl_enc_match.cuh)
    std::vector<u16> tabS((size_t)1 << hlogS, 0), tabL(hlogL ? (size_t)1 << hlogL : 1, 0);
    for (u32 p = 0; p + 8 <= n; p++) {
        const u32 lo = rd32(src + p), hi = rd32(src + p + 4);
        const u32 hS = zl_hash_short(lo, hi, P.mls, hlogS);
        const i32 qS = zl_cand_pos(tabS[hS], p); tabS[hS] = (u16)p;
        u32 bestLen = 0, bestOff = 0;
        if (P.hlogL) {
            const u32 hL = zl_hash_long(lo, hi, hlogL);
            const i32 qL = zl_cand_pos(tabL[hL], p); tabL[hL] = (u16)p;
            if (qL >= 0) { const u32 l = match_len_capped(src, n, p, qL); if (l >= P.mls) { bestLen = l; bestOff = p - (u32)qL; } }
        }
        if (qS >= 0 && bestLen < 8) { const u32 l = match_len_capped(src, n, p, qS); if (l >= P.mls && l > bestLen) { bestLen = l; bestOff = p - (u32)qS; } }   // (a long-hash match of >= 8 bytes is taken as it is)
        u32 lim = n - p; if (lim > ZL_M_CAP) lim = ZL_M_CAP;
        if (D && bestLen < lim) {
            u32 dOff = 0;
            if (P.hlogL) { const u32 l = dict_match(*D, D->tabL, zl_hash_long(lo, hi, D->d.hlogL), src, n, p, P.mls, &dOff); if (l > bestLen) { bestLen = l; bestOff = dOff; } }
            if (bestLen < lim) { const u32 l = dict_match(*D, D->tabS, zl_hash_short(lo, hi, P.mls, D->d.hlogS), src, n, p, P.mls, &dOff); if (l > bestLen) { bestLen = l; bestOff = dOff; } }
        }
        if (g_far.tab) {                                     // far candidates (zl_k_match): own region, then the one before it
            const u32 pos = g_far.blockOff + p, hx = zl_far_hash(lo, hi, g_far.log), hF = hx >> 8, reg = pos >> ZL_FAR_REGION_LOG;
            u32 limF = n - p; if (limF > ZL_M_CAP) limF = ZL_M_CAP;
            bool got = false;
            u32 e = g_far.tab[((size_t)reg << g_far.log) + hF];
            u32 q = (reg << ZL_FAR_REGION_LOG) + (e >> 8);
            if (e != ZL_FAR_EMPTY && (e & 255u) == (hx & 255u) && q < pos && pos - q > 65535u && pos - q < P.farMaxOff) {
                u32 l = 0;
                while (l < limF && g_far.frame[q + l] == src[p + l]) l++;
                if (zl_far_better(l, bestLen, P.mls)) { bestLen = l; bestOff = pos - q; got = true; }
            }
            if (!got && zl_far_use_prev(pos)) {
                e = g_far.tab[((size_t)(reg - 1) << g_far.log) + hF];
                q = ((reg - 1) << ZL_FAR_REGION_LOG) + (e >> 8);
                if (e != ZL_FAR_EMPTY && (e & 255u) == (hx & 255u) && pos - q > 65535u && pos - q < P.farMaxOff) {
                    u32 l = 0;
                    while (l < limF && g_far.frame[q + l] == src[p + l]) l++;
                    if (zl_far_better(l, bestLen, P.mls)) { bestLen = l; bestOff = pos - q; }
                }
            }
        }
        M[p] = bestLen ? ((bestOff << 8) | bestLen) : 0;
    }
}
// stage 2: greedy walk -> records, literals, histogram; the block is walked as independent segments of ZL_PARSE_SEG bytes (zl_k_parse)
struct EmulSeq { u32 ll, ml, off; };                         // explicit form, as ZSTD_Sequence carries it (src/zstd/zstd.h:1501-1530)
static std::vector<EmulSeq>* g_seqOut = nullptr;             // stage-level tests: the walk's sequences with explicit offsets
static const EmulSeq* g_seqIn = nullptr; static u32 g_seqInCount = 0;     // stage-level tests: sequences to encode instead of the walk's
static void emul_parse(const u8* src, u32 n, const std::vector<u32>& M, std::vector<u64>& recs, std::vector<u8>& lit, u32* hist, bool firstBlock,
                       const EmulDict* D)
{
    recs.clear(); lit.clear();
    for (u32 i = 0; i < 256; i++) hist[i] = 0;
    const bool repPref = firstBlock && D;
    u32 carry = 0;                                           // literals since the last match of the segments before
    for (u32 s0 = 0; s0 < n; s0 += ZL_PARSE_SEG) {
        const u32 s1 = n - s0 < ZL_PARSE_SEG ? n : s0 + ZL_PARSE_SEG;
        // blocks are compressed independently and walked in independent segments: only the first segment of a frame knows the
        // decoder's repeat offsets (1, 4, 8; zstd.c:15416); the others start with an unknown history (0 never matches an offset)
        const bool first = firstBlock && s0 == 0;
        ZlReps reps = {first ? 1u : 0u, first ? 4u : 0u, first ? 8u : 0u};
        if (first && D && D->d.hasEntropy) { reps.r0 = D->d.rep[0]; reps.r1 = D->d.rep[1]; reps.r2 = D->d.rep[2]; }
        u32 p = s0, anchor = s0; bool firstSeq = true;
        while (p < s1) {
            const u32 m = M[p];
            if (!m) { p++; continue; }
            u32 len = m & 0xFF; u32 off = m >> 8;
            if (len >= ZL_M_VERIFY && off <= p) while (p + len < s1 && src[p + len] == src[p + len - off]) len++;
            if (p + len > s1) len = s1 - p;                  // a match ends with its segment
            if (len < 3) { p++; continue; }
            // Dictionary mode only: repeat-offset preference (cf. the repcode checks at ip+1 / ip+2 of zstd.c:29989, 30801).  A match
            // at the most recent offset starting at p, p+1 or p+2 (inside the same 32-position window) costs no offset bits; it
            // is taken when it is at most 4 bytes shorter.  Small dictionary-compressed inputs are dominated by offset cost.
            if (repPref && reps.r0 && off != reps.r0) {
                for (u32 k = 0; k < 3; k++) {
                    const u32 q = p + k;
                    if ((p & 31) + k >= 32 || reps.r0 > q || q + 4 > s1) continue;
                    u32 lim = s1 - q; if (lim > ZL_M_CAP) lim = ZL_M_CAP;
                    u32 rl = 0; while (rl < lim && src[q + rl] == src[q + rl - reps.r0]) rl++;
                    if (rl >= 4 && rl + 4 >= len) {
                        if (rl == ZL_M_CAP) while (q + rl < s1 && src[q + rl] == src[q + rl - reps.r0]) rl++;
                        p = q; len = rl; off = reps.r0; break;
                    }
                }
            }
            const u32 ll = p - anchor;
            for (u32 k = anchor; k < p; k++) { lit.push_back(src[k]); hist[src[k]]++; }
            const u32 ob = zl_rep_encode(reps, off, ll);
            if (g_seqOut) g_seqOut->push_back({ll + (firstSeq ? carry : 0u), len, off});
            recs.push_back(zl_enc_rec(ll + (firstSeq ? carry : 0u), len, ob));
            firstSeq = false;
            p += len; anchor = p;
        }
        for (u32 k = anchor; k < s1; k++) { lit.push_back(src[k]); hist[src[k]]++; }
        carry = firstSeq ? carry + (s1 - anchor) : s1 - anchor;
    }
}

// one block -> payload bytes (without the 3-byte block header); returns 0 when the block must be stored raw
static u32 emul_block(const u8* src, u32 n, const ZlEncParams& P, const ZlEncConst& K, std::vector<u8>& out, bool firstBlock, const EmulDict* D)
{
    if (!firstBlock) D = nullptr;
    const ZlEncDictDev* De = (D && D->d.hasEntropy) ? &D->d : nullptr;
    out.clear();
    if (n < 7) return 0;                                     // zstd.c:25725
    std::vector<u32> M; std::vector<u64> recs; std::vector<u8> lit;
    static ZlHufSm hs; static ZlSeqEncSm ss; static ZlEncBlockOut o;
    if (g_seqIn) {
        // sequences from outside (the reference's ZSTD_generateSequences): only the entropy stage below is ours
        for (u32 i = 0; i < 256; i++) hs.count[i] = 0;
        ZlReps reps = {firstBlock ? 1u : 0u, firstBlock ? 4u : 0u, firstBlock ? 8u : 0u};
        u32 p = 0;
        for (u32 i = 0; i < g_seqInCount; i++) {
            const EmulSeq q = g_seqIn[i];
            for (u32 k = 0; k < q.ll; k++) { lit.push_back(src[p + k]); hs.count[src[p + k]]++; }
            p += q.ll;
            if (!q.ml) continue;                             // (a block delimiter: its literals are the last literals)
            recs.push_back(zl_enc_rec(q.ll, q.ml, zl_rep_encode(reps, q.off, q.ll)));
            p += q.ml;
        }
        for (; p < n; p++) { lit.push_back(src[p]); hs.count[src[p]]++; }
    } else {
    emul_match(src, n, P, M, D);
    emul_parse(src, n, M, recs, lit, hs.count, firstBlock, D);
    }
    const u32 nLit = (u32)lit.size(), nbSeq = (u32)recs.size();
    std::vector<u8> litPad(nLit + 16); if (nLit) memcpy(litPad.data(), lit.data(), nLit);
    // literals kernel
    std::vector<u32> sbuf[4];
    zl_lit_plan(hs, o, litPad.data(), nLit, De);
    if (hs.ctl.mode == 2) {
        // as the CUDA kernel does it: 32 lanes, 32 / nStreams chunks per stream, exclusive scan of the chunk sizes
        const u32 per = 32 / hs.ctl.nStreams;
        for (u32 q = 0; q < hs.ctl.nStreams; q++) {
            const u32 cnt = hs.ctl.sEnd[q] - hs.ctl.sBeg[q];
            sbuf[q].assign(cnt * 11 / 32 + 4, 0);
            u32 bits[32], cb[32], ce[32], total = 0;
            for (u32 k = 0; k < per; k++) { zl_huf_chunk_range(hs.ctl.sBeg[q], hs.ctl.sEnd[q], k, per, &cb[k], &ce[k]); bits[k] = zl_huf_chunk_bits(hs.nbBits, litPad.data(), cb[k], ce[k]); total += bits[k]; }
            u32 off = 0;
            for (i32 k = (i32)per - 1; k >= 0; k--) { zl_huf_encode_chunk(hs.code, litPad.data(), cb[k], ce[k], sbuf[q].data(), (u32)sbuf[q].size(), off, k == 0, &hs.ctl.ovf); off += bits[k]; }
            hs.ctl.sBytes[q] = (total + 1 + 7) >> 3;
            // cross-check against the single-lane form
            std::vector<u32> ref1(sbuf[q].size(), 0); u32 ovf1 = 0;
            const u32 b1 = zl_huf_encode_stream(hs.code, litPad.data(), hs.ctl.sBeg[q], hs.ctl.sEnd[q], ref1.data(), (u32)ref1.size(), &ovf1);
            if (b1 != hs.ctl.sBytes[q] || memcmp(ref1.data(), sbuf[q].data(), b1)) return 0xFFFFFFFFu;
        }
        zl_lit_finish(hs, o);
    }
    // sequences kernel
    std::vector<u32> seqBits(n / 4 + 16, 0);
    ss.ctl.nbSeq = nbSeq;
    o.seqBitsSize = 0; o.seqOvf = 0;
    if (nbSeq) {
        for (u32 t = 0; t < 3; t++) zl_seq_build_table(ss, t, recs.data(), nbSeq, K, De);
        zl_seq_write_head(ss, o);
        u32 ovf = 0;
        o.seqBitsSize = zl_seq_encode(ss, K, recs.data(), nbSeq, seqBits.data(), (u32)seqBits.size(), &ovf);
        o.seqOvf = (ovf || !o.seqBitsSize) ? 1u : 0u;
    } else zl_seq_write_head(ss, o);
    const u32 payload = zl_enc_block_payload(o, n, nbSeq);
    if (!payload) return 0;
    out.insert(out.end(), o.litHead, o.litHead + o.litHeadSize);
    if (o.litBodyMode == 1) out.insert(out.end(), lit.begin(), lit.end());
    else if (o.litBodyMode == 2) for (u32 q = 0; q < o.nStreams; q++) { const u8* b = (const u8*)sbuf[q].data(); out.insert(out.end(), b, b + o.sBytes[q]); }
    out.insert(out.end(), o.seqHead, o.seqHead + o.seqHeadSize);
    { const u8* b = (const u8*)seqBits.data(); out.insert(out.end(), b, b + o.seqBitsSize); }
    return (u32)out.size() == payload ? payload : 0xFFFFFFFFu;
}

// whole frame (zstd.c:27007 frame chunk loop, 27733 epilogue); xxh32 = low 32 bits of XXH64(content) supplied by the caller
static size_t emul_compress(void* dstv, size_t cap, const void* srcv, size_t size, int level, int checksumFlag, unsigned xxh32, const EmulDict* D);
extern "C" size_t zl_emul_compress_frame(void* dstv, size_t cap, const void* srcv, size_t size, int level, int checksumFlag, unsigned xxh32)
{
    return emul_compress(dstv, cap, srcv, size, level, checksumFlag, xxh32, nullptr);
}
// same with a dictionary (digested per call: test infrastructure)
extern "C" size_t zl_emul_compress_frame_dict(void* dstv, size_t cap, const void* srcv, size_t size, int level, int checksumFlag, unsigned xxh32,
                                              const void* dict, size_t dictSize)
{
    static EmulDict D; static std::vector<u8> last; static int lastLevel = -100;
    const u8* dp = (const u8*)dict;
    if (lastLevel != level || last.size() != dictSize || memcmp(last.data(), dp, dictSize)) {
        if (zl_dict_digest_host(dp, dictSize, zl_enc_params(level), &D.d, D.content, D.tabS, D.tabL)) return (size_t)0 - 30;
        last.assign(dp, dp + dictSize); lastLevel = level;
    }
    return emul_compress(dstv, cap, srcv, size, level, checksumFlag, xxh32, &D);
}
static size_t emul_compress(void* dstv, size_t cap, const void* srcv, size_t size, int level, int checksumFlag, unsigned xxh32, const EmulDict* D)
{
    static ZlEncConst K; static bool init = false;
    if (!init) { zl_enc_const_init(&K); init = true; }
    const ZlEncParams P = zl_enc_params(level);
    const u8* src = (const u8*)srcv;
    std::vector<u8> frame(32);
    // far candidates (zl_enc_match.cuh): frames of more than one block get the table of earliest occurrences
    const bool far = size > ZL_BLOCKSIZE_MAX && size < 0xFFFFFF00ull && !getenv("ZL_EMUL_NOFAR");
    std::vector<u32> farTab;
    if (far) {
        const u32 flog = zl_far_log(size);
        farTab.assign((size_t)zl_far_entries(size), 0xFFFFFFFFu);
        for (size_t q = 0; q + 8 <= size; q++) {
            const u32 hx = zl_far_hash(rd32(src + q), rd32(src + q + 4), flog);
            const size_t h = ((q >> ZL_FAR_REGION_LOG) << flog) + (hx >> 8);
            const u32 e = (((u32)q & ((1u << ZL_FAR_REGION_LOG) - 1)) << 8) | (hx & 255u);
            if (e < farTab[h]) farTab[h] = e;
        }
        g_far.frame = src; g_far.frameSize = (u32)size; g_far.log = flog; g_far.tab = farTab.data();
    }
    frame.resize(zl_write_frame_header(frame.data(), size, D ? D->d.dictID : 0u, checksumFlag ? 1u : 0u, far ? P.farMaxOff : 0u));
    size_t pos = 0; bool first = true;
    std::vector<u8> payload;
    do {
        const u32 n = (u32)(size - pos < ZL_BLOCKSIZE_MAX ? size - pos : ZL_BLOCKSIZE_MAX);
        const u32 last = pos + n == size ? 1u : 0u;
        u8 bh[3];
        // (no RLE blocks: the reference only emits them for non-first blocks under 25 bytes of output, zstd.c:26873-26884;
        //  a run compresses to ~10 bytes as one sequence, so the product skips the special case)
        g_far.blockOff = (u32)pos;
        u32 ps = emul_block(src + pos, n, P, K, payload, first, D);
        if (ps == 0xFFFFFFFFu) { g_far = EmulFar(); return (size_t)0 - 1; }
        if (!ps) { zl_write_block_header(bh, last, 0, n); frame.insert(frame.end(), bh, bh + 3); frame.insert(frame.end(), src + pos, src + pos + n); }
        else { zl_write_block_header(bh, last, 2, ps); frame.insert(frame.end(), bh, bh + 3); frame.insert(frame.end(), payload.begin(), payload.end()); }
        pos += n; first = false;
    } while (pos < size);
    g_far = EmulFar();
    if (checksumFlag) for (u32 i = 0; i < 4; i++) frame.push_back((u8)(xxh32 >> (8 * i)));
    if (frame.size() > cap) return (size_t)0 - 70;
    memcpy(dstv, frame.data(), frame.size());
    return frame.size();
}

// ---- stage-level entry points (tests/test_stage_level.py) ----------------------------------------------------------------------
// (a) the emulated match finder + walk of ONE block (<= 128 KiB) -> explicit (litLength, matchLength, offset) triples; returns the count
extern "C" size_t zl_emul_sequences(const void* srcv, size_t size, int level, unsigned* triples, size_t cap)
{
    if (size > ZL_BLOCKSIZE_MAX) return 0;
    const ZlEncParams P = zl_enc_params(level < 1 ? 1 : (level > 3 ? 3 : level));
    std::vector<u32> M; std::vector<u64> recs; std::vector<u8> lit; u32 hist[256];
    std::vector<EmulSeq> out;
    emul_match((const u8*)srcv, (u32)size, P, M, nullptr);
    g_seqOut = &out;
    emul_parse((const u8*)srcv, (u32)size, M, recs, lit, hist, true, nullptr);
    g_seqOut = nullptr;
    if (out.size() > cap) return 0;
    for (size_t i = 0; i < out.size(); i++) { triples[3 * i] = out[i].ll; triples[3 * i + 1] = out[i].ml; triples[3 * i + 2] = out[i].off; }
    return out.size();
}
// (b) ONE block's frame from GIVEN sequences: the product's literal / sequence entropy stage (zl_enc_entropy.cuh) in isolation
extern "C" size_t zl_emul_encode_sequences(void* dstv, size_t cap, const void* srcv, size_t size, int level, const unsigned* triples, size_t nseq)
{
    if (size > ZL_BLOCKSIZE_MAX) return 0;
    std::vector<EmulSeq> in(nseq);
    for (size_t i = 0; i < nseq; i++) in[i] = {triples[3 * i], triples[3 * i + 1], triples[3 * i + 2]};
    g_seqIn = in.data(); g_seqInCount = (u32)nseq;
    const size_t r = emul_compress(dstv, cap, srcv, size, level, 0, 0, nullptr);
    g_seqIn = nullptr; g_seqInCount = 0;
    return r;
}
