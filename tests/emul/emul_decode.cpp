// TEST INFRASTRUCTURE ONLY.  Compiles the product's ZL_HD device logic (zl_dec_entropy.cuh) with g++
// and drives the quad phases of kernels K1a / K1b sequentially on the CPU, so format bugs can be found
// without a GPU.  The sequence records are then executed by a trivial serial loop (stands in for K2).
// Never linked into the product.
#include <vector>
#include <cstring>
#include "../../zstdlite_b200/csrc/zl_dec_entropy.cuh"
#include "../../zstdlite_b200/csrc/zl_plan.h"

static const ZlConstTables g_ct = {ZL_LL_BASE_INIT, ZL_ML_BASE_INIT, ZL_LL_BITS_INIT, ZL_ML_BITS_INIT,
                                   ZL_LL_DEFNORM_INIT, ZL_ML_DEFNORM_INIT, ZL_OF_DEFNORM_INIT};

extern "C" size_t zl_emul_decompress_frame(void* dstv, size_t cap, const void* srcv, size_t size, unsigned* nrecOut)
{
    // copy src into a padded buffer at a deliberately odd offset to exercise `bias`
    std::vector<u8> sbuf(size + 64, 0xA5);
    u8* src = sbuf.data() + 32 + 1;
    memcpy(src, srcv, size);
    ZlFrameDesc d; memset(&d, 0, sizeof(d));
    d.src = src; d.dst = (u8*)dstv; d.srcSize = (u32)size; d.dstCap = (u32)cap;
    zl_plan_frame(d.srcSize, d.dstCap, true, &d.litCap, &d.recCap, &d.hdrCap);
    std::vector<u8> lits(d.litCap + 16);
    std::vector<u64> recs(d.recCap);
    std::vector<ZlBlockHdr> hdrs(d.hdrCap);
    std::vector<ZlUnit> units(d.hdrCap);
    u32 nunits = 0;
    ZlFrameInfo info;
    const u32 bias = (u32)(((size_t)d.src) & 3);
    const u32* wbase = (const u32*)(d.src - bias);
    zl_index_frame(d, info, hdrs.data(), 0, 0, 0, units.data(), &nunits, d.hdrCap);          // K0
    // K1a / K1b: every unit on its own, in an arbitrary order (here: last block first) -- blocks are independent
    for (u32 ui = nunits; !info.err && ui-- > 0;) {
        const u32 blk = units[ui].block;
        if (((hdrs[blk].flags >> 4) & 3) == 2) {   // K1a
            static ZlLitSm f;
            u32 useDict = 0;
            zl_lit_unit_head(f, d, hdrs.data(), blk, wbase, bias, &useDict);
            if (useDict) f.ctl.err = ZL_E_dictionary_corrupted;
            if (!f.ctl.err) {
                if (f.ctl.needHufFill) for (u32 q = 0; q < 4; q++) zl_huf_fill(f, q);
                for (u32 q = 0; q < f.ctl.nStreams; q++)
                    f.ctl.sErr[q] = zl_huf_stream(f.huf, f.ctl.hufLog, wbase, bias, f.ctl.sBeg[q], f.ctl.sEnd[q],
                                                  lits.data() + f.ctl.sOut[q], f.ctl.sLen[q]);
            }
            const u32 e = zl_lit_unit_finish(f);
            if (e) info.err = e;
        }
        if (!info.err && hdrs[blk].nbSeq) {        // K1b
            static ZlSeqSm f;
            static i16 norm[3 * ZL_NORM_STRIDE];
            zl_seq_head(f, d, hdrs.data(), blk, g_ct, norm);
            if (!f.ctl.err && f.ctl.useDict) f.ctl.err = ZL_E_corruption_detected;
            if (!f.ctl.err && f.ctl.needBuild) for (u32 q = 0; q < 3; q++) zl_seq_fse_build(f, q, norm);
            hdrs[blk].nrec = zl_seq_decode(f, recs.data() + hdrs[blk].recOff, wbase, bias, g_ct, nullptr);
            if (f.ctl.err) info.err = f.ctl.err;
        }
    }
    if (info.err) return (size_t)0 - (size_t)info.err;
    // K2 restated serially: records -> (ll, ml, offBase), repeat-offset history, the checks of ZSTD_execSequence, the copies
    u8* out = d.dst; u32 op = 0; unsigned nrecTotal = 0;
    u32 hist[3] = {1, 4, 8};
    for (u32 b = 0; b < info.nblocks; b++) {
        const ZlBlockHdr& h = hdrs[b];
        u32 type = h.flags & 3;
        const u32 room = d.dstCap - op;
        if (type != 2) {
            if (h.regenSize > room) return (size_t)0 - (size_t)ZL_E_dstSize_tooSmall;
            if (type == 0) memcpy(out + op, d.src + h.srcOff, h.regenSize); else memset(out + op, (h.flags >> 8) & 0xFF, h.regenSize);
            op += h.regenSize;
            continue;
        }
        u32 cap = room, capErr = ZL_E_dstSize_tooSmall;
        if (cap > ZL_BLOCKSIZE_MAX) { cap = ZL_BLOCKSIZE_MAX; capErr = ZL_E_corruption_detected; }
        u32 litMode = (h.flags >> 4) & 3, o = op, lp = 0;
        const u8* lit = litMode == 0 ? d.src + h.srcOff : lits.data() + h.litOff;
        u8 rle = (u8)((h.flags >> 8) & 0xFF);
        for (u32 i = 0; i < h.nrec; i++) {
            u32 ll, ml, ob;
            zl_rec_decode(recs[h.recOff + i], g_ct.llBase, g_ct.llBits, g_ct.mlBase, g_ct.mlBits, ll, ml, ob);
            const u32 off = zl_rep_resolve(hist, ll, ml, ob);
            if (ll > h.litSize - lp) return (size_t)0 - (size_t)ZL_E_corruption_detected;
            if ((u64)(o - op) + ll + ml > cap) return (size_t)0 - (size_t)capErr;
            if (ml && (off == 0 || off > o + ll)) return (size_t)0 - (size_t)ZL_E_corruption_detected;
            for (u32 k = 0; k < ll; k++) out[o + k] = litMode == 1 ? rle : lit[lp + k];
            o += ll; lp += ll;
            for (u32 k = 0; k < ml; k++) { out[o] = out[o - off]; o++; }
        }
        if ((o - op) + (h.litSize - lp) > cap) return (size_t)0 - (size_t)capErr;
        for (u32 k = lp; k < h.litSize; k++) out[o++] = litMode == 1 ? rle : lit[k];
        nrecTotal += h.nrec;
        op = o;
    }
    if (info.contentSize != ~0ull && info.contentSize != (u64)op) return (size_t)0 - (size_t)ZL_E_corruption_detected;
    if (nrecOut) *nrecOut = nrecTotal;
    return op;
}
