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
    zl_plan_frame(d.srcSize, d.dstCap, &d.litCap, &d.recCap, &d.hdrCap, &d.ckCap);
    std::vector<u8> lits(d.litCap + 16);
    std::vector<u64> recs(d.recCap);
    std::vector<ZlBlockHdr> hdrs(d.hdrCap);
    ZlFrameInfo info;
    const u32 bias = (u32)(((size_t)d.src) & 3);
    const u32* wbase = (const u32*)(d.src - bias);
    {   // K1a
        static ZlLitSm f;
        zl_lit_begin_frame(f, d, info, 0);
        for (;;) {
            zl_lit_block_head(f, d, info, hdrs.data(), wbase, bias);
            if (f.ctl.done) break;
            if (f.ctl.needHufFill) for (u32 q = 0; q < 4; q++) zl_huf_fill(f, q);
            for (u32 q = 0; q < f.ctl.nStreams; q++)
                f.ctl.sErr[q] = zl_huf_stream(f.huf, f.ctl.hufLog, wbase, bias, f.ctl.sBeg[q], f.ctl.sEnd[q],
                                              lits.data() + f.ctl.sOut[q], f.ctl.sLen[q]);
        }
    }
    {   // K1b
        static ZlSeqSm f;
        static i16 norm[3 * ZL_NORM_STRIDE];
        zl_seq_begin_frame(f, info, 0);
        for (u32 b = 0; b < info.nblocks; b++) {
            ZlBlockHdr h = hdrs[b];
            if ((h.flags & 3) != 2) { zl_seq_plain_block(f, d, h); continue; }
            zl_seq_head(f, d, h, g_ct, norm);
            if (!f.ctl.err && f.ctl.needBuild) for (u32 q = 0; q < 3; q++) zl_seq_fse_build(f, q, norm);
            zl_seq_decode(f, d, h, recs.data(), wbase, bias, g_ct, nullptr);
            hdrs[b] = h;
        }
        zl_seq_finish_frame(f, info);
    }
    if (info.err) return (size_t)0 - (size_t)info.err;
    u8* out = d.dst; u32 op = 0; unsigned nrecTotal = 0;
    for (u32 b = 0; b < info.nblocks; b++) {
        const ZlBlockHdr& h = hdrs[b];
        u32 type = h.flags & 3;
        if (type == 0) memcpy(out + op, d.src + h.srcOff, h.regenSize);
        else if (type == 1) memset(out + op, (h.flags >> 8) & 0xFF, h.regenSize);
        else {
            u32 litMode = (h.flags >> 4) & 3, o = op, lp = 0;
            const u8* lit = litMode == 0 ? d.src + h.srcOff : lits.data() + h.litOff;
            u8 rle = (u8)((h.flags >> 8) & 0xFF);
            for (u32 i = 0; i < h.nrec; i++) {
                u64 r = recs[h.recOff + i];
                u32 ll = (u32)(r & 0xFFFF), ml = (u32)((r >> 16) & 0xFFFF), off = (u32)(r >> 32);
                for (u32 k = 0; k < ll; k++) out[o + k] = litMode == 1 ? rle : lit[lp + k];
                o += ll; lp += ll;
                for (u32 k = 0; k < ml; k++) { out[o] = out[o - off]; o++; }
            }
            for (u32 k = lp; k < h.litSize; k++) out[o++] = litMode == 1 ? rle : lit[k];
            if (o - op != h.regenSize) return (size_t)0 - 998;
            nrecTotal += h.nrec;
        }
        op += h.regenSize;
    }
    if (nrecOut) *nrecOut = nrecTotal;
    return op;
}
