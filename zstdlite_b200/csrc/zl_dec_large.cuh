// zl_dec_large.cuh -- sequence execution for LARGE frames (one multi-megabyte frame with a window that
// spans many blocks, the shape zstd_unserialize() of a real data.frame has).
//
// The warp-per-frame execute kernel (zl_dec_exec.cuh) walks a frame block after block; for a 16 MiB frame that is one
// warp and ~230 ms.  Frames of ZL_LARGE_FRAME_BYTES and more take this path instead, which is parallel over blocks and
// then over bytes:
//   L1 zl_k_lblock_scan   warp per CHUNK of ZL_LCHUNK_RECS records of a block: sums the lengths of its records and composes their
//                         repeat-offset transforms into one transform of the history (zstd.c:44290-44326)
//   L2 a zl_k_lblock_compose (thread per block: sums and composed transform of its chunks), b zl_k_lframe_prefix (lane per frame:
//                         prefix over the blocks -> output offset and incoming history of every block, destination / content-size
//                         checks, zstd.c:41497-41520, 41646), c zl_k_lblock_spread (thread per block: the same for every chunk)
//   L3 zl_k_lblock_emit   warp per chunk: offsets and positions of its sequences (as K2), the checks of ZSTD_execSequence,
//                         literals written to the output, and for every MATCH byte p a parent pointer parent[p] = p - offset
//                         (bytes that come from the dictionary are copied at once)
//   L4 zl_k_ljump         thread per output byte, repeated: pointer jumping.  A byte whose parent is final copies it and
//                         becomes final; otherwise it adopts its parent's parent.  Chains of length n resolve in
//                         ceil(log2 n) + 1 passes, every pass fully parallel (ZSTD_execSequence's byte-serial match copy,
//                         zstd.c:44013, restated as a parallel prefix over the copy graph).
// parent[p] >= ZL_PAR_DONE marks a final byte; the low 8 bits hold the pass that made it final, so that a byte finalised
// in the running pass is not read before the kernel boundary makes it visible.
#pragma once
#include "zl_dec_exec.cuh"
#include "zl_launch.h"            // ZlLBlock, ZlSymSlot, ZL_PAR_DONE

#if defined(__CUDACC__)

ZL_D ZlSymSlot zl_sym_minus1(ZlSymSlot s) { if (s.kind == 3) s.val -= 1; else s.val += 1; return s; }
ZL_D u32 zl_sym_apply(const ZlSymSlot& s, u32 h0, u32 h1, u32 h2)
{
    if (s.kind == 3) return s.val;
    const u32 h = s.kind == 0 ? h0 : (s.kind == 1 ? h1 : h2);
    return h > s.val ? h - s.val : 0u;               // 0 = invalid offset (caught where it is used)
}

// parent[0 .. n) = ZL_PAR_DONE, 16 bytes per lane where aligned
ZL_D void zl_mark_done(u32* parent, u32 n, u32 lane)
{
    u32 head = (u32)((4 - ((((size_t)parent) >> 2) & 3)) & 3);
    if (head > n) head = n;
    if (lane < head) parent[lane] = ZL_PAR_DONE;
    const u32 body = (n - head) >> 2;
    uint4* p4 = (uint4*)(parent + head);
    const uint4 v = make_uint4(ZL_PAR_DONE, ZL_PAR_DONE, ZL_PAR_DONE, ZL_PAR_DONE);
    for (u32 i = lane; i < body; i += 32) p4[i] = v;
    const u32 done = head + (body << 2);
    if (done + lane < n) parent[done + lane] = ZL_PAR_DONE;
}

// ---- L1 ------------------------------------------------------------------------------------------------------------
// records [r0, r1) of a compressed block
ZL_D void zl_lchunk_scan(const u64* __restrict__ recs, u32 r0, u32 r1, const u32* xtab, u32 lane, ZlLChunk& outC)
{
    ZlSymSlot t0 = {0, 0}, t1 = {1, 0}, t2 = {2, 0};
    u32 sumL = 0, sumO = 0;
    for (u32 base = r0; base < r1; base += 32) {
        const bool valid = base + lane < r1;
        const u64 rec = valid ? __ldg(recs + base + lane) : 0ull;
        u32 ll, ml, ob;
        zl_lane_record(rec, valid, xtab, ll, ml, ob);
        u32 sl, so, totalL, totalO;
        zl_batch_positions(ll, ml, lane, sl, so, totalL, totalO);
        sumL += totalL; sumO += totalO;
        const bool isM = ml != 0, isNew = isM && ob >= 4;
        const u32 idx = zl_rep_idx(ll, ml, ob);
        if (__ballot_sync(ZL_FULL, idx == 3)) {      // rare: symbolic, record by record
            for (u32 l = 0; l < 32; l++) {
                const u32 lll = __shfl_sync(ZL_FULL, ll, l), lml = __shfl_sync(ZL_FULL, ml, l), lob = __shfl_sync(ZL_FULL, ob, l);
                if (!lml || !lob) continue;
                if (lob >= 4) { t2 = t1; t1 = t0; t0.kind = 3; t0.val = lob - 3; continue; }
                const u32 ix = lob - 1 + (lll == 0 ? 1u : 0u);
                if (ix == 0) continue;
                const ZlSymSlot n = ix == 1 ? t1 : (ix == 2 ? t2 : zl_sym_minus1(t0));
                if (ix >= 2) t2 = t1;
                t1 = t0; t0 = n;
            }
        } else {
            const u32 fresh = ob - 3;
            const u32 T = zl_rept_scan(zl_rept_of(isNew, idx, lane), lane);
            const u32 Lt = __shfl_sync(ZL_FULL, T, 31);
            ZlSymSlot n[3];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const u32 b = (Lt >> (8 * j)) & 0xFFu;
                const u32 f = __shfl_sync(ZL_FULL, fresh, b & 31u);
                if (b & 0x80u) { n[j].kind = 3; n[j].val = f; }
                else n[j] = (b & 3u) == 0 ? t0 : ((b & 3u) == 1 ? t1 : t2);
            }
            t0 = n[0]; t1 = n[1]; t2 = n[2];
        }
    }
    if (lane == 0) { outC.sumL = sumL; outC.sumO = sumO; outC.t[0] = t0; outC.t[1] = t1; outC.t[2] = t2; }
}

// ---- L2: prefix over the chunks of a frame in three steps ------------------------------------------------------------------
// transform B after transform A, symbolically (zl_sym_apply of the result == apply B to the outputs of A)
ZL_D ZlSymSlot zl_sym_after(const ZlSymSlot& b, const ZlSymSlot& a0, const ZlSymSlot& a1, const ZlSymSlot& a2)
{
    if (b.kind == 3) return b;
    ZlSymSlot r = b.kind == 0 ? a0 : (b.kind == 1 ? a1 : a2);
    if (r.kind == 3) r.val = r.val > b.val ? r.val - b.val : 0u;      // (0 = invalid offset, as zl_sym_apply)
    else r.val += b.val;
    return r;
}
// L2a (thread per block): sums and composed transform of the block's chunks
ZL_D void zl_lblock_compose(const ZlBlockHdr& h, const ZlLChunk* __restrict__ C, ZlLBlock& B)
{
    ZlSymSlot t0 = {0, 0}, t1 = {1, 0}, t2 = {2, 0};
    u32 sumL = 0, sumO = 0, err = 0;
    if ((h.flags & 3) == 2 && h.nrec) {
        const u32 nch = zl_lchunks(h.nrec);
        for (u32 c = 0; c < nch; c++) {
            const ZlLChunk k = C[c];
            const ZlSymSlot n0 = zl_sym_after(k.t[0], t0, t1, t2), n1 = zl_sym_after(k.t[1], t0, t1, t2), n2 = zl_sym_after(k.t[2], t0, t1, t2);
            t0 = n0; t1 = n1; t2 = n2;
            sumL += k.sumL; sumO += k.sumO;
            if (sumO > ZL_BLOCKSIZE_MAX || sumL > ZL_BLOCKSIZE_MAX) { err = ZL_E_corruption_detected; break; }
        }
    }
    B.sumL = sumL; B.sumO = sumO; B.t[0] = t0; B.t[1] = t1; B.t[2] = t2; B.err = err;
}
// L2b (one lane per frame): prefix over the blocks -> output offset and incoming history of every block, size checks
ZL_D u32 zl_lframe_prefix(const ZlFrameDesc& d, const ZlFrameInfo& info, const ZlBlockHdr* __restrict__ hdrs, ZlLBlock* lb, u32 r0, u32 r1, u32 r2, u32* totalOut)
{
    u32 h0 = r0, h1 = r1, h2 = r2;
    u64 op = 0;
    for (u32 b = 0; b < info.nblocks; b++) {
        ZlLBlock& B = lb[b];
        const ZlBlockHdr& h = hdrs[b];
        if (B.err) return B.err;
        u32 regen;
        if ((h.flags & 3) != 2) regen = h.regenSize;
        else {
            if (B.sumL > h.litSize) return ZL_E_corruption_detected;
            regen = B.sumO + (h.litSize - B.sumL);
            if (regen > ZL_BLOCKSIZE_MAX) return ZL_E_corruption_detected;          // zstd.c:44066 (a block regenerates <= 128 KiB)
        }
        if (op + regen > d.dstCap) return ZL_E_dstSize_tooSmall;
        B.regen = regen; B.outOff = (u32)op; B.h[0] = h0; B.h[1] = h1; B.h[2] = h2;
        const u32 n0 = zl_sym_apply(B.t[0], h0, h1, h2), n1 = zl_sym_apply(B.t[1], h0, h1, h2), n2 = zl_sym_apply(B.t[2], h0, h1, h2);
        h0 = n0; h1 = n1; h2 = n2;
        op += regen;
    }
    if (info.contentSize != ~0ull && info.contentSize != op) return ZL_E_corruption_detected;      // zstd.c:41646
    *totalOut = (u32)op;
    return 0;
}
// L2c (thread per block): output / literal offset and incoming history of every chunk of the block
ZL_D void zl_lblock_spread(const ZlBlockHdr& h, const ZlLBlock& B, ZlLChunk* C)
{
    if ((h.flags & 3) != 2) return;
    const u32 nch = zl_lchunks(h.nrec);
    u32 h0 = B.h[0], h1 = B.h[1], h2 = B.h[2], sumL = 0, sumO = 0;
    for (u32 c = 0; c < nch; c++) {
        C[c].outOff = B.outOff + sumO; C[c].litOff = sumL; C[c].h[0] = h0; C[c].h[1] = h1; C[c].h[2] = h2;
        if (h.nrec) {
            const u32 n0 = zl_sym_apply(C[c].t[0], h0, h1, h2), n1 = zl_sym_apply(C[c].t[1], h0, h1, h2), n2 = zl_sym_apply(C[c].t[2], h0, h1, h2);
            h0 = n0; h1 = n1; h2 = n2;
            sumL += C[c].sumL; sumO += C[c].sumO;
        }
    }
}

// ---- L3 ------------------------------------------------------------------------------------------------------------
// chunk c of block B: records [r0, r1); the last chunk of a block also writes the literals that follow the last sequence
template <bool kDict>
ZL_D u32 zl_lchunk_emit(u8* out, u32* parent, const ZlFrameDesc& d, const ZlBlockHdr& h, const ZlLBlock& B, const ZlLChunk& C, u32 r0, u32 r1, bool lastChunk,
                        const u8* __restrict__ lit, const u64* __restrict__ recs, const u8* __restrict__ dict, u32 dictSize, const u32* xtab, u32 lane)
{
    const u32 op = B.outOff;
    const u32 type = h.flags & 3;
    if (type != 2) {                                               // raw / RLE block: final bytes
        if (type == 0) zl_warp_copy(out + op, d.src + h.srcOff, h.regenSize, lane);
        else zl_warp_fill(out + op, (h.flags >> 8) & 0xFF, h.regenSize, lane);
        zl_mark_done(parent + op, h.regenSize, lane);
        return 0;
    }
    const u32 litMode = (h.flags >> 4) & 3, rleByte = (h.flags >> 8) & 0xFF;
    const u32 litSize = h.litSize;
    u32 outPos = C.outOff, litPos = C.litOff;
    u32 h0 = C.h[0], h1 = C.h[1], h2 = C.h[2];
    for (u32 base = r0; base < r1; base += 32) {
        const bool valid = base + lane < r1;
        const u64 rec = valid ? __ldcs(recs + base + lane) : 0ull;
        u32 ll, ml, ob;
        zl_lane_record(rec, valid, xtab, ll, ml, ob);
        const bool isM = ml != 0;
        const u32 off = zl_batch_offsets(ll, ml, ob, lane, h0, h1, h2);
        u32 sl, so, totalL, totalO;
        zl_batch_positions(ll, ml, lane, sl, so, totalL, totalO);
        const u32 litExcl = sl - ll, dstLit = outPos + so - ll - ml, dm = dstLit + ll;
        if (totalL > litSize - litPos) return ZL_E_corruption_detected;
        if ((outPos - op) + totalO > B.regen) return ZL_E_corruption_detected;                       // (sizes were summed by L1 / L2)
        if (__ballot_sync(ZL_FULL, isM && (off == 0 || off > dm + dictSize))) return ZL_E_corruption_detected;   // zstd.c:44066
        // literals: final bytes.  Long runs (literal-heavy blocks: few sequences, tens of kilobytes of literals) go segment by
        // segment through the vectorised copy; short ones through the flat byte-parallel loop
        if (totalL >= 1024) {
            for (u32 k = 0; k < 32; k++) {
                const u32 kll = __shfl_sync(ZL_FULL, ll, k);
                if (!kll) continue;
                const u32 kDst = __shfl_sync(ZL_FULL, dstLit, k), kEx = __shfl_sync(ZL_FULL, litExcl, k);
                if (litMode == 1) zl_warp_fill(out + kDst, rleByte, kll, lane);
                else zl_warp_copy(out + kDst, lit + litPos + kEx, kll, lane);
                zl_mark_done(parent + kDst, kll, lane);
            }
        } else
        for (u32 j0 = 0; j0 < totalL; j0 += 32) {
            const u32 j = j0 + lane;
            const u32 k = zl_flat_owner(sl, j);
            const u32 kDst = __shfl_sync(ZL_FULL, dstLit, k), kEx = __shfl_sync(ZL_FULL, litExcl, k);
            if (j < totalL) {
                const u32 p = kDst + (j - kEx);
                out[p] = litMode == 1 ? (u8)rleByte : __ldg(lit + litPos + j);
                parent[p] = ZL_PAR_DONE;
            }
        }
        // matches: one parent pointer per byte (dictionary bytes are final at once)
        u32 rs = ml;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { const u32 a = __shfl_up_sync(ZL_FULL, rs, dd); if ((int)lane >= dd) rs += a; }
        const u32 totalM = __shfl_sync(ZL_FULL, rs, 31);
        const u32 rsExcl = rs - ml;
        for (u32 j0 = 0; j0 < totalM; j0 += 32) {
            const u32 j = j0 + lane;
            const u32 k = zl_flat_owner(rs, j);
            const u32 kDm = __shfl_sync(ZL_FULL, dm, k), kEx = __shfl_sync(ZL_FULL, rsExcl, k), kOff = __shfl_sync(ZL_FULL, off, k);
            if (j < totalM) {
                const u32 p = kDm + (j - kEx);
                if (kDict && kOff > p) { out[p] = dict[dictSize - (kOff - p)]; parent[p] = ZL_PAR_DONE; }
                else parent[p] = p - kOff;
            }
        }
        outPos += totalO; litPos += totalL;
    }
    if (!lastChunk) return 0;
    // last literals (zstd.c:44692-44698)
    const u32 lastLL = litSize - litPos;
    if ((outPos - op) + lastLL != B.regen) return ZL_E_corruption_detected;
    if (litMode == 1) zl_warp_fill(out + outPos, rleByte, lastLL, lane);
    else zl_warp_copy(out + outPos, lit + litPos, lastLL, lane);
    zl_mark_done(parent + outPos, lastLL, lane);
    return 0;
}

#endif  // __CUDACC__
