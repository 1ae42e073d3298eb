// zl_dec_exec.cuh -- sequence execution stage of the B200 Zstandard decoder (kernel K2) and the
// XXH64 content-checksum kernel (K3).
//
// K2: one warp per frame.  Sequence records produced by the entropy kernel are consumed 32 at a
// time (one per lane).  A warp prefix sum gives every lane its literal source, literal
// destination and match destination; literals of the whole batch are copied with a flat
// byte-parallel loop, then matches are resolved in rounds: a match may run as soon as every byte
// of its source lies below the destination of the first still-pending match of the batch
// ("high-water mark"), which is always true for the first pending match itself.  Overlapping
// matches (offset < matchLength) are expanded with the periodic formula dst[k] = src[k mod offset],
// so they never read bytes they write.  This restates ZSTD_execSequence / ZSTD_overlapCopy8 /
// ZSTD_wildcopy (zstd.c:44013, 43816, 15567) for a SIMT machine; the bounds checks of
// zstd.c:44024-44066 were already applied by the entropy kernel when it produced the records.
#pragma once
#include "zl_common.cuh"
#include "zl_dec_entropy.cuh"     // record forms, zl_rep_resolve

#if defined(__CUDACC__)

#define ZL_FULL 0xFFFFFFFFu

ZL_D void zl_warp_copy(u8* dst, const u8* src, u32 n, u32 lane)
{
    // cooperative copy of non-overlapping ranges: 16 bytes per lane when source and destination are co-aligned, otherwise
    // aligned 4-byte stores fed by funnel-shifted aligned loads (every word touched holds at least one valid byte)
    if (n < 64) { for (u32 i = lane; i < n; i += 32) dst[i] = src[i]; return; }
    if (((((size_t)dst) ^ ((size_t)src)) & 15) == 0) {
        u32 head = (u32)((16 - (((size_t)dst) & 15)) & 15);
        if (lane < head) dst[lane] = src[lane];
        u32 body = (n - head) >> 4;
        const uint4* s4 = (const uint4*)(src + head);
        uint4* d4 = (uint4*)(dst + head);
        u32 i = lane;
        for (; i + 96 < body; i += 128) {                 // four independent loads in flight per lane
            const uint4 a = s4[i], b = s4[i + 32], c = s4[i + 64], e = s4[i + 96];
            d4[i] = a; d4[i + 32] = b; d4[i + 64] = c; d4[i + 96] = e;
        }
        for (; i < body; i += 32) d4[i] = s4[i];
        u32 done = head + (body << 4);
        for (u32 i = done + lane; i < n; i += 32) dst[i] = src[i];
        return;
    }
    const u32 head = (u32)((4 - (((size_t)dst) & 3)) & 3);
    if (lane < head) dst[lane] = src[lane];
    const u8* s = src + head;
    u32* d4 = (u32*)(dst + head);
    const u32 words = (n - head) >> 2;
    const u32 sh = (u32)(((size_t)s) & 3) * 8;
    const u32* s4 = (const u32*)(((size_t)s) & ~(size_t)3);
    if (sh == 0) for (u32 i = lane; i < words; i += 32) d4[i] = s4[i];
    else {
        u32 i = lane;
        for (; i + 96 < words; i += 128) {
            const u32 a0 = s4[i], a1 = s4[i + 1], b0 = s4[i + 32], b1 = s4[i + 33], c0 = s4[i + 64], c1 = s4[i + 65], e0 = s4[i + 96], e1 = s4[i + 97];
            d4[i] = __funnelshift_r(a0, a1, sh); d4[i + 32] = __funnelshift_r(b0, b1, sh); d4[i + 64] = __funnelshift_r(c0, c1, sh); d4[i + 96] = __funnelshift_r(e0, e1, sh);
        }
        for (; i < words; i += 32) d4[i] = __funnelshift_r(s4[i], s4[i + 1], sh);
    }
    const u32 done = head + (words << 2);
    if (done + lane < n) dst[done + lane] = src[done + lane];
}

ZL_D void zl_warp_fill(u8* dst, u32 byte, u32 n, u32 lane)
{
    u32 head = (u32)((4 - (((size_t)dst) & 3)) & 3);
    if (head > n) head = n;
    if (lane < head) dst[lane] = (u8)byte;
    u32 body = (n - head) >> 2;
    u32 w = byte * 0x01010101u;
    u32* d4 = (u32*)(dst + head);
    for (u32 i = lane; i < body; i += 32) d4[i] = w;
    u32 done = head + (body << 2);
    for (u32 i = done + lane; i < n; i += 32) dst[i] = (u8)byte;
}

// ---- repeat-offset history as a composable transform ------------------------------------------------------------------
// A record maps the history (h0, h1, h2) to a new one.  Except for the rare "rep0 - 1" code, every output slot is either
// one of the three input slots or the fresh offset of some record of the batch, so a transform is three bytes:
// 0..2 = input slot, 0x80 | lane = the offset carried by that lane.  Two transforms compose with a single byte permute
// (PRMT), and a warp scan over the 32 records of a batch gives every lane the history that precedes its record
// (zstd.c:44290-44326 restated for 32 lanes); actual offsets are fetched from the named lane with one shuffle.
#define ZL_REPT_ID 0x00020100u
ZL_D u32 zl_rept_compose(u32 A, u32 B)          // A first, then B
{
    const u32 m = (B >> 7) & 0x00010101u;                               // 1 in the bytes of B that name a lane
    const u32 mm = m * 0xFFu;
    const u32 r = (B & 0x00030303u & ~mm) | (0x00060504u & mm);         // per byte: input slot of A, or 4 + j (keep B's byte)
    const u32 sel = (r & 0xFu) | ((r >> 4) & 0xF0u) | ((r >> 8) & 0xF00u) | 0x3000u;
    return __byte_perm(A, B, sel) & 0x00FFFFFFu;
}

// ---- pieces shared by the warp-per-frame execute kernel and the large-frame kernels -----------------------------------------
// one record per lane -> (litLength, matchLength, offBase); lanes without a record get zeros
ZL_D void zl_lane_record(u64 rec, bool valid, const u32* xtab, u32& ll, u32& ml, u32& ob)
{
    ll = 0; ml = 0; ob = 0;
    if (!valid) return;
    if (rec & ZL_REC_B) { ll = (u32)rec & 0xFFFFu; ml = ((u32)rec >> 16) & 0xFFFFu; ob = (u32)(rec >> 32) & 0x7FFFFFFFu; return; }
    const u32 snap = (u32)rec, c = (u32)(rec >> 32);
    const u32 llCode = c & 63u, mlCode = (c >> 6) & 63u, aOF = (c >> 12) & 31u;
    const u32 xl = xtab[llCode], xm = xtab[36 + mlCode];
    const u32 aLL = xl >> 24, aML = xm >> 24;
    ob = (1u << aOF) + zl_shr(snap, 32u - aOF);
    ml = (xm & 0xFFFFFFu) + zl_shr(zl_shl(snap, aOF), 32u - aML);
    ll = (xl & 0xFFFFFFu) + zl_shr(zl_shl(snap, aOF + aML), 32u - aLL);
}
// history slot (0..3, 3 = "rep0 - 1") a record reads; 0 also for records that do not read the history
ZL_D u32 zl_rep_idx(u32 ll, u32 ml, u32 ob) { return (ml != 0 && ob >= 1 && ob <= 3) ? ob - 1 + (ll == 0 ? 1u : 0u) : 0u; }
// the transform of one record in the byte form (never called for idx == 3)
ZL_D u32 zl_rept_of(bool isNew, u32 idx, u32 tag)       // tag < 128: names the record that carries the fresh offset
{
    if (isNew) return 0x00010000u | 0x80u | tag;                                 // (mine, h0, h1)
    if (idx == 1) return 0x00020001u;                                             // (h1, h0, h2)
    if (idx == 2) return 0x00010002u;                                             // (h2, h0, h1)
    return ZL_REPT_ID;
}
ZL_D u32 zl_rept_scan(u32 T, u32 lane)                                            // inclusive warp scan by composition
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 A = __shfl_up_sync(ZL_FULL, T, d);
        if ((int)lane >= d) T = zl_rept_compose(A, T);
    }
    return T;
}
// Offsets of the 32 records of a batch and the history after it (zstd.c:44290-44326); h0..h2 in/out.
ZL_D u32 zl_batch_offsets(u32 ll, u32 ml, u32 ob, u32 lane, u32& h0, u32& h1, u32& h2)
{
    const bool isM = ml != 0, isNew = isM && ob >= 4;
    const u32 idx = zl_rep_idx(ll, ml, ob);
    u32 off;
    if (__ballot_sync(ZL_FULL, idx == 3)) {
        // rare "rep0 - 1" code somewhere in the batch: resolve the 32 records one after the other (uniform loop)
        off = 0;
        u32 hh[3] = {h0, h1, h2};
        for (u32 l = 0; l < 32; l++) {
            const u32 lll = __shfl_sync(ZL_FULL, ll, l), lml = __shfl_sync(ZL_FULL, ml, l), lob = __shfl_sync(ZL_FULL, ob, l);
            const u32 o = zl_rep_resolve(hh, lll, lml, lob);
            if (l == lane) off = o;
        }
        h0 = hh[0]; h1 = hh[1]; h2 = hh[2];
        return off;
    }
    const u32 fresh = ob - 3;                                               // meaningful on isNew lanes
    const u32 T = zl_rept_scan(zl_rept_of(isNew, idx, lane), lane);
    u32 E = __shfl_up_sync(ZL_FULL, T, 1);                                  // exclusive prefix: history before this record
    if (lane == 0) E = ZL_REPT_ID;
    // the slot of the incoming history this record reads (repeat codes and continuations): byte idx of E
    const u32 eb = (E >> (8 * idx)) & 0xFFu;
    const u32 fromLane = __shfl_sync(ZL_FULL, fresh, eb & 31u);
    const u32 fromHist = (eb & 3u) == 0 ? h0 : ((eb & 3u) == 1 ? h1 : h2);
    off = isNew ? fresh : ((eb & 0x80u) ? fromLane : fromHist);
    // history after the batch: the inclusive prefix of lane 31
    const u32 Lt = __shfl_sync(ZL_FULL, T, 31);
    const u32 b0 = Lt & 0xFFu, b1 = (Lt >> 8) & 0xFFu, b2 = (Lt >> 16) & 0xFFu;
    const u32 f0 = __shfl_sync(ZL_FULL, fresh, b0 & 31u), f1 = __shfl_sync(ZL_FULL, fresh, b1 & 31u), f2 = __shfl_sync(ZL_FULL, fresh, b2 & 31u);
    const u32 n0 = (b0 & 0x80u) ? f0 : ((b0 & 3u) == 0 ? h0 : ((b0 & 3u) == 1 ? h1 : h2));
    const u32 n1 = (b1 & 0x80u) ? f1 : ((b1 & 3u) == 0 ? h0 : ((b1 & 3u) == 1 ? h1 : h2));
    const u32 n2 = (b2 & 0x80u) ? f2 : ((b2 & 3u) == 0 ? h0 : ((b2 & 3u) == 1 ? h1 : h2));
    h0 = n0; h1 = n1; h2 = n2;
    return off;
}
// inclusive warp scans of ll and ll + ml
ZL_D void zl_batch_positions(u32 ll, u32 ml, u32 lane, u32& sl, u32& so, u32& totalL, u32& totalO)
{
    sl = ll; so = ll + ml;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 a = __shfl_up_sync(ZL_FULL, sl, d), b = __shfl_up_sync(ZL_FULL, so, d);
        if ((int)lane >= d) { sl += a; so += b; }
    }
    totalL = __shfl_sync(ZL_FULL, sl, 31); totalO = __shfl_sync(ZL_FULL, so, 31);
}
// owner of flat index j: first lane whose inclusive sum `rs` exceeds j
ZL_D u32 zl_flat_owner(u32 rs, u32 j)
{
    u32 k = 0;
#pragma unroll
    for (int st = 16; st >= 1; st >>= 1) {
        const u32 v = __shfl_sync(ZL_FULL, rs, (k + st - 1) & 31);
        if (v <= j) k += st;
    }
    return k & 31;
}

// ---- one-pass execute ---------------------------------------------------------------------------------------------------
// A batch is 128 records, four consecutive ones per lane.  Its (up to 256) non-empty segments -- literal run, match, literal
// run, match ... -- tile the batch's output bytes in order, so the warp writes that output ONCE, front to back, 64 bytes (two
// 32-byte rows) per step; rows are aligned to 32 bytes of the destination address, so a row is one full sector.
//   planning   record -> (litLength, matchLength, offBase) per lane; one warp scan over the lane sums gives every record its
//              place in the output and in the literal stream; the repeat-offset history (zstd.c:44290-44326) is one warp scan over
//              the composed transforms of the lanes' four records, fresh offsets are fetched through shared memory;
//              the bounds checks of ZSTD_execSequence (zstd.c:44024-44066) for the whole batch, before anything is written
//   tables     `seg`: the compacted segment table (source address of flat index 0, flat start, match offset or the literal
//              flag); `row`: for every 32-byte row of a 4 KiB window of the output the bit mask of the segments starting in it
//              (shared-memory atomics) and the number of segments before it (a scan over the rows)
//   steps      owner of an output byte = segments before its row + popcount of the row's start bits up to its own bit.  The steps
//              run in output order, so every byte below the current step is final when the step loads its sources; only sources
//              INSIDE the step (offset < 64) are not in memory yet -- those bytes are resolved between the lanes with shuffles
//              (value + "resolved" flag of both rows packed in one register; a source always lies earlier in the step, so the
//              lowest unresolved byte resolves every iteration; overlapping matches first fold their source below their own
//              start with the periodic form src = start - offset + (k mod offset), zstd.c:43816 ZSTD_overlapCopy8).  Steps that
//              lie inside one segment (no start bit in either row: long literal runs, long or run-length matches) take a
//              uniform path without the owner lookup.
// Round 1 handled 32 records at a time, copied their literals in one pass and resolved the matches in rounds against a
// high-water mark (2.4 - 4.4 rounds per batch, each with its own scan and segment table): ~900 warp instructions per 32 sequences.
#define ZL_SEG_LIT 0x80000000u
#ifndef ZL_XB
#define ZL_XB 4                          // records per lane (2 or 4)
#endif
#define ZL_XBATCH (32 * ZL_XB)           // records per batch
#define ZL_XROWS (32 * ZL_XB)            // rows per window of the row table (4 KiB of output at 4 records per lane)
struct ZlExecSm {                        // per warp: 5 KB
    uint4 seg[2 * ZL_XBATCH];            // x, y = P: address of the segment's source byte for flat index 0 (source of byte j = P + j);
                                         // z = flat start; w = match offset, or ZL_SEG_LIT for a literal run
    union { u32 fresh[ZL_XBATCH]; uint2 row[ZL_XROWS]; } u;      // fresh offsets while planning / row table while stepping
};
ZL_D uint2 zl_lds64(u32 a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
ZL_D uint4 zl_lds128(u32 a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
ZL_D u32 zl_lds32s(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
// k mod d for d < 64 with the reciprocal table rcp[d] = floor((2^32 - 1) / d) in shared memory (k < 2^18: the estimate is at most one short)
ZL_D u32 zl_mod_small(u32 k, u32 d, u32 rcpA)
{
    const u32 q = __umulhi(k, zl_lds32s(rcpA + (d << 2)));
    u32 r = k - q * d;
    if (r >= d) r -= d;
    return r;
}
ZL_D const u8* zl_ptr_add(u32 lo, u32 hi, u32 j) { return reinterpret_cast<const u8*>((((u64)hi << 32) | lo) + j); }

// Offsets of the 128 records of a batch (record m of lane l carries the tag 4 l + m) and the history after it; h0..h2 in/out.
ZL_D void zl_batch_offsets4(const u32 (&ll)[ZL_XB], const u32 (&ml)[ZL_XB], const u32 (&ob)[ZL_XB], u32 lane, u32* fresh,
                            u32& h0, u32& h1, u32& h2, u32 (&off)[ZL_XB])
{
    u32 idx[ZL_XB], t[ZL_XB];
    bool minus1 = false;
#pragma unroll
    for (int m = 0; m < ZL_XB; m++) {
        idx[m] = zl_rep_idx(ll[m], ml[m], ob[m]);
        minus1 |= idx[m] == 3;
        t[m] = zl_rept_of(ml[m] != 0 && ob[m] >= 4, idx[m], ZL_XB * lane + (u32)m);
    }
    if (__any_sync(ZL_FULL, minus1)) {
        // rare "rep0 - 1" code somewhere in the batch: resolve the records one after the other (uniform loop)
        u32 hh[3] = {h0, h1, h2};
        for (u32 l = 0; l < 32; l++) {
#pragma unroll
            for (int m = 0; m < ZL_XB; m++) {
                const u32 lll = __shfl_sync(ZL_FULL, ll[m], l), lml = __shfl_sync(ZL_FULL, ml[m], l), lob = __shfl_sync(ZL_FULL, ob[m], l);
                const u32 o = zl_rep_resolve(hh, lll, lml, lob);
                if (l == lane) off[m] = o;
            }
        }
        h0 = hh[0]; h1 = hh[1]; h2 = hh[2];
        return;
    }
#if ZL_XB == 4
    *reinterpret_cast<uint4*>(fresh + 4u * lane) = make_uint4(ob[0] - 3u, ob[1] - 3u, ob[2] - 3u, ob[3] - 3u);
    const u32 t01 = zl_rept_compose(t[0], t[1]), t012 = zl_rept_compose(t01, t[2]);
    const u32 T = zl_rept_scan(zl_rept_compose(t012, t[3]), lane);
#else
    *reinterpret_cast<uint2*>(fresh + 2u * lane) = make_uint2(ob[0] - 3u, ob[1] - 3u);
    const u32 t01 = zl_rept_compose(t[0], t[1]);
    const u32 T = zl_rept_scan(t01, lane);
#endif
    u32 E = __shfl_up_sync(ZL_FULL, T, 1);                                  // exclusive prefix: history before this lane's records
    if (lane == 0) E = ZL_REPT_ID;
    __syncwarp();
#if ZL_XB == 4
    const u32 e[ZL_XB] = {E, zl_rept_compose(E, t[0]), zl_rept_compose(E, t01), zl_rept_compose(E, t012)};
#else
    const u32 e[ZL_XB] = {E, zl_rept_compose(E, t[0])};
#endif
#pragma unroll
    for (int m = 0; m < ZL_XB; m++) {
        const u32 eb = (e[m] >> (8u * idx[m])) & 0xFFu;                     // the slot of the incoming history this record reads
        const u32 fromTag = fresh[eb & (ZL_XBATCH - 1)];
        const u32 fromHist = (eb & 3u) == 0 ? h0 : ((eb & 3u) == 1 ? h1 : h2);
        off[m] = (ml[m] != 0 && ob[m] >= 4) ? ob[m] - 3u : ((eb & 0x80u) ? fromTag : fromHist);
    }
    const u32 Lt = __shfl_sync(ZL_FULL, T, 31);                            // history after the batch
    const u32 b0 = Lt & 0xFFu, b1 = (Lt >> 8) & 0xFFu, b2 = (Lt >> 16) & 0xFFu;
    const u32 f0 = fresh[b0 & (ZL_XBATCH - 1)], f1 = fresh[b1 & (ZL_XBATCH - 1)], f2 = fresh[b2 & (ZL_XBATCH - 1)];
    const u32 n0 = (b0 & 0x80u) ? f0 : ((b0 & 3u) == 0 ? h0 : ((b0 & 3u) == 1 ? h1 : h2));
    const u32 n1 = (b1 & 0x80u) ? f1 : ((b1 & 3u) == 0 ? h0 : ((b1 & 3u) == 1 ? h1 : h2));
    const u32 n2 = (b2 & 0x80u) ? f2 : ((b2 & 3u) == 0 ? h0 : ((b2 & 3u) == 1 ? h1 : h2));
    h0 = n0; h1 = n1; h2 = n2;
}

template <bool kDict>
ZL_D u32 zl_exec_block(u8* out, u32 op, u32 cap, u32 capErr, const ZlBlockHdr& h, const u8* __restrict__ lit, u32 rleByte, u32 litMode,
                       const u64* __restrict__ recs, const u8* __restrict__ dict, u32 dictSize, u32 (&hist)[3], const u32* xtab, const u32* rcp,
                       ZlExecSm* sm, u32 lane, u32& regen)
{
    // `recs` is 16-byte aligned (the index kernel hands out even record offsets); rcp: the 64 reciprocals of zl_mod_small
    const u32 nrec = h.nrec, litSize = h.litSize;
    const u32 leMask = 0xFFFFFFFFu >> (31u - lane);
    const u32 segA = (u32)__cvta_generic_to_shared(sm->seg), rowA = (u32)__cvta_generic_to_shared(sm->u.row), rcpA = (u32)__cvta_generic_to_shared(rcp);
    const bool rleLits = litMode == 1;
    u32 outPos = op, litPos = 0;
    u32 h0 = hist[0], h1 = hist[1], h2 = hist[2];
    // output byte at frame position sp (negative: dictionary)
    auto outPtr = [&](i32 sp) -> const u8* { return (kDict && sp < 0) ? dict + ((i32)dictSize + sp) : out + sp; };
    for (u32 base = 0; base < nrec; base += ZL_XBATCH) {
        const u32 i0 = base + ZL_XB * lane;
        ulonglong2 ra = make_ulonglong2(0ull, 0ull), rb = ra;
        if (i0 < nrec) ra = __ldcs(reinterpret_cast<const ulonglong2*>(recs + i0));
        if (ZL_XB == 4 && i0 + 2 < nrec) rb = __ldcs(reinterpret_cast<const ulonglong2*>(recs + i0 + 2));
        if (i0 + ZL_XBATCH < nrec) asm volatile("prefetch.global.L1 [%0];" ::"l"(recs + i0 + ZL_XBATCH));
        u32 ll[ZL_XB], ml[ZL_XB], ob[ZL_XB], off[ZL_XB];
        zl_lane_record(ra.x, i0 < nrec, xtab, ll[0], ml[0], ob[0]);
        zl_lane_record(ra.y, i0 + 1 < nrec, xtab, ll[1], ml[1], ob[1]);
#if ZL_XB == 4
        zl_lane_record(rb.x, i0 + 2 < nrec, xtab, ll[2], ml[2], ob[2]);
        zl_lane_record(rb.y, i0 + 3 < nrec, xtab, ll[3], ml[3], ob[3]);
#endif
        // ---- places: one scan over (literal bytes | non-empty segments << 23, output bytes) of the lanes
        u32 L = 0, O = 0, nseg = 0;                           // L < 2^23: lengths are < 65536
#pragma unroll
        for (int m = 0; m < ZL_XB; m++) { L += ll[m]; O += ll[m] + ml[m]; nseg += (ll[m] ? 1u : 0u) + (ml[m] ? 1u : 0u); }
        u32 sl = L | (nseg << 23), so = O;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 a = __shfl_up_sync(ZL_FULL, sl, d), b = __shfl_up_sync(ZL_FULL, so, d);
            if ((int)lane >= d) { sl += a; so += b; }
        }
        const u32 tot = __shfl_sync(ZL_FULL, sl, 31), totalO = __shfl_sync(ZL_FULL, so, 31);
        const u32 totalL = tot & 0x7FFFFFu;
        // ---- the checks of ZSTD_execSequence (zstd.c:44024-44066, 44320) for the whole batch, before anything is written
        if (totalL > litSize - litPos) return ZL_E_corruption_detected;
        if ((outPos - op) + totalO > cap) return capErr;
        zl_batch_offsets4(ll, ml, ob, lane, sm->u.fresh, h0, h1, h2, off);
        // ---- compacted segment table
        const u32 ord0 = (sl >> 23) - nseg;                       // segments of earlier lanes
        u8* const ob8 = out + outPos;                        // flat index 0 of the batch
        {
            u32 f = so - O, li = litPos + (sl & 0x7FFFFFu) - L, ord = ord0;
            bool bad = false;
#pragma unroll
            for (int m = 0; m < ZL_XB; m++) {
                if (ll[m]) { const u64 P = (u64)(size_t)lit + li - f; sm->seg[ord++] = make_uint4((u32)P, (u32)(P >> 32), f, ZL_SEG_LIT); }
                f += ll[m]; li += ll[m];
                bad |= ml[m] != 0 && (off[m] == 0 || off[m] > outPos + f + dictSize);
                if (ml[m]) { const u64 P = (u64)(size_t)ob8 - off[m]; sm->seg[ord++] = make_uint4((u32)P, (u32)(P >> 32), f, off[m]); }
                f += ml[m];
            }
            if (__any_sync(ZL_FULL, bad)) return ZL_E_corruption_detected;
        }
        // rows are aligned to 32 bytes of the destination: aligned flat index a = flat index + mis
        const u32 mis = (u32)((size_t)ob8 & 31u);
        const u32 totalA = totalO + mis;
        u32 segsBefore = 0;                                  // segments that start in earlier windows
        for (u32 w0 = 0; w0 < totalA; w0 += ZL_XROWS * 32) {
            // ---- row table of the window [w0, w0 + 4 KiB): start masks by atomics, then the exclusive count over the rows
            __syncwarp();
#if ZL_XB == 4
            reinterpret_cast<uint4*>(sm->u.row)[2 * lane] = make_uint4(0u, 0u, 0u, 0u);
            reinterpret_cast<uint4*>(sm->u.row)[2 * lane + 1] = make_uint4(0u, 0u, 0u, 0u);
#else
            reinterpret_cast<uint4*>(sm->u.row)[lane] = make_uint4(0u, 0u, 0u, 0u);
#endif
            __syncwarp();
#pragma unroll
            for (u32 k = 0; k < 2 * ZL_XB; k++) {
                const u32 a = zl_lds32s(segA + (((ord0 + k) & (2 * ZL_XBATCH - 1)) << 4) + 8u) + mis - w0;
                if (k < nseg && a < ZL_XROWS * 32) atomicOr(&sm->u.row[a >> 5].x, 1u << (a & 31u));
            }
            __syncwarp();
            {
#if ZL_XB == 4
                uint4 qa = zl_lds128(rowA + lane * 32u), qb = zl_lds128(rowA + lane * 32u + 16u);      // rows 4 lane .. 4 lane + 3
                const u32 c0 = __popc(qa.x), c1 = __popc(qa.z), c2 = __popc(qb.x), c3 = __popc(qb.z);
#else
                uint4 qa = zl_lds128(rowA + lane * 16u);                                                  // rows 2 lane, 2 lane + 1
                const u32 c0 = __popc(qa.x), c1 = __popc(qa.z), c2 = 0, c3 = 0;
#endif
                u32 inc = c0 + c1 + c2 + c3;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const u32 a = __shfl_up_sync(ZL_FULL, inc, d); if ((int)lane >= d) inc += a; }
                const u32 ex = segsBefore + inc - (c0 + c1 + c2 + c3);
                qa.y = ex; qa.w = ex + c0;
#if ZL_XB == 4
                qb.y = ex + c0 + c1; qb.w = ex + c0 + c1 + c2;
                reinterpret_cast<uint4*>(sm->u.row)[2 * lane] = qa;
                reinterpret_cast<uint4*>(sm->u.row)[2 * lane + 1] = qb;
#else
                reinterpret_cast<uint4*>(sm->u.row)[lane] = qa;
#endif
                segsBefore += __shfl_sync(ZL_FULL, inc, 31);
            }
            __syncwarp();
            const u32 wEnd = min(totalA, w0 + ZL_XROWS * 32);
            for (u32 a0 = w0; a0 < wEnd; a0 += 64) {
                const uint4 ri = zl_lds128(rowA + ((a0 - w0) >> 5) * 8u);    // {mask, segments before} of the step's two rows
                const u32 ja = a0 + lane - mis, jb = ja + 32;                  // flat indices of this lane's two bytes (first step: ja may wrap below 0)
                const u32 d0 = a0 < mis ? lane - (mis - a0) : lane;            // distance of byte a from the first flat index of the step
                if ((ri.x | ri.z) == 0u && a0 + 64 <= totalA) {
                    // ---- the whole step lies inside one segment, which began before it: no owner lookup, no in-step sources
                    const uint4 s = zl_lds128(segA + (((ri.y - 1u) & (2 * ZL_XBATCH - 1)) << 4));
                    u32 va = rleByte, vb = rleByte;
                    if (s.w < 64u) {                                           // sources inside the match itself: fold below its start
                        u32 ta = ja - s.w, tb = jb - s.w;
                        if ((i32)ta >= (i32)s.z) ta = s.z - s.w + zl_mod_small(ja - s.z, s.w, rcpA);
                        if ((i32)tb >= (i32)s.z) tb = s.z - s.w + zl_mod_small(jb - s.z, s.w, rcpA);
                        va = *outPtr((i32)(outPos + ta)); vb = *outPtr((i32)(outPos + tb));
                    } else if (!(s.w == ZL_SEG_LIT && rleLits)) {
                        const u8* pa = zl_ptr_add(s.x, s.y, ja);
                        const u8* pb = zl_ptr_add(s.x, s.y, jb);
                        if (kDict && s.w != ZL_SEG_LIT) { pa = outPtr((i32)(outPos + ja - s.w)); pb = outPtr((i32)(outPos + jb - s.w)); }
                        va = *pa; vb = *pb;
                    }
                    ob8[ja] = (u8)va; ob8[jb] = (u8)vb;
                    __syncwarp();
                    continue;
                }
                const u32 c0 = ri.y + __popc(ri.x & leMask) - 1u, c1 = ri.w + __popc(ri.z & leMask) - 1u;
                const uint4 sa = zl_lds128(segA + ((c0 & (2 * ZL_XBATCH - 1)) << 4)), sb = zl_lds128(segA + ((c1 & (2 * ZL_XBATCH - 1)) << 4));
                const bool aa = ja < totalO, ab = jb < totalO;
                bool na = aa && sa.w <= d0, nb = ab && sb.w <= d0 + 32u;       // source inside this step (offset <= distance from its start): not in memory yet
                const u8* pa = zl_ptr_add(sa.x, sa.y, ja);
                const u8* pb = zl_ptr_add(sb.x, sb.y, jb);
                if (kDict) {
                    if (sa.w != ZL_SEG_LIT) pa = outPtr((i32)(outPos + ja - sa.w));
                    if (sb.w != ZL_SEG_LIT) pb = outPtr((i32)(outPos + jb - sb.w));
                }
                u32 va = rleByte, vb = rleByte;
                if (aa && !na && !(sa.w == ZL_SEG_LIT && rleLits)) va = *pa;
                if (ab && !nb && !(sb.w == ZL_SEG_LIT && rleLits)) vb = *pb;
                if (__any_sync(ZL_FULL, na || nb)) {
                    const u32 j0 = a0 < mis ? 0u : a0 - mis;                   // first flat index of the step
                    u32 ta = ja - sa.w, tb = jb - sb.w;                        // flat source index (>= j0 where na / nb)
                    // overlapping matches: fold the source below the match's own start (it may then lie below the step)
                    if (na && ta >= sa.z) { ta = sa.z - sa.w + zl_mod_small(ja - sa.z, sa.w, rcpA); if ((i32)ta < (i32)j0) { na = false; va = *outPtr((i32)(outPos + ta)); } }
                    if (nb && tb >= sb.z) { tb = sb.z - sb.w + zl_mod_small(jb - sb.z, sb.w, rcpA); if ((i32)tb < (i32)j0) { nb = false; vb = *outPtr((i32)(outPos + tb)); } }
                    const u32 ra_ = ta + mis - a0, rb_ = tb + mis - a0;          // position of the source inside the step (0..63)
                    while (__any_sync(ZL_FULL, na || nb)) {
                        const u32 w = (va & 0xFFu) | (na ? 0u : 0x100u) | ((vb & 0xFFu) << 16) | (nb ? 0u : 0x1000000u);
                        const u32 wa = __shfl_sync(ZL_FULL, w, ra_ & 31u) >> ((ra_ & 32u) >> 1), wb = __shfl_sync(ZL_FULL, w, rb_ & 31u) >> ((rb_ & 32u) >> 1);
                        if (na && (wa & 0x100u)) { va = wa & 0xFFu; na = false; }
                        if (nb && (wb & 0x100u)) { vb = wb & 0xFFu; nb = false; }
                    }
                }
                if (aa) ob8[ja] = (u8)va;
                if (ab) ob8[jb] = (u8)vb;
                __syncwarp();
            }
        }
        outPos += totalO; litPos += totalL;
    }
    // last literals (zstd.c:44692-44698)
    const u32 lastLL = litSize - litPos;
    if ((outPos - op) + lastLL > cap) return capErr;
    if (litMode == 1) zl_warp_fill(out + outPos, rleByte, lastLL, lane);
    else zl_warp_copy(out + outPos, lit + litPos, lastLL, lane);
    hist[0] = h0; hist[1] = h1; hist[2] = h2;
    regen = outPos - op + lastLL;
    return 0;
}

// ---- XXH64 (zstd.c:11509-11664), one quad per buffer: lane a of the quad owns accumulator a -----------
#define ZL_P1 0x9E3779B185EBCA87ULL
#define ZL_P2 0xC2B2AE3D27D4EB4FULL
#define ZL_P3 0x165667B19E3779F9ULL
#define ZL_P4 0x85EBCA77C2B2AE63ULL
#define ZL_P5 0x27D4EB2F165667C5ULL
ZL_D u64 zl_rotl64(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
ZL_D u64 zl_xround(u64 acc, u64 in) { return zl_rotl64(acc + in * ZL_P2, 31) * ZL_P1; }
ZL_D u64 zl_xmerge(u64 h, u64 v) { return (h ^ zl_xround(0, v)) * ZL_P1 + ZL_P4; }
ZL_D u64 zl_ld64u(const u8* p)
{
    if ((((size_t)p) & 7) == 0) return *(const u64*)p;
    if ((((size_t)p) & 3) == 0) return (u64)(*(const u32*)p) | ((u64)(*(const u32*)(p + 4)) << 32);
    return zl_rd64(p);
}
// all 4 lanes of the quad call this; result valid on every lane of the quad.  `v` / `s0`: accumulator of lane q after the first
// s0 stripes (zl_xxh_seed(q), 0 to start from the beginning)
ZL_D u64 zl_xxh_seed(u32 q) { return q == 0 ? (ZL_P1 + ZL_P2) : (q == 1 ? ZL_P2 : (q == 2 ? 0ull : (0ull - ZL_P1))); }
ZL_D u64 zl_quad_xxh64_from(const u8* p, u32 len, u64 v, u32 s0, u32 q, u32 qmask, u32 qbase)
{
    u64 h;
    u32 done = 0;
    if (len >= 32) {
        const u32 stripes = len >> 5;
        const u8* pp = p + 8 * q;
        for (u32 s = s0; s < stripes; s++) v = zl_xround(v, zl_ld64u(pp + 32 * s));
        const u64 v1 = __shfl_sync(qmask, v, qbase + 0), v2 = __shfl_sync(qmask, v, qbase + 1);
        const u64 v3 = __shfl_sync(qmask, v, qbase + 2), v4 = __shfl_sync(qmask, v, qbase + 3);
        h = zl_rotl64(v1, 1) + zl_rotl64(v2, 7) + zl_rotl64(v3, 12) + zl_rotl64(v4, 18);
        h = zl_xmerge(h, v1); h = zl_xmerge(h, v2); h = zl_xmerge(h, v3); h = zl_xmerge(h, v4);
        done = stripes << 5;
    } else h = ZL_P5;
    h += (u64)len;
    const u8* t = p + done; u32 r = len - done;
    while (r >= 8) { h ^= zl_xround(0, zl_rd64(t)); h = zl_rotl64(h, 27) * ZL_P1 + ZL_P4; t += 8; r -= 8; }
    if (r >= 4) { h ^= (u64)zl_rd32(t) * ZL_P1; h = zl_rotl64(h, 23) * ZL_P2 + ZL_P3; t += 4; r -= 4; }
    while (r) { h ^= (u64)(*t++) * ZL_P5; h = zl_rotl64(h, 11) * ZL_P1; r--; }
    h ^= h >> 33; h *= ZL_P2; h ^= h >> 29; h *= ZL_P3; h ^= h >> 32;
    return h;
}
ZL_D u64 zl_quad_xxh64(const u8* p, u32 len, u32 q, u32 qmask, u32 qbase) { return zl_quad_xxh64_from(p, len, zl_xxh_seed(q), 0, q, qmask, qbase); }

// One WARP per large buffer (16-byte aligned): the accumulator chains are serial (rotate and multiply do not commute) and run on
// lanes 0-3, but their inputs are not -- all 32 lanes stream the buffer through shared memory in 2 KiB chunks (64 stripes), the
// next chunk in flight in registers while the chains run over the current one, so the chains never wait for memory.
#define ZL_XXH_CHUNK 2048u
ZL_D u64 zl_warp_xxh64(const u8* p, u32 len, u8 (*buf)[ZL_XXH_CHUNK], u32 lane)
{
    const u32 nchunks = len / ZL_XXH_CHUNK;
    u64 v = zl_xxh_seed(lane & 3);
    if (nchunks) {
        uint4 r[4];
#pragma unroll
        for (int i = 0; i < 4; i++) r[i] = __ldg((const uint4*)p + i * 32 + lane);
        for (u32 c = 0; c < nchunks; c++) {
            uint4* b4 = (uint4*)buf[c & 1];
#pragma unroll
            for (int i = 0; i < 4; i++) b4[i * 32 + lane] = r[i];
            __syncwarp();
            if (c + 1 < nchunks) {
                const uint4* g = (const uint4*)(p + (size_t)(c + 1) * ZL_XXH_CHUNK);
#pragma unroll
                for (int i = 0; i < 4; i++) r[i] = __ldg(g + i * 32 + lane);
            }
            if (lane < 4) {
                const u64* b8 = (const u64*)buf[c & 1] + lane;
#pragma unroll 8
                for (u32 s = 0; s < ZL_XXH_CHUNK / 32; s++) v = zl_xround(v, b8[4 * s]);
            }
        }
        __syncwarp();
    }
    u64 h = 0;
    if (lane < 4) h = zl_quad_xxh64_from(p, len, v, nchunks * (ZL_XXH_CHUNK / 32), lane, 0xFu, 0);
    return __shfl_sync(ZL_FULL, h, 0);
}
#endif  // __CUDACC__
