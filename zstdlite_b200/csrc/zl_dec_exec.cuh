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
        for (u32 i = lane; i < body; i += 32) d4[i] = s4[i];
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
    else for (u32 i = lane; i < words; i += 32) d4[i] = __funnelshift_r(s4[i], s4[i + 1], sh);
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
ZL_D u32 zl_rept_of(bool isNew, u32 idx, u32 lane)
{
    if (isNew) return 0x00010000u | 0x80u | lane;                                 // (mine, h0, h1)
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

// Flat byte-parallel copy over the (up to 32) segments of a batch, two rows of 32 bytes per step, both loads of a step before
// its first store (4 rows and a higher occupancy were measured slower).  Lane k owns flat indices [start, start + len).  Owner of a flat index: the segments that hold bytes
// were compacted into `seg` (x = destination - flat start, y = source - flat start); per row one REDUX ORs the start bits of the
// segments beginning in it and a popcount up to the lane's own bit ranks the lane among them -- no per-byte search.
template <typename LoadF>
ZL_D void zl_flat_copy(u8* out, const uint2* seg, u32 start, u32 len, u32 total, u32 lane, u32 leMask, LoadF load)
{
    const u32 srow = start >> 5, sbit = len ? 1u << (start & 31u) : 0u;
    u32 cnt = 0;
    for (u32 j0 = 0; j0 < total; j0 += 64) {
        const u32 r = j0 >> 5;
        const u32 m0 = __reduce_or_sync(ZL_FULL, srow == r ? sbit : 0u);
        const u32 m1 = __reduce_or_sync(ZL_FULL, srow == r + 1 ? sbit : 0u);
        const u32 c0 = cnt + __popc(m0 & leMask) - 1u;
        cnt += __popc(m0);
        const u32 c1 = cnt + __popc(m1 & leMask) - 1u;
        cnt += __popc(m1);
        const u32 ja = j0 + lane, jb = ja + 32;
        const bool aa = ja < total, ab = jb < total;
        const uint2 sa = seg[c0 & 31u], sb = seg[c1 & 31u];
        u32 va = 0, vb = 0;
        if (aa) va = load(ja, sa.y);
        if (ab) vb = load(jb, sb.y);
        if (aa) out[ja + sa.x] = (u8)va;
        if (ab) out[jb + sb.x] = (u8)vb;
    }
}

// Execute one compressed block: `out` = frame output base, `op` = frame-relative position of the block, `cap` = bytes the
// block may still regenerate (destination room, at most one block size), `capErr` the error to report beyond it.
// hist[3] is the repeat-offset history carried from block to block.  Returns 0 and sets `regen`, or a ZlErr.
template <bool kDict>
ZL_D u32 zl_exec_block(u8* out, u32 op, u32 cap, u32 capErr, const ZlBlockHdr& h, const u8* __restrict__ lit, u32 rleByte, u32 litMode,
                       const u64* __restrict__ recs, const u8* __restrict__ dict, u32 dictSize, u32 (&hist)[3], const u32* xtab, uint2* seg, u32 lane, u32& regen)
{
    // `seg`: 32 x uint2 of shared memory owned by this warp -- the compacted segment table of the flat copies (below)
    const u32 nrec = h.nrec, litSize = h.litSize;
    const u32 ltMask = (1u << lane) - 1u, leMask = 0xFFFFFFFFu >> (31u - lane);
    u32 outPos = op, litPos = 0;
    u32 h0 = hist[0], h1 = hist[1], h2 = hist[2];
    u64 recNext = lane < nrec ? __ldcs(recs + lane) : 0ull;
    for (u32 base = 0; base < nrec; base += 32) {
        const u64 rec = recNext;
        const bool valid = base + lane < nrec;
        {   const u32 in = base + 32 + lane; recNext = in < nrec ? __ldcs(recs + in) : 0ull; }
        u32 ll, ml, ob;
        zl_lane_record(rec, valid, xtab, ll, ml, ob);
        const bool isM = ml != 0;
        const u32 off = zl_batch_offsets(ll, ml, ob, lane, h0, h1, h2);
        u32 sl, so, totalL, totalO;
        zl_batch_positions(ll, ml, lane, sl, so, totalL, totalO);
        const u32 litExcl = sl - ll;                 // literal bytes of earlier lanes in this batch
        const u32 dstLit = outPos + so - ll - ml;    // where this lane's literals go
        const u32 dm = dstLit + ll;                  // where this lane's match goes
        // ---- the checks of ZSTD_execSequence (zstd.c:44024-44066, 44320) for the whole batch, before anything is written
        if (totalL > litSize - litPos) return ZL_E_corruption_detected;
        if ((outPos - op) + totalO > cap) return capErr;
        if (__ballot_sync(ZL_FULL, isM && (off == 0 || off > dm + dictSize))) return ZL_E_corruption_detected;
        // ---- literals: one flat copy over the batch (zl_flat_copy)
        if (totalL) {
            const u32 ci = __popc(__ballot_sync(ZL_FULL, ll != 0) & ltMask);
            if (ll) seg[ci].x = dstLit - litExcl;
            __syncwarp();
            const u8* lsrc = lit + litPos;
            if (litMode == 1) zl_flat_copy(out, seg, litExcl, ll, totalL, lane, leMask, [&](u32, u32) { return rleByte; });
            else zl_flat_copy(out, seg, litExcl, ll, totalL, lane, leMask, [&](u32 j, u32) { return (u32)__ldg(lsrc + j); });
        }
        __syncwarp();
        // ---- matches: rounds against the high-water mark (signed positions: negative = dictionary)
        u32 pending = __ballot_sync(ZL_FULL, isM);
        const i32 srcBeg = (i32)dm - (i32)off;
        const bool overlap = off < ml;
        const i32 needEnd = srcBeg + (i32)(overlap ? off : ml);       // end of the source bytes actually read, <= dm
        while (pending) {
            const u32 f = (u32)__ffs((int)pending) - 1;
            const i32 hwm = (i32)__shfl_sync(ZL_FULL, dm, f);
            const bool ready = ((pending >> lane) & 1) && (lane == f || needEnd <= hwm);
            // (a) ready matches that do not overlap their own output: one flat byte-parallel copy over all of them, so the
            //     lanes share the bytes evenly whatever the individual lengths are; two rows per step, loads before stores
            const u32 fl = (ready && !overlap) ? ml : 0u;
            u32 rs = fl;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 a = __shfl_up_sync(ZL_FULL, rs, d); if ((int)lane >= d) rs += a; }
            const u32 totalM = __shfl_sync(ZL_FULL, rs, 31);
            const u32 rsExcl = rs - fl;
            if (totalM) {
                const u32 ci = __popc(__ballot_sync(ZL_FULL, fl != 0) & ltMask);
                if (fl) seg[ci] = make_uint2(dm - rsExcl, (u32)srcBeg - rsExcl);
                __syncwarp();
                zl_flat_copy(out, seg, rsExcl, fl, totalM, lane, leMask, [&](u32 j, u32 sd) {
                    const i32 sp = (i32)(j + sd);
                    return (kDict && sp < 0) ? (u32)dict[(i32)dictSize + sp] : (u32)out[sp];
                });
            }
            // (b) ready matches that overlap their own output (offset < length): the periodic form dst[k] = src[k mod offset]
            //     only reads bytes below the match, one match at a time across the warp
            u32 ov = __ballot_sync(ZL_FULL, ready && overlap);
            while (ov) {
                const u32 L = (u32)__ffs((int)ov) - 1;
                ov &= ov - 1;
                const u32 bdm = __shfl_sync(ZL_FULL, dm, L), boff = __shfl_sync(ZL_FULL, off, L), bml = __shfl_sync(ZL_FULL, ml, L);
                const i32 bsrc = (i32)bdm - (i32)boff;
                for (u32 k = lane; k < bml; k += 32) {
                    const i32 sp = bsrc + (i32)(k % boff);
                    out[bdm + k] = (kDict && sp < 0) ? dict[(i32)dictSize + sp] : out[sp];
                }
            }
            __syncwarp();
            pending &= ~__ballot_sync(ZL_FULL, ready);
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
