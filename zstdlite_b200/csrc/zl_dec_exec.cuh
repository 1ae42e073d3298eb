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

#if defined(__CUDACC__)

#define ZL_FULL 0xFFFFFFFFu

ZL_D void zl_warp_copy(u8* dst, const u8* src, u32 n, u32 lane)
{
    // generic byte-granular cooperative copy with a 16-byte fast path when co-aligned
    if (n >= 64 && ((((size_t)dst) ^ ((size_t)src)) & 15) == 0) {
        u32 head = (u32)((16 - (((size_t)dst) & 15)) & 15);
        if (lane < head) dst[lane] = src[lane];
        u32 body = (n - head) >> 4;
        const uint4* s4 = (const uint4*)(src + head);
        uint4* d4 = (uint4*)(dst + head);
        for (u32 i = lane; i < body; i += 32) d4[i] = s4[i];
        u32 done = head + (body << 4);
        for (u32 i = done + lane; i < n; i += 32) dst[i] = src[i];
    } else {
        for (u32 i = lane; i < n; i += 32) dst[i] = src[i];
    }
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

// Execute one compressed block.  `out` = frame output base, `op` = frame-relative position of the block.
template <bool kDict>
ZL_D void zl_exec_block(u8* out, u32 op, const ZlBlockHdr& h, const u8* lit, u32 rleByte, u32 litMode,
                        const u64* __restrict__ recs, const u8* dict, u32 dictSize, u32 lane)
{
    const u32 nrec = h.nrec;
    u32 outPos = op, litPos = 0;
    u64 recNext = lane < nrec ? recs[lane] : 0ull;
    for (u32 base = 0; base < nrec; base += 32) {
        const u64 rec = recNext;
        {   const u32 in = base + 32 + lane; recNext = in < nrec ? recs[in] : 0ull; }
        const u32 ll = (u32)(rec & 0xFFFF), ml = (u32)((rec >> 16) & 0xFFFF), off = (u32)(rec >> 32);
        // inclusive scans of ll and ll+ml
        u32 sl = ll, so = ll + ml;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 a = __shfl_up_sync(ZL_FULL, sl, d), b = __shfl_up_sync(ZL_FULL, so, d);
            if ((int)lane >= d) { sl += a; so += b; }
        }
        const u32 totalL = __shfl_sync(ZL_FULL, sl, 31), totalO = __shfl_sync(ZL_FULL, so, 31);
        const u32 litExcl = sl - ll;                 // literal bytes of earlier lanes in this batch
        const u32 dstLit = outPos + so - ll - ml;    // where this lane's literals go
        const u32 dm = dstLit + ll;                  // where this lane's match goes
        // ---- literals: flat byte-parallel copy over the batch (all lanes run every iteration: shuffles)
        for (u32 j0 = 0; j0 < totalL; j0 += 32) {
            const u32 j = j0 + lane;
            u32 k = 0;                               // owner = first lane whose inclusive sum exceeds j
#pragma unroll
            for (int st = 16; st >= 1; st >>= 1) {
                const u32 v = __shfl_sync(ZL_FULL, sl, (k + st - 1) & 31);
                if (v <= j) k += st;
            }
            k &= 31;
            const u32 kDst = __shfl_sync(ZL_FULL, dstLit, k), kEx = __shfl_sync(ZL_FULL, litExcl, k);
            if (j < totalL) out[kDst + (j - kEx)] = litMode == 1 ? (u8)rleByte : lit[litPos + j];
        }
        __syncwarp();
        // ---- matches: rounds against the high-water mark (signed positions: negative = dictionary)
        u32 pending = __ballot_sync(ZL_FULL, ml != 0);
        const i32 srcBeg = (i32)dm - (i32)off;
        const i32 needEnd = srcBeg + (i32)(off < ml ? off : ml);      // end of the source bytes actually read, <= dm
        while (pending) {
            const u32 f = (u32)__ffs((int)pending) - 1;
            const i32 hwm = (i32)__shfl_sync(ZL_FULL, dm, f);
            const bool ready = ((pending >> lane) & 1) && (lane == f || needEnd <= hwm);
            if (ready && ml <= 32) {
                u32 s = 0;
                for (u32 k = 0; k < ml; k += 4) {
                    u8 t[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (k + j < ml) {
                            const i32 sp = srcBeg + (i32)s;
                            if (kDict && sp < 0) t[j] = dict[(i32)dictSize + sp]; else t[j] = out[sp];
                            if (++s == off) s = 0;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) if (k + j < ml) out[dm + k + j] = t[j];
                }
            }
            u32 big = __ballot_sync(ZL_FULL, ready && ml > 32);
            while (big) {
                const u32 L = (u32)__ffs((int)big) - 1;
                big &= big - 1;
                const u32 bdm = __shfl_sync(ZL_FULL, dm, L), boff = __shfl_sync(ZL_FULL, off, L), bml = __shfl_sync(ZL_FULL, ml, L);
                const i32 bsrc = (i32)bdm - (i32)boff;
                if (boff >= bml && bsrc >= 0) zl_warp_copy(out + bdm, out + bsrc, bml, lane);
                else {
                    for (u32 k = lane; k < bml; k += 32) {
                        const i32 sp = bsrc + (i32)(boff >= bml ? k : k % boff);
                        out[bdm + k] = (kDict && sp < 0) ? dict[(i32)dictSize + sp] : out[sp];
                    }
                }
            }
            __syncwarp();
            pending &= ~__ballot_sync(ZL_FULL, ready);
        }
        outPos += totalO; litPos += totalL;
    }
    // last literals (zstd.c:44692-44698)
    const u32 lastLL = h.litSize - litPos;
    if (litMode == 1) zl_warp_fill(out + outPos, rleByte, lastLL, lane);
    else zl_warp_copy(out + outPos, lit + litPos, lastLL, lane);
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
// all 4 lanes of the quad call this; result valid on every lane of the quad
ZL_D u64 zl_quad_xxh64(const u8* p, u32 len, u32 q, u32 qmask, u32 qbase)
{
    u64 h;
    u32 done = 0;
    if (len >= 32) {
        u64 v = q == 0 ? (ZL_P1 + ZL_P2) : (q == 1 ? ZL_P2 : (q == 2 ? 0ull : (0ull - ZL_P1)));
        const u32 stripes = len >> 5;
        const u8* pp = p + 8 * q;
        for (u32 s = 0; s < stripes; s++) v = zl_xround(v, zl_ld64u(pp + 32 * s));
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
#endif  // __CUDACC__
