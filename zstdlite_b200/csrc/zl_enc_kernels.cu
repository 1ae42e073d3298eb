// zl_enc_kernels.cu -- the compression kernels and their launcher (sm_100a).  Work unit: one BLOCK of <= 128 KiB.
//   E1 zl_k_match          CTA of ZL_MATCH_WARPS warps per block: hash every position, look up / insert into the
//                          shared-memory tables in position order (warps take 32-position groups round-robin and pass
//                          a token through named barriers for the table section only), verify candidates -> M[p]
//   E2 zl_k_parse          warp per block: greedy walk over M in 32-position windows (ballot + ffs jumps), repeat-offset
//                          coding, literal gather + histogram -> sequence records, literal buffer
//   E3 zl_k_enc_literals   warp per block: Huffman code construction (lane 0), tree description, streams in 32 chunks
//   E4 zl_k_enc_sequences  warp per block: histograms (all lanes), three FSE tables and state chains (lanes 0-2),
//                          bit packing by all lanes (scan + shared staging window)
//   E5 zl_k_enc_plan       thread per frame: block types (compressed / raw), sizes, offsets, capacity check
//   E6 zl_k_enc_assemble   warp per block: frame header, block header, section copies into the destination, checksum
// The serial per-lane logic lives in zl_enc_entropy.cuh / zl_enc_match.cuh (shared with the CPU emulation in tests/emul).
#include "zl_enc_entropy.cuh"
#include "zl_enc_match.cuh"
#include "zl_dec_exec.cuh"        // zl_warp_copy
#include "zl_launch.h"

__constant__ ZlEncConst c_enc;
__device__ ZlEncDefTables g_encDef;                 // predefined LL / OF / ML encoding tables, built once (zl_enc_entropy.cuh)

cudaError_t zl_enc_upload_const()
{
    ZlEncConst h;
    zl_enc_const_init(&h);
    cudaError_t e = cudaMemcpyToSymbol(c_enc, &h, sizeof(h));
    if (e != cudaSuccess) return e;
    static ZlEncDefTables d;                        // (the same bytes for every device)
    static ZlSeqEncSm scratch;
    for (u32 t = 0; t < 3; t++) {
        const i16* def = t == 0 ? h.llDef : (t == 1 ? h.ofDef : h.mlDef);
        const u32 defSyms = t == 0 ? 36u : (t == 1 ? 29u : 53u), defLog = t == 1 ? 5u : 6u;
        i16 norm[64] = {};
        for (u32 s = 0; s < defSyms; s++) norm[s] = def[s];
        for (u32 s = 0; s < 64; s++) { d.dNb[t][s] = 0; d.dFS[t][s] = 0; }
        zl_fse_build_ctable(d.state[t], d.dNb[t], d.dFS[t], norm, defSyms - 1, defLog, scratch.symOf[t], scratch.cumul[t]);
    }
    return cudaMemcpyToSymbol(g_encDef, &d, sizeof(d));
}

// ---------------------------------------------------------------------------------------------- helpers
__device__ __forceinline__ void zl_bar_sync(u32 id, u32 count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void zl_bar_arrive(u32 id, u32 count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// shared-memory mbarriers (one arrival per phase): the table token of zl_k_match when a CTA has more warps than named barriers
__device__ __forceinline__ void zl_mbar_init(u64* bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void zl_mbar_arrive(u64* bar) { asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"((u32)__cvta_generic_to_shared(bar)) : "memory"); }
__device__ __forceinline__ void zl_mbar_wait(u64* bar, u32 parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tZL_MBAR_WAIT:\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n\t@p bra ZL_MBAR_DONE;\n\tbra ZL_MBAR_WAIT;\n\tZL_MBAR_DONE:\n\t}"
                 ::"r"((u32)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

// 8 bytes at byte offset `bo` from the 4-aligned base; words past `lastWord` are clamped (callers bound lengths by the block size)
__device__ __forceinline__ void zl_ld8(const u32* __restrict__ wbase, u32 bo, u32 lastWord, u32& lo, u32& hi)
{
    const u32 wi = bo >> 2, sh = (bo & 3) * 8;
    const u32 a = __ldg(wbase + min(wi, lastWord)), b = __ldg(wbase + min(wi + 1, lastWord)), c = __ldg(wbase + min(wi + 2, lastWord));
    lo = __funnelshift_r(a, b, sh); hi = __funnelshift_r(b, c, sh);
}
// The same from an 8-ALIGNED base, for candidate positions: every lane reads somewhere else, and the L1 pipe spends a wavefront per lane and
// load instruction whatever the width -- so two 8-byte loads instead of three 4-byte loads, without a branch (a 16-byte load with a
// conditional second one was measured slower: the divergent selection cost more than the loads it saved).  lastVec: last 8-byte unit.
__device__ __forceinline__ void zl_ld8v(const u32* __restrict__ wbase, u32 bo, u32 lastVec, u32& lo, u32& hi)
{
    const uint2* __restrict__ vb = reinterpret_cast<const uint2*>(wbase);
    const u32 vi = bo >> 3;
    const uint2 a = __ldg(vb + min(vi, lastVec)), b = __ldg(vb + min(vi + 1, lastVec));
    const bool up = (bo & 4) != 0;
    const u32 w0 = up ? a.y : a.x, w1 = up ? b.x : a.y, w2 = up ? b.y : b.x;
    const u32 sh = (bo & 3) * 8;
    lo = __funnelshift_r(w0, w1, sh); hi = __funnelshift_r(w1, w2, sh);
}
// the rest of a match whose first 8 bytes are equal (zl_match_len below), capped at lim
__device__ __forceinline__ u32 zl_match_more(const u32* __restrict__ wbase, u32 bias, u32 lastWord, u32 p, u32 q, u32 lim)
{
    u32 len = 8;
    for (u32 k = 8; k < lim; k += 8) {                        // (16 bytes a step, four loads in flight, was measured slower: +8 % at level 1)
        u32 alo, ahi, blo, bhi;
        zl_ld8(wbase, bias + p + k, lastWord, alo, ahi);
        zl_ld8v(wbase, bias + q + k, lastWord >> 1, blo, bhi);
        const u32 c = zl_common8(alo, ahi, blo, bhi);
        len += c;
        if (c < 8) break;
    }
    return len < lim ? len : lim;
}
// length of the match between positions p and q (< p), capped at lim (<= ZL_M_CAP); (lo, hi) = the 8 bytes at p.  wbase is 8-aligned here.
__device__ __forceinline__ u32 zl_match_len(const u32* __restrict__ wbase, u32 bias, u32 lastWord, u32 p, u32 q, u32 lo, u32 hi, u32 lim)
{
    u32 blo, bhi;
    zl_ld8v(wbase, bias + q, lastWord >> 1, blo, bhi);
    u32 len = zl_common8(lo, hi, blo, bhi);
    if (len == 8) {
        for (u32 k = 8; k < lim; k += 8) {
            u32 alo, ahi;
            zl_ld8(wbase, bias + p + k, lastWord, alo, ahi);
            zl_ld8v(wbase, bias + q + k, lastWord >> 1, blo, bhi);
            const u32 c = zl_common8(alo, ahi, blo, bhi);
            len += c;
            if (c < 8) break;
        }
    }
    return len < lim ? len : lim;
}

// ---------------------------------------------------------------------------------------------- E1: match candidates
// dictionary candidate for position p (first block of a frame): last occurrence of the hash in the prebuilt table,
// verified against the dictionary content; the match may not run past the end of the content.  Returns the length
// (0 = none) and the offset (p + distance to the end of the dictionary).
__device__ __forceinline__ u32 zl_dict_match(const ZlEncDictDev& D, const u32* __restrict__ tab, u32 h, const u32* __restrict__ wbase, u32 bias,
                                             u32 lastWord, u32 p, u32 lo, u32 hi, u32 lim, u32 mls, u32& offOut)
{
    const u32 e = __ldg(tab + h);
    if (!e) return 0;
    const u32 q = e - 1, room = D.contentSize - q;
    const u32* __restrict__ dw = reinterpret_cast<const u32*>(D.content);
    const u32 dLast = (D.contentSize + 15) >> 2;                     // the content is followed by >= 16 zero bytes
    if (lim > room) lim = room;
    u32 blo, bhi;
    zl_ld8(dw, q, dLast, blo, bhi);
    u32 len = zl_common8(lo, hi, blo, bhi);
    if (len == 8) {
        for (u32 k = 8; k < lim; k += 8) {
            u32 alo, ahi;
            zl_ld8(wbase, bias + p + k, lastWord, alo, ahi);
            zl_ld8(dw, q + k, dLast, blo, bhi);
            const u32 c = zl_common8(alo, ahi, blo, bhi);
            len += c;
            if (c < 8) break;
        }
    }
    if (len > lim) len = lim;
    if (len < mls) return 0;
    offOut = p + room;
    return len;
}

// ---------------------------------------------------------------------------------------------- E0: far table (frames of > 1 block)
// zl_enc_match.cuh, "far candidates": tab[region][hash8(position)] = min(position) over the region (8 MiB of the frame).  ZlEncFrame::pad = offset of the frame's
// table in the far arena (u32 units, low 56 bits) | log2 entries << 56 (0: the frame has no table); ZlEncBlock::pad = the block's offset
// inside its frame.  A CTA takes 8 KiB of a block (ZL_FAR_CHUNKS CTAs per block), so the CTAs in flight (148 SMs x 8) cover about 9 MiB of the
// frame: one or two regions, whose tables (32 MB each) then live in L2 -- with a CTA per block 148 MiB were in flight, 18 regions, and the
// atomics went to DRAM (ncu: L2 hit rate 33 %, 6.9 ms for a 256 MiB frame).  The 8 bytes of a position may reach into the next block
// (the frame is contiguous).
#define ZL_FAR_OFF_MASK 0x00FFFFFFFFFFFFFFull
#define ZL_FAR_CHUNKS 16u
__global__ void __launch_bounds__(256)
zl_k_far_build(const ZlEncBlock* __restrict__ blocks, const ZlEncFrame* __restrict__ frames, u32* __restrict__ farArena)
{
    const ZlEncBlock b = blocks[blockIdx.x / ZL_FAR_CHUNKS];
    const u64 fpad = frames[b.frame].pad;
    const u32 flog = (u32)(fpad >> 56);
    if (!flog || b.srcSize == 0) return;
    const u32 p0 = (blockIdx.x % ZL_FAR_CHUNKS) * (ZL_BLOCKSIZE_MAX / ZL_FAR_CHUNKS);
    u32* __restrict__ tab = farArena + (fpad & ZL_FAR_OFF_MASK) + ((size_t)(b.pad >> ZL_FAR_REGION_LOG) << flog);      // the block's region
    const u32 n = b.srcSize;
    const u32 bias = (u32)(((size_t)b.src) & 3);
    const u32* __restrict__ wbase = reinterpret_cast<const u32*>(b.src - bias);
    const bool last = (b.flags & ZL_BLK_LAST) != 0;
    const u32 npos = last ? (n >= 8 ? n - 7 : 0u) : n;                      // positions whose 8 bytes lie inside the frame
    const u32 lastWord = (bias + (last ? n - 1 : n + 7)) >> 2;
    const u32 p1 = min(npos, p0 + ZL_BLOCKSIZE_MAX / ZL_FAR_CHUNKS);
    for (u32 p = p0 + threadIdx.x; p < p1; p += 256) {
        u32 lo, hi;
        zl_ld8(wbase, bias + p, lastWord, lo, hi);
        const u32 hx = zl_far_hash(lo, hi, flog);
        const u32 val = (((b.pad + p) & ((1u << ZL_FAR_REGION_LOG) - 1)) << 8) | (hx & 255u);
        // entries only ever decrease: one that is already lower needs no atomic (text repeats its common strings thousands of times per
        // region, and those atomics serialise on one L2 line)
        if (__ldcg(tab + (hx >> 8)) > val) atomicMin(tab + (hx >> 8), val);
    }
}
// length of the match between block position p (bytes lo:hi, words of the block) and FRAME position q (words of the frame), <= lim
__device__ __forceinline__ u32 zl_match_len_far(const u32* __restrict__ wbase, u32 bias, u32 lastWord, u32 p, const u32* __restrict__ fw, u32 fbias, u32 q,
                                                u32 lo, u32 hi, u32 lim)
{
    u32 blo, bhi;
    zl_ld8(fw, fbias + q, 0xFFFFFFFFu, blo, bhi);
    u32 len = zl_common8(lo, hi, blo, bhi);
    if (len == 8) {
        for (u32 k = 8; k < lim; k += 8) {
            u32 alo, ahi;
            zl_ld8(wbase, bias + p + k, lastWord, alo, ahi);
            zl_ld8(fw, fbias + q + k, 0xFFFFFFFFu, blo, bhi);
            const u32 c = zl_common8(alo, ahi, blo, bhi);
            len += c;
            if (c < 8) break;
        }
    }
    return len < lim ? len : lim;
}

// E1b: the far candidate of every position of a multi-block frame, after E1 (zl_enc_match.cuh, "far candidates").  Its own kernel: in E1 it
// cost 20 registers (two CTAs a SM instead of three) and its lookups went to DRAM, because the blocks in flight there cover more regions than
// L2 holds tables for.  Here a CTA takes 8 KiB like the build, so the lookups of the CTAs in flight fall into one or two region tables, and
// the high occupancy hides what latency is left.  M[p] is replaced when the far match is clearly better than what E1 found (zl_far_better).
__global__ void __launch_bounds__(256)
zl_k_far_match(const ZlEncBlock* __restrict__ blocks, const ZlEncFrame* __restrict__ frames, const u32* __restrict__ farArena, u32* __restrict__ Marena, u32 slotM, ZlEncParams P)
{
    const u32 blk = blockIdx.x / ZL_FAR_CHUNKS;
    const ZlEncBlock b = blocks[blk];
    const u64 fpad = frames[b.frame].pad;
    const u32 flog = (u32)(fpad >> 56);
    if (!flog || b.srcSize < 8) return;
    const u32* __restrict__ ftab = farArena + (fpad & ZL_FAR_OFF_MASK);
    u32* __restrict__ M = Marena + (size_t)blk * slotM;
    const u32 n = b.srcSize;
    const u32 bias = (u32)(((size_t)b.src) & 3);
    const u32* __restrict__ wbase = reinterpret_cast<const u32*>(b.src - bias);
    const u32 lastWord = (bias + n - 1) >> 2;
    const u8* fbase = b.src - b.pad;                                          // the frame is contiguous
    const u32 fbias = (u32)(((size_t)fbase) & 3);
    const u32* __restrict__ fw = reinterpret_cast<const u32*>(fbase - fbias);
    const u32 p0 = (blockIdx.x % ZL_FAR_CHUNKS) * (ZL_BLOCKSIZE_MAX / ZL_FAR_CHUNKS);
    const u32 p1 = min(n - 7, p0 + ZL_BLOCKSIZE_MAX / ZL_FAR_CHUNKS);         // positions with 8 bytes inside the block, as in E1
    for (u32 p = p0 + threadIdx.x; p < p1; p += 256) {
        u32 lo, hi;
        zl_ld8(wbase, bias + p, lastWord, lo, hi);
        const u32 m = M[p];
        u32 bestLen = m & 0xFFu, bestOff = 0;
        const u32 lim = min(n - p, ZL_M_CAP);
        const u32 pos = b.pad + p, hx = zl_far_hash(lo, hi, flog), hF = hx >> 8, reg = pos >> ZL_FAR_REGION_LOG;
        bool got = false;                                                     // own region first, then the one before it
        u32 e = __ldg(ftab + ((size_t)reg << flog) + hF);
        u32 q = (reg << ZL_FAR_REGION_LOG) + (e >> 8);
        if ((e & 255u) == (hx & 255u) && q < pos && pos - q > 65535u && pos - q < P.farMaxOff) {      // (an empty entry has q >= pos)
            const u32 l = zl_match_len_far(wbase, bias, lastWord, p, fw, fbias, q, lo, hi, lim);
            if (zl_far_better(l, bestLen, P.mls)) { bestLen = l; bestOff = pos - q; got = true; }
        }
        if (!got && zl_far_use_prev(pos)) {
            e = __ldg(ftab + ((size_t)(reg - 1) << flog) + hF);
            q = ((reg - 1) << ZL_FAR_REGION_LOG) + (e >> 8);
            if (e != ZL_FAR_EMPTY && (e & 255u) == (hx & 255u) && pos - q > 65535u && pos - q < P.farMaxOff) {
                const u32 l = zl_match_len_far(wbase, bias, lastWord, p, fw, fbias, q, lo, hi, lim);
                if (zl_far_better(l, bestLen, P.mls)) { bestLen = l; bestOff = pos - q; }
            }
        }
        if (bestOff) M[p] = (bestOff << 8) | bestLen;
    }
}

// M[p] of one position (E1, both kernels): its candidates are the highest lower lane of its group with the same hash (prevS / prevL, -1 = none),
// else the table entry from before the group (eS / eL); first block of a frame compressed with a dictionary: the dictionary's tables as well.
template <bool kLong, bool kDict>
__device__ __forceinline__ u32 zl_match_verify(const u32* __restrict__ wbase, u32 bias, u32 lastWord, u32 n, u32 p, u32 gbase, u32 lo, u32 hi, u32 eS, u32 eL,
                                               i32 prevS, i32 prevL, u32 mls, bool firstBlock, const ZlEncDictDev* __restrict__ dict)
{
    const u32 lim = min(n - p, ZL_M_CAP), limV = min(n - p, ZL_M_VERIFY);     // dictionary / in-block candidates
    u32 bestLen = 0, bestOff = 0;
    // the first 8 bytes of BOTH candidates are requested before either is looked at: the two loads are a cache miss each more often
    // than not, and one after the other they were the longest wait of a warp's turn
    const i32 qL = !kLong ? -1 : (prevL >= 0 ? (i32)(gbase + (u32)prevL) : zl_cand_pos(eL, p));
    i32 qS = prevS >= 0 ? (i32)(gbase + (u32)prevS) : zl_cand_pos(eS, p);
    if (kLong && qS == qL) qS = -1;                  // both tables name the same position (the usual case of a real match): one look is enough
    u32 cLlo = 0, cLhi = 0, cSlo = 0, cShi = 0;
    if (kLong && qL >= 0) zl_ld8v(wbase, bias + (u32)qL, lastWord >> 1, cLlo, cLhi);
    if (qS >= 0) zl_ld8v(wbase, bias + (u32)qS, lastWord >> 1, cSlo, cShi);
    if (kLong && qL >= 0) {
        u32 l = zl_common8(lo, hi, cLlo, cLhi);
        if (l == 8) l = zl_match_more(wbase, bias, lastWord, p, (u32)qL, limV);
        if (l > limV) l = limV;
        if (l >= mls) { bestLen = l; bestOff = p - (u32)qL; }
    }
    // a verified long-hash candidate (>= 8 bytes) is taken as it is, like the reference's double-fast search (zstd.c:29989);
    // and a longer match is impossible once the limit is reached
    if (qS >= 0 && bestLen < 8 && bestLen < limV) {
        u32 l = zl_common8(lo, hi, cSlo, cShi);
        if (l == 8) l = zl_match_more(wbase, bias, lastWord, p, (u32)qS, limV);
        if (l > limV) l = limV;
        if (l >= mls && l > bestLen) { bestLen = l; bestOff = p - (u32)qS; }
    }
    if (kDict && firstBlock && bestLen < lim) {      // dictionary content precedes the first block
        u32 dOff = 0;
        if (kLong) {
            const u32 l = zl_dict_match(*dict, dict->tabL, zl_hash_long(lo, hi, dict->hlogL), wbase, bias, lastWord, p, lo, hi, lim, mls, dOff);
            if (l > bestLen) { bestLen = l; bestOff = dOff; }
        }
        if (bestLen < lim) {
            const u32 l = zl_dict_match(*dict, dict->tabS, zl_hash_short(lo, hi, mls, dict->hlogS), wbase, bias, lastWord, p, lo, hi, lim, mls, dOff);
            if (l > bestLen) { bestLen = l; bestOff = dOff; }
        }
    }
    return bestLen ? ((bestOff << 8) | bestLen) : 0u;
}

#define ZL_MATCH_SCRATCH 4096u      // u16 slots of the duplicate-detection scratch (8 KB per CTA)
template <bool kLong, bool kDict>
__global__ void __launch_bounds__(ZL_MATCH_WARPS * 32)
zl_k_match(const ZlEncBlock* __restrict__ blocks, u32* __restrict__ Marena, u32 slotM, ZlEncParams P, const ZlEncDictDev* __restrict__ dict, u32 skipSmall)
{
    extern __shared__ __align__(16) u8 smraw[];
    u16* tabS = reinterpret_cast<u16*>(smraw);
    u16* tabL = tabS + (1u << P.hlogS);
    volatile u16* scr = tabL + (kLong ? (1u << P.hlogL) : 0u);            // duplicate detection (below); never initialised, never trusted
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ZlEncBlock b = blocks[blockIdx.x];
    const u32 n = b.srcSize;
    if (skipSmall && n <= ZL_SMALL_BLOCK) return;                        // zl_k_match_small's
    u32 hlogS, hlogL;
    zl_block_hlog(P, n, hlogS, hlogL);                                   // (the layout keeps the level's sizes; a small block uses less of it)
    u32* __restrict__ M = Marena + (size_t)blockIdx.x * slotM;
#if ZL_MATCH_WARPS > 15
    __shared__ u64 token[ZL_MATCH_WARPS];
    if (tid < ZL_MATCH_WARPS) zl_mbar_init(&token[tid], 1);
    u32 nwait = 0;
#endif
    {   uint4* z = reinterpret_cast<uint4*>(tabS);
        for (u32 i = tid; i < (2u << hlogS) / 16; i += ZL_MATCH_WARPS * 32) z[i] = make_uint4(0, 0, 0, 0);
        z = reinterpret_cast<uint4*>(tabL);
        if (kLong) for (u32 i = tid; i < (2u << hlogL) / 16; i += ZL_MATCH_WARPS * 32) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    if (n == 0) return;
    const u32 bias = (u32)(((size_t)b.src) & 7);      // 8-aligned base: zl_ld8v
    const u32* __restrict__ wbase = reinterpret_cast<const u32*>(b.src - bias);
    const u32 lastWord = (bias + n - 1) >> 2;
    const u32 ngroups = (n + 31) >> 5;
    const u32 ltMask = (1u << lane) - 1;
    // Each warp takes TWO consecutive 32-position groups per turn (64 positions): their loads and verifications overlap,
    // and the table token is passed half as often.
    const u32 npairs = (ngroups + 1) >> 1;
    // the pair's own bytes are fetched one turn ahead: a warp used to spend a quarter of its stall samples waiting for this load (ncu)
    u32 wq[2];
#pragma unroll
    for (int h = 0; h < 2; h++) wq[h] = lane < 11 ? __ldg(wbase + min(((bias + ((2 * warp + h) << 5)) >> 2) + lane, lastWord)) : 0u;
    for (u32 pr = warp; pr < npairs; pr += ZL_MATCH_WARPS) {
        u32 wc[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            wc[h] = wq[h];
            wq[h] = lane < 11 ? __ldg(wbase + min(((bias + ((2 * (pr + ZL_MATCH_WARPS) + h) << 5)) >> 2) + lane, lastWord)) : 0u;
        }
        u32 lo[2], hi[2], hS[2], hL[2], eS[2], eL[2];
        i32 prevS[2], prevL[2];
        bool valid[2], lastS[2], lastL[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const u32 g = 2 * pr + h;
            const u32 p = (g << 5) + lane;
            // the 8 bytes at every position of the group from 11 coalesced words
            const u32 w = wc[h];
            const u32 bo = (bias & 3) + lane, j = bo >> 2, sh = (bo & 3) * 8;
            const u32 w0 = __shfl_sync(ZL_FULL, w, j), w1 = __shfl_sync(ZL_FULL, w, j + 1), w2 = __shfl_sync(ZL_FULL, w, j + 2);
            lo[h] = __funnelshift_r(w0, w1, sh); hi[h] = __funnelshift_r(w1, w2, sh);
            valid[h] = p + 8 <= n;
            hS[h] = valid[h] ? zl_hash_short(lo[h], hi[h], P.mls, hlogS) : (0x10000u + lane);
            hL[h] = 0; prevS[h] = -1; prevL[h] = -1; lastS[h] = true; lastL[h] = true;
            if (kLong) hL[h] = valid[h] ? zl_hash_long(lo[h], hi[h], hlogL) : (0x10000u + lane);
            // The tables are updated as if the 32 positions of a group were inserted one after the other: an entry ends up with the HIGHEST
            // lane of its hash (lastS), and a lane's candidate is the highest LOWER lane with its hash (prevS), else the entry from before the
            // group.  __match_any_sync answers both, but costs about 64 cycles of the ADU pipe an instruction, and with one or two of them
            // per group that pipe was 78 % busy and bounded the kernel (ncu, profiles/r02_ncu_full_compress_kernels_raw.csv).  Groups in which
            // two lanes share a hash are the minority (measured: text 17 % for 5 bytes and 4 % for 8, columnar 68 % / 26 %, noise 0), so
            // every lane first drops its position into a small scratch table under its hash and reads the slot back: a lane that finds
            // another value shared the slot with someone (a duplicate, a slot collision or another warp -- the scratch is the CTA's), and
            // only then does the warp ask match.any.  Of two lanes with one hash at least one reads the other's value, so no duplicate is missed.
            const u32 p16 = p & 0xFFFFu;
            const u32 sS = hS[h] & (ZL_MATCH_SCRATCH - 1);
            if (valid[h]) scr[sS] = (u16)p16;
            __syncwarp();
            const u32 rS = valid[h] ? (u32)scr[sS] : p16;
            if (__ballot_sync(ZL_FULL, rS != p16)) {
                const u32 mS = __match_any_sync(ZL_FULL, hS[h]);
                lastS[h] = (mS >> lane) == 1u;                    // no higher lane shares the hash: this lane's insert survives
                prevS[h] = (mS & ltMask) ? (31 - __clz((int)(mS & ltMask))) : -1;
            }
            if (kLong) {
                const u32 sL = (hL[h] ^ 0x555u) & (ZL_MATCH_SCRATCH - 1);
                __syncwarp();
                if (valid[h]) scr[sL] = (u16)p16;
                __syncwarp();
                const u32 rL = valid[h] ? (u32)scr[sL] : p16;
                if (__ballot_sync(ZL_FULL, rL != p16)) {
                    const u32 mL = __match_any_sync(ZL_FULL, hL[h]);
                    lastL[h] = (mL >> lane) == 1u;
                    prevL[h] = (mL & ltMask) ? (31 - __clz((int)(mL & ltMask))) : -1;
                }
            }
        }
        // ---- table section, in position order across warps (and across the two groups: same-warp shared-memory order)
#if ZL_MATCH_WARPS > 15
        if (pr > 0) { zl_mbar_wait(&token[warp], nwait & 1u); nwait++; }
#else
#ifndef ZL_EXP_NOTOKEN      /* (timing experiment only: the warps run free and the tables are updated in any order) */
        if (ZL_MATCH_WARPS > 1 && pr > 0) zl_bar_sync(1 + warp, 64);
#endif
#endif
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const u32 p = ((2 * pr + h) << 5) + lane;
            eS[h] = 0; eL[h] = 0;
            if (valid[h]) {
                eS[h] = tabS[hS[h]];
                if (kLong) eL[h] = tabL[hL[h]];
            }
            __syncwarp();                                     // (every lane has read the old entries before any lane writes)
            if (valid[h]) {
                if (lastS[h]) tabS[hS[h]] = (u16)p;
                if (kLong && lastL[h]) tabL[hL[h]] = (u16)p;
            }
            __syncwarp();
        }
#if ZL_MATCH_WARPS > 15
        if (pr + 1 < npairs) { __syncwarp(); if (lane == 0) zl_mbar_arrive(&token[(warp + 1) % ZL_MATCH_WARPS]); }
#else
        // (no fence before the arrive: st.shared; bar.arrive | bar.sync; ld.shared is the producer / consumer pattern the PTX manual gives
        //  for named barriers, and a membar here also waited for the global loads this warp has in flight -- the prefetch above)
#ifndef ZL_EXP_NOTOKEN
        if (ZL_MATCH_WARPS > 1 && pr + 1 < npairs) zl_bar_arrive(1 + (warp + 1) % ZL_MATCH_WARPS, 64);
#endif
#endif
        // ---- verify
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const u32 g = 2 * pr + h;
            const u32 p = (g << 5) + lane;
            const u32 m = valid[h] ? zl_match_verify<kLong, kDict>(wbase, bias, lastWord, n, p, g << 5, lo[h], hi[h], eS[h], eL[h], prevS[h], prevL[h], P.mls,
                                                                   (b.flags & ZL_BLK_FIRST) != 0, dict) : 0u;
            if (p < n) M[p] = m;
        }
    }
}

// E1 for small blocks (<= ZL_SMALL_BLOCK bytes, zl_enc_match.cuh): a WARP per block, ZL_SMALL_WARPS blocks per CTA, 2 + 4 KB of tables per warp.
// The groups of a block are taken in order by one warp, so there is no token; duplicates inside a group are found through the tables
// themselves: every lane reads the old entry, stores its position, reads the entry back -- a lane that reads another position shares its hash
// with a later store, and only then match.any sorts the group out and the highest lane of each hash stores again.
#define ZL_SMALL_WARPS 8
template <bool kLong, bool kDict>
__global__ void __launch_bounds__(ZL_SMALL_WARPS * 32)
zl_k_match_small(const ZlEncBlock* __restrict__ blocks, u32 nblocks, u32* __restrict__ Marena, u32 slotM, ZlEncParams P, const ZlEncDictDev* __restrict__ dict)
{
    constexpr u32 kPer = (2u << ZL_SMALL_HLOG_S) + (kLong ? (2u << ZL_SMALL_HLOG_L) : 0u);
    __shared__ __align__(16) u8 sm[ZL_SMALL_WARPS * kPer];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 blk = blockIdx.x * ZL_SMALL_WARPS + warp;
    if (blk >= nblocks) return;
    const ZlEncBlock b = blocks[blk];
    const u32 n = b.srcSize;
    if (n == 0 || n > ZL_SMALL_BLOCK) return;                            // zl_k_match's
    volatile u16* tabS = reinterpret_cast<u16*>(sm + warp * kPer);
    volatile u16* tabL = tabS + (1u << ZL_SMALL_HLOG_S);
    {   uint4* z = reinterpret_cast<uint4*>(sm + warp * kPer);
        for (u32 i = lane; i < kPer / 16; i += 32) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    u32* __restrict__ M = Marena + (size_t)blk * slotM;
    const u32 bias = (u32)(((size_t)b.src) & 7);
    const u32* __restrict__ wbase = reinterpret_cast<const u32*>(b.src - bias);
    const u32 lastWord = (bias + n - 1) >> 2;
    const u32 ngroups = (n + 31) >> 5;
    const u32 ltMask = (1u << lane) - 1;
    const bool firstBlock = (b.flags & ZL_BLK_FIRST) != 0;
    u32 wq = lane < 11 ? __ldg(wbase + min((bias >> 2) + lane, lastWord)) : 0u;
    for (u32 g = 0; g < ngroups; g++) {
        const u32 w = wq;
        wq = lane < 11 ? __ldg(wbase + min(((bias + ((g + 1) << 5)) >> 2) + lane, lastWord)) : 0u;
        const u32 p = (g << 5) + lane;
        const u32 bo = (bias & 3) + lane, j = bo >> 2, sh = (bo & 3) * 8;
        const u32 w0 = __shfl_sync(ZL_FULL, w, j), w1 = __shfl_sync(ZL_FULL, w, j + 1), w2 = __shfl_sync(ZL_FULL, w, j + 2);
        const u32 lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
        const bool valid = p + 8 <= n;
        const u32 hS = valid ? zl_hash_short(lo, hi, P.mls, ZL_SMALL_HLOG_S) : (0x10000u + lane);
        const u32 hL = !kLong ? 0u : (valid ? zl_hash_long(lo, hi, ZL_SMALL_HLOG_L) : (0x10000u + lane));
        u32 eS = 0, eL = 0;
        if (valid) { eS = tabS[hS]; if (kLong) eL = tabL[hL]; }
        __syncwarp();
        if (valid) { tabS[hS] = (u16)p; if (kLong) tabL[hL] = (u16)p; }
        __syncwarp();
        u32 rS = p, rL = p;
        if (valid) { rS = tabS[hS]; if (kLong) rL = tabL[hL]; }
        i32 prevS = -1, prevL = -1;
        if (__ballot_sync(ZL_FULL, rS != p)) {
            const u32 mS = __match_any_sync(ZL_FULL, hS);
            prevS = (mS & ltMask) ? (31 - __clz((int)(mS & ltMask))) : -1;
            if (valid && (mS >> lane) == 1u) tabS[hS] = (u16)p;          // no higher lane shares the hash: this lane's insert survives
        }
        if (kLong && __ballot_sync(ZL_FULL, rL != p)) {
            const u32 mL = __match_any_sync(ZL_FULL, hL);
            prevL = (mL & ltMask) ? (31 - __clz((int)(mL & ltMask))) : -1;
            if (valid && (mL >> lane) == 1u) tabL[hL] = (u16)p;
        }
        __syncwarp();
        const u32 m = valid ? zl_match_verify<kLong, kDict>(wbase, bias, lastWord, n, p, g << 5, lo, hi, eS, eL, prevS, prevL, P.mls, firstBlock, dict) : 0u;
        if (p < n) M[p] = m;
    }
}

// ---------------------------------------------------------------------------------------------- E2: greedy walk
// cooperative extension of a match that hit the cap: compares 256 bytes per round
__device__ __forceinline__ u32 zl_extend_match(const u32* __restrict__ wbase, u32 bias, u32 lastWord, u32 n, u32 pos, u32 off, u32 lane, u32 len0)
{
    u32 len = len0;
    for (;;) {
        const u32 q = pos + len + 8 * lane;
        u32 c = 0;
        if (q < n) {
            u32 alo, ahi, blo, bhi;
            zl_ld8(wbase, bias + q, lastWord, alo, ahi);
            zl_ld8(wbase, bias + q - off, lastWord, blo, bhi);
            c = zl_common8(alo, ahi, blo, bhi);
            if (c > n - q) c = n - q;
        }
        const bool full = c == 8 && q + 8 <= n;
        const u32 stop = __ballot_sync(ZL_FULL, !full);
        if (!stop) { len += 256; continue; }
        const u32 first = (u32)__ffs((int)stop) - 1;
        return len + 8 * first + __shfl_sync(ZL_FULL, c, first);
    }
}

// One CTA per block, one warp per SEGMENT of ZL_PARSE_SEG bytes (zl_enc_match.cuh): the greedy walk is a serial chain, so a
// 128 KiB block is walked as eight independent 16 KiB pieces (each starts with an unknown repeat-offset history and clips its
// matches at its end, exactly like blocks do inside a frame) and the pieces are then packed into one record / literal array.
__global__ void __launch_bounds__(ZL_PARSE_WARPS * 32)
zl_k_parse(const ZlEncBlock* __restrict__ blocks, u32 nblocks, const u32* __restrict__ Marena, u32 slotM, u64* __restrict__ recArena,
           u32 slotRec, u8* __restrict__ litArena, u32 slotLit, u32* __restrict__ histArena, ZlEncBlockMeta* __restrict__ metas,
           const ZlEncDictDev* __restrict__ dict, u32 segmented)
{
    // segmented == 0: no block of the wave is longer than one segment; a CTA then takes ZL_PARSE_WARPS blocks, one per warp
    __shared__ u32 hist[ZL_PARSE_WARPS][256];
    __shared__ u32 segSeq[ZL_PARSE_WARPS], segLit[ZL_PARSE_WARPS], segTail[ZL_PARSE_WARPS];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 blk = segmented ? blockIdx.x : blockIdx.x * ZL_PARSE_WARPS + warp;
    if (blk >= nblocks) return;                              // (only without segments: no CTA-wide barrier follows)
    const u32 sg = segmented ? warp : 0u;
    for (u32 i = lane; i < 256; i += 32) hist[warp][i] = 0;
    __syncwarp();
    const ZlEncBlock b = blocks[blk];
    const u32 n = b.srcSize;
    const u32* __restrict__ M = Marena + (size_t)blk * slotM;
    u64* __restrict__ recsAll = recArena + (size_t)blk * slotRec;
    u8* __restrict__ litAll = litArena + (size_t)blk * slotLit;
    const u32 segBeg = sg * ZL_PARSE_SEG, segEnd = min(n, segBeg + ZL_PARSE_SEG);
    u64* __restrict__ recs = recsAll + (size_t)sg * ZL_PARSE_SEG_RECS;
    u8* __restrict__ lit = litAll + segBeg;
    const u32 bias = (u32)(((size_t)b.src) & 7);      // 8-aligned base: zl_ld8v
    const u32* __restrict__ wbase = reinterpret_cast<const u32*>(b.src - bias);
    const u32 lastWord = n ? (bias + n - 1) >> 2 : 0;
    const u32 ltMask = (1u << lane) - 1;
    const bool firstBlk = (b.flags & ZL_BLK_FIRST) != 0;
    const bool first = firstBlk && sg == 0;                  // only the first segment of a frame knows the decoder's history
    ZlReps reps;                                             // zl_enc_match.cuh: unknown history (0) otherwise
    reps.r0 = first ? 1u : 0u; reps.r1 = first ? 4u : 0u; reps.r2 = first ? 8u : 0u;
    if (first && dict && dict->hasEntropy) { reps.r0 = dict->rep[0]; reps.r1 = dict->rep[1]; reps.r2 = dict->rep[2]; }    // zstd.c:27450 (dictionary repcodes)
    const bool repPref = firstBlk && dict != nullptr;
    u32 p = segBeg, anchor = segBeg, nseq = 0, nlit = 0;
    if (segBeg < n) {
    // M and the source bytes are fetched one 128-position super-window ahead (the walk itself never waits on memory)
    u32 mq[4], bq[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { const u32 x = segBeg + 32 * k + lane; mq[k] = x < n ? __ldcs(M + x) : 0u; bq[k] = x < n ? (u32)b.src[x] : 0u; }
    for (u32 sw = segBeg; sw < segEnd; sw += 128) {
        u32 mc[4], bc[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { mc[k] = mq[k]; bc[k] = bq[k]; }
#pragma unroll
        for (int k = 0; k < 4; k++) { const u32 x = sw + 128 + 32 * k + lane; mq[k] = x < n ? __ldcs(M + x) : 0u; bq[k] = x < n ? (u32)b.src[x] : 0u; }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const u32 w0 = sw + 32 * k;
            if (w0 >= segEnd || p >= w0 + 32) continue;          // past the end / window entirely inside a match
            const u32 m = mc[k], byte = bc[k];
            const u32 pos = w0 + lane;
            // a match ends with its segment; one clipped below the format's minimum of 3 is no match (its bytes are literals)
            u32 len = min(m & 0xFFu, pos < segEnd ? segEnd - pos : 0u);
            const u32 off = m >> 8;
            u32 c = p > w0 ? p - w0 : 0;
            const u32 cstart = c;
            const u32 matchMask = __ballot_sync(ZL_FULL, len >= 3);
            if (!matchMask && cstart == 0 && w0 + 32 <= segEnd) {          // a window without any match (noise, low-entropy data): 32 literals
                lit[nlit + lane] = (u8)byte; atomicAdd(&hist[warp][byte], 1u);
                nlit += 32; p = w0 + 32;
                continue;
            }
            u32 takenMask = 0, myLL = 0, myOB = 0;
            for (;;) {
                const u32 mm = c < 32 ? (matchMask >> c) << c : 0u;
                if (!mm) break;
                const u32 c1 = (u32)__ffs((int)mm) - 1;
                u32 l = __shfl_sync(ZL_FULL, len, c1);
                u32 o = __shfl_sync(ZL_FULL, off, c1);
                u32 pos1 = w0 + c1, c2 = c1;
                if (l >= ZL_M_VERIFY && o <= pos1) l = zl_extend_match(wbase, bias, lastWord, segEnd, pos1, o, lane, ZL_M_VERIFY);   // (stops at segEnd; matches into the dictionary are not extended)
                if (repPref && reps.r0 && o != reps.r0) {
                    // Dictionary mode only: lanes 0..2 probe the most recent offset at pos1, pos1+1, pos1+2 (same window).  Such a
                    // match costs no offset bits; it is taken when it is at most 4 bytes shorter (cf. the repcode checks at
                    // ip+1 / ip+2 of zstd.c:29989, 30801).  Small dictionary-compressed inputs are dominated by offset cost.
                    const u32 q = pos1 + lane;
                    u32 rl = 0;
                    if (lane < 3 && c1 + lane < 32 && reps.r0 <= q && q + 4 <= segEnd) {
                        u32 qlo, qhi;
                        zl_ld8(wbase, bias + q, lastWord, qlo, qhi);
                        rl = zl_match_len(wbase, bias, lastWord, q, q - reps.r0, qlo, qhi, min(segEnd - q, ZL_M_CAP));
                    }
                    const u32 hit = __ballot_sync(ZL_FULL, rl >= 4 && rl + 4 >= l) & 7u;
                    if (hit) {
                        const u32 k2 = (u32)__ffs((int)hit) - 1;
                        c2 = c1 + k2; pos1 += k2; o = reps.r0;
                        l = __shfl_sync(ZL_FULL, rl, k2);
                        if (l == ZL_M_CAP) l = zl_extend_match(wbase, bias, lastWord, segEnd, pos1, o, lane, ZL_M_CAP);
                    }
                }
                const u32 ll = pos1 - anchor;
                u32 ob;                                          // most offsets are new: skip the repeat-offset cases with one (uniform) branch
                if (o != reps.r0 && o != reps.r1 && o != reps.r2 && o + 1 != reps.r0) { ob = o + 3; reps.r2 = reps.r1; reps.r1 = reps.r0; reps.r0 = o; }
                else ob = zl_rep_encode(reps, o, ll);
                if (lane == c2) { len = l; myLL = ll; myOB = ob; }
                takenMask |= 1u << c2;
                anchor = pos1 + l;
                c = c2 + l;
            }
            p = w0 + (c < 32 ? 32 : c);
            // literals of this window: positions from cstart on that no taken match covers
            const u32 below = takenMask & (ltMask | (1u << lane));
            const u32 t = below ? 31 - __clz((int)below) : 0;
            const u32 endRel = lane + len;                       // meaningful on taken lanes
            const u32 e = __shfl_sync(ZL_FULL, endRel, t);
            const bool covered = below != 0 && e > lane;
            const bool isLit = pos < segEnd && lane >= cstart && !covered;
            const u32 litMask = __ballot_sync(ZL_FULL, isLit);
            if (isLit) { lit[nlit + __popc(litMask & ltMask)] = (u8)byte; atomicAdd(&hist[warp][byte], 1u); }
            nlit += __popc(litMask);
            if ((takenMask >> lane) & 1) recs[nseq + __popc(takenMask & ltMask)] = zl_enc_rec(myLL, len, myOB);
            nseq += __popc(takenMask);
        }
    }
    }
    if (!segmented) {
        __syncwarp();
        for (u32 i = lane; i < 256; i += 32) histArena[(size_t)blk * 256 + i] = hist[warp][i];
        if (lane == 0) { metas[blk].nseq = nseq; metas[blk].nlit = nlit; }
        return;
    }
    if (lane == 0) { segSeq[warp] = nseq; segLit[warp] = nlit; segTail[warp] = segBeg < n ? segEnd - anchor : 0u; }
    __syncthreads();
    // ---- pack the segments: records and literals of segment w move down behind those of the segments before it (the CTA moves
    // one segment after the other, a step at a time, reads before writes: source and destination may overlap -- but only inside a
    // step, because data moves DOWN: what a step overwrites is below everything later steps read, so one barrier a step is enough);
    // the literals a segment ends with belong to the first sequence that follows (its litLength grows by that count)
    const u32 tid = threadIdx.x;
    u32 dSeq = segSeq[0], dLit = segLit[0], carry = segTail[0];
    for (u32 w = 1; w < ZL_PARSE_WARPS; w++) {
        const u32 ns = segSeq[w], nl = segLit[w];
        const u64* __restrict__ rs = recsAll + (size_t)w * ZL_PARSE_SEG_RECS;
        for (u32 i0 = 0; i0 < ns; i0 += ZL_PARSE_WARPS * 32) {
            const u32 i = i0 + tid;
            u64 r = i < ns ? rs[i] : 0ull;
            if (i == 0) r += carry;                              // litLength is the low field of a record (zl_enc_rec)
            __syncthreads();
            if (i < ns) recsAll[dSeq + i] = r;
        }
        const u8* __restrict__ ls = litAll + (size_t)w * ZL_PARSE_SEG;
        for (u32 i0 = 0; i0 < nl; i0 += 4 * ZL_PARSE_WARPS * 32) {
            u8 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { const u32 i = i0 + k * ZL_PARSE_WARPS * 32 + tid; v[k] = i < nl ? ls[i] : (u8)0; }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 4; k++) { const u32 i = i0 + k * ZL_PARSE_WARPS * 32 + tid; if (i < nl) litAll[dLit + i] = v[k]; }
        }
        carry = ns ? segTail[w] : carry + segTail[w];
        dSeq += ns; dLit += nl;
    }
    for (u32 i = tid; i < 256; i += ZL_PARSE_WARPS * 32) {
        u32 v = 0;
#pragma unroll
        for (int w = 0; w < ZL_PARSE_WARPS; w++) v += hist[w][i];
        histArena[(size_t)blk * 256 + i] = v;
    }
    if (tid == 0) { metas[blk].nseq = dSeq; metas[blk].nlit = dLit; }
}

// ---------------------------------------------------------------------------------------------- E3: literals
// warp per block.  Lane 0 builds the Huffman code from the histogram (serial, a few thousand steps); the streams are then
// encoded by all 32 lanes: 32 / nStreams chunks per stream, bit offsets from a segmented scan of the chunk bit counts.
__global__ void __launch_bounds__(ZL_ENT_WARPS * 32)
zl_k_enc_literals(const ZlEncBlock* __restrict__ blocks, u32 nblocks, const u8* __restrict__ litArena, u32 slotLit,
                  const u32* __restrict__ histArena, const ZlEncBlockMeta* __restrict__ metas, u32* __restrict__ streamArena,
                  u32 slotStreamWords, u32 streamCapWords, ZlEncBlockOut* __restrict__ outs, const ZlEncDictDev* __restrict__ dict)
{
    extern __shared__ __align__(16) u8 smraw[];
    ZlHufSm* fs = reinterpret_cast<ZlHufSm*>(smraw);
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 blk = blockIdx.x * ZL_ENT_WARPS + warp;
    if (blk >= nblocks) return;
    ZlHufSm& f = fs[warp];
    ZlEncBlockOut& o = outs[blk];
    const u8* lit = litArena + (size_t)blk * slotLit;
    const u32 nLit = metas[blk].nlit;
    for (u32 i = lane; i < 256; i += 32) f.count[i] = histArena[(size_t)blk * 256 + i];
    __syncwarp();
    if (lane == 0) zl_lit_plan(f, o, lit, nLit, (dict && dict->hasEntropy && (blocks[blk].flags & ZL_BLK_FIRST)) ? dict : nullptr);
    __syncwarp();
    if (f.ctl.mode != 2) return;
    const u32 ns = f.ctl.nStreams, per = 32 / ns, sIdx = lane / per, k = lane % per;
    u32* sOut = streamArena + (size_t)blk * slotStreamWords + (size_t)sIdx * streamCapWords;
    u32 cbeg, cend;
    zl_huf_chunk_range(f.ctl.sBeg[sIdx], f.ctl.sEnd[sIdx], k, per, &cbeg, &cend);
    const u32 bits = zl_huf_chunk_bits(f.nbBits, lit, cbeg, cend);
    // inclusive scan inside the group of `per` lanes; chunks with a larger k are written first
    u32 inc = bits;
    for (u32 d = 1; d < per; d <<= 1) { const u32 v = __shfl_up_sync(ZL_FULL, inc, d, per); if (k >= d) inc += v; }
    const u32 total = __shfl_sync(ZL_FULL, inc, per - 1, per);
    const u32 off = total - inc;
    const u32 words = (total + 1 + 31) >> 5;
    for (u32 i = k; i < words && i < streamCapWords; i += per) sOut[i] = 0;
    __syncwarp();
    u32 ovf = 0;
    zl_huf_encode_chunk(f.code, lit, cbeg, cend, sOut, streamCapWords, off, k == 0, &ovf);
    if (words > streamCapWords) ovf = 1;
    if (ovf) f.ctl.ovf = 1;
    if (k == 0) f.ctl.sBytes[sIdx] = (total + 1 + 7) >> 3;
    __syncwarp();
    if (lane == 0) zl_lit_finish(f, o);
}

// ---------------------------------------------------------------------------------------------- E4: sequences
// warp per block.  (a) all lanes: codes + three histograms; (b) lanes 0-2: table choice/construction, one table each;
// (c) rounds of 32 sequences, last sequence first: lanes 0-2 advance the three FSE state chains through shared memory,
// (d) then all lanes pack one sequence each: warp scan of the bit counts, OR into a shared staging window, whole words
// flushed coalesced.
#define ZL_ENC_CT_BYTES ((sizeof(ZlEncConst) + 15) & ~(size_t)15)
struct ZlSeqWarpSm {
    ZlSeqEncSm f;
    u32 stage[96];
    uint2 dd[3][32];         // (deltaNbBits, deltaFindState) of the round's sequences per table: looked up by all lanes, consumed by the chains
    u16 sbv[3][32];          // nbBits<<10 | bits emitted by each chain for the round's sequences
    u32 finalState[3];
    u32 pad;
};
__device__ __forceinline__ void zl_stage_put(u32* stage, u32 pos, u64 v, u32 nb)
{
    if (!nb) return;
    const u32 w = pos >> 5, s = pos & 31;
    const u64 lo = v << s;
    atomicOr(&stage[w], (u32)lo);
    if (s + nb > 32) atomicOr(&stage[w + 1], (u32)(lo >> 32));
    if (s + nb > 64) atomicOr(&stage[w + 2], (u32)(v >> (64 - s)));
}
__global__ void __launch_bounds__(ZL_ENT_WARPS * 32)
zl_k_enc_sequences(const ZlEncBlock* __restrict__ blocks, u32 nblocks, const u64* __restrict__ recArena, u32 slotRec,
                   const ZlEncBlockMeta* __restrict__ metas, u32* __restrict__ seqBitsArena, u32 slotSeqWords, u32 seqCapWords,
                   u32 sbitsWordOff, ZlEncBlockOut* __restrict__ outs, const ZlEncDictDev* __restrict__ dict)
{
    extern __shared__ __align__(16) u8 smraw[];
    ZlEncConst& K = *reinterpret_cast<ZlEncConst*>(smraw);
    ZlSeqWarpSm* ws = reinterpret_cast<ZlSeqWarpSm*>(smraw + ZL_ENC_CT_BYTES);
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {   const u32* s = reinterpret_cast<const u32*>(&c_enc);
        u32* d = reinterpret_cast<u32*>(&K);
        for (u32 i = threadIdx.x; i < sizeof(ZlEncConst) / 4; i += ZL_ENT_WARPS * 32) d[i] = s[i];
    }
    __syncthreads();
    const u32 blk = blockIdx.x * ZL_ENT_WARPS + warp;
    if (blk >= nblocks) return;
    ZlSeqWarpSm& w = ws[warp];
    ZlSeqEncSm& f = w.f;
    ZlEncBlockOut& o = outs[blk];
    const u64* __restrict__ recs = recArena + (size_t)blk * slotRec;
    const u32 nbSeq = metas[blk].nseq;
    u32* __restrict__ out = seqBitsArena + (size_t)blk * slotSeqWords;
    if (lane == 0) f.ctl.nbSeq = nbSeq;
    if (nbSeq == 0) { if (lane == 0) { zl_seq_write_head(f, o); o.seqBitsSize = 0; o.seqOvf = 0; } return; }
    // ---- (a) histograms
    for (u32 i = lane; i < 192; i += 32) (&f.count[0][0])[i] = 0;
    __syncwarp();
    for (u32 i = lane; i < nbSeq; i += 32) {
        const u64 r = recs[i];
        atomicAdd(&f.count[0][zl_seq_code(K, 0, r)], 1u);
        atomicAdd(&f.count[1][zl_seq_code(K, 1, r)], 1u);
        atomicAdd(&f.count[2][zl_seq_code(K, 2, r)], 1u);
    }
    __syncwarp();
    u32 maxSym[3];
#pragma unroll
    for (u32 t = 0; t < 3; t++) {
        const u32 m0 = __ballot_sync(ZL_FULL, f.count[t][lane] != 0), m1 = __ballot_sync(ZL_FULL, f.count[t][lane + 32] != 0);
        maxSym[t] = m1 ? 63 - __clz((int)m1) : 31 - __clz((int)m0);
    }
    // ---- (b) tables
    if (lane < 3) {
        const u32 t = lane;
        const u32 ms = t == 0 ? maxSym[0] : (t == 1 ? maxSym[1] : maxSym[2]);
        zl_seq_build_from_hist(f, t, nbSeq, ms, zl_seq_code(K, t, recs[nbSeq - 1]), K,
                               (dict && dict->hasEntropy && (blocks[blk].flags & ZL_BLK_FIRST)) ? dict : nullptr);
    }
    __syncwarp();
    if (lane == 0) zl_seq_write_head(f, o);
    // ---- (c)+(d) rounds of 32 sequences in writing order (last sequence first): every lane loads one record and
    // publishes its three codes; lanes 0-2 advance their state chain over the 32 codes (shared-memory reads only);
    // every lane then packs the bits of its own sequence.
    for (u32 i = lane; i < 96; i += 32) w.stage[i] = 0;
    u32 st = 0;                                             // lanes 0-2: running FSE state of their table
    const u32 myT = lane < 3 ? lane : 0;
    // predefined and dictionary tables are not built per block: all 32 lanes copy the prebuilt ones into the block's slots (a lane building
    // or copying one alone was most of the time of a 30-sequence block)
#pragma unroll
    for (u32 t = 0; t < 3; t++) {
        const u32 md = f.ctl.mode[t];
        if (md == 0 || md == 3) {
            const u16* sS = md == 0 ? g_encDef.state[t] : dict->state[t];
            const u32* sN = md == 0 ? g_encDef.dNb[t] : dict->dNb[t];
            const i32* sF = md == 0 ? g_encDef.dFS[t] : dict->dFS[t];
            const u32 sz = 1u << f.ctl.log[t];
            for (u32 i = lane; i < sz; i += 32) f.state[t][i] = sS[i];
            for (u32 i = lane; i < 64; i += 32) { f.dNb[t][i] = sN[i]; f.dFS[t][i] = sF[i]; }
        }
    }
    __syncwarp();
    const u16* tbl = f.state[myT];
    u32 bitPos = 0, ovf = 0;
    const u32 rounds = (nbSeq + 31) >> 5;
    u64 rNext = lane < nbSeq ? recs[nbSeq - 1 - lane] : 0;
    for (u32 rd = 0; rd <= rounds; rd++) {
        u64 a = 0; u32 na = 0, b = 0, nb2 = 0;
        if (rd < rounds) {
            const u32 j = (rd << 5) + lane;
            const u64 r = rNext;
            { const u32 jn = j + 32; rNext = jn < nbSeq ? recs[nbSeq - 1 - jn] : 0; }
            const u32 cLL = zl_seq_code(K, 0, r), cOF = zl_seq_code(K, 1, r), cML = zl_seq_code(K, 2, r);
            // everything about a step that does not depend on the running state is fetched here, by the sequence's own lane:
            // the chain lanes below then do one 8-byte load, the transition and one store per sequence
            if (j < nbSeq) {                                 // (lanes past the last sequence hold an empty record: its codes are no table indices)
                w.dd[0][lane] = make_uint2(f.dNb[0][cLL], (u32)f.dFS[0][cLL]);
                w.dd[1][lane] = make_uint2(f.dNb[1][cOF], (u32)f.dFS[1][cOF]);
                w.dd[2][lane] = make_uint2(f.dNb[2][cML], (u32)f.dFS[2][cML]);
            }
            __syncwarp();
            if (lane < 3) {
                const u32 cnt = min(32u, nbSeq - (rd << 5));
                u32 s0 = 0;
                const uint2* dd = w.dd[myT];
                if (rd == 0) {                              // the first sequence written only initialises the state (zstd.c:21166-21170)
                    st = zl_fse_init_state(tbl, dd[0].x, (i32)dd[0].y);
                    w.sbv[myT][0] = 0; s0 = 1;
                }
#pragma unroll 4
                for (u32 s = s0; s < cnt; s++) {
                    const uint2 d = dd[s];
                    const u32 e = zl_fse_step(tbl, d.x, (i32)d.y, st);
                    w.sbv[myT][s] = (u16)(((e >> 16) << 10) | (e & 0x3FF));
                }
            }
            __syncwarp();
            if (j < nbSeq) {
                const u32 eLL = w.sbv[0][lane], eOF = w.sbv[1][lane], eML = w.sbv[2][lane];
                a = eOF & 0x3FF; na = eOF >> 10;
                a |= (u64)(eML & 0x3FF) << na; na += eML >> 10;
                a |= (u64)(eLL & 0x3FF) << na; na += eLL >> 10;
                u32 x, v;
                v = zl_seq_extra(K, 0, r, cLL, &x); a |= (u64)v << na; na += x;
                v = zl_seq_extra(K, 2, r, cML, &x); a |= (u64)v << na; na += x;
                b = zl_seq_extra(K, 1, r, cOF, &nb2);
            }
        } else {                                            // closing: ML, OF, LL states then the end mark (zstd.c:21227-21231, 2334)
            if (lane < 3) w.finalState[lane] = st;
            __syncwarp();
            if (lane == 0) {
                const u32 lML = f.ctl.log[2], lOF = f.ctl.log[1], lLL = f.ctl.log[0];
                a = w.finalState[2] & ((1u << lML) - 1); na = lML;
                a |= (u64)(w.finalState[1] & ((1u << lOF) - 1)) << na; na += lOF;
                a |= (u64)(w.finalState[0] & ((1u << lLL) - 1)) << na; na += lLL;
                a |= (u64)1 << na; na += 1;
            }
        }
        const u32 mine = na + nb2;
        u32 inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 v = __shfl_up_sync(ZL_FULL, inc, d); if ((int)lane >= d) inc += v; }
        const u32 roundBits = __shfl_sync(ZL_FULL, inc, 31);
        const u32 start = (bitPos & 31) + inc - mine;
        zl_stage_put(w.stage, start, a, na);
        zl_stage_put(w.stage, start + na, (u64)b, nb2);
        __syncwarp();
        const u32 endPos = (bitPos & 31) + roundBits, nFull = endPos >> 5, wbase = bitPos >> 5;
        for (u32 kx = lane; kx < nFull; kx += 32) { if (wbase + kx < seqCapWords) out[wbase + kx] = w.stage[kx]; else ovf = 1; }
        __syncwarp();
        const u32 carry = w.stage[nFull];
        __syncwarp();
        for (u32 kx = lane; kx <= nFull; kx += 32) w.stage[kx] = 0;
        __syncwarp();
        if (lane == 0) w.stage[0] = carry;
        __syncwarp();
        bitPos += roundBits;
    }
    if (lane == 0 && (bitPos & 31)) { if ((bitPos >> 5) < seqCapWords) out[bitPos >> 5] = w.stage[0]; else ovf = 1; }
    ovf = __any_sync(ZL_FULL, ovf != 0) ? 1u : 0u;
    if (lane == 0) {
        o.seqOvf = ovf;
        o.seqBitsSize = ovf ? 0u : (bitPos + 7) >> 3;
    }
}

// ---------------------------------------------------------------------------------------------- E5: plan
__global__ void __launch_bounds__(128)
zl_k_enc_plan(const ZlEncFrame* __restrict__ frames, u32 nframes, const ZlEncBlock* __restrict__ blocks,
              const ZlEncBlockMeta* __restrict__ metas, const ZlEncBlockOut* __restrict__ outs, ZlEncBlockPlan* __restrict__ plans,
              u64* __restrict__ results)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nframes) return;
    const ZlEncFrame fr = frames[i];
    u64 pos = fr.hdrSize;
    for (u32 k = 0; k < fr.nblocks; k++) {
        const u32 bi = fr.firstBlock + k;
        const u32 n = blocks[bi].srcSize;
        const u32 payload = n >= 7 ? zl_enc_block_payload(outs[bi], n, metas[bi].nseq) : 0u;      // zstd.c:25725
        ZlEncBlockPlan pl;
        pl.dstOff = pos; pl.type = payload ? 2u : 0u; pl.size = payload ? payload : n;
        plans[bi] = pl;
        pos += 3 + pl.size;
    }
    if (fr.checksumFlag) pos += 4;
    results[i] = pos <= fr.dstCap ? pos : (u64)0 - (u64)ZL_E_dstSize_tooSmall;
}

// ---------------------------------------------------------------------------------------------- E6: assemble
__device__ __forceinline__ void zl_copy_small(u8* dst, const u8* src, u32 n, u32 lane)
{
    for (u32 i = lane; i < n; i += 32) dst[i] = src[i];
}
__global__ void __launch_bounds__(ZL_ASM_WARPS * 32)
zl_k_enc_assemble(const ZlEncFrame* __restrict__ frames, const ZlEncBlock* __restrict__ blocks, u32 nblocks,
                  const ZlEncBlockPlan* __restrict__ plans, const ZlEncBlockOut* __restrict__ outs, const u8* __restrict__ litArena,
                  u32 slotLit, const u32* __restrict__ streamArena, u32 slotStreamWords, u32 streamCapWords,
                  const u32* __restrict__ seqBitsArena, u32 slotSeqWords, const u64* __restrict__ results, const u64* __restrict__ xxh)
{
    const u32 lane = threadIdx.x & 31;
    const u32 blk = blockIdx.x * ZL_ASM_WARPS + (threadIdx.x >> 5);
    if (blk >= nblocks) return;
    const ZlEncBlock b = blocks[blk];
    const ZlEncFrame& fr = frames[b.frame];
    if (results[b.frame] > (u64)0 - (u64)ZL_E_maxCode) return;        // destination too small: nothing is written
    const ZlEncBlockPlan pl = plans[blk];
    u8* dst = fr.dst + pl.dstOff;
    const bool last = (b.flags & ZL_BLK_LAST) != 0;
    if (b.flags & ZL_BLK_FIRST) zl_copy_small(fr.dst, fr.hdr, fr.hdrSize, lane);
    if (lane == 0) zl_write_block_header(dst, last ? 1u : 0u, pl.type, pl.size);
    if (last && fr.checksumFlag && lane < 4) dst[3 + pl.size + lane] = (u8)((u32)xxh[b.frame] >> (8 * lane));   // zstd.c:27760-27766
    u8* op = dst + 3;
    if (pl.type == 0) { zl_warp_copy(op, b.src, b.srcSize, lane); return; }
    const ZlEncBlockOut& o = outs[blk];
    zl_copy_small(op, o.litHead, o.litHeadSize, lane); op += o.litHeadSize;
    if (o.litBodyMode == 1) { zl_warp_copy(op, litArena + (size_t)blk * slotLit, o.nLit, lane); op += o.nLit; }
    else if (o.litBodyMode == 2) {
        for (u32 k = 0; k < o.nStreams; k++) {
            zl_warp_copy(op, reinterpret_cast<const u8*>(streamArena + (size_t)blk * slotStreamWords + (size_t)k * streamCapWords), o.sBytes[k], lane);
            op += o.sBytes[k];
        }
    }
    zl_copy_small(op, o.seqHead, o.seqHeadSize, lane); op += o.seqHeadSize;
    zl_warp_copy(op, reinterpret_cast<const u8*>(seqBitsArena + (size_t)blk * slotSeqWords), o.seqBitsSize, lane);
}

// gather of variable-size frames into one contiguous stream (zl_compress_split)
__global__ void __launch_bounds__(ZL_ASM_WARPS * 32)
zl_k_gather(const u8* const* __restrict__ srcs, const u64* __restrict__ sizes, const u64* __restrict__ offs, u8* __restrict__ dst, u32 n)
{
    const u32 lane = threadIdx.x & 31;
    const u32 i = blockIdx.x * ZL_ASM_WARPS + (threadIdx.x >> 5);
    if (i >= n) return;
    zl_warp_copy(dst + offs[i], srcs[i], (u32)sizes[i], lane);
}

// ---------------------------------------------------------------------------------------------- dictionary training
// warp per parsed block: literal histogram and LL / ML / OF code counts summed into one table of 512 counters
// ([0,256) literals, [256,292) LL, [292,345) ML, [345,377) OF) -- the statistics of ZDICT_countEStats (zstd.c:50440-50510)
__global__ void __launch_bounds__(ZL_PARSE_WARPS * 32)
zl_k_dict_stats(u32 nblocks, const u64* __restrict__ recArena, u32 slotRec, const u32* __restrict__ histArena,
                const ZlEncBlockMeta* __restrict__ metas, u32* stats)
{
    __shared__ u32 sh[ZL_PARSE_WARPS][128];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 blk = blockIdx.x * ZL_PARSE_WARPS + warp;
    if (blk >= nblocks) return;
    for (u32 i = lane; i < 128; i += 32) sh[warp][i] = 0;
    __syncwarp();
    for (u32 i = lane; i < 256; i += 32) { const u32 v = histArena[(size_t)blk * 256 + i]; if (v) atomicAdd(stats + i, v); }
    const u64* __restrict__ recs = recArena + (size_t)blk * slotRec;
    const u32 nseq = metas[blk].nseq;
    for (u32 r = lane; r < nseq; r += 32) {
        const u64 rec = recs[r];
        atomicAdd(&sh[warp][zl_seq_code(c_enc, 0, rec)], 1u);
        atomicAdd(&sh[warp][36 + zl_seq_code(c_enc, 2, rec)], 1u);
        atomicAdd(&sh[warp][89 + (zl_seq_code(c_enc, 1, rec) & 31u)], 1u);
    }
    __syncwarp();
    for (u32 i = lane; i < 121; i += 32) { const u32 v = sh[warp][i]; if (v) atomicAdd(stats + 256 + i, v); }
}

// ---------------------------------------------------------------------------------------------- launcher
size_t zl_enc_match_smem(const ZlEncParams& P) { return ((size_t)2 << P.hlogS) + (P.hlogL ? ((size_t)2 << P.hlogL) : 0) + 2 * ZL_MATCH_SCRATCH; }

cudaError_t zl_launch_encode(const ZlEncodeLaunch& L, cudaStream_t st)
{
    if (L.nblocks == 0 && L.nframes == 0) return cudaSuccess;
    cudaEvent_t* ev = L.stageEv;
    const size_t smM = zl_enc_match_smem(L.params);
    const size_t smL = ZL_ENT_WARPS * sizeof(ZlHufSm);
    const size_t smS = ZL_ENC_CT_BYTES + ZL_ENT_WARPS * sizeof(ZlSeqWarpSm);
    cudaError_t e;
    const bool useDict = L.dict != nullptr;
    if (L.params.hlogL) e = useDict ? cudaFuncSetAttribute(zl_k_match<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smM)
                                    : cudaFuncSetAttribute(zl_k_match<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smM);
    else e = useDict ? cudaFuncSetAttribute(zl_k_match<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smM)
                     : cudaFuncSetAttribute(zl_k_match<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(zl_k_enc_literals, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smL);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(zl_k_enc_sequences, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smS);
    if (e != cudaSuccess) return e;
    const u32 nb = L.nblocks;
    if (ev) cudaEventRecord(ev[0], st);
    if (nb && L.far) zl_k_far_build<<<nb * ZL_FAR_CHUNKS, 256, 0, st>>>(L.blocks, L.frames, L.far);
    if (nb) {
        // blocks of at most ZL_SMALL_BLOCK bytes go to the warp-per-block kernel, the others to the CTA-per-block one; each skips the other's
        const u32 skipSmall = L.nSmall ? 1u : 0u;
        if (L.nSmall) {
            const u32 gs = (nb + ZL_SMALL_WARPS - 1) / ZL_SMALL_WARPS;
            if (L.params.hlogL) {
                if (useDict) zl_k_match_small<true, true><<<gs, ZL_SMALL_WARPS * 32, 0, st>>>(L.blocks, nb, L.M, L.slotM, L.params, L.dict);
                else zl_k_match_small<true, false><<<gs, ZL_SMALL_WARPS * 32, 0, st>>>(L.blocks, nb, L.M, L.slotM, L.params, nullptr);
            } else {
                if (useDict) zl_k_match_small<false, true><<<gs, ZL_SMALL_WARPS * 32, 0, st>>>(L.blocks, nb, L.M, L.slotM, L.params, L.dict);
                else zl_k_match_small<false, false><<<gs, ZL_SMALL_WARPS * 32, 0, st>>>(L.blocks, nb, L.M, L.slotM, L.params, nullptr);
            }
        }
        if (L.nSmall < nb) {
            if (L.params.hlogL) {
                if (useDict) zl_k_match<true, true><<<nb, ZL_MATCH_WARPS * 32, smM, st>>>(L.blocks, L.M, L.slotM, L.params, L.dict, skipSmall);
                else zl_k_match<true, false><<<nb, ZL_MATCH_WARPS * 32, smM, st>>>(L.blocks, L.M, L.slotM, L.params, nullptr, skipSmall);
            } else {
                if (useDict) zl_k_match<false, true><<<nb, ZL_MATCH_WARPS * 32, smM, st>>>(L.blocks, L.M, L.slotM, L.params, L.dict, skipSmall);
                else zl_k_match<false, false><<<nb, ZL_MATCH_WARPS * 32, smM, st>>>(L.blocks, L.M, L.slotM, L.params, nullptr, skipSmall);
            }
        }
        if (L.far) zl_k_far_match<<<nb * ZL_FAR_CHUNKS, 256, 0, st>>>(L.blocks, L.frames, L.far, L.M, L.slotM, L.params);
    }
    if (ev) cudaEventRecord(ev[1], st);
    {   const u32 segmented = L.maxBlock > ZL_PARSE_SEG ? 1u : 0u;
        if (nb) zl_k_parse<<<segmented ? nb : (nb + ZL_PARSE_WARPS - 1) / ZL_PARSE_WARPS, ZL_PARSE_WARPS * 32, 0, st>>>(L.blocks, nb, L.M, L.slotM, L.recs, L.slotRec, L.lit, L.slotLit,
                                                                                                                    L.hist, L.metas, L.dict, segmented); }
    if (ev) cudaEventRecord(ev[2], st);
    if (L.stats) {                                                  // dictionary training stops after the parse
        if (nb) zl_k_dict_stats<<<(nb + ZL_PARSE_WARPS - 1) / ZL_PARSE_WARPS, ZL_PARSE_WARPS * 32, 0, st>>>(nb, L.recs, L.slotRec, L.hist, L.metas, L.stats);
        if (ev) for (int k = 3; k <= 5; k++) cudaEventRecord(ev[k], st);
        return cudaGetLastError();
    }
    const u32 gq = (nb + ZL_ENT_WARPS - 1) / ZL_ENT_WARPS;
    // the stream / bitstream buffers reuse the M arena (dead after the parse): [streams | sequence bits] per block slot
    u32* streamArena = L.M;
    u32* seqArena = L.M + L.streamWordsPerBlock;
    // the two entropy kernels write disjoint fields of `outs` and disjoint halves of the M arena: with a side stream they run together
    const bool useSide = L.side != nullptr && nb;
    if (useSide) { cudaEventRecord(L.sideFork, st); cudaStreamWaitEvent(L.side, L.sideFork, 0); }
    if (nb) zl_k_enc_literals<<<gq, ZL_ENT_WARPS * 32, smL, st>>>(L.blocks, nb, L.lit, L.slotLit, L.hist, L.metas, streamArena, L.slotM, L.streamCapWords, L.outs, L.dict);
    if (ev) cudaEventRecord(ev[3], st);
    if (nb) zl_k_enc_sequences<<<gq, ZL_ENT_WARPS * 32, smS, useSide ? L.side : st>>>(L.blocks, nb, L.recs, L.slotRec, L.metas, seqArena, L.slotM, L.seqCapWords, L.seqCapWords, L.outs, L.dict);
    if (useSide) { cudaEventRecord(L.sideJoin, L.side); cudaStreamWaitEvent(st, L.sideJoin, 0); }
    if (ev) cudaEventRecord(ev[4], st);
    zl_k_enc_plan<<<(L.nframes + 127) / 128, 128, 0, st>>>(L.frames, L.nframes, L.blocks, L.metas, L.outs, L.plans, L.results);
    if (nb) zl_k_enc_assemble<<<(nb + ZL_ASM_WARPS - 1) / ZL_ASM_WARPS, ZL_ASM_WARPS * 32, 0, st>>>(L.frames, L.blocks, nb, L.plans, L.outs, L.lit, L.slotLit, streamArena, L.slotM,
                                                                                             L.streamCapWords, seqArena, L.slotM, L.results, L.xxh);
    if (ev) cudaEventRecord(ev[5], st);
    return cudaGetLastError();
}

// The per-frame results go back to the host through a kernel that stores into the pinned (device-mapped) result array, not through a
// cudaMemcpyAsync: a copy engine serves its requests in submission order, and in the pipelined calls the few KB of results queued up behind
// the previous chunk's hundreds of MB of output (measured: the last, short chunk of a 4 GiB call waited 8 ms for them).
__global__ void zl_k_results_out(const u64* __restrict__ src, u64* __restrict__ hostDst, u32 n)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hostDst[i] = src[i];
}
cudaError_t zl_launch_results_out(const u64* src, u64* hostDst, u32 n, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    zl_k_results_out<<<(n + 255) / 256, 256, 0, st>>>(src, hostDst, n);
    return cudaGetLastError();
}

cudaError_t zl_launch_gather(const u8* const* srcs, const u64* sizes, const u64* offs, u8* dst, u32 n, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    zl_k_gather<<<(n + ZL_ASM_WARPS - 1) / ZL_ASM_WARPS, ZL_ASM_WARPS * 32, 0, st>>>(srcs, sizes, offs, dst, n);
    return cudaGetLastError();
}
