// zl_dec_kernels.cu -- the decode kernels and their launchers (sm_100a).
//   K1a zl_k_literals / K1b zl_k_sequences   quad-per-frame entropy decode -> literal / record / header arenas
//   K2 zl_k_execute   warp-per-frame sequence execution -> frame output
//   K3 zl_k_checksum  quad-per-frame XXH64 of the output, compared with the frame trailer
#include "zl_dec_entropy.cuh"
#include "zl_dec_exec.cuh"
#include "zl_launch.h"

__constant__ ZlConstTables c_tables = {
    ZL_LL_BASE_INIT, ZL_ML_BASE_INIT, ZL_LL_BITS_INIT, ZL_ML_BITS_INIT,
    ZL_LL_DEFNORM_INIT, ZL_ML_DEFNORM_INIT, ZL_OF_DEFNORM_INIT};

// ---- K1a: literals ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
zl_k_literals(const ZlFrameDesc* __restrict__ descs, ZlFrameInfo* __restrict__ infos, ZlBlockHdr* hdrArena,
              u8* litArena, u32 nframes, const ZlDictDev* dict)
{
    extern __shared__ __align__(16) u8 smraw[];
    ZlLitSm* fs = reinterpret_cast<ZlLitSm*>(smraw);
    const u32 lane = threadIdx.x, quad = lane >> 2, q = lane & 3;
    const u32 qmask = 0xFu << (quad * 4);
    const u32 frame = blockIdx.x * ZL_QUADS_PER_WARP + quad;
    if (frame >= nframes) return;
    ZlLitSm& f = fs[quad];
    const ZlFrameDesc d = descs[frame];
    ZlFrameInfo& info = infos[frame];
    ZlBlockHdr* hdrs = hdrArena + d.hdrBase;
    u8* lits = litArena + d.litBase;
    const u32 bias = (u32)(((size_t)d.src) & 3);
    const u32* wbase = reinterpret_cast<const u32*>(d.src - bias);
    const bool seed = dict && dict->hasEntropy;
    if (q == 0) {
        zl_lit_begin_frame(f, d, info, dict ? dict->dictID : 0u);
        if (seed && !f.ctl.err) { f.ctl.hufValid = 1; f.ctl.hufLog = dict->hufLog; }      // zstd.c:42140-42159
    }
    if (seed) for (u32 i = q; i < 2048; i += 4) f.huf[i] = dict->huf[i];
    // Lane 0 writes f.ctl between quad barriers; the other lanes snapshot what they need right after a
    // barrier and a second barrier keeps lane 0 from overwriting it before everyone has read it.
    for (;;) {
        __syncwarp(qmask);
        if (q == 0) zl_lit_block_head(f, d, info, hdrs, wbase, bias);
        __syncwarp(qmask);
        const u32 done = f.ctl.done, fill = f.ctl.needHufFill, ns = f.ctl.nStreams;
        __syncwarp(qmask);
        if (done) break;
        if (fill) { zl_huf_fill(f, q); __syncwarp(qmask); }
        if (q < ns)
            f.ctl.sErr[q] = zl_huf_stream(f.huf, f.ctl.hufLog, wbase, bias, f.ctl.sBeg[q], f.ctl.sEnd[q],
                                          lits + f.ctl.sOut[q], f.ctl.sLen[q]);
    }
}

// ---- K1b: sequences -----------------------------------------------------------------------------------------
#define ZL_XTAB_BYTES ((ZL_XTAB_WORDS * 4 + 15) & ~15)
__global__ void __launch_bounds__(32)
zl_k_sequences(const ZlFrameDesc* __restrict__ descs, ZlFrameInfo* __restrict__ infos, ZlBlockHdr* hdrArena,
               u64* recArena, i16* normArena, u32 nframes, const ZlDictDev* dict)
{
    extern __shared__ __align__(16) u8 smraw[];
    u32* xtab = reinterpret_cast<u32*>(smraw);
    ZlSeqSm* fs = reinterpret_cast<ZlSeqSm*>(smraw + ZL_XTAB_BYTES);
    const ZlConstTables& ct = c_tables;          // constant memory: only the header parser and the rare generic step index it
    const u32 lane = threadIdx.x, quad = lane >> 2, q = lane & 3;
    const u32 qmask = 0xFu << (quad * 4);
    for (u32 i = lane; i < ZL_XTAB_WORDS; i += 32)
        xtab[i] = i < 36 ? (c_tables.llBase[i] | ((u32)c_tables.llBits[i] << 24)) : (c_tables.mlBase[i - 36] | ((u32)c_tables.mlBits[i - 36] << 24));
    __syncwarp();
    const u32 frame = blockIdx.x * ZL_QUADS_PER_WARP + quad;
    if (frame >= nframes) return;
    ZlSeqSm& f = fs[quad];
    const ZlFrameDesc d = descs[frame];
    ZlFrameInfo& info = infos[frame];
    ZlBlockHdr* hdrs = hdrArena + d.hdrBase;
    u64* recs = recArena + d.recBase;
    i16* norm = normArena + (size_t)frame * (3 * ZL_NORM_STRIDE);
    const u32 bias = (u32)(((size_t)d.src) & 3);
    const u32* wbase = reinterpret_cast<const u32*>(d.src - bias);
    const u32 nblocks = info.nblocks;
    const bool seed = dict && dict->hasEntropy;
    if (q == 0) {
        zl_seq_begin_frame(f, info);
        if (seed) {                                                   // zstd.c:42140-42159: tables from the dictionary (repcodes: K2)
            f.ctl.fseValid = 7;
            f.ctl.tlog[0] = dict->tlog[0]; f.ctl.tlog[1] = dict->tlog[1]; f.ctl.tlog[2] = dict->tlog[2];
        }
    }
    if (seed) {
        for (u32 i = q; i < 512; i += 4) { f.fseLL[i] = dict->fseLL[i]; f.fseML[i] = dict->fseML[i]; }
        for (u32 i = q; i < 256; i += 4) f.fseOF[i] = dict->fseOF[i];
    }
    __syncwarp(qmask);
    for (u32 b = 0; b < nblocks; b++) {
        ZlBlockHdr h = hdrs[b];
        if ((h.flags & 3) != 2) continue;
        if (q == 0) zl_seq_head(f, d, h, ct, norm);
        __syncwarp(qmask);
        const u32 build = f.ctl.err ? 0u : f.ctl.needBuild;
        __syncwarp(qmask);
        if (build) { if (q < 3) zl_seq_fse_build(f, q, norm); __syncwarp(qmask); }
        if (q == 0) { zl_seq_decode(f, d, h, recs, wbase, bias, ct, xtab); hdrs[b] = h; }
    }
    if (q == 0) zl_seq_finish_frame(f, info);
}

template <bool kDict>
__global__ void __launch_bounds__(ZL_EXEC_WARPS * 32)
zl_k_execute(const ZlFrameDesc* __restrict__ descs, ZlFrameInfo* __restrict__ infos,
             const ZlBlockHdr* __restrict__ hdrArena, const u64* __restrict__ recArena,
             const u8* __restrict__ litArena, u64* __restrict__ results, u32 nframes, const ZlDictDev* dict)
{
    __shared__ u32 xtab[ZL_XTAB_WORDS];
    for (u32 i = threadIdx.x; i < ZL_XTAB_WORDS; i += ZL_EXEC_WARPS * 32)
        xtab[i] = i < 36 ? (c_tables.llBase[i] | ((u32)c_tables.llBits[i] << 24)) : (c_tables.mlBase[i - 36] | ((u32)c_tables.mlBits[i - 36] << 24));
    __syncthreads();
    const u32 lane = threadIdx.x & 31;
    const u32 frame = blockIdx.x * ZL_EXEC_WARPS + (threadIdx.x >> 5);
    if (frame >= nframes) return;
    const ZlFrameInfo info = infos[frame];
    if (info.err) { if (lane == 0) results[frame] = (u64)0 - (u64)info.err; return; }
    const ZlFrameDesc d = descs[frame];
    const ZlBlockHdr* hdrs = hdrArena + d.hdrBase;
    const u8* dictContent = kDict ? dict->content : nullptr;
    const u32 dictSize = kDict ? dict->contentSize : 0u;
    u32 hist[3] = {1, 4, 8};                                          // zstd.c:15416
    if (kDict && dict->hasEntropy) { hist[0] = dict->rep[0]; hist[1] = dict->rep[1]; hist[2] = dict->rep[2]; }   // zstd.c:42140-42159
    u32 op = 0, err = 0;
    for (u32 b = 0; b < info.nblocks && !err; b++) {
        const ZlBlockHdr h = hdrs[b];
        const u32 type = h.flags & 3;
        const u32 room = d.dstCap - op;
        if (type != 2) {                                              // raw / RLE block (zstd.c:41497-41520)
            if (h.regenSize > room) { err = ZL_E_dstSize_tooSmall; break; }
            if (type == 0) zl_warp_copy(d.dst + op, d.src + h.srcOff, h.regenSize, lane);
            else zl_warp_fill(d.dst + op, (h.flags >> 8) & 0xFF, h.regenSize, lane);
            op += h.regenSize;
        } else {
            const u32 litMode = (h.flags >> 4) & 3;
            const u8* lit = litMode == 0 ? d.src + h.srcOff : litArena + d.litBase + h.litOff;
            u32 cap = room, capErr = ZL_E_dstSize_tooSmall, regen = 0;
            if (cap > ZL_BLOCKSIZE_MAX) { cap = ZL_BLOCKSIZE_MAX; capErr = ZL_E_corruption_detected; }
            err = zl_exec_block<kDict>(d.dst, op, cap, capErr, h, lit, (h.flags >> 8) & 0xFF, litMode, recArena + d.recBase + h.recOff,
                                       dictContent, dictSize, hist, xtab, lane, regen);
            op += regen;
        }
        __syncwarp();
    }
    if (!err && info.contentSize != ~0ull && info.contentSize != (u64)op) err = ZL_E_corruption_detected;      // zstd.c:41646
    if (lane == 0) {
        results[frame] = err ? (u64)0 - (u64)err : (u64)op;
        infos[frame].err = err; infos[frame].totalOut = op;
    }
}

__global__ void __launch_bounds__(128)
zl_k_checksum(const ZlFrameDesc* __restrict__ descs, const ZlFrameInfo* __restrict__ infos, u64* __restrict__ results, u32 nframes)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 frame = t >> 2, q = t & 3, lane = threadIdx.x & 31;
    const u32 qbase = lane & ~3u, qmask = 0xFu << qbase;
    if (frame >= nframes) return;
    const ZlFrameInfo info = infos[frame];
    if (info.err || !info.checksumFlag) return;
    const u64 h = zl_quad_xxh64(descs[frame].dst, info.totalOut, q, qmask, qbase);
    if (q == 0 && (u32)h != info.checksum) results[frame] = (u64)0 - (u64)ZL_E_checksum_wrong;   // zstd.c:41650-41657
}

// generic XXH64 of independent buffers (used by the compressor for frame trailers)
__global__ void __launch_bounds__(128)
zl_k_xxh64(const u8* const* __restrict__ ptrs, const u32* __restrict__ sizes, u64* __restrict__ out, u32 n)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 i = t >> 2, q = t & 3, lane = threadIdx.x & 31;
    const u32 qbase = lane & ~3u, qmask = 0xFu << qbase;
    if (i >= n) return;
    const u64 h = zl_quad_xxh64(ptrs[i], sizes[i], q, qmask, qbase);
    if (q == 0) out[i] = h;
}

// ---- launchers ---------------------------------------------------------------------------------------
size_t zl_literals_smem_bytes() { return ZL_QUADS_PER_WARP * sizeof(ZlLitSm); }
size_t zl_sequences_smem_bytes() { return ZL_XTAB_BYTES + ZL_QUADS_PER_WARP * sizeof(ZlSeqSm); }

cudaError_t zl_launch_decode(const ZlDecodeLaunch& L, cudaStream_t st)
{
    if (L.nframes == 0) return cudaSuccess;
    const size_t smA = zl_literals_smem_bytes(), smB = zl_sequences_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(zl_k_literals, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(zl_k_sequences, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smB);
    if (e != cudaSuccess) return e;
    const u32 g1 = (L.nframes + ZL_QUADS_PER_WARP - 1) / ZL_QUADS_PER_WARP;
    cudaEvent_t* ev = L.stageEv;
    if (ev) cudaEventRecord(ev[0], st);
    zl_k_literals<<<g1, 32, smA, st>>>(L.descs, L.infos, L.hdrArena, L.litArena, L.nframes, L.dict);
    if (ev) cudaEventRecord(ev[1], st);
    zl_k_sequences<<<g1, 32, smB, st>>>(L.descs, L.infos, L.hdrArena, L.recArena, L.normArena, L.nframes, L.dict);
    if (ev) cudaEventRecord(ev[2], st);
    const u32 g2 = (L.nframes + ZL_EXEC_WARPS - 1) / ZL_EXEC_WARPS;
    if (L.dict)
        zl_k_execute<true><<<g2, ZL_EXEC_WARPS * 32, 0, st>>>(L.descs, L.infos, L.hdrArena, L.recArena, L.litArena, L.results, L.nframes, L.dict);
    else
        zl_k_execute<false><<<g2, ZL_EXEC_WARPS * 32, 0, st>>>(L.descs, L.infos, L.hdrArena, L.recArena, L.litArena, L.results, L.nframes, nullptr);
    if (ev) cudaEventRecord(ev[3], st);
    if (L.verifyChecksum) {
        const u32 g3 = (L.nframes * 4 + 127) / 128;
        zl_k_checksum<<<g3, 128, 0, st>>>(L.descs, L.infos, L.results, L.nframes);
    }
    if (ev) cudaEventRecord(ev[4], st);
    return cudaGetLastError();
}

cudaError_t zl_launch_xxh64(const u8* const* ptrs, const u32* sizes, u64* out, u32 n, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    zl_k_xxh64<<<(n * 4 + 127) / 128, 128, 0, st>>>(ptrs, sizes, out, n);
    return cudaGetLastError();
}
