// zl_launch.h -- host-visible launcher interface between zl_api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include "zl_common.cuh"

#define ZL_QUADS_PER_WARP 8
#define ZL_EXEC_WARPS 4
#define ZL_DEC_STAGES 4      // literals, sequences, execute, checksum

// digested dictionary in device memory (zstd.c:42053-42159, ZSTD_DDict): entropy tables in the packed
// cell formats of zl_dec_entropy.cuh plus the raw content that acts as history before the frame.
struct ZlDictDev {
    u32 fseLL[512];
    u32 fseML[512];
    u32 fseOF[256];
    u16 huf[2048];
    u32 hufLog;
    u32 tlog[3];
    u32 rep[3];
    u32 hasEntropy;
    u32 dictID;
    u32 contentSize;
    const u8* content;     // device pointer
};

struct ZlDecodeLaunch {
    const ZlFrameDesc* descs;
    ZlFrameInfo* infos;
    ZlBlockHdr* hdrArena;
    u64* recArena;
    u8* litArena;
    u64* results;
    u32 nframes;
    int verifyChecksum;
    const ZlDictDev* dict;   // device pointer or null
    cudaEvent_t* stageEv;    // null, or ZL_DEC_STAGES + 1 events recorded around each kernel (per-kernel timing for bench.py)
};

size_t zl_literals_smem_bytes();
size_t zl_sequences_smem_bytes();
cudaError_t zl_launch_decode(const ZlDecodeLaunch& L, cudaStream_t st);
cudaError_t zl_launch_xxh64(const u8* const* ptrs, const u32* sizes, u64* out, u32 n, cudaStream_t st);
