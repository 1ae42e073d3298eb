// zl_host.h -- host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <vector>
#include "zl_common.cuh"
#include "zl_launch.h"

#define ZL_ERROR(name) ((size_t)0 - (size_t)ZL_E_##name)
static inline bool zl_is_error(size_t r) { return r > (size_t)0 - (size_t)ZL_E_maxCode; }

// grow-only device / pinned-host buffers owned by a context
struct ZlDevBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool reserve(size_t n)
    {
        if (n <= cap) return true;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + (n >> 3) + 4096;
        want = (want + 0xFFFFF) & ~(size_t)0xFFFFF;
        if (cudaMalloc(&p, want) != cudaSuccess) { (void)cudaGetLastError(); if (cudaMalloc(&p, n + 256) != cudaSuccess) { (void)cudaGetLastError(); p = nullptr; return false; } want = n + 256; }
        cap = want;
        return true;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct ZlPinBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool reserve(size_t n)
    {
        if (n <= cap) return true;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + (n >> 2) + 4096;
        if (cudaMallocHost(&p, want) != cudaSuccess) { (void)cudaGetLastError(); p = nullptr; return false; }
        cap = want;
        return true;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// host copy of the frame header logic (zstd.c:41050-41152); returns 0, an error, or bytes wanted
struct ZlHostFrameHeader {
    unsigned long long contentSize, windowSize;
    unsigned blockSizeMax, skippable, headerSize, dictID, checksumFlag;
};
size_t zl_host_frame_header(ZlHostFrameHeader* h, const void* src, size_t srcSize);
size_t zl_host_find_frame_size(const void* src, size_t srcSize, unsigned* nblocksOut);

// contiguous-run staging between host buffers and a device arena
struct ZlRun { size_t first, count; const uint8_t* hbase; size_t bytes; size_t devOff; };
