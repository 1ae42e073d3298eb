// zl_dict_train.cuh -- dictionary training on the GPU (included at the end of zl_api_compress.cu).
//
// Replaces what src/dictionaries.c:60-208 calls: ZDICT_trainFromBuffer (zstd.c:50979 -> fastCOVER with d = 8, f = 20,
// steps = 4, split 75/25, zstd.c:49849) and ZDICT_optimizeTrainFromBuffer_cover (zstd.c:46957; here run with the same
// hashed d-mer machinery).  The method is fastCOVER's, re-cut for a data-parallel machine:
//   T1 zl_k_cover_hash    thread per position: the d-mer hash of every position of the training buffer (zstd.c:49196)
//   T2 zl_k_cover_count   thread per position: freq[hash]++ for positions whose 8 bytes lie inside one sample (zstd.c:49389-49410)
//   T3 zl_k_cover_prev    thread per position: distance to the previous position with the same hash (tile of hashes in shared
//                         memory).  A position p adds freq[h(p)] to the score of the windows that contain p but not that
//                         previous occurrence -- the reference's "first occurrence in the active segment" rule
//                         (segmentFreqs, zstd.c:49285-49307) -- so the windows it scores are a RANGE of window starts.
//   per epoch (sequential, because a chosen segment zeroes the frequencies it covers, zstd.c:49322-49327):
//   T4 zl_k_cover_score   thread per position of the epoch: +freq at the first window start of its range, -freq after the
//                         last (a difference array, u64 atomics) -- the sliding window of FASTCOVER_selectSegment as a scan
//   T5 zl_k_cover_select  one CTA: prefix sum of the difference array = score of every window, arg max (earliest wins),
//                         copy the segment to the back of the dictionary, zero the frequencies it covers (zstd.c:49536-49566)
//   entropy tables (ZDICT_analyzeEntropy, zstd.c:50531): the training samples are parsed by the compressor's own match + parse
//   kernels against the raw content; zl_k_dict_stats (zl_enc_kernels.cu) sums literal and LL/ML/OF code counts; the host
//   builds the Huffman / FSE descriptions with the encoder's ZL_HD routines and writes the header (zstd.c:50730-50830).
//   Every candidate k is scored like COVER_checkTotalCompressedSize (zstd.c:46672): the test samples are compressed on the
//   GPU with the finished dictionary; the smallest total wins (zstd.c:46755).
// No CPU fallback: without a CUDA device the entry points return an error code.
#pragma once

struct ZlTrainState { u32 tail, zeroRun, done, nsel; };

__global__ void __launch_bounds__(256)
zl_k_cover_hash(const u8* __restrict__ samples, u32 nbDmers, u32 d, u32 f, u32* __restrict__ h)
{
    const u32 p = blockIdx.x * 256 + threadIdx.x;
    if (p >= nbDmers) return;
    const u32 lo = zl_rd32(samples + p), hi = zl_rd32(samples + p + 4);
    h[p] = d == 8 ? zl_hash_long(lo, hi, f) : zl_hash_short(lo, hi, 6, f);
}
// offsets[0..nb] = start of every training sample (offsets[nb] = end); counted like FASTCOVER_computeFrequency (skip = 0)
__global__ void __launch_bounds__(256)
zl_k_cover_count(const u32* __restrict__ h, u32 nbDmers, const u32* __restrict__ offsets, u32 nb, u32 readLen, u32* freq)
{
    const u32 p = blockIdx.x * 256 + threadIdx.x;
    if (p >= nbDmers) return;
    u32 a = 0, b = nb;                                   // the sample that holds p: offsets[a] <= p < offsets[a + 1]
    while (b - a > 1) { const u32 m = (a + b) >> 1; if (offsets[m] <= p) a = m; else b = m; }
    if (p + readLen <= offsets[a + 1]) atomicAdd(freq + h[p], 1u);
}
#define ZL_COVER_TILE 1024
#define ZL_COVER_MAXSPAN 8191u
__global__ void __launch_bounds__(ZL_COVER_TILE)
zl_k_cover_prev(const u32* __restrict__ h, u32 n, u32 span, u16* __restrict__ prevDist)
{
    extern __shared__ u32 sh[];                          // span + tile hashes
    const u32 t0 = blockIdx.x * ZL_COVER_TILE;
    for (u32 i = threadIdx.x; i < span + ZL_COVER_TILE; i += ZL_COVER_TILE) {
        const long long pos = (long long)t0 + i - span;
        sh[i] = (pos >= 0 && pos < (long long)n) ? h[pos] : 0xFFFFFFFFu;
    }
    __syncthreads();
    const u32 p = t0 + threadIdx.x;
    if (p >= n) return;
    const u32 me = sh[span + threadIdx.x];
    u32 dist = 0;
    for (u32 j = 1; j <= span; j++) if (sh[span + threadIdx.x - j] == me) { dist = j; break; }
    prevDist[p] = (u16)dist;                             // 0: none within `span`
}
__global__ void __launch_bounds__(256)
zl_k_cover_score(const u32* __restrict__ h, const u16* __restrict__ prevDist, const u32* __restrict__ freq, u32 eb, u32 ee, u32 dmersInK,
                 unsigned long long* diff, const ZlTrainState* __restrict__ st)
{
    if (st->done) return;
    const u32 p = eb + blockIdx.x * 256 + threadIdx.x;
    if (p >= ee) return;
    const u32 v = freq[h[p]];
    if (!v) return;
    long long lo = (long long)p + 1 - dmersInK;          // window starts b with b <= p < b + dmersInK ...
    const u32 pd = prevDist[p];
    if (pd && pd < dmersInK) lo = (long long)p - pd + 1;    // ... that do not contain the previous occurrence of this hash
    if (lo < (long long)eb) lo = eb;
    long long hi = p;
    if (hi > (long long)ee - dmersInK) hi = (long long)ee - dmersInK;      // the last full window of the epoch
    if (lo > hi) return;
    atomicAdd(diff + (lo - eb), (unsigned long long)v);
    atomicAdd(diff + (hi + 1 - eb), 0ull - (unsigned long long)v);
}
__global__ void __launch_bounds__(1024)
zl_k_cover_select(const u8* __restrict__ samples, const u32* __restrict__ h, u32* freq, u32 eb, u32 ee, u32 dmersInK, u32 d,
                  unsigned long long* diff, u8* dict, ZlTrainState* st)
{
    __shared__ unsigned long long wsum[32], carry, bScore[32];
    __shared__ u32 bPos[32], segBegin, segBytes, segTail;
    if (st->done) return;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 nB = ee - dmersInK - eb + 1;               // window starts
    if (tid == 0) carry = 0;
    __syncthreads();
    unsigned long long best = 0; u32 bestPos = 0xFFFFFFFFu;
    for (u32 c0 = 0; c0 < nB; c0 += 1024) {
        const u32 i = c0 + tid;
        unsigned long long v = i < nB ? diff[i] : 0ull;
        if (i <= nB) diff[i] = 0;                        // (clean for the next epoch; entry nB only ever receives subtractions)
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const unsigned long long a = __shfl_up_sync(0xFFFFFFFFu, v, s); if ((int)lane >= s) v += a; }
        if (lane == 31) wsum[warp] = v;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = wsum[lane];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) { const unsigned long long a = __shfl_up_sync(0xFFFFFFFFu, w, s); if ((int)lane >= s) w += a; }
            wsum[lane] = w;
        }
        __syncthreads();
        const unsigned long long score = v + (warp ? wsum[warp - 1] : 0ull) + carry;
        if (i < nB && score > best) { best = score; bestPos = i; }       // strictly greater: the earliest window wins inside a thread
        __syncthreads();
        if (tid == 1023) carry = score;
        __syncthreads();
    }
    if (tid == 0 && nB % 1024 == 0) diff[nB] = 0;
    // arg max over the CTA, earliest position on ties
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const unsigned long long ob = __shfl_down_sync(0xFFFFFFFFu, best, s); const u32 op = __shfl_down_sync(0xFFFFFFFFu, bestPos, s);
        if (ob > best || (ob == best && op < bestPos)) { best = ob; bestPos = op; }
    }
    if (lane == 0) { bScore[warp] = best; bPos[warp] = bestPos; }
    __syncthreads();
    if (tid == 0) {
        for (u32 w = 1; w < 32; w++) if (bScore[w] > best || (bScore[w] == best && bPos[w] < bestPos)) { best = bScore[w]; bestPos = bPos[w]; }
        segBytes = 0;
        if (best == 0) { if (++st->zeroRun >= 10) st->done = 1; }                    // zstd.c:49547-49552
        else {
            st->zeroRun = 0;
            u32 sz = dmersInK + d - 1;                                              // = k bytes
            if (sz > st->tail) sz = st->tail;
            if (sz < d) st->done = 1;                                               // zstd.c:49557
            else { st->tail -= sz; st->nsel++; segBegin = eb + bestPos; segBytes = sz; segTail = st->tail; if (st->tail == 0) st->done = 1; }
        }
    }
    __syncthreads();
    if (!segBytes) return;
    for (u32 i = tid; i < segBytes; i += 1024) dict[segTail + i] = samples[segBegin + i];
    for (u32 i = tid; i < dmersInK; i += 1024) freq[h[segBegin + i]] = 0;          // zstd.c:49322-49327
}

// ---------------------------------------------------------------------------------------------- host side
struct ZlTrainBufs {
    ZlDevBuf samples, offsets, h, freq0, freq, prev, diff, dict, state, stats, evalDst;
    void release() { ZlDevBuf* b[] = {&samples, &offsets, &h, &freq0, &freq, &prev, &diff, &dict, &state, &stats, &evalDst}; for (ZlDevBuf* x : b) x->release(); }
};
#define ZL_STAT_WORDS 512      // [0,256) literals, [256,292) LL codes, [292,345) ML codes, [345,377) OF codes

// entropy section of a dictionary from the statistics (zstd.c:50560-50700): Huffman description, OF / ML / LL NCounts, 1 4 8
static size_t zl_dict_entropy(u8* out, size_t cap, const u32* stats, u32 contentSize)
{
    ZlHufSm* f = new (std::nothrow) ZlHufSm();
    if (!f) return ZL_ERROR(memory_allocation);
    size_t pos = 0;
    bool ok = false;
    for (u32 shift = 0; shift < 32 && !ok; shift++) {     // (flatter counts until the tree can be described, cf. ZDICT_flatLit)
        memset(f, 0, sizeof(*f));
        for (u32 s = 0; s < 256; s++) f->count[s] = 1 + (stats[s] >> shift);                    // "any character must be described"
        zl_huf_build(*f);
        ok = zl_huf_write_desc(*f) && f->ctl.descSize <= cap;
    }
    if (!ok) { delete f; return ZL_ERROR(dictionaryCreation_failed); }
    memcpy(out, f->desc, f->ctl.descSize); pos = f->ctl.descSize;
    delete f;
    u32 offMax = zl_highbit(contentSize + (128u << 10));
    if (offMax > 30) return ZL_ERROR(dictionaryCreation_failed);
    const u32 base[3] = {345, 292, 256}, maxSym[3] = {offMax, 52, 35}, logs[3] = {8, 9, 9};     // OF, ML, LL: the order in the file
    for (int t = 0; t < 3; t++) {
        u32 cnt[64]; i16 norm[64]; u32 total = 0;
        for (u32 s = 0; s <= maxSym[t]; s++) { cnt[s] = 1 + stats[base[t] + s]; total += cnt[s]; }
        if (!zl_fse_normalize(norm, logs[t], cnt, total, maxSym[t], true)) return ZL_ERROR(dictionaryCreation_failed);
        const u32 n = zl_fse_write_ncount(out + pos, (u32)(cap - pos), norm, maxSym[t], logs[t]);
        if (!n) return ZL_ERROR(dictionaryCreation_failed);
        pos += n;
    }
    if (cap - pos < 12) return ZL_ERROR(dstSize_tooSmall);
    const u32 rep[3] = {1, 4, 8};
    for (int i = 0; i < 3; i++) for (int b = 0; b < 4; b++) out[pos++] = (u8)(rep[i] >> (8 * b));
    return pos;
}

static u64 zl_host_xxh64(const u8* p, size_t len)        // zstd.c:11509-11664, seed 0 (dictionary ID)
{
    const u64 P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL, P4 = 0x85EBCA77C2B2AE63ULL, P5 = 0x27D4EB2F165667C5ULL;
    auto rotl = [](u64 x, int r) { return (x << r) | (x >> (64 - r)); };
    auto rd64 = [](const u8* q) { u64 v; memcpy(&v, q, 8); return v; };
    auto round = [&](u64 acc, u64 in) { return rotl(acc + in * P2, 31) * P1; };
    auto merge = [&](u64 h, u64 v) { return (h ^ round(0, v)) * P1 + P4; };
    const u8* end = p + len; u64 h;
    if (len >= 32) {
        u64 v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0ull - P1;
        do { v1 = round(v1, rd64(p)); v2 = round(v2, rd64(p + 8)); v3 = round(v3, rd64(p + 16)); v4 = round(v4, rd64(p + 24)); p += 32; } while (p + 32 <= end);
        h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
        h = merge(h, v1); h = merge(h, v2); h = merge(h, v3); h = merge(h, v4);
    } else h = P5;
    h += (u64)len;
    while (p + 8 <= end) { h ^= round(0, rd64(p)); h = rotl(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= end) { u32 v; memcpy(&v, p, 4); h ^= (u64)v * P1; h = rotl(h, 23) * P2 + P3; p += 4; }
    while (p < end) { h ^= (u64)(*p++) * P5; h = rotl(h, 11) * P1; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

#include <chrono>
#include <mutex>
static double zl_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct ZlTrainJob {
    double tSelect = 0, tStats = 0, tEval = 0;
    const u8* hostSamples; const size_t* sizes; u32 nb, nbTrain, testFirst, nbTest;
    size_t trainBytes, totalBytes;
    int level; u32 dictID;
    ZlTrainBufs B;
    ZSTD_CCtx* statCtx; ZSTD_CCtx* evalCtx;
    std::vector<const void*> srcPtr; std::vector<void*> dstPtr; std::vector<size_t> dstCap, res, statSizes;
    cudaStream_t st;
};

// one candidate (d is fixed by the hash arrays, k given): select the content, finalize, score.  `out` receives the dictionary.
static size_t zl_train_candidate(ZlTrainJob& J, u32 k, u32 d, u32 nbDmers, size_t dictCap, std::vector<u8>& out, u64* score)
{
    cudaStream_t st = J.st;
    const u32 dmersInK = k - d + 1;
    // COVER_computeEpochs (zstd.c:46524), passes = 1
    u32 epochNum = (u32)(dictCap / k); if (epochNum < 1) epochNum = 1;
    u32 epochSize = nbDmers / epochNum;
    if (epochSize < k * 10) { epochSize = k * 10 < nbDmers ? k * 10 : nbDmers; epochNum = nbDmers / epochSize; }
    if (epochSize < dmersInK) return ZL_ERROR(srcSize_wrong);
    if (!J.B.diff.reserve(((size_t)epochSize + 2) * 8) || !J.B.dict.reserve(dictCap + 64)) return ZL_ERROR(memory_allocation);
    cudaMemcpyAsync(J.B.freq.p, J.B.freq0.p, (size_t)4 << 20, cudaMemcpyDeviceToDevice, st);
    cudaMemsetAsync(J.B.diff.p, 0, ((size_t)epochSize + 2) * 8, st);
    ZlTrainState hs = {(u32)dictCap, 0, 0, 0};
    cudaMemcpyAsync(J.B.state.p, &hs, sizeof(hs), cudaMemcpyHostToDevice, st);
    const u32 maxIter = 2 * (u32)(dictCap / d + 16);
    const double t0 = zl_now_ms();
    u32 iter = 0;
    while (iter < maxIter) {
        for (u32 e = 0; e < epochNum && iter < maxIter; e++, iter++) {
            const u32 eb = e * epochSize, ee = eb + epochSize;
            zl_k_cover_score<<<(epochSize + 255) / 256, 256, 0, st>>>(J.B.h.as<u32>(), J.B.prev.as<u16>(), J.B.freq.as<u32>(), eb, ee, dmersInK,
                                                                      J.B.diff.as<unsigned long long>(), J.B.state.as<ZlTrainState>());
            zl_k_cover_select<<<1, 1024, 0, st>>>(J.B.samples.as<u8>(), J.B.h.as<u32>(), J.B.freq.as<u32>(), eb, ee, dmersInK, d,
                                                  J.B.diff.as<unsigned long long>(), J.B.dict.as<u8>(), J.B.state.as<ZlTrainState>());
        }
        cudaMemcpyAsync(&hs, J.B.state.p, sizeof(hs), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { (void)cudaGetLastError(); return ZL_ERROR(GENERIC); }
        if (hs.done) break;
    }
    const double t1 = zl_now_ms();
    const size_t contentSize0 = dictCap - hs.tail;
    if (contentSize0 < 128) return ZL_ERROR(dictionaryCreation_failed);             // ZDICT_CONTENTSIZE_MIN
    std::vector<u8> content(contentSize0);
    if (cudaMemcpy(content.data(), J.B.dict.as<u8>() + hs.tail, contentSize0, cudaMemcpyDeviceToHost) != cudaSuccess) { (void)cudaGetLastError(); return ZL_ERROR(GENERIC); }
    // ---- statistics: the training samples against the raw content (ZDICT_countEStats, zstd.c:50440)
    ZSTD_CCtx* sc = J.statCtx;
    ZSTD_CCtx_loadDictionary(sc, content.data(), content.size());
    cudaMemsetAsync(J.B.stats.p, 0, ZL_STAT_WORDS * 4, sc->stream);
    sc->statsDev = J.B.stats.as<u32>();
    // samples of any size take part in the selection above (every d-mer counts) and in the scoring below (whole samples); the
    // statistics look at the first block of each, as ZDICT_countEStats does (zstd.c:50440-50460)
    size_t r = zl_compress_batch(sc, J.srcPtr.data(), J.statSizes.data(), J.dstPtr.data(), J.dstCap.data(), J.res.data(), J.nbTrain, 1);
    sc->statsDev = nullptr;
    if (zl_is_error(r)) return r;
    u32 stats[ZL_STAT_WORDS];
    if (cudaMemcpy(stats, J.B.stats.p, sizeof(stats), cudaMemcpyDeviceToHost) != cudaSuccess) { (void)cudaGetLastError(); return ZL_ERROR(GENERIC); }
    const double t2 = zl_now_ms();
    // ---- header (zstd.c:50730-50830)
    u8 header[320];
    const u32 magic = ZL_MAGIC_DICT;
    const u64 rid = zl_host_xxh64(content.data(), content.size());
    const u32 id = J.dictID ? J.dictID : (u32)(rid % ((1u << 31) - 32768)) + 32768;
    for (int b = 0; b < 4; b++) { header[b] = (u8)(magic >> (8 * b)); header[4 + b] = (u8)(id >> (8 * b)); }
    const size_t es = zl_dict_entropy(header + 8, sizeof(header) - 8, stats, (u32)content.size());
    if (zl_is_error(es)) return es;
    const size_t hSize = 8 + es;
    if (hSize + 128 > dictCap) return ZL_ERROR(dstSize_tooSmall);
    size_t contentSize = content.size();
    if (hSize + contentSize > dictCap) contentSize = dictCap - hSize;              // like the reference, the FIRST bytes are kept (zstd.c:50773, 50803)
    out.assign(header, header + hSize);
    out.insert(out.end(), content.begin(), content.begin() + (ptrdiff_t)contentSize);
    // ---- score: test samples compressed with this dictionary + its size (zstd.c:46672-46720)
    ZSTD_CCtx* ec = J.evalCtx;
    ZSTD_CCtx_loadDictionary(ec, out.data(), out.size());
    r = zl_compress_batch(ec, J.srcPtr.data() + J.testFirst, J.sizes + J.testFirst, J.dstPtr.data() + J.testFirst, J.dstCap.data() + J.testFirst,
                          J.res.data(), J.nbTest, 1);
    if (zl_is_error(r)) return r;
    u64 total = out.size();
    for (u32 i = 0; i < J.nbTest; i++) { if (zl_is_error(J.res[i])) return J.res[i]; total += J.res[i]; }
    *score = total;
    J.tSelect += t1 - t0; J.tStats += t2 - t1; J.tEval += zl_now_ms() - t2;
    return 0;
}

static size_t zl_train(void* dictBuffer, size_t dictCap, const void* samplesBuffer, const size_t* sizes, unsigned nb,
                       unsigned kPar, unsigned dPar, unsigned steps, double splitPoint, int level, unsigned dictID, unsigned* kOut, unsigned* dOut)
{
    if (!dictBuffer || !samplesBuffer || !sizes) return ZL_ERROR(GENERIC);
    if (dictCap < 256) return ZL_ERROR(dstSize_tooSmall);                           // ZDICT_DICTSIZE_MIN
    if (splitPoint <= 0.0 || splitPoint > 1.0) return ZL_ERROR(parameter_outOfBound);
    if (dPar != 0 && dPar != 6 && dPar != 8) return ZL_ERROR(parameter_outOfBound);  // zstd.c:49344
    if (kPar != 0 && (kPar > dictCap || kPar < (dPar ? dPar : 8))) return ZL_ERROR(parameter_outOfBound);
    // the two compression contexts and the device buffers of the trainer are kept for the process (grow-only, like a context's
    // arenas): a second training does not pay the allocations again.  One training at a time.
    static std::mutex trainLock;
    static ZSTD_CCtx* keepStat = nullptr; static ZSTD_CCtx* keepEval = nullptr; static ZlTrainBufs keepBufs;
    std::lock_guard<std::mutex> guard(trainLock);
    ZlTrainJob J;
    J.B = keepBufs;
    J.hostSamples = (const u8*)samplesBuffer; J.sizes = sizes; J.nb = nb; J.level = level ? level : 3; J.dictID = dictID;
    J.nbTrain = splitPoint < 1.0 ? (u32)((double)nb * splitPoint) : nb;             // zstd.c:49443-49446
    J.nbTest = splitPoint < 1.0 ? nb - J.nbTrain : nb;
    J.testFirst = splitPoint < 1.0 ? J.nbTrain : 0;
    if (J.nbTrain < 5 || J.nbTest < 1) return ZL_ERROR(srcSize_wrong);              // zstd.c:49461-49470
    J.totalBytes = 0; J.trainBytes = 0;
    std::vector<u32> offs(nb + 1);
    for (u32 i = 0; i < nb; i++) {
        offs[i] = (u32)J.totalBytes;
        J.totalBytes += sizes[i];
        if (i + 1 == J.nbTrain) J.trainBytes = J.totalBytes;
        if (J.totalBytes > 0xF0000000ull) return ZL_ERROR(srcSize_wrong);           // as the reference: FASTCOVER_MAX_SAMPLES_SIZE (4 GiB), zstd.c:49474
    }
    offs[nb] = (u32)J.totalBytes;
    if (J.trainBytes < 16) return ZL_ERROR(srcSize_wrong);
    const u32 f = 20;
    const bool verbose = getenv("ZL_TRAIN_VERBOSE") != nullptr;      // (the reference reports through notificationLevel)
    if (!keepStat) keepStat = ZSTD_createCCtx();
    if (!keepEval) keepEval = ZSTD_createCCtx();
    J.statCtx = keepStat; J.evalCtx = keepEval;
    size_t result = ZL_ERROR(memory_allocation);
    std::vector<u8> best, cand;
    u64 bestScore = ~0ull; u32 bestK = 0, bestD = 0;
    do {
        if (!J.statCtx || !J.evalCtx || !zl_cctx_ready(J.statCtx) || !zl_cctx_ready(J.evalCtx)) break;
        ZSTD_CCtx_setParameter(J.statCtx, ZSTD_c_compressionLevel, J.level);
        ZSTD_CCtx_setParameter(J.evalCtx, ZSTD_c_compressionLevel, J.level);
        J.st = J.statCtx->stream;
        ZlTrainBufs& B = J.B;
        size_t dstBytes = 0;
        J.dstCap.resize(nb); J.srcPtr.resize(nb); J.dstPtr.resize(nb); J.res.resize(nb); J.statSizes.resize(nb);
        for (u32 i = 0; i < nb; i++) J.statSizes[i] = sizes[i] < ZL_BLOCKSIZE_MAX ? sizes[i] : ZL_BLOCKSIZE_MAX;
        for (u32 i = 0; i < nb; i++) { J.dstCap[i] = ZSTD_compressBound(sizes[i]) + 32; dstBytes += (J.dstCap[i] + 15) & ~(size_t)15; }
        if (!B.samples.reserve(J.totalBytes + 64) || !B.offsets.reserve(((size_t)nb + 1) * 4) || !B.h.reserve(J.trainBytes * 4 + 64) ||
            !B.freq0.reserve((size_t)4 << f) || !B.freq.reserve((size_t)4 << f) || !B.prev.reserve(J.trainBytes * 2 + 64) ||
            !B.state.reserve(64) || !B.stats.reserve(ZL_STAT_WORDS * 4) || !B.evalDst.reserve(dstBytes + 64)) break;
        cudaMemsetAsync(B.samples.as<u8>() + J.totalBytes, 0, 64, J.st);
        cudaMemcpyAsync(B.samples.p, samplesBuffer, J.totalBytes, cudaMemcpyHostToDevice, J.st);
        cudaMemcpyAsync(B.offsets.p, offs.data(), ((size_t)nb + 1) * 4, cudaMemcpyHostToDevice, J.st);
        {   size_t o = 0;
            for (u32 i = 0; i < nb; i++) { J.srcPtr[i] = B.samples.as<u8>() + offs[i]; J.dstPtr[i] = B.evalDst.as<u8>() + o; o += (J.dstCap[i] + 15) & ~(size_t)15; } }
        // candidate lists (zstd.c:49760-49790): d given or {8} (fastCOVER default) / {6, 8} (optimize_cover); k given or `steps` values in [50, 2000]
        std::vector<u32> dList, kList;
        if (dPar) dList.push_back(dPar); else { if (steps > 4) dList.push_back(6); dList.push_back(8); }
        if (kPar) kList.push_back(kPar);
        else {
            const u32 kMin = 50, kMax = 2000, n = steps ? steps : 4, stepSize = (kMax - kMin) / n > 1 ? (kMax - kMin) / n : 1;
            for (u32 k = kMin; k <= kMax; k += stepSize) if (k <= dictCap) kList.push_back(k);
            if (kList.empty()) kList.push_back((u32)dictCap);
        }
        u32 kLargest = 0; for (u32 k : kList) if (k > kLargest) kLargest = k;
        result = ZL_ERROR(dictionaryCreation_failed);
        bool fatal = false;
        for (u32 d : dList) {
            const u32 readLen = d > 8 ? d : 8;
            if (J.trainBytes < readLen) continue;
            const u32 nbDmers = (u32)(J.trainBytes - readLen + 1);                  // zstd.c:49487
            u32 span = kLargest - d; if (span > ZL_COVER_MAXSPAN) span = ZL_COVER_MAXSPAN; if (span < 1) span = 1;
            cudaMemsetAsync(B.freq0.p, 0, (size_t)4 << f, J.st);
            zl_k_cover_hash<<<(nbDmers + 255) / 256, 256, 0, J.st>>>(B.samples.as<u8>(), nbDmers, d, f, B.h.as<u32>());
            zl_k_cover_count<<<(nbDmers + 255) / 256, 256, 0, J.st>>>(B.h.as<u32>(), nbDmers, B.offsets.as<u32>(), J.nbTrain, readLen, B.freq0.as<u32>());
            zl_k_cover_prev<<<(nbDmers + ZL_COVER_TILE - 1) / ZL_COVER_TILE, ZL_COVER_TILE, (span + ZL_COVER_TILE) * 4, J.st>>>(B.h.as<u32>(), nbDmers, span, B.prev.as<u16>());
            if (cudaGetLastError() != cudaSuccess) { fatal = true; result = ZL_ERROR(GENERIC); break; }
            if (verbose) { const double ta = zl_now_ms(); cudaStreamSynchronize(J.st); fprintf(stderr, "zstdlite_gpu train: d=%u hash + count + previous-occurrence kernels %.1f ms (%u d-mers, span %u)\n", d, zl_now_ms() - ta, nbDmers, span); }
            for (u32 k : kList) {
                if (k < d || k > dictCap) continue;
                u64 score = 0;
                const size_t r = zl_train_candidate(J, k, d, nbDmers, dictCap, cand, &score);
                if (verbose) fprintf(stderr, "zstdlite_gpu train: d=%u k=%u -> %s, dictionary %zu B, test total %llu\n", d, k,
                                     zl_is_error(r) ? ZSTD_getErrorName(r) : "ok", cand.size(), (unsigned long long)score);
                if (zl_is_error(r)) { if (result == ZL_ERROR(dictionaryCreation_failed)) result = r; continue; }
                if (score < bestScore) { bestScore = score; best.swap(cand); bestK = k; bestD = d; }
            }
        }
        if (verbose) fprintf(stderr, "zstdlite_gpu train: select %.1f ms, statistics %.1f ms, header + scoring %.1f ms\n", J.tSelect, J.tStats, J.tEval);
        if (!fatal && !best.empty()) { memcpy(dictBuffer, best.data(), best.size()); result = best.size(); if (kOut) *kOut = bestK; if (dOut) *dOut = bestD; }
    } while (0);
    keepBufs = J.B;                                       // (the buffers may have grown)
    if (J.statCtx) { J.statCtx->statsDev = nullptr; ZSTD_CCtx_loadDictionary(J.statCtx, nullptr, 0); }
    if (J.evalCtx) ZSTD_CCtx_loadDictionary(J.evalCtx, nullptr, 0);
    return result;
}

// ---- the entry points src/dictionaries.c:60-208 binds -------------------------------------------------------------------

ZL_EXPORT size_t ZDICT_trainFromBuffer(void* dictBuffer, size_t dictBufferCapacity, const void* samplesBuffer, const size_t* samplesSizes, unsigned nbSamples)
{   // zstd.c:50979: fastCOVER, d = 8, steps = 4, split point 0.75 (zstd.c:49810), level 3
    return zl_train(dictBuffer, dictBufferCapacity, samplesBuffer, samplesSizes, nbSamples, 0, 8, 4, 0.75, 3, 0, nullptr, nullptr);
}
ZL_EXPORT size_t ZDICT_optimizeTrainFromBuffer_cover(void* dictBuffer, size_t dictBufferCapacity, const void* samplesBuffer, const size_t* samplesSizes,
                                                     unsigned nbSamples, ZDICT_cover_params_t* parameters)
{   // zstd.c:46957: steps defaults to 40, d to {6, 8}, the split point to 1.0 (train and test on every sample); the best k / d are written back.
    // shrinkDict is accepted and ignored (the dictionary is never larger than the capacity).
    ZDICT_cover_params_t* p = parameters;
    if (!p) return ZL_ERROR(GENERIC);
    unsigned k = 0, d = 0;
    const size_t r = zl_train(dictBuffer, dictBufferCapacity, samplesBuffer, samplesSizes, nbSamples, p->k, p->d, p->steps ? p->steps : 40,
                              p->splitPoint <= 0.0 ? 1.0 : p->splitPoint, p->zParams.compressionLevel, p->zParams.dictID, &k, &d);
    if (!zl_is_error(r)) { p->k = k; p->d = d; if (!p->steps) p->steps = 40; }
    return r;
}
ZL_EXPORT unsigned ZDICT_isError(size_t code) { return zl_is_error(code); }                      // zstd.c:50101
ZL_EXPORT const char* ZDICT_getErrorName(size_t code) { return ZSTD_getErrorName(code); }       // zstd.c:50103
ZL_ALIAS(size_t, ZDICT_trainFromBuffer, (void*, size_t, const void*, const size_t*, unsigned))
ZL_ALIAS(size_t, ZDICT_optimizeTrainFromBuffer_cover, (void*, size_t, const void*, const size_t*, unsigned, ZDICT_cover_params_t*))
ZL_ALIAS(unsigned, ZDICT_isError, (size_t))
ZL_ALIAS(const char*, ZDICT_getErrorName, (size_t))
