// Test infrastructure: exercises the host copy pool of the C-ABI library (zstdlite_b200/csrc/zl_host.h) on the CPU -- the streaming-store
// piece copy at every source / destination alignment and length class, and ZlCopyPool::run over many segments on several threads.
// Built and run by tests/test_host_copy.py; no CUDA call is made (the header's device helpers are unused inline functions).
#include "zl_host.h"
#include "zl_plan.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

int main()
{
    std::vector<unsigned char> a(8u << 20), b(8u << 20);
    for (size_t i = 0; i < a.size(); i++) a[i] = (unsigned char)((i * 2654435761u) >> 13);
    int bad = 0;
    const size_t lens[] = {0, 1, 15, 16, 17, 127, 128, 129, 255, 1000, 4096, 70001};
    for (int so = 0; so < 33; so++)
        for (int dof = 0; dof < 33; dof++)
            for (size_t n : lens) {
                memset(b.data(), 0xEE, 80000);
                zl_copy_stream(b.data() + 64 + dof, a.data() + so, n);
                if (memcmp(b.data() + 64 + dof, a.data() + so, n)) bad++;
                if (b[63 + dof] != 0xEE || b[64 + dof + n] != 0xEE) bad++;        // nothing outside the range
            }
    // the pool: a list of segments of mixed sizes (above the 4 MiB threshold in total, so the workers take part), then a short list (inline path)
    setenv("ZL_COPY_THREADS", "5", 1);
    for (int round = 0; round < 3; round++) {
        std::vector<ZlCopySeg> segs;
        memset(b.data(), 0, b.size());
        size_t so = 3, dof = 5 + (size_t)round;
        const size_t sizes[] = {1, 300, 5000, 65536, 262144, 262145, 700000, 1u << 20, 2500001};
        for (int rep = 0; rep < (round == 2 ? 1 : 2); rep++)
            for (size_t n : sizes) {
                if (round == 2 && n > 5000) continue;
                if (so + n > a.size() || dof + n > b.size()) break;
                segs.push_back({b.data() + dof, a.data() + so, n});
                so += n + 7; dof += n + 7;
            }
        ZlCopyPool::get().run(segs);
        for (const ZlCopySeg& s : segs) if (memcmp(s.dst, s.src, s.bytes)) bad++;
        size_t covered = 0;
        for (const ZlCopySeg& s : segs) covered += s.bytes;
        size_t nonzeroOutside = 0;                                                  // the 7-byte gaps between segments stay zero
        for (size_t i = 1; i < segs.size(); i++) {
            const unsigned char* gap = (const unsigned char*)segs[i - 1].dst + segs[i - 1].bytes;
            for (int k = 0; k < 7; k++) nonzeroOutside += gap[k] != 0;
        }
        if (nonzeroOutside) bad++;
        printf("round %d: %zu segments, %zu bytes, threads %u\n", round, segs.size(), covered, ZlCopyPool::get().threads());
    }
    // contiguous ranges of equal weight (batches spread over GPUs)
    {
        std::vector<size_t> w(1000);
        for (size_t i = 0; i < w.size(); i++) w[i] = 100 + (i % 7) * 1000;
        const std::vector<size_t> cut = zl_split_ranges(w.data(), w.size(), 4);
        if (cut.size() != 5 || cut[0] != 0 || cut[4] != w.size()) bad++;
        for (int k = 0; k < 4; k++) if (cut[k] > cut[k + 1]) bad++;
    }
    // arena shares of a slice whose descriptors are built on the device: the bound from the slice's totals must cover the per-frame sums
    {
        unsigned long long x = 88172645463325252ull;
        auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
        for (int trial = 0; trial < 2000; trial++) {
            const unsigned m = 1 + (unsigned)(rnd() % 3000);
            const bool worst = trial & 1;
            const unsigned smax = trial % 5 == 0 ? 4096u : 1u + (unsigned)(rnd() % 4096), dmax = trial % 7 == 0 ? 262144u : 1u + (unsigned)(rnd() % 262144);
            unsigned long long S = 0, D = 0, lit = 0, rec = 0, hdr = 0;
            for (unsigned i = 0; i < m; i++) {
                const u32 sz = (u32)(rnd() % (smax + 1)), dc = (u32)(rnd() % (dmax + 1));
                u32 lc, rc, hc;
                zl_plan_frame(sz, dc, worst, &lc, &rc, &hc);
                S += sz; D += dc; lit += ((unsigned long long)lc + 15) & ~15ull; rec += rc; hdr += hc;
                if (rc & 1) bad++;
            }
            unsigned long long bl, br, bh;
            zl_plan_slice_bound(S, D, m, worst, &bl, &br, &bh);
            if (bl < lit || br < rec || bh < hdr || (bl & 15) || (br & 1)) bad++;
        }
    }
    printf("bad=%d\n", bad);
    return bad ? 1 : 0;
}
