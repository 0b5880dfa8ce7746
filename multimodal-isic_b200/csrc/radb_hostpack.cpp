// Host side of the packed-mask transfer path (include/radb.h: radb_pack_mask_host).  The end-to-end rate of
// the extraction is bound by the host-to-device link (4096 pixel bytes + 4096 mask bytes per 64x64 patch at
// ~50 GB/s), and a mask carries one bit of information per pixel (ROI = mask == label,
// RadiomicExtractor.py:33-38 / params.yml:93).  Packing it on the host -- AVX2 compare + movemask over a few
// threads, memory bound -- and unpacking it on the device (radb_unpack_mask) takes 7/16 of the bytes off the
// link.  Plain host C++ (compiled by the host compiler, no CUDA), no CPU path of the engine itself.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define RADB_X86 1
#endif

static void pack_scalar(const uint8_t* m, int64_t n, uint8_t lab, uint8_t* out)
{
    // n is a multiple of 8 except possibly for the last call (tail bits are zero)
    for (int64_t i = 0; i < n; i += 8) {
        unsigned b = 0;
        const int64_t e = n - i < 8 ? n - i : 8;
        for (int64_t k = 0; k < e; k++) b |= (unsigned)(m[i + k] == lab) << k;
        out[i >> 3] = (uint8_t)b;
    }
}
#ifdef RADB_X86
__attribute__((target("avx2"))) static void pack_avx2(const uint8_t* m, int64_t n, uint8_t lab, uint8_t* out)
{
    const __m256i l = _mm256_set1_epi8((char)lab);
    int64_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i v = _mm256_loadu_si256((const __m256i*)(m + i));
        const uint32_t bits = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, l));  // bit j <-> byte i + j
        memcpy(out + (i >> 3), &bits, 4);
    }
    if (i < n) pack_scalar(m + i, n - i, lab, out + (i >> 3));
}
#endif

#ifdef RADB_X86
// 128 mask bytes per step (four independent loads, one 16-byte store) with a software prefetch a few lines ahead:
// one packing thread is bound by the memory latency its few outstanding line fills can cover (~10 GB/s), not by the
// compare; a non-temporal prefetch hint was measured at half the rate
__attribute__((target("avx2"))) static void pack_avx2_x4(const uint8_t* m, int64_t n, uint8_t lab, uint8_t* out)
{
    const __m256i l = _mm256_set1_epi8((char)lab);
    int64_t i = 0;
    for (; i + 128 <= n; i += 128) {
        _mm_prefetch((const char*)(m + i + 1536), _MM_HINT_T0);
        _mm_prefetch((const char*)(m + i + 1536 + 64), _MM_HINT_T0);
        const __m256i v0 = _mm256_loadu_si256((const __m256i*)(m + i));
        const __m256i v1 = _mm256_loadu_si256((const __m256i*)(m + i + 32));
        const __m256i v2 = _mm256_loadu_si256((const __m256i*)(m + i + 64));
        const __m256i v3 = _mm256_loadu_si256((const __m256i*)(m + i + 96));
        uint32_t bits[4];
        bits[0] = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v0, l));
        bits[1] = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v1, l));
        bits[2] = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v2, l));
        bits[3] = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v3, l));
        memcpy(out + (i >> 3), bits, 16);
    }
    if (i < n) pack_avx2(m + i, n - i, lab, out + (i >> 3));
}
#endif

// Persistent packing threads.  The pipeline packs one ~0.5 ms chunk at a time; spawning and joining eight threads per
// chunk cost a fifth of that.  Workers sleep on a condition variable between jobs; a job is a number of blocks handed
// out through an atomic counter (the caller works too, and only waits for workers that actually joined the job, not
// for one that wakes up late).  The pool is created on first use and never torn down.
namespace {
struct PackPool {
    std::mutex mu, call_mu;
    std::condition_variable cv_go, cv_done;
    std::vector<std::thread> workers;
    const std::function<void(int64_t)>* job = nullptr;
    std::atomic<int64_t> next{0};
    int64_t nblocks = 0;
    int active = 0, running = 0;
    uint64_t gen = 0;

    void drain()
    {
        for (;;) {
            const int64_t b = next.fetch_add(1, std::memory_order_relaxed);
            if (b >= nblocks) break;
            (*job)(b);
        }
    }
    void worker(int id)
    {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_go.wait(lk, [&] { return gen != seen; });
            seen = gen;
            if (id >= active || !job) continue;  // this job asked for fewer threads / is already over (late wake-up)
            running++;
            lk.unlock();
            drain();
            lk.lock();
            if (--running == 0) cv_done.notify_one();
        }
    }
    void run(int threads, int64_t nb, const std::function<void(int64_t)>& f)
    {
        std::lock_guard<std::mutex> one(call_mu);  // one job at a time
        const int want = threads - 1;
        {
            std::unique_lock<std::mutex> lk(mu);
            while ((int)workers.size() < want) {
                const int id = (int)workers.size();
                workers.emplace_back([this, id] { worker(id); });
                workers.back().detach();
            }
            job = &f;
            nblocks = nb;
            next.store(0, std::memory_order_relaxed);
            active = want;  // (running == 0 here: the previous job waited for its workers)
            gen++;
        }
        if (want > 0) cv_go.notify_all();
        drain();
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return running == 0; });
        job = nullptr;
    }
};
PackPool* pack_pool()
{
    static PackPool* p = new PackPool();  // leaked on purpose: its threads outlive main()
    return p;
}
// A/B switch (scripts/e2e_probe.py): RADB_PACK_SPAWN=1 keeps the round-1 packer (threads spawned per call, 32 bytes per step)
bool pack_spawn() { const char* e = getenv("RADB_PACK_SPAWN"); return e && e[0] == '1'; }
}  // namespace

extern "C" int radb_pack_mask_host(const uint8_t* mask, int64_t n_bytes, int label, uint8_t* packed, int threads)
{
    if (!mask || !packed || n_bytes < 0) return -1;
    if (label < 0 || label > 255) {  // no uint8 value can match: empty ROI everywhere
        memset(packed, 0, (size_t)((n_bytes + 7) / 8));
        return 0;
    }
    void (*fn)(const uint8_t*, int64_t, uint8_t, uint8_t*) = pack_scalar;
    const bool spawn = pack_spawn();
#ifdef RADB_X86
    if (__builtin_cpu_supports("avx2")) fn = spawn ? pack_avx2 : pack_avx2_x4;
#endif
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    const int64_t block = spawn ? (1 << 20) : (1 << 18);  // bytes per work item (a multiple of 128)
    if (threads == 1 || n_bytes <= block) {
        fn(mask, n_bytes, (uint8_t)label, packed);
        return 0;
    }
    const int64_t nblocks = (n_bytes + block - 1) / block;
    if (!spawn) {
        const uint8_t lab = (uint8_t)label;
        const std::function<void(int64_t)> job = [=](int64_t b) {
            const int64_t lo = b * block, len = (n_bytes - lo < block) ? n_bytes - lo : block;
            fn(mask + lo, len, lab, packed + (lo >> 3));
        };
        pack_pool()->run(threads, nblocks, job);
        return 0;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([=]() {
            for (int64_t b = t; b < nblocks; b += threads) {
                const int64_t lo = b * block, len = (n_bytes - lo < block) ? n_bytes - lo : block;
                fn(mask + lo, len, (uint8_t)label, packed + (lo >> 3));
            }
        });
    for (auto& th : pool) th.join();
    return 0;
}

// Per-patch variant for radb_extract_packed: n_patches masks of `hw` bytes each, back to back; patch b's bit stream
// starts at packed + b * stride_b (stride_b >= ceil(hw / 8); padding bits/bytes are zeroed), bit i <=> mask byte i
// equals `label`.  With hw a multiple of 32 and stride_b == hw / 8 this is the same stream radb_pack_mask_host
// writes, and the whole buffer is packed in large blocks; otherwise patches are dealt to the threads one by one.
extern "C" int radb_pack_masks_host(const uint8_t* mask, int64_t n_patches, int64_t hw, int label, uint8_t* packed,
                                    int64_t stride_b, int threads)
{
    if (!mask || !packed || n_patches < 0 || hw < 1 || stride_b < (hw + 7) / 8) return -1;
    if (hw % 32 == 0 && stride_b == hw / 8) return radb_pack_mask_host(mask, n_patches * hw, label, packed, threads);
    if (label < 0 || label > 255) {
        memset(packed, 0, (size_t)(n_patches * stride_b));
        return 0;
    }
    void (*fn)(const uint8_t*, int64_t, uint8_t, uint8_t*) = pack_scalar;
#ifdef RADB_X86
    if (__builtin_cpu_supports("avx2")) fn = pack_avx2;
#endif
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    const int64_t used = (hw + 7) / 8;
    auto work = [=](int t, int nt) {
        const int64_t per = (n_patches + nt - 1) / nt;
        const int64_t lo = t * per, hi = lo + per < n_patches ? lo + per : n_patches;
        for (int64_t b = lo; b < hi; b++) {
            uint8_t* o = packed + b * stride_b;
            fn(mask + b * hw, hw, (uint8_t)label, o);
            if (stride_b > used) memset(o + used, 0, (size_t)(stride_b - used));
        }
    };
    if (threads == 1 || n_patches * hw <= (1 << 20)) {
        work(0, 1);
        return 0;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(work, t, threads);
    for (auto& th : pool) th.join();
    return 0;
}
