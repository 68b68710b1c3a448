// maz_hostrng.cu -- host code only (include/maz_hostrng.h): numpy's legacy Dirichlet draw, bit for bit, in parallel.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/maz_hostrng.h"

namespace maz { int set_last_error(int code, const std::string &msg); }

namespace {

// ---- MT19937 exactly as numpy/random/src/mt19937/mt19937.c (state = key[624] + pos) ---------------------------
struct MT {
    uint32_t key[624];
    int pos;
    void gen()
    {
        constexpr uint32_t UP = 0x80000000u, LO = 0x7fffffffu, MA = 0x9908b0dfu;
        int i;
        uint32_t y;
        for (i = 0; i < 624 - 397; ++i) {
            y = (key[i] & UP) | (key[i + 1] & LO);
            key[i] = key[i + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MA);
        }
        for (; i < 623; ++i) {
            y = (key[i] & UP) | (key[i + 1] & LO);
            key[i] = key[i + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MA);
        }
        y = (key[623] & UP) | (key[0] & LO);
        key[623] = key[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MA);
        pos = 0;
    }
    // next n tempered words
    void fill(uint32_t *out, size_t n)
    {
        size_t i = 0;
        while (i < n) {
            if (pos == 624) gen();
            const size_t take = std::min(n - i, (size_t)(624 - pos));
            for (size_t k = 0; k < take; ++k) {
                uint32_t y = key[pos + k];
                y ^= (y >> 11);
                y ^= (y << 7) & 0x9d2c5680u;
                y ^= (y << 15) & 0xefc60000u;
                y ^= (y >> 18);
                out[i + k] = y;
            }
            pos += (int)take;
            i += take;
        }
    }
    void skip(size_t n)
    {
        while (n) {
            if (pos == 624) gen();
            const size_t take = std::min(n, (size_t)(624 - pos));
            pos += (int)take;
            n -= take;
        }
    }
};

inline double to_double(uint32_t w0, uint32_t w1)   // mt19937_next_double
{
    const int32_t a = (int32_t)(w0 >> 5), b = (int32_t)(w1 >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

// one attempt of legacy_standard_gamma(shape < 1): returns true + X when accepted
inline bool gamma_attempt(const uint32_t *w, double shape, double &X)
{
    const double U = to_double(w[0], w[1]);
    const double V = -std::log(1.0 - to_double(w[2], w[3]));   // legacy_standard_exponential
    if (U <= 1.0 - shape) {
        X = std::pow(U, 1. / shape);
        return X <= V;
    }
    const double Y = -std::log((1 - U) / shape);
    X = std::pow(1.0 - shape + shape * Y, 1. / shape);
    return X <= (V + Y);
}

// ---- a small persistent worker pool -----------------------------------------------------------------------------
// The draw happens once per search, so thread start-up would cost more than it saves.  The caller never waits for a
// worker to WAKE UP (that can take milliseconds on a busy host): a job is a set of blocks claimed with an atomic
// counter, the caller claims blocks like everybody else and returns as soon as all blocks are done; workers that wake
// late find nothing to claim.  Worst case the caller evaluates everything itself (= numpy's own cost).
struct Job {
    double alpha = 0;
    size_t M = 0, nblk = 0, blk = 0;
    std::vector<uint32_t> words;
    std::vector<double> X;
    std::vector<uint8_t> ok;
    std::atomic<size_t> ready{0}, next{0};
    std::unique_ptr<std::atomic<uint8_t>[]> fin;   // per block: evaluated
    size_t fin_cap = 0;
    void eval(size_t b)
    {
        const size_t a0 = b * blk, a1 = std::min(M, a0 + blk);
        for (size_t j = a0; j < a1; ++j) {
            double x;
            ok[j] = gamma_attempt(&words[4 * j], alpha, x) ? 1 : 0;
            X[j] = x;
        }
        fin[b].store(1, std::memory_order_release);
    }
    std::atomic<int> active{0};                     // workers currently inside consume()
    void consume()
    {
        for (;;) {
            const size_t b = next.fetch_add(1);
            if (b >= nblk) return;
            while (ready.load(std::memory_order_acquire) <= b) std::this_thread::yield();
            eval(b);
        }
    }
    // caller only: blocks that a (possibly descheduled) worker claimed but has not finished are simply evaluated
    // again -- the result is a pure function of the words, both writers store identical values
    void finish()
    {
        for (size_t b = 0; b < nblk; ++b)
            if (!fin[b].load(std::memory_order_acquire)) eval(b);
    }
};

class Pool {
public:
    explicit Pool(int n)
    {
        for (int i = 0; i < n; ++i) th_.emplace_back([this] { loop(); });
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
            ++epoch_;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return (int)th_.size(); }
    void post(Job *job)
    {
        {
            std::lock_guard<std::mutex> l(m_);
            job_ = job;
            ++epoch_;
        }
        cv_.notify_all();
    }

private:
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            Job *job;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
                job = job_;
            }
            if (job) {
                job->active.fetch_add(1);
                job->consume();
                job->active.fetch_sub(1);
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_;
    Job *job_ = nullptr;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

Pool *pool_for(int threads)
{
    static Pool *pool = nullptr;           // (guarded by the caller's scratch mutex)
    if (!pool || pool->size() != threads - 1) {
        delete pool;
        pool = new Pool(threads - 1);      // the caller is the last worker
    }
    return pool;
}

}   // namespace

extern "C" int maz_legacy_dirichlet(unsigned int *key, int *pos, double alpha, int rows, int A, float *out, double *out64,
                                    int threads)
{
    if (!key || !pos || rows <= 0 || A <= 0 || *pos < 0 || *pos > 624 || (!out && !out64))
        return maz::set_last_error(1, "maz_legacy_dirichlet: bad arguments");
    if (!(alpha > 0.0) || !(alpha < 1.0))
        return maz::set_last_error(3, "maz_legacy_dirichlet: only 0 < alpha < 1 (numpy's rejection branch) is restated");
    if (threads <= 0) threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    const size_t total = (size_t)rows * A;
    MT mt;
    std::copy(key, key + 624, mt.key);
    mt.pos = *pos;

    // one call at a time; scratch lives across calls (fresh 1 MB vectors cost more in page faults than the whole draw)
    static std::mutex call_mutex;
    static std::vector<double> g;
    static std::vector<MT> snaps;            // generator state at the start of every block (to rewind cheaply)
    static Job &J = *new Job;                // ONE job object for the life of the process, never destroyed (a worker
                                             // that wakes up late may look at it even while the process exits)
    std::lock_guard<std::mutex> call_lock(call_mutex);
    if (g.size() < total) g.resize(total);
    size_t have = 0;
    Pool *pool = threads > 1 ? pool_for(threads) : nullptr;
    constexpr size_t BLK = 1024;            // attempts per block (4096 words)
    while (have < total) {
        // acceptance of this sampler is ~0.7-0.9 for small shapes: ask for a little more than what is missing
        const size_t M = std::max<size_t>(64, (size_t)((total - have) * 1.45) + 32);
        const size_t nblk = (M + BLK - 1) / BLK;
        // a worker that woke up late may still be evaluating a block of the PREVIOUS job (whose result the caller has
        // already recomputed): wait for it before the buffers are reused.  Reset order matters: `ready` before `next`,
        // so that a block claimed from the new counter can never pass the ready check on the old value.
        while (J.active.load() != 0) std::this_thread::yield();
        if (J.words.size() < 4 * M) J.words.resize(4 * M);
        if (J.X.size() < M) { J.X.resize(M); J.ok.resize(M); }
        if (snaps.size() < nblk) snaps.resize(nblk);
        J.alpha = alpha; J.M = M; J.nblk = nblk; J.blk = BLK;
        if (J.fin_cap < nblk) { J.fin.reset(new std::atomic<uint8_t>[nblk]); J.fin_cap = nblk; }
        for (size_t b = 0; b < nblk; ++b) J.fin[b].store(0, std::memory_order_relaxed);
        J.ready.store(0);
        J.next.store(0);
        if (pool) pool->post(&J);
        // the caller generates the word stream block by block (the only serial part) while the workers already
        // evaluate the blocks that are ready; then it joins them
        for (size_t b = 0; b < nblk; ++b) {
            const size_t a0 = b * BLK, a1 = std::min(M, a0 + BLK);
            snaps[b] = mt;
            mt.fill(J.words.data() + 4 * a0, 4 * (a1 - a0));
            J.ready.store(b + 1, std::memory_order_release);
        }
        J.consume();
        J.finish();
        size_t j = 0;
        for (; j < M && have < total; ++j)
            if (J.ok[j]) g[have++] = J.X[j];
        if (have == total && j < M) {       // the stream position is right after the last attempt actually made
            mt = snaps[j / BLK];
            mt.skip(4 * (j % BLK));
        }
    }
    // RandomState.dirichlet: acc = sum in order; val *= 1/acc
    for (int r = 0; r < rows; ++r) {
        double *v = &g[(size_t)r * A];
        double acc = 0.0;
        for (int a = 0; a < A; ++a) acc = acc + v[a];
        const double invacc = 1 / acc;
        for (int a = 0; a < A; ++a) {
            v[a] = v[a] * invacc;
            if (out) out[(size_t)r * A + a] = (float)v[a];
            if (out64) out64[(size_t)r * A + a] = v[a];
        }
    }
    std::copy(mt.key, mt.key + 624, key);
    *pos = mt.pos;
    return 0;
}
