// oc_rng.h -- numpy's LEGACY random stream (np.random.seed / RandomState: MT19937), restated for the host side of the
// library so that the per-step randomness of the GCFM sweep and the crowd placement can be drawn in C with exactly the
// reference's stream (simulations.py:130-131,140,271,303):
//   next_double      genrand_res53
//   interval(max)    random_interval: masked rejection on 32-bit draws (max < 2^32)
//   permutation(n)   RandomState.permutation(n) == choice(arange(n), n, replace=False): Fisher-Yates from the top
//   gauss()          legacy_gauss: polar Box-Muller with one cached value (has_gauss / cached_gaussian of get_state())
// Pinned against numpy itself by tests/test_cpu_host.py (bit-identical draws and generator state).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

namespace ocrng {

// Persistent helper threads for the Box-Muller transform of gauss_fill.  The rejection loop that advances the generator
// is inherently serial; f = sqrt(-2 log(r2) / r2) of the accepted events is not, and costs twice as much.  The producer
// (the thread inside gauss_fill) publishes how many events it has accepted so far; the helpers -- parked on a condition
// variable between calls, so that a call costs a wake-up and not a thread creation -- transform blocks of EV_BLOCK events
// behind it, the producer joins in once the generator is done.  Every value is computed by the same scalar expression
// whoever evaluates it: the stream is bit-identical to the serial one (tests/test_cpu_host.py).
// Threads: OC_RNG_THREADS (default: min(3, hardware threads per local process - 2)); 0 = serial.
class TransformPool {
  public:
    static constexpr long long EV_BLOCK = 512;
    static TransformPool &get() {
        // never destroyed: the helpers stay parked on the condition variable until the process exits, and destroying a
        // condition variable with waiters blocks (glibc) -- the process would hang in its static destructors
        static TransformPool *p = new TransformPool();
        return *p;
    }
    int threads() const { return (int)workers_.size(); }
    // returns false if the pool is absent or busy (another generator is using it): the caller transforms inline
    bool begin(const double *X1, const double *X2, const double *R2, double *out, long long idx, long long n, long long n_ev) {
        if (workers_.empty() || !busy_.try_lock()) return false;
        // A helper that has just finished the previous job may still be on its way to its next ticket: everything it
        // could read for a ticket of THIS job is in place before the ticket counter is reset (release / acq_rel pair),
        // and the progress counter is zeroed first, so that it cannot run ahead of the producer.
        produced_.store(0, std::memory_order_relaxed);
        done_.store(0, std::memory_order_relaxed);
        X1_ = X1; X2_ = X2; R2_ = R2; out_ = out; idx_ = idx; n_ = n; n_ev_ = n_ev;
        n_blocks_ = (n_ev + EV_BLOCK - 1) / EV_BLOCK;
        // tickets and the job descriptor carry the job's generation: a ticket drawn from the previous job's counter can
        // never be mistaken for a block of this one (it fails the generation test, or -- against the old descriptor --
        // the range test)
        gen_ = (gen_ + 1) & 0x7fffffffull;
        desc_.store((gen_ << 32) | (unsigned long long)n_blocks_, std::memory_order_release);
        next_.store((long long)(gen_ << 32), std::memory_order_release);
        {
            std::lock_guard<std::mutex> lk(m_);
            job_++;
        }
        cv_.notify_all();
        return true;
    }
    void publish(long long produced) { produced_.store(produced, std::memory_order_release); }
    void finish() {  // producer: all events are accepted; help, then wait for the helpers' last blocks
        produced_.store(n_ev_, std::memory_order_release);
        work();
        while (done_.load(std::memory_order_acquire) < n_blocks_) std::this_thread::yield();
        busy_.unlock();
    }

  private:
    TransformPool() {
        int n = -1;
        if (const char *e = std::getenv("OC_RNG_THREADS")) n = std::atoi(e);
        if (n < 0) {
            // leave two cores per process to its main thread and the look-ahead thread that calls gauss_fill; under
            // torchrun (LOCAL_WORLD_SIZE processes share the host) the cores are divided first
            int procs = 1;
            if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) procs = std::max(1, std::atoi(e));
            n = std::max(0, std::min(3, (int)std::thread::hardware_concurrency() / procs - 2));
        }
        for (int t = 0; t < n; t++) workers_.emplace_back([this] { loop(); });
        for (auto &w : workers_) w.detach();  // parked for the life of the process
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return job_ != seen; });
                seen = job_;
            }
            work();
        }
    }
    void work() {
        for (;;) {
            const unsigned long long t = (unsigned long long)next_.fetch_add(1, std::memory_order_acq_rel);
            const unsigned long long d = desc_.load(std::memory_order_acquire);
            if ((t >> 32) != (d >> 32) || (t & 0xffffffffull) >= (d & 0xffffffffull)) return;
            const long long b = (long long)(t & 0xffffffffull);
            const long long e0 = b * EV_BLOCK, e1 = std::min(n_ev_, e0 + EV_BLOCK);
            while (produced_.load(std::memory_order_acquire) < e1) std::this_thread::yield();
            for (long long e = e0; e < e1; e++) {
                const double f = std::sqrt(-2.0 * std::log(R2_[e]) / R2_[e]);
                out_[idx_ + 2 * e] = f * X2_[e];
                if (idx_ + 2 * e + 1 < n_) out_[idx_ + 2 * e + 1] = f * X1_[e];
            }
            done_.fetch_add(1, std::memory_order_release);
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_, busy_;
    std::condition_variable cv_;
    unsigned long long job_ = 0;
    const double *X1_ = nullptr, *X2_ = nullptr, *R2_ = nullptr;
    double *out_ = nullptr;
    long long idx_ = 0, n_ = 0, n_ev_ = 0, n_blocks_ = 0;
    std::atomic<long long> next_{0}, done_{0}, produced_{0};
    std::atomic<unsigned long long> desc_{0};   // (generation << 32) | number of blocks of the current job
    unsigned long long gen_ = 0;                // written by the producer that holds busy_
};


struct Mt {
    uint32_t *key;   // 624 words (np.random.get_state()[1]), advanced in place
    int pos;         // get_state()[2]
    int has_gauss;   // get_state()[3]
    double gauss_;   // get_state()[4]

    void gen() {  // numpy's mt19937_gen == the reference genrand
        const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
        int i;
        uint32_t y;
        for (i = 0; i < 624 - 397; i++) {
            y = (key[i] & UPPER) | (key[i + 1] & LOWER);
            key[i] = key[i + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT);
        }
        for (; i < 623; i++) {
            y = (key[i] & UPPER) | (key[i + 1] & LOWER);
            key[i] = key[i + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT);
        }
        y = (key[623] & UPPER) | (key[0] & LOWER);
        key[623] = key[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT);
        pos = 0;
    }
    // tempered copy of key[from .. 624): the output transform of a whole block in one vectorisable loop instead of word
    // by word (the serial generator is what bounds a step's draw: 5 words per agent)
    uint32_t tb[624] = {};
    bool tb_valid = false;
    void temper_block(int from) {
        for (int i = from; i < 624; i++) {
            uint32_t y = key[i];
            y ^= (y >> 11);
            y ^= (y << 7) & 0x9d2c5680u;
            y ^= (y << 15) & 0xefc60000u;
            y ^= (y >> 18);
            tb[i] = y;
        }
        tb_valid = true;
    }
    uint32_t next32() {
        if (pos == 624) { gen(); temper_block(0); }
        else if (!tb_valid) temper_block(pos);
        return tb[pos++];
    }
    double next_double() {
        const int32_t a = next32() >> 5, b = next32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    uint32_t interval(uint32_t max) {
        if (max == 0) return 0;
        uint32_t mask = max;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        while ((v = (next32() & mask)) > max) {}
        return v;
    }
    void permutation(int n, int *out) {
        for (int i = 0; i < n; i++) out[i] = i;
        for (int i = n - 1; i > 0; i--) {
            const int j = (int)interval((uint32_t)i);
            const int t = out[i]; out[i] = out[j]; out[j] = t;
        }
    }
    double gauss() {
        if (has_gauss) {
            const double t = gauss_;
            has_gauss = 0;
            gauss_ = 0.0;
            return t;
        }
        double f, x1, x2, r2;
        do {
            x1 = 2.0 * next_double() - 1.0;
            x2 = 2.0 * next_double() - 1.0;
            r2 = x1 * x1 + x2 * x2;
        } while (r2 >= 1.0 || r2 == 0.0);
        f = std::sqrt(-2.0 * std::log(r2) / r2);  // libm log, as numpy's legacy_gauss
        gauss_ = f * x1;
        has_gauss = 1;
        return f * x2;
    }
    // n consecutive gauss() values into out[], same values and same final state as n calls (the generator is advanced
    // first -- the rejection test needs only x1, x2 -- then f = sqrt(-2 log(r2) / r2) is evaluated).  ckpt (optional): snapshots of the generator state as it is after
    // value number n - 2q, q = 0 .. n_ckpt-1 (q-th snapshot: key at ckpt_key + 624 q, ...), for a caller that drew more
    // values than it turns out to need (look-ahead of the GCFM step: values come in pairs, n even).
    void gauss_fill(double *out, long long n, int n_ckpt = 0, uint32_t *ckpt_key = nullptr, int *ckpt_pos = nullptr,
                    int *ckpt_has = nullptr, double *ckpt_cached = nullptr) {
        long long idx = 0;
        const int had = has_gauss;
        const double had_val = gauss_;
        if (n > 0 && has_gauss) { out[idx++] = gauss_; has_gauss = 0; gauss_ = 0.0; }
        const long long rem = n - idx, n_ev = (rem + 1) / 2;
        std::vector<double> X1(n_ev), X2(n_ev), R2(n_ev);
        // snapshot q is wanted after value n - 2q = after event e_q = (n - 2q - idx + 1) / 2 - 1 (or the start state)
        auto snap = [&](int q) {
            if (!ckpt_key) return;
            std::memcpy(ckpt_key + (size_t)624 * q, key, 624 * sizeof(uint32_t));
            ckpt_pos[q] = pos;
        };
        std::vector<long long> ev_of(n_ckpt, -2);
        for (int q = 0; q < n_ckpt; q++) {
            const long long v = n - 2ll * q;            // values drawn at that point
            if (v < 0) continue;
            ev_of[q] = (v <= idx) ? -1 : (v - idx + 1) / 2 - 1;
            if (ev_of[q] == -1) snap(q);                // the state before any new event (only the cached value was used)
        }
        long long first_ck_ev = n_ev;
        for (int q = 0; q < n_ckpt; q++)
            if (ev_of[q] >= 0) first_ck_ev = std::min(first_ck_ev, ev_of[q]);
        // the transform runs behind the generator on the pool's helper threads (large draws only: a wake-up costs ~30 us)
        TransformPool *pool = nullptr;
        if (n_ev >= 8 * TransformPool::EV_BLOCK && TransformPool::get().begin(X1.data(), X2.data(), R2.data(), out, idx, n, n_ev))
            pool = &TransformPool::get();
        for (long long e = 0; e < n_ev; e++) {
            double x1, x2, r2;
            do {
                x1 = 2.0 * next_double() - 1.0;
                x2 = 2.0 * next_double() - 1.0;
                r2 = x1 * x1 + x2 * x2;
            } while (r2 >= 1.0 || r2 == 0.0);
            X1[e] = x1; X2[e] = x2; R2[e] = r2;
            if (pool && ((e + 1) & (TransformPool::EV_BLOCK - 1)) == 0) pool->publish(e + 1);
            if (e >= first_ck_ev)   // the snapshots sit at the last n_ckpt events
                for (int q = 0; q < n_ckpt; q++)
                    if (ev_of[q] == e) snap(q);
        }
        if (pool) {
            pool->finish();
        } else {
            for (long long e = 0; e < n_ev; e++) {
                const double f = std::sqrt(-2.0 * std::log(R2[e]) / R2[e]);
                out[idx + 2 * e] = f * X2[e];
                if (idx + 2 * e + 1 < n) out[idx + 2 * e + 1] = f * X1[e];
            }
        }
        // cache flag / value as they are after v values (v = n for the generator itself, n - 2q for the snapshots)
        auto cached_after = [&](long long v, int *has, double *cv) {
            if (v == 0) { *has = had; *cv = had_val; return; }
            if (v <= idx || (v - idx) % 2 == 0) { *has = 0; *cv = 0.0; return; }
            const long long e = (v - idx + 1) / 2 - 1;      // the event whose second value is still cached
            *has = 1;
            *cv = std::sqrt(-2.0 * std::log(R2[e]) / R2[e]) * X1[e];
        };
        cached_after(n, &has_gauss, &gauss_);
        for (int q = 0; q < n_ckpt; q++) {
            const long long v = n - 2ll * q;
            if (v < 0 || !ckpt_key) continue;
            cached_after(v, &ckpt_has[q], &ckpt_cached[q]);
        }
    }
};

}  // namespace ocrng
