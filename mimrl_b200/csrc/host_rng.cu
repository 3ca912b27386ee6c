// Host side of the k-NN sampler's query draw (Model.py:81: np.random.choice(range(N), size=m, replace=False)).
// numpy's legacy RandomState.choice without replacement is permutation(N)[:m]: arange(N), then a Fisher-Yates pass
// from i = N - 1 down to 1 with j = random_interval(i) (masked rejection on MT19937 32-bit outputs) -- N - 1 draws of
// a variable number of outputs each, so neither the first m entries nor the generator state afterwards can be had
// without walking the whole stream.  This file restates that published algorithm (numpy/random/mtrand.pyx
// RandomState.shuffle/_shuffle_raw, src/distributions/distributions.c random_interval, src/mt19937/mt19937.c) on the
// caller's copy of the generator state, so that the draw and the state afterwards are numpy's bit for bit, and runs it
// faster than numpy does (11.3 -> 2.5 ms at N = 2^20 on the B200 host): the j's depend on the stream only, not on the
// array, so they are produced ahead (by a second thread for large pools) and their cache lines prefetched before the
// swaps touch them; the array is int32.
// No GPU work here; it lives in this library so that the reference-facing sampler needs nothing else.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>

#include "common.cuh"

namespace {

constexpr int kMtN = 624, kMtM = 397;

inline void mt_regenerate(uint32_t *mt) {
  int kk = 0;
  uint32_t y;
  for (; kk < kMtN - kMtM; ++kk) {
    y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
    mt[kk] = mt[kk + kMtM] ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
  }
  for (; kk < kMtN - 1; ++kk) {
    y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
    mt[kk] = mt[kk + (kMtM - kMtN)] ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
  }
  y = (mt[kMtN - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
  mt[kMtN - 1] = mt[kMtM - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
}

struct Mt {
  uint32_t *key;
  int pos;
  uint32_t out[kMtN];          // the tempered words of the current block (tempering a whole block vectorises)
  inline void temper() {
    for (int k = 0; k < kMtN; ++k) {
      uint32_t y = key[k];
      y ^= y >> 11;
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= y >> 18;
      out[k] = y;
    }
  }
  inline uint32_t next() {
    if (pos == kMtN) mt_regenerate(key), temper(), pos = 0;
    return out[pos++];
  }
};

// `want` swap targets for i, i - 1, ..., i - want + 1 (numpy's random_interval: masked rejection on 32-bit outputs).
// Branch-free: a rejected value is overwritten by the next draw.
struct Drawer {
  Mt g;
  void fill(uint32_t *dst, int64_t i, int want) {
    int cnt = 0;
    if (want > 0 && __builtin_clz((uint32_t)i) == __builtin_clz((uint32_t)(i - want + 1))) {
      // one mask for the whole block (i crosses a power of two ~20 times per draw); v <= lo is accepted and v > top
      // rejected whatever the count, only the `want` values in between need it -- so the chain between consecutive
      // draws is the count alone
      const uint32_t mask = 0xffffffffu >> __builtin_clz((uint32_t)i);
      const uint32_t top = (uint32_t)i, lo = top - (uint32_t)want;
      while (cnt < want) {
        if (g.pos == kMtN) mt_regenerate(g.key), g.temper(), g.pos = 0;
        int k = g.pos;
        for (; k < kMtN && cnt < want; ++k) {
          const uint32_t v = g.out[k] & mask;
          dst[cnt] = v;
          uint32_t ok = v <= lo ? 1u : 0u;
          if (__builtin_expect(v - lo - 1u < (uint32_t)want, 0)) ok = v <= top - (uint32_t)cnt ? 1u : 0u;
          cnt += (int)ok;
        }
        g.pos = k;
      }
    } else {
      while (cnt < want) {
        const uint32_t mx = (uint32_t)(i - cnt);
        const uint32_t v = g.next() & (0xffffffffu >> __builtin_clz(mx));
        dst[cnt] = v;
        cnt += v <= mx ? 1 : 0;
      }
    }
  }
};

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#else
  std::this_thread::yield();
#endif
}

struct Scratch {
  int32_t *p = nullptr;
  size_t n = 0;
  int32_t *get(size_t want) {
    if (want > n) {
      free(p);
      p = static_cast<int32_t *>(malloc(want * sizeof(int32_t)));
      n = p ? want : 0;
    }
    return p;
  }
  ~Scratch() { free(p); }
};

}  // namespace

// key[624], *pos: numpy's np.random.get_state()[1:3]; updated in place.  out[m] = np.random.permutation(n)[:m].
extern "C" int mimrl_legacy_permutation_head(uint32_t *key, int *pos, int64_t n, int64_t m, int64_t *out) {
  MIMRL_REQUIRE(key && pos && out && n >= 1 && m >= 0 && m <= n && n <= 0x7fffffff && *pos >= 0 && *pos <= kMtN,
                "legacy_permutation_head: bad arguments");
  static thread_local Scratch scratch;          // kept between calls: no page faults on a fresh 4 MB block per draw
  int32_t *arr = scratch.get((size_t)n);
  MIMRL_REQUIRE(arr, "legacy_permutation_head: out of memory");
  for (int64_t i = 0; i < n; ++i) arr[i] = (int32_t)i;
  Drawer d;
  d.g.key = key, d.g.pos = *pos;
  d.g.temper();
  static const int two_threads_from = getenv("MIMRL_RNG_THREADS_FROM") ? atoi(getenv("MIMRL_RNG_THREADS_FROM")) : (1 << 18);
  bool done = false;
  if (n >= two_threads_from) {
    // Large pools: the targets are produced by a second thread (generator + rejection, ~60 % of the time) while this
    // one swaps; blocks of 2048 targets travel through a ring of 8 slots.
    constexpr int kBig = 2048, kRing = 8;
    struct Slot {
      uint32_t js[kBig];
      int64_t i0;
      int cnt;
    };
    Slot *ring = static_cast<Slot *>(malloc(sizeof(Slot) * kRing));
    if (ring) {
      std::atomic<int64_t> produced{0}, consumed{0};
      const int64_t n_blocks = (n - 1 + kBig - 1) / kBig;
      try {
        std::thread producer([&] {
          int64_t i = n - 1;
          for (int64_t b = 0; b < n_blocks; ++b) {
            while (b - consumed.load(std::memory_order_acquire) >= kRing) cpu_relax();
            Slot &s = ring[b % kRing];
            const int want = i >= kBig ? kBig : (int)i;
            d.fill(s.js, i, want);
            s.i0 = i, s.cnt = want;
            i -= want;
            produced.store(b + 1, std::memory_order_release);
          }
        });
        for (int64_t b = 0; b < n_blocks; ++b) {
          while (produced.load(std::memory_order_acquire) <= b) cpu_relax();
          const Slot &s = ring[b % kRing];
          const int64_t i0 = s.i0;
          const int cnt = s.cnt;
          for (int t = 0; t < 32 && t < cnt; ++t) __builtin_prefetch(arr + s.js[t], 1);
          for (int t = 0; t < cnt; ++t) {
            if (t + 32 < cnt) __builtin_prefetch(arr + s.js[t + 32], 1);
            const uint32_t j = s.js[t];
            const int32_t v = arr[j];
            arr[j] = arr[i0 - t];
            arr[i0 - t] = v;
          }
          consumed.store(b + 1, std::memory_order_release);
        }
        producer.join();
        done = true;
      } catch (...) {          // no second thread available: nothing has been consumed from the generator yet
        done = false;
      }
      free(ring);
    }
  }
  if (!done) {
    // The swap targets depend on the stream only, not on the array: they are drawn one block ahead and prefetched, then
    // the previous block is swapped.
    constexpr int kBlock = 64;
    uint32_t js[2][kBlock];
    int cnts[2] = {0, 0};
    int64_t starts[2] = {0, 0};
    int cur = 0;
    int64_t i = n - 1;
    auto draw_block = [&](int slot) {
      const int want = i >= kBlock ? kBlock : (int)(i > 0 ? i : 0);
      d.fill(js[slot], i, want);
      for (int b = 0; b < want; ++b) __builtin_prefetch(arr + js[slot][b], 1);
      starts[slot] = i, cnts[slot] = want;
      i -= want;
    };
    draw_block(cur);
    while (cnts[cur] > 0) {
      draw_block(cur ^ 1);
      const int64_t i0 = starts[cur];
      for (int b = 0; b < cnts[cur]; ++b) {
        const uint32_t j = js[cur][b];
        const int32_t t = arr[j];
        arr[j] = arr[i0 - b];
        arr[i0 - b] = t;
      }
      cur ^= 1;
    }
  }
  for (int64_t t = 0; t < m; ++t) out[t] = arr[t];
  *pos = d.g.pos;
  return 0;
}
