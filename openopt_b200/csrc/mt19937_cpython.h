// CPython's `random` module stream, restated for the host side of OODE_RNG_CPYTHON.
//
// The reference's live RNG is the stdlib `random` (de_oo.py:16-53: `from np import random` at :12
// always raises, so Seed/Rand/Randint are random.seed / a per-element random.random() loop /
// a per-element random.randint(0, low-1) loop).  `random` is a third-party dependency of the
// reference (the interpreter's stdlib, not under /root/reference); its published algorithm
// (CPython Modules/_randommodule.c and Lib/random.py, unchanged since 3.2) is:
//   seed(int n)   : init_by_array(key = 32-bit little-endian limbs of abs(n))   [Matsumoto & Nishimura]
//   random()      : a = genrand_uint32() >> 5; b = genrand_uint32() >> 6;
//                   return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)
//   randint(0,n-1): _randbelow(n): k = n.bit_length(); r = getrandbits(k);
//                   while r >= n: r = getrandbits(k)      with getrandbits(k<=32) = genrand_uint32() >> (32-k)
// tests/test_host_logic.py checks this file (through liboode's test hook) against the
// interpreter's own `random`.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>

class MT19937CPython {
    static constexpr int N = 624, M = 397;
    uint32_t mt[N];
    int mti;

    void init_genrand(uint32_t s) {
        mt[0] = s;
        for (mti = 1; mti < N; mti++) mt[mti] = 1812433253u * (mt[mti - 1] ^ (mt[mti - 1] >> 30)) + (uint32_t)mti;
    }

    void init_by_array(const std::vector<uint32_t> &key) {
        init_genrand(19650218u);
        size_t i = 1, j = 0;
        const size_t klen = key.size();
        size_t k = (N > klen ? (size_t)N : klen);
        for (; k; k--) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            i++; j++;
            if (i >= (size_t)N) { mt[0] = mt[N - 1]; i = 1; }
            if (j >= klen) j = 0;
        }
        for (k = N - 1; k; k--) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            i++;
            if (i >= (size_t)N) { mt[0] = mt[N - 1]; i = 1; }
        }
        mt[0] = 0x80000000u;
    }

public:
    explicit MT19937CPython(uint64_t seed) {
        std::vector<uint32_t> key;
        key.push_back((uint32_t)(seed & 0xFFFFFFFFull));
        if (seed >> 32) key.push_back((uint32_t)(seed >> 32));
        init_by_array(key);
    }

    uint32_t genrand_uint32() {
        static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
        uint32_t y;
        if (mti >= N) {
            int kk;
            for (kk = 0; kk < N - M; kk++) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + M] ^ (y >> 1) ^ mag01[y & 1u];
            }
            for (; kk < N - 1; kk++) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ mag01[y & 1u];
            }
            y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ mag01[y & 1u];
            mti = 0;
        }
        y = mt[mti++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }

    double random() {
        const uint32_t a = genrand_uint32() >> 5, b = genrand_uint32() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }

    // random.randint(0, n-1) for 1 <= n < 2^32
    uint32_t randbelow(uint32_t n) {
        int k = 0;
        for (uint32_t t = n; t; t >>= 1) ++k;
        uint32_t r = genrand_uint32() >> (32 - k);
        while (r >= n) r = genrand_uint32() >> (32 - k);
        return r;
    }

    void fill_random(double *out, size_t n) {
        for (size_t i = 0; i < n; ++i) out[i] = random();
    }

    // generator state for checkpoints: 624 words + the position
    static constexpr size_t STATE_WORDS = N + 1;
    void save(uint32_t *out) const {
        for (int i = 0; i < N; ++i) out[i] = mt[i];
        out[N] = (uint32_t)mti;
    }
    void load(const uint32_t *in) {
        for (int i = 0; i < N; ++i) mt[i] = in[i];
        mti = (int)in[N];
    }
};
