// pqperm_rng.cpp -- the raw 64-bit streams of many numpy generators at once (host).
//
// The reference sampler gives shot idx its own np.random.default_rng(seed + idx)
// (piquasso/_simulators/passive/sampling.py:149-194).  default_rng(s) is
// Generator(PCG64(SeedSequence(s))); creating 10^4 of them from Python and pulling
// their raw outputs costs 75 ms per run -- more than the GPU needs for the first
// fifteen photons.  This restates, for ALL shots in one call,
//   * numpy's SeedSequence (numpy/random/bit_generator.pyx: mix_entropy, generate_state)
//     for an integer seed below 2^64 and an empty spawn key,
//   * PCG64's seeding (pcg64_set_seed -> pcg_setseq_128_srandom_r) and its output
//     function (XSL-RR 128/64, step first),
// so that out[i * draws + k] == np.random.PCG64(seed0 + i).random_raw(draws)[k].
// tests/test_host.py::test_shot_streams_replay_numpy_generators pins it against real
// numpy generators.
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/pqperm.h"

namespace {

constexpr uint32_t kInitA = 0x43b0d7e5u, kMultA = 0x931e8875u;
constexpr uint32_t kInitB = 0x8b51f9ddu, kMultB = 0x58f38dedu;
constexpr uint32_t kMixL = 0xca01f9ddu, kMixR = 0x4973f715u;
constexpr int kPool = 4;

inline uint32_t hashmix(uint32_t value, uint32_t &hash_const)
{
    value ^= hash_const;
    hash_const *= kMultA;
    value *= hash_const;
    value ^= value >> 16;
    return value;
}

inline uint32_t mix(uint32_t x, uint32_t y)
{
    uint32_t r = kMixL * x - kMixR * y;
    r ^= r >> 16;
    return r;
}

// SeedSequence(seed).generate_state(4, uint64) for an integer seed < 2^64
void seed_state(uint64_t seed, uint64_t out[4])
{
    uint32_t entropy[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const int nent = entropy[1] ? 2 : 1; // little-endian 32-bit words, at least one
    uint32_t pool[kPool];
    uint32_t hc = kInitA;
    for (int i = 0; i < kPool; i++)
        pool[i] = hashmix(i < nent ? entropy[i] : 0u, hc);
    for (int src = 0; src < kPool; src++)
        for (int dst = 0; dst < kPool; dst++)
            if (src != dst)
                pool[dst] = mix(pool[dst], hashmix(pool[src], hc));
    uint32_t words[8];
    hc = kInitB;
    for (int i = 0; i < 8; i++) {
        uint32_t v = pool[i % kPool];
        v ^= hc;
        hc *= kMultB;
        v *= hc;
        v ^= v >> 16;
        words[i] = v;
    }
    for (int k = 0; k < 4; k++)
        out[k] = (uint64_t)words[2 * k] | ((uint64_t)words[2 * k + 1] << 32);
}

typedef unsigned __int128 u128;
const u128 kPcgMult = ((u128)0x2360ED051FC65DA4ull << 64) | (u128)0x4385DF649FCCF645ull;

inline uint64_t rotr64(uint64_t v, unsigned r) { return (v >> r) | (v << ((-r) & 63)); }

void stream(uint64_t seed, int draws, uint64_t *out)
{
    uint64_t val[4];
    seed_state(seed, val);
    // pcg64_set_seed: seed = (high val[0], low val[1]), inc = (high val[2], low val[3])
    const u128 initstate = ((u128)val[0] << 64) | val[1];
    const u128 initseq = ((u128)val[2] << 64) | val[3];
    u128 inc = (initseq << 1) | 1;
    u128 state = 0;
    state = state * kPcgMult + inc;
    state += initstate;
    state = state * kPcgMult + inc;
    for (int k = 0; k < draws; k++) {
        state = state * kPcgMult + inc;
        out[k] = rotr64((uint64_t)(state >> 64) ^ (uint64_t)state, (unsigned)(state >> 122));
    }
}

} // namespace

extern "C" int pq_pcg64_streams(uint64_t seed0, int64_t n, int draws, uint64_t *out)
{
    if (n < 0 || draws < 0 || (n > 0 && draws > 0 && !out))
        return PQ_ERR_BAD_ARG;
    const int nthreads = (int)std::max<int64_t>(
        1, std::min<int64_t>({(int64_t)8, n / 2048, (int64_t)std::thread::hardware_concurrency()}));
    auto work = [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; i++)
            stream(seed0 + (uint64_t)i, draws, out + (size_t)i * draws);
    };
    if (nthreads == 1) {
        work(0, n);
        return PQ_OK;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++)
        pool.emplace_back(work, n * t / nthreads, n * (t + 1) / nthreads);
    for (std::thread &t : pool)
        t.join();
    return PQ_OK;
}
