// philox.cuh -- Philox4x32-10 counter-based RNG (Salmon et al., SC'11), one stream
// per energy packet.  Replaces the reference's wall-clock-seeded random_number
// (photon_mod.f90:68-87): key = 64-bit run seed, counter = (packet id lo, packet id
// hi, draw block, source index).  A uniform is (word >> 8) * 2^-24, the same 24-bit
// [0,1) grid a real(4) random_number produces.  Because the stream depends only on
// the global packet id, results are independent of how packets are sharded over
// threads, CTAs or GPUs.
#pragma once
#include <cstdint>

namespace mcb {

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]);
        uint32_t lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]);
        uint32_t lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0;
        uint32_t n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

struct Rng {
    uint32_t k0, k1, p0, p1, stream, n;
    uint32_t blk;            // block index held in buf (0xffffffff = none): lets a stream be
    uint32_t buf[4];         // resumed from its draw counter n alone (wave-front kernels)

    __device__ __forceinline__ void init(uint64_t seed, uint64_t pid, uint32_t s, uint32_t n0 = 0)
    {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        p0 = (uint32_t)pid;  p1 = (uint32_t)(pid >> 32);
        stream = s; n = n0; blk = 0xffffffffu;
    }
    __device__ __forceinline__ float uniform()
    {
        uint32_t lane = n & 3u;
        if ((n >> 2) != blk) {
            blk = n >> 2;
            buf[0] = p0; buf[1] = p1; buf[2] = blk; buf[3] = stream;
            philox4x32_10(buf, k0, k1);
        }
        ++n;
        uint32_t w = lane == 0 ? buf[0] : lane == 1 ? buf[1] : lane == 2 ? buf[2] : buf[3];
        return (float)(w >> 8) * 5.9604644775390625e-08f;
    }
    // The same uniform from the draw counter alone: key and counter words are passed in and the
    // block is recomputed on every call instead of being cached.  For code that draws rarely but
    // is short of registers (the FLY kernel: one draw per interaction, ~1 per 70 cell crossings):
    // only `n` of this struct stays live.
    __device__ __forceinline__ float uniform_at(uint64_t seed, uint64_t pid, uint32_t s)
    {
        uint32_t c[4] = {(uint32_t)pid, (uint32_t)(pid >> 32), n >> 2, s};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        uint32_t lane = n & 3u;
        ++n;
        uint32_t w = lane == 0 ? c[0] : lane == 1 ? c[1] : lane == 2 ? c[2] : c[3];
        return (float)(w >> 8) * 5.9604644775390625e-08f;
    }
};

}  // namespace mcb
