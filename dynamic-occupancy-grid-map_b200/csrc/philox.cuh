// Counter-based Philox4x32-10 stream (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11).
// Replaces the reference's per-thread XORWOW states (curandState, 48 B each, dogm.cu:68,101; init.cu:16-22),
// whose stream depends on the launch geometry.  Here a draw is a pure function of
// (seed, particle slot, stage, cycle), so no RNG state is ever loaded or stored.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dogm_b200
{

enum PhiloxStage : uint32_t
{
    STAGE_INIT = 1,     // first-cycle velocities (init_new_particles.cu:116-117)
    STAGE_PREDICT = 2,  // process noise (predict.cu:27-30)
    STAGE_BIRTH = 3,    // birth velocities (init_new_particles.cu:178-187)
    STAGE_RESAMPLE = 4, // resampling offsets (resampling.cu:28)
};

struct Philox4
{
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
{
#ifdef __CUDA_ARCH__
    hi = __umulhi(a, b);
    lo = a * b;
#else
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                            uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++)
    {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(M0, c0, hi0, lo0);
        philox_mulhilo(M1, c2, hi1, lo1);
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n1 = lo1;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        const uint32_t n3 = lo0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += W0;
        k1 += W1;
    }
    Philox4 out = {c0, c1, c2, c3};
    return out;
}

// 24-bit uniforms: [0,1) and (0,1]
__host__ __device__ __forceinline__ float u01_half_open(uint32_t v)
{
    return (float)(v >> 8) * (1.0f / 16777216.0f);
}
__host__ __device__ __forceinline__ float u01_open_low(uint32_t v)
{
    return (float)((v >> 8) + 1u) * (1.0f / 16777216.0f);
}

#ifdef __CUDACC__
// Box-Muller pair from two 32-bit words
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b)
{
    const float u1 = u01_open_low(a);
    const float u2 = u01_half_open(b);
    // SFU intrinsics (the reference's curand_normal does the same, curand_normal.h:70-87): the noise is a pure
    // function of the counter on the device; a checker gets the values through dogm_export_philox_noise
    const float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

// four standard normals for (slot, stage, cycle)
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t slot, uint32_t stage, uint32_t cycle)
{
    const Philox4 p = philox4x32_10(slot, stage, cycle, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float2 a = box_muller(p.x, p.y);
    const float2 b = box_muller(p.z, p.w);
    return make_float4(a.x, a.y, b.x, b.y);
}
#endif

} // namespace dogm_b200
