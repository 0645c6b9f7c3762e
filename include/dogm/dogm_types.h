// dogm_types.h — the reference's POD layouts (reference dogm/include/dogm/dogm_types.h:13-144) for users of the
// drop-in C++ class in dogm.h.  Same field order, sizes and semantics; GLM is not needed: `dogm::vec4` stands in for
// glm::vec4 (four floats, operator[], .x/.y/.z/.w, operator==, operator+), and when <glm/vec4.hpp> is available define
// DOGM_B200_USE_GLM before including this header to get the real type back.
#pragma once

#include "../dogm_b200.h"

#include <cassert>
#include <cstdlib>
#include <cstring>

#ifdef DOGM_B200_USE_GLM
#include <glm/vec4.hpp>
#endif

namespace dogm
{

#ifdef DOGM_B200_USE_GLM
using vec4 = glm::vec4;
#else
struct vec4
{
    float x, y, z, w;
    vec4() = default;
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
inline bool operator==(const vec4& a, const vec4& b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
inline bool operator!=(const vec4& a, const vec4& b) { return !(a == b); }
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator*(float s, const vec4& a) { return vec4(s * a.x, s * a.y, s * a.z, s * a.w); }
#endif

using GridCell = ::dogm_grid_cell;        // dogm_types.h:13-33, 64 bytes
using MeasurementCell = ::dogm_meas_cell; // dogm_types.h:35-41, 16 bytes

struct Particle // dogm_types.h:43-49 (sizing only: sizeof == 28)
{
    int grid_cell_idx;
    float weight;
    bool associated;
    vec4 state;
};
static_assert(sizeof(GridCell) == 64 && sizeof(MeasurementCell) == 16 && sizeof(Particle) == 28, "reference layouts");

// ParticlesSoA, dogm_types.h:51-144: one block of size * sizeof(Particle) bytes,
//   state[size] | grid_cell_idx[size] | weight[size] | associated[size].
// Host-side sets own malloc'ed memory and must be free()d by the caller exactly like the reference's
// (dogm_spec.cpp:62-65); device-side sets handed out by DOGM are views of buffers the DOGM object owns.
struct ParticlesSoA
{
    vec4* state = nullptr;
    int* grid_cell_idx = nullptr;
    float* weight = nullptr;
    bool* associated = nullptr;

    void* memory_block = nullptr;
    int size = 0;
    bool device = true;
    bool owns = false;

    ParticlesSoA() = default;
    ParticlesSoA(int new_size, bool is_device) { init(new_size, is_device); }

    void init(int new_size, bool is_device)
    {
        size = new_size;
        device = is_device;
        const size_t num_bytes = DOGM_PARTICLE_BLOCK_BYTES(size);
        if (device)
        {
            void* p = nullptr;
            dogm_device_alloc(&p, num_bytes);
            memory_block = p;
        }
        else
        {
            memory_block = std::malloc(num_bytes ? num_bytes : 1);
        }
        owns = true;
        assignPointers();
    }

    // view of a block owned by somebody else (DOGM's public particle_array members)
    static ParticlesSoA view(void* block, int count, bool is_device)
    {
        ParticlesSoA p;
        p.memory_block = block;
        p.size = count;
        p.device = is_device;
        p.owns = false;
        p.assignPointers();
        return p;
    }

    void free()
    {
        assert(size);
        if (owns)
        {
            if (device)
                dogm_device_free(memory_block);
            else
                std::free(memory_block);
        }
        memory_block = nullptr;
        size = 0;
    }

    void copy(const ParticlesSoA& other) // dogm_types.h:103-116
    {
        assert(size && size == other.size);
        const size_t num_bytes = DOGM_PARTICLE_BLOCK_BYTES(size);
        if (!device && !other.device)
            std::memcpy(memory_block, other.memory_block, num_bytes);
        else if (device && !other.device)
            dogm_memcpy_h2d(memory_block, other.memory_block, num_bytes);
        else if (!device && other.device)
            dogm_memcpy_d2h(memory_block, other.memory_block, num_bytes);
        else
            dogm_memcpy_d2d(memory_block, other.memory_block, num_bytes); // device to device, also across GPUs of this process
    }

    // dogm_types.h:118-126: assignment COPIES THE CONTENTS into this set's own block (sizes must match, like the reference's
    // assert).  Only an empty set (size 0, nothing to copy into) adopts the other's block as a non-owning view - the
    // reference would assert there.  The copy constructor stays memberwise, as in the reference (returning a set by value
    // hands the block on).
    ParticlesSoA(const ParticlesSoA&) = default;
    ParticlesSoA& operator=(const ParticlesSoA& other)
    {
        if (this == &other)
            return *this;
        if (size == 0 || memory_block == nullptr)
            rebind(other.memory_block, other.size, other.device);
        else
            copy(other);
        return *this;
    }

    // points this set at a block owned by somebody else (what DOGM does with its public members after every call)
    void rebind(void* block, int count, bool is_device)
    {
        memory_block = block;
        size = count;
        device = is_device;
        owns = false;
        assignPointers();
    }

  private:
    void assignPointers()
    {
        char* b = static_cast<char*>(memory_block);
        state = reinterpret_cast<vec4*>(b + DOGM_PARTICLE_STATE_OFFSET(size));
        grid_cell_idx = reinterpret_cast<int*>(b + DOGM_PARTICLE_IDX_OFFSET(size));
        weight = reinterpret_cast<float*>(b + DOGM_PARTICLE_WEIGHT_OFFSET(size));
        associated = reinterpret_cast<bool*>(b + DOGM_PARTICLE_ASSOC_OFFSET(size));
    }
};

} // namespace dogm
