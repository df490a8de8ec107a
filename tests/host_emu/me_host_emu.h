// me_host_emu.h -- TEST INFRASTRUCTURE: lets the device source of the motion search (csrc/me_device.cuh, thread-only
// build) compile and run on the HOST so that `-m "not gpu"` tests can execute the very code the frame-search kernel runs
// (tests/test_me_host_emu_cpu.py compares it with the reference's MotionEstimate).  CUDA qualifiers become no-ops, the
// handful of intrinsics the header uses are restated below, shared memory is a host buffer (emu::smem_base), and the
// lanes of one PU run as host threads that meet at a barrier for every shuffle -- they follow identical control flow by
// construction, exactly what the kernel relies on.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <type_traits>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __constant__ static const
#define __builtin_assume(x) ((void)0)
#define __syncwarp(...) ((void)0)

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct int4 { int x, y, z, w; };
inline int4 make_int4(int x, int y, int z, int w) { int4 r = { x, y, z, w }; return r; }

namespace x265b200 {
namespace emu {
struct Group { pthread_barrier_t bar; int n; int slot[32]; };
extern unsigned char* smem_base;                 // start of the emulated shared-memory allocation
extern thread_local Group* t_group;              // lanes of the PU this host thread belongs to
extern thread_local int t_q;                     // this lane's index inside the group
}

inline bool __isShared(const void*) { return true; }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)((const unsigned char*)p - emu::smem_base); }
template<typename T> inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }

inline int __shfl_xor_sync(unsigned, int v, int o)
{
    emu::Group* g = emu::t_group;
    if (!g || g->n == 1) return v;
    g->slot[emu::t_q] = v;
    pthread_barrier_wait(&g->bar);
    const int r = g->slot[emu::t_q ^ o];
    pthread_barrier_wait(&g->bar);
    return r;
}
inline int __shfl_sync(unsigned, int v, int) { abort(); return v; }     // warp-cooperative paths are not emulated

inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
    const uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++)
    {
        const uint32_t n = (sel >> (4 * i)) & 0xf;
        uint32_t byte = (uint32_t)(src >> (8 * (n & 7))) & 0xff;
        if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31)); }
inline uint32_t __vsadu4(uint32_t a, uint32_t b)
{
    uint32_t s = 0;
    for (int i = 0; i < 4; i++) { int d = (int)((a >> (8 * i)) & 0xff) - (int)((b >> (8 * i)) & 0xff); s += d < 0 ? -d : d; }
    return s;
}
inline uint32_t __vsadu2(uint32_t a, uint32_t b)
{
    uint32_t s = 0;
    for (int i = 0; i < 2; i++) { int d = (int)((a >> (16 * i)) & 0xffff) - (int)((b >> (16 * i)) & 0xffff); s += d < 0 ? -d : d; }
    return s;
}
inline uint32_t sad_u16x2(uint32_t a, uint32_t b) { return __vsadu2(a, b); }      // common.cuh's 16-bit SAD helper (same value)
inline uint32_t __vavgu4(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= ((((a >> (8 * i)) & 0xff) + ((b >> (8 * i)) & 0xff) + 1) >> 1) << (8 * i);
    return r;
}
inline uint32_t __vavgu2(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 2; i++) r |= ((((a >> (16 * i)) & 0xffff) + ((b >> (16 * i)) & 0xffff) + 1) >> 1) << (16 * i);
    return r;
}
inline int __dp2a_lo(int a, int b, int c) { return c + (int)(int16_t)(a & 0xffff) * (int)(int8_t)(b & 0xff) + (int)(int16_t)((uint32_t)a >> 16) * (int)(int8_t)((b >> 8) & 0xff); }
inline int __vimin_s32_relu(int a, int b) { int v = a < b ? a : b; return v < 0 ? 0 : v; }

inline int min(int a, int b) { return a < b ? a : b; }                   // CUDA's global integer min / max
inline int max(int a, int b) { return a > b ? a : b; }
inline int abs(int a) { return a < 0 ? -a : a; }

// the two helpers of common.cuh the search uses
inline int warp_sum(int v) { abort(); return v; }                          // warp-cooperative paths are not emulated
inline int clip3i(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

} // namespace x265b200
