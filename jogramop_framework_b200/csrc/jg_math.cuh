// jg_math.cuh -- small fp32 vector / rigid-transform helpers shared by the kernels.
#pragma once
#include <stdint.h>
#ifdef JG_HOST_EMU
// host emulation shim: lets tests compile the per-thread device math (GJK, simplex solver) with g++ so the
// fp32 numerics can be checked against the oracle on a machine without a GPU (tests/emu/).  Never defined
// in the shipped library.
#include <math.h>
#define __host__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __uint_as_float(uint32_t u) { union { uint32_t u; float f; } c; c.u = u; return c.f; }
static inline uint32_t __float_as_uint(float f) { union { uint32_t u; float f; } c; c.f = f; return c.u; }
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
#else
#include <cuda_runtime.h>
#endif

namespace jg {

// Work counters of the INSTRUMENTED build (-DJG_COUNT_WORK, libjogramop_b200_count.so): executed support dot products,
// GJK iterations, polytope expansions.  bench.py uses them to audit the nominal flops-per-check figure of the roofline
// (SURVEY.md §8d); the shipped library compiles them away.
enum { JG_WORK_DOTS = 0, JG_WORK_GJK_ITERS = 1, JG_WORK_EPA_ITERS = 2, JG_WORK_TRI_TESTS = 3, JG_WORK_N = 4 };
#if defined(JG_COUNT_WORK) && !defined(JG_HOST_EMU)
__device__ unsigned long long g_jg_work[JG_WORK_N];
__device__ __forceinline__ void work_add(int slot, unsigned v) {
    const unsigned m = __activemask();
    const unsigned sum = __reduce_add_sync(m, v);
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&g_jg_work[slot], (unsigned long long)sum);
}
#define JG_WORK(slot, v) work_add(slot, v)
#else
#define JG_WORK(slot, v) ((void)0)
#endif

struct V3 {
    float x, y, z;
};
__host__ __device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float norm2(V3 a) { return dot(a, a); }

// rigid transform, rotation row-major
struct Tf {
    float r00, r01, r02, r10, r11, r12, r20, r21, r22, tx, ty, tz;
};

__device__ __forceinline__ Tf tf_identity() {
    Tf T;
    T.r00 = 1.f; T.r01 = 0.f; T.r02 = 0.f;
    T.r10 = 0.f; T.r11 = 1.f; T.r12 = 0.f;
    T.r20 = 0.f; T.r21 = 0.f; T.r22 = 1.f;
    T.tx = T.ty = T.tz = 0.f;
    return T;
}

__device__ __forceinline__ Tf tf_mul(const Tf& A, const Tf& B) {
    Tf C;
    C.r00 = fmaf(A.r00, B.r00, fmaf(A.r01, B.r10, A.r02 * B.r20));
    C.r01 = fmaf(A.r00, B.r01, fmaf(A.r01, B.r11, A.r02 * B.r21));
    C.r02 = fmaf(A.r00, B.r02, fmaf(A.r01, B.r12, A.r02 * B.r22));
    C.r10 = fmaf(A.r10, B.r00, fmaf(A.r11, B.r10, A.r12 * B.r20));
    C.r11 = fmaf(A.r10, B.r01, fmaf(A.r11, B.r11, A.r12 * B.r21));
    C.r12 = fmaf(A.r10, B.r02, fmaf(A.r11, B.r12, A.r12 * B.r22));
    C.r20 = fmaf(A.r20, B.r00, fmaf(A.r21, B.r10, A.r22 * B.r20));
    C.r21 = fmaf(A.r20, B.r01, fmaf(A.r21, B.r11, A.r22 * B.r21));
    C.r22 = fmaf(A.r20, B.r02, fmaf(A.r21, B.r12, A.r22 * B.r22));
    C.tx = fmaf(A.r00, B.tx, fmaf(A.r01, B.ty, fmaf(A.r02, B.tz, A.tx)));
    C.ty = fmaf(A.r10, B.tx, fmaf(A.r11, B.ty, fmaf(A.r12, B.tz, A.ty)));
    C.tz = fmaf(A.r20, B.tx, fmaf(A.r21, B.ty, fmaf(A.r22, B.tz, A.tz)));
    return C;
}

// inv(A) * B for rigid transforms
__device__ __forceinline__ Tf tf_inv_mul(const Tf& A, const Tf& B) {
    Tf C;
    C.r00 = fmaf(A.r00, B.r00, fmaf(A.r10, B.r10, A.r20 * B.r20));
    C.r01 = fmaf(A.r00, B.r01, fmaf(A.r10, B.r11, A.r20 * B.r21));
    C.r02 = fmaf(A.r00, B.r02, fmaf(A.r10, B.r12, A.r20 * B.r22));
    C.r10 = fmaf(A.r01, B.r00, fmaf(A.r11, B.r10, A.r21 * B.r20));
    C.r11 = fmaf(A.r01, B.r01, fmaf(A.r11, B.r11, A.r21 * B.r21));
    C.r12 = fmaf(A.r01, B.r02, fmaf(A.r11, B.r12, A.r21 * B.r22));
    C.r20 = fmaf(A.r02, B.r00, fmaf(A.r12, B.r10, A.r22 * B.r20));
    C.r21 = fmaf(A.r02, B.r01, fmaf(A.r12, B.r11, A.r22 * B.r21));
    C.r22 = fmaf(A.r02, B.r02, fmaf(A.r12, B.r12, A.r22 * B.r22));
    float dx = B.tx - A.tx, dy = B.ty - A.ty, dz = B.tz - A.tz;
    C.tx = fmaf(A.r00, dx, fmaf(A.r10, dy, A.r20 * dz));
    C.ty = fmaf(A.r01, dx, fmaf(A.r11, dy, A.r21 * dz));
    C.tz = fmaf(A.r02, dx, fmaf(A.r12, dy, A.r22 * dz));
    return C;
}

__device__ __forceinline__ V3 tf_apply(const Tf& T, V3 p) {
    return mk(fmaf(T.r00, p.x, fmaf(T.r01, p.y, fmaf(T.r02, p.z, T.tx))),
              fmaf(T.r10, p.x, fmaf(T.r11, p.y, fmaf(T.r12, p.z, T.ty))),
              fmaf(T.r20, p.x, fmaf(T.r21, p.y, fmaf(T.r22, p.z, T.tz))));
}
// R^T d
__device__ __forceinline__ V3 tf_rot_t(const Tf& T, V3 d) {
    return mk(fmaf(T.r00, d.x, fmaf(T.r10, d.y, T.r20 * d.z)), fmaf(T.r01, d.x, fmaf(T.r11, d.y, T.r21 * d.z)),
              fmaf(T.r02, d.x, fmaf(T.r12, d.y, T.r22 * d.z)));
}

// order-preserving float <-> int map so that atomicMin(int) is a float min
__host__ __device__ __forceinline__ int enc_float(float f) {
#ifdef __CUDA_ARCH__
    int i = __float_as_int(f);
#else
    union { float f; int i; } u; u.f = f; int i = u.i;
#endif
    return i >= 0 ? i : (i ^ 0x7FFFFFFF);
}
__host__ __device__ __forceinline__ float dec_float(int i) {
    i = i >= 0 ? i : (i ^ 0x7FFFFFFF);
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    union { float f; int i; } u; u.i = i; return u.f;
#endif
}
#define JG_ENC_INF 0x7F800000

}  // namespace jg
